// b2_sigma.h — the SigmaPlan: a flat, per-site list of contraction terms that reproduces Heff::makeHeff.
//
// Reference: Heff.cpp:43-248 calls ~70 addDiagram* functions per target block and per matvec; each re-derives the
// neighbour block, the Wigner prefactor and the operator block and issues 1-2 dgemm_.  Here that derivation is done
// ONCE per site: every (target block, source block, left operator block, right operator block, factor) tuple becomes
// one SigmaTerm.  The device kernels (b2_kernels.cu) execute the list for every Davidson matvec.
//
//   sigma[dst] += factor * opL(A) * S[src] * opR(B)           (A or B may be absent = identity)
//
// opL(A): A is a block of a renormalized operator of the left boundary;  ltrans=1 means the block is stored as
// (src-left-sector -> dst-left-sector) and enters transposed.  opR(B): block of a right-boundary operator; rtrans=0
// means stored as (src-right-sector -> dst-right-sector) and enters as is, rtrans=1 stored the other way (enters ^T).
#pragma once
#include <string>

#include "b2_ops.h"

namespace b2 {

enum OpSrc : int8_t { SRC_NONE = 0, SRC_LEFT = 1, SRC_RIGHT = 2, SRC_PRESUM = 3 };

struct BRef {          // one operator block used by a term
   int8_t src = SRC_NONE;
   int8_t trans = 0;
   int op = -1;        // index in the left/right OpSet, or presum index
   int blk = -1;       // block index in that operator's OpLayout
};

struct SigmaTerm {
   int dst = -1, src = -1;
   BRef l, r;
   int owner = 0;      // GPU that owns this term under the reference's ownership maps (MPIchemps2.h:158-231)
   double factor = 0.0;
};

// Integral-weighted operator pre-sum  O~ = sum_l coef_l * O_l  (HeffDiagrams3.cpp:64-75 etc).  The operators are
// fixed during a Davidson solve, so it is materialised once per site instead of once per term per matvec.
struct Presum {
   int side = SRC_LEFT;                          // which OpSet the parts come from
   std::shared_ptr<const OpLayout> lay;
   std::vector<std::pair<double, int>> parts;    // (coefficient, op index in that OpSet)
   int64_t off = 0;                              // offset in the plan's presum arena
};

struct SigmaPlan {
   int site = 0;
   bool at_left = false, at_right = false;
   SLayout S;
   BigVec<SigmaTerm> terms;        // ~40 bytes x up to millions per site: blocks recycled through the host block cache
   std::vector<Presum> presums;
   int64_t presum_size = 0;
   long long skipped_zero = 0;                   // terms dropped because their prefactor is exactly 0
   // algorithmic FLOPs exactly as SURVEY.md 8(d) defines them: 2mnk per reference dgemm_, 2n per daxpy_
   double flops_ref = 0.0;
};

int plan_threads(int nblocks);   // host threads for plan building (B2_PLAN_THREADS)
void set_plan_local_ranks(int n);   // processes sharing this host (one per GPU): the host threads are divided among them

// world = number of GPUs the ownership maps are evaluated for (1 = everything owned by GPU 0)
void build_sigma_plan(SigmaPlan& plan, const Bookkeeper& bk, const Problem& prob, const OpSet* left, const OpSet* right,
                      int site, int world);

}   // namespace b2
