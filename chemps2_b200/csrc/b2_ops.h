// b2_ops.h — the set of renormalized operators living at one boundary (host metadata only).
//
// Replaces the reference's per-boundary pointer tables Ltensors[b][k], S0tensors[b][c2][c3], Atensors[b][c2][c3],
// Qtensors[b][c2], Xtensors[b] (DMRG.h:211-226, index conventions DMRGoperators.cpp:909-1140) by ONE arena per
// (boundary, direction) addressed by (kind, site_i, site_j).  Every operator keeps the reference's packed block
// layout (OpLayout == TensorOperator.cpp:29-102) so host copies are interchangeable with gStorage().
#pragma once
#include <map>
#include <memory>

#include "b2_core.h"

namespace b2 {

enum OpKind { K_L = 0, K_S0, K_S1, K_F0, K_F1, K_A, K_B, K_C, K_D, K_Q, K_X, K_G, K_Y, K_Z, K_K, K_M, K_NKINDS };

// (two_j, n_elec) per kind: TensorL.cpp:29-37, TensorS0.cpp:27-35, TensorS1.cpp:28-36, TensorF0.cpp:27-35,
// TensorF1.cpp:28-36, DMRGoperators.cpp:977-987 (A,B,C,D), TensorQ.cpp:28-36, TensorX.cpp:28-36
// G, Y, Z (TensorGYZ.cpp:26-36: two_j 0, n_elec 0, irrep 0) and K, M (TensorKM.cpp:26-36: two_j 1, n_elec 1, irrep of the site)
// are the helper operators of the two-orbital correlation functions; they never carry a Jordan-Wigner phase.
inline int kind_two_j(int k) { static const int t[K_NKINDS] = {1, 0, 2, 0, 2, 0, 2, 0, 2, 1, 0, 0, 0, 0, 1, 1}; return t[k]; }
inline int kind_nelec(int k) { static const int t[K_NKINDS] = {1, 2, 2, 0, 0, 2, 2, 0, 0, 1, 0, 0, 0, 0, 1, 1}; return t[k]; }
inline bool kind_jw(int k) { return k == K_L || k == K_Q; }
inline const char* kind_name(int k) {
   static const char* n[K_NKINDS] = {"L", "S0", "S1", "F0", "F1", "A", "B", "C", "D", "Q", "X", "G", "Y", "Z", "K", "M"};
   return n[k];
}

struct OpTensor {
   int kind = 0, i = -1, j = -1;             // site indices (i <= j); i == j for L, Q; -1 for X
   int irrep = 0;
   bool prime_last = true;                   // F1 and D: prime_last = moving_right (TensorF1.cpp:33, DMRGoperators.cpp:987)
   std::shared_ptr<const OpLayout> lay;      // shared between operators with equal quantum numbers
   int64_t off = 0;                          // offset (doubles) inside the OpSet arena, 16-double aligned
};

struct OpSet {
   int boundary = 0;
   bool moving_right = true;                 // true: covers sites < boundary; false: covers sites >= boundary
   std::vector<OpTensor> ops;
   std::map<int64_t, int> index;
   std::map<int64_t, std::shared_ptr<const OpLayout>> layouts;
   int64_t size = 0;                         // arena length in doubles

   static int64_t key(int kind, int i, int j) { return ((int64_t)kind << 40) | ((int64_t)(i + 1) << 20) | (int64_t)(j + 1); }
   int find(int kind, int i, int j) const { auto it = index.find(key(kind, i, j)); return it == index.end() ? -1 : it->second; }
   std::shared_ptr<const OpLayout> layout(const Bookkeeper& bk, int two_j, int n_elec, int irrep);
   int add(const Bookkeeper& bk, int kind, int i, int j);
   // allocate the full complement the reference keeps at this boundary (DMRGoperators.cpp:909-1140)
   void build_all(const Bookkeeper& bk, int boundary, bool moving_right);
   // the G, Y, Z, K, M tensors of every site left of `boundary` (DMRG::update_correlations_tensors, DMRGoperators3RDM.cpp:415-479)
   void build_correlation(const Bookkeeper& bk, int boundary);
   // reduced complement for the 2-RDM chain (DMRG::updateMovingLeftSafe2DM keeps L, S0, S1, F0, F1 only; the left blocks need L only)
   void build_reduced(const Bookkeeper& bk, int boundary, bool moving_right, bool only_L);
   bool reduced = false;                     // not a full sweep set: b2_heff_create refuses it
};

}   // namespace b2
