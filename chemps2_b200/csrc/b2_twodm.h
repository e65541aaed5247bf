// b2_twodm.h — plan of the site contribution to the 2-RDM (TwoDM::FillSite); see b2_twodm.cpp.
#pragma once
#include <memory>
#include <vector>

#include "b2_compile.h"
#include "b2_core.h"
#include "b2_ops.h"

namespace b2 {

struct TwoDMPlan {
   struct MOp {                      // one effective operator  M = sum f * op(T) [L_g] op(T)
      int tag = 0, g = -1, kind = 0, irrep = 0;
      bool left_side = false;        // lives on the left boundary (diagram 7) instead of the right one
      std::shared_ptr<const OpLayout> lay;
      int64_t off = 0;               // in the M arena
      int group = 0, col = 0;
   };
   struct Group {                    // all M of one (side, kind, irrep): a dense [stride x members] matrix in the M arena
      bool left_side = false;
      int kind = 0, irrep = 0;
      int64_t off = 0, stride = 0;
      std::vector<int> members;      // indices into mops
      std::vector<int> partner_kinds;   // kinds of stored operators this group is paired with (empty: just `kind`)
      std::vector<int> partners;     // operator indices in the left / right OpSet with a matching kind and irrep: the Gram columns
   };
   int site = 0;
   TLayout T;
   std::vector<MOp> mops;
   std::vector<Group> groups;
   int64_t m_size = 0;
   std::vector<Term3> terms;         // spaces: SP_RIGHT = T, SP_LEFT = left operator arena, dst = M arena
   std::vector<DstBlock> dst;
   std::vector<int> block_base;
   std::vector<double> d1_scale;     // per T block: (2SL+1) for the doubly-occupied blocks, else 0 (doD1)
};

void build_twodm_plan(TwoDMPlan& plan, const Bookkeeper& bk, int site, const OpSet* left, const OpSet* right);

// Correlations::FillSite (Correlations.cpp:212-351): the five diagram functions (:353-560) between the MPS tensor of `site` and the
// G / Y / Z / K / M tensors of every previous site, as Gram matrices of the effective operators N = f * T_a T_b^T (left boundary) with
// the tensors of the correlation operator set `corr` (boundary `site`).  Same plan structure as the 2-RDM (mops tags CORR_D1..D5).
enum { CORR_D1 = 100, CORR_D2, CORR_D3, CORR_D4, CORR_D5 };
void build_corr_plan(TwoDMPlan& plan, const Bookkeeper& bk, int site, const OpSet& corr);
// < N(tag) , tensor (kind, p) of the correlation set >
double corr_value(const TwoDMPlan& plan, const OpSet& corr, const std::vector<std::vector<double>>& gram, int tag, int kind, int p);
// gram[group][member + members * partner] = < M_member , partner operator >
void twodm_scatter(const TwoDMPlan& plan, const Bookkeeper& bk, const OpSet* left, const OpSet* right, double d1,
                   const std::vector<std::vector<double>>& gram, double* A, double* B);

}   // namespace b2
