// b2_update.h — the UpdatePlan: every renormalized operator of the NEW boundary expressed as a flat list of
// three-factor contractions over the OLD boundary's operators and the MPS site tensor T that was just optimised.
//
// Replaces DMRG::updateMovingRight / updateMovingLeft (DMRGoperators.cpp:243-907) and the per-tensor algebra they call
// (TensorOperator::update, TensorL::create, TensorS0/S1/F0/F1::makenew, TensorQ::AddTerm*, TensorX::update): the
// reference walks tensor by tensor, block by block, re-deriving sectors and Wigner factors and issuing two dgemm_ per
// (block, case).  Here the same terms are enumerated once into
//      new_block += f * op(T_up) * [old block | pre-summed old blocks | identity] * op(T_down)
// and executed by the grouped DMMA kernels, batched over ALL operators of the boundary (they share T).
//
// moving right (T = MPS[index], old boundary index, new boundary index+1):  op(T_up) = T_up^T, op(T_down) = T_down
// moving left  (T = MPS[index], old boundary index+1, new boundary index):  op(T_up) = T_up,   op(T_down) = T_down^T
#pragma once
#include "b2_compile.h"
#include "b2_ops.h"
#include "b2_sigma.h"

namespace b2 {

struct UpdatePlan {
   int index = 0;
   bool moving_right = true;
   TLayout T;
   std::vector<Term3> terms;          // dst = global block id over the new OpSet (see block_base)
   std::vector<DstBlock> dst;         // offsets in the new arena
   std::vector<int> block_base;       // first global block id of new operator i
   std::vector<Presum> presums;       // linear combinations of OLD operators (side = SRC_LEFT means "old arena")
   int64_t presum_size = 0;
   // Mixing (reads the NEW arena): A/B/C/D += integral-weighted S0/S1/F0/F1 that have a leg on the new site
   // (DMRGoperators.cpp:367-405 / :700-738).  TensorOperator::daxpy (:407-414) adds operators of IDENTICAL layout: whole-operator
   // axpys `mix_flat` (dst op += coef * src op, element by element).  daxpy_transpose_tensorCD (:416-455) adds transposed blocks with
   // a spin factor that depends on the block only: the transposed, factor-scaled copy of a source operator in the destination layout
   // (`mix_temps`, kept behind the pre-sums in the pre-sum arena) is built ONCE by the block terms `mix_terms` (destination blocks
   // `mix_dst`, space SP_PRESUM) and then enters `mix_flat` like a plain source.  The device runs mix_flat as one tall-skinny
   // GEMM-like kernel per layout: Dst[elem, pair] += Src[elem, partner] * Coef[partner, pair].
   struct MixFlat { int dst_op; int src_op; int temp; double coef; };   // temp >= 0: source = mix_temps[temp] instead of new op src_op
   struct MixTemp { int src_op; const OpLayout* lay; int64_t off; int64_t size; };
   std::vector<MixFlat> mix_flat;
   std::vector<MixTemp> mix_temps;
   std::vector<Term3> mix_terms;
   std::vector<DstBlock> mix_dst;
   double flops_ref = 0.0;            // 2mnk per reference dgemm_ (F_upd)
};

// old_set may be null at the chain ends (index == 0 moving right / index == L-1 moving left).
void build_update_plan(UpdatePlan& plan, const Bookkeeper& bk, const Problem& prob, const OpSet* old_set, const OpSet& new_set,
                       int index, bool moving_right);

}   // namespace b2
