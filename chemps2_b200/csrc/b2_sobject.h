// b2_sobject.h — two-site object: Join (device contraction terms) and Split (host recoupling/truncation + batched device SVD).  See b2_sobject.cpp.
#pragma once
#include <vector>

#include <functional>

#include "b2_compile.h"
#include "b2_core.h"
#include "b2_svd.h"

namespace b2 {

// S[kappa] = sum_jM f * T_left[L -> M] * T_right[M -> R]; spaces: SP_LEFT = T_left storage, SP_RIGHT = T_right storage, dst = S
void join_terms(std::vector<Term3>& terms, std::vector<DstBlock>& dst, const Bookkeeper& bk, const SLayout& S, const TLayout& TL, const TLayout& TR);

// Sobject::Split. s_storage: S in program convention (layout S, dims of bk at entry).  On exit bk holds the new dimensions of
// boundary ix+1 (when change) and t_left / t_right the new site tensors in the new layouts.  Returns the discarded weight.
// The recoupling into centre-sector matrices, the global truncation rule (Sobject.cpp:451-486) and the scatter into the new site
// tensors are index work on the host; the decompositions go to `svd_batch` (dev_svd_batch on the GPU: there is no host SVD).
using SvdBatchFn = std::function<int(std::vector<SvdJob>&)>;
double split_host(Bookkeeper& bk, int ix, const SLayout& S, const double* s_storage, int D, bool moving_right, bool change,
                  std::vector<double>& t_left, std::vector<double>& t_right, const SvdBatchFn& svd_batch);
void left_normalize_host(const Bookkeeper& bk, const TLayout& T, double* t);

}   // namespace b2
