// b2_sobject.h — two-site object: Join (device contraction terms) and Split (host SVD + truncation).  See b2_sobject.cpp.
#pragma once
#include <vector>

#include "b2_compile.h"
#include "b2_core.h"

namespace b2 {

// S[kappa] = sum_jM f * T_left[L -> M] * T_right[M -> R]; spaces: SP_LEFT = T_left storage, SP_RIGHT = T_right storage, dst = S
void join_terms(std::vector<Term3>& terms, std::vector<DstBlock>& dst, const Bookkeeper& bk, const SLayout& S, const TLayout& TL, const TLayout& TR);

// Sobject::Split. s_storage: S in program convention (layout S, dims of bk at entry).  On exit bk holds the new dimensions of
// boundary ix+1 (when change) and t_left / t_right the new site tensors in the new layouts.  Returns the discarded weight.
double split_host(Bookkeeper& bk, int ix, const SLayout& S, const double* s_storage, int D, bool moving_right, bool change,
                  std::vector<double>& t_left, std::vector<double>& t_right);

void jacobi_svd(int m, int n, const double* a, double* s, double* u, double* vt);
void left_normalize_host(const Bookkeeper& bk, const TLayout& T, double* t);

}   // namespace b2
