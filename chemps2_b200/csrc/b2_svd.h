// b2_svd.h — batched device SVD used by Split (b2_svd.cu).
#pragma once
#include <vector>

namespace b2 {

// thin SVD of the column-major m x n matrix a (ld = m): a = U diag(s) V^T, k = min(m, n); u is m x k (ld m), vt is k x n (ld k);
// singular values in decreasing order.  All pointers are HOST memory owned by the caller.
struct SvdJob {
   int m = 0, n = 0;
   const double* a = nullptr;
   double *s = nullptr, *u = nullptr, *vt = nullptr;
};

// decomposes all jobs together on the GPU (one-sided Jacobi, round-robin pair order); 0 on success, message in err otherwise
int dev_svd_batch(std::vector<SvdJob>& jobs, void* stream, char* err, int errlen);

}   // namespace b2
