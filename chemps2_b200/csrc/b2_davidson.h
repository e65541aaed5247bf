// b2_davidson.h — Davidson eigensolver with device-resident vectors (replaces CheMPS2::Davidson for the DMRG path,
// Davidson.cpp:31-538, as driven by Heff::SolveDAVIDSON, Heff.cpp:317-386).
#pragma once
#include <cstdint>
#include <functional>

#include "b2_device.h"

namespace b2 {

struct DavidsonParams {
   int max_vec = 32;          // DAVIDSON_NUM_VEC       (Options.h:70)
   int keep_vec = 3;          // DAVIDSON_NUM_VEC_KEEP  (Options.h:71)
   double rtol = 1e-5;        // per-instruction residual tolerance
   double cutoff = 1e-12;     // DAVIDSON_PRECOND_CUTOFF (Options.h:72)
   int max_matvec = 5000;     // safety net (the reference loops until convergence)
};

// out_dev = H * in_dev, asynchronous on `stream`; returns 0 on success
typedef std::function<int(const double* in_dev, double* out_dev)> MatVec;

// x_dev: initial guess on entry, lowest eigenvector (unit norm) on exit.  diag_dev: diagonal of H (preconditioner).
// Returns 0 on success; *eigenvalue and *n_matvec are filled.  All vectors live on the device; per iteration the host
// only sees one column of the projected matrix and the residual norm.
int davidson_solve(void* stream, int64_t n, const MatVec& matvec, double* x_dev, const double* diag_dev, const DavidsonParams& prm,
                   double* eigenvalue, int* n_matvec, char* errbuf, int errlen);

// symmetric eigenproblem of a small dense matrix (n <= 32): cyclic Jacobi; eigenvalues ascending, eigenvectors in the
// columns of evec (column-major, ld = n).  Replaces the dsyev_ calls of Davidson.cpp:276,383.
void small_symmetric_eig(int n, const double* a, int lda, double* eval, double* evec);

}   // namespace b2
