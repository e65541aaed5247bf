// b2_davidson.cpp — host control flow of the device Davidson solver.  See b2_davidson.h.
//
// The algorithm is the reference's (Davidson.cpp): modified Gram-Schmidt against the current basis (:214-222), one
// matrix-vector product per new basis vector, Rayleigh-Ritz in the basis (:242-320), diagonally preconditioned residual
// with the Olsen projection (:322-350), and, when 32 vectors are reached, deflation to the 3 lowest Ritz vectors
// re-orthonormalised with a Loewdin transform (:352-412) followed by fresh products for the kept vectors (:510-536).
#include "b2_davidson.h"

#include <cuda_runtime.h>

#include "b2_pool.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace b2 {

void small_symmetric_eig(int n, const double* a, int lda, double* eval, double* evec) {
   std::vector<double> A((size_t)n * n);
   for (int j = 0; j < n; j++)
      for (int i = 0; i < n; i++) A[i + (size_t)n * j] = 0.5 * (a[i + (size_t)lda * j] + a[j + (size_t)lda * i]);
   std::vector<double> V((size_t)n * n, 0.0);
   for (int i = 0; i < n; i++) V[i + (size_t)n * i] = 1.0;
   for (int sweep = 0; sweep < 100; sweep++) {
      double off = 0.0, dia = 0.0;
      for (int j = 0; j < n; j++)
         for (int i = 0; i < n; i++) (i == j ? dia : off) += A[i + (size_t)n * j] * A[i + (size_t)n * j];
      if (off <= 1e-32 * (dia + 1e-300)) break;
      for (int p = 0; p < n - 1; p++)
         for (int q = p + 1; q < n; q++) {
            const double apq = A[p + (size_t)n * q];
            if (apq == 0.0) continue;
            const double app = A[p + (size_t)n * p], aqq = A[q + (size_t)n * q];
            const double tau = (aqq - app) / (2.0 * apq);
            const double t = (tau >= 0.0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
            const double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
            for (int k = 0; k < n; k++) {   // A <- A * J
               const double akp = A[k + (size_t)n * p], akq = A[k + (size_t)n * q];
               A[k + (size_t)n * p] = c * akp - s * akq;
               A[k + (size_t)n * q] = s * akp + c * akq;
            }
            for (int k = 0; k < n; k++) {   // A <- J^T * A
               const double apk = A[p + (size_t)n * k], aqk = A[q + (size_t)n * k];
               A[p + (size_t)n * k] = c * apk - s * aqk;
               A[q + (size_t)n * k] = s * apk + c * aqk;
            }
            for (int k = 0; k < n; k++) {
               const double vkp = V[k + (size_t)n * p], vkq = V[k + (size_t)n * q];
               V[k + (size_t)n * p] = c * vkp - s * vkq;
               V[k + (size_t)n * q] = s * vkp + c * vkq;
            }
         }
   }
   std::vector<int> idx(n);
   for (int i = 0; i < n; i++) idx[i] = i;
   std::sort(idx.begin(), idx.end(), [&](int x, int y) { return A[x + (size_t)n * x] < A[y + (size_t)n * y]; });
   for (int j = 0; j < n; j++) {
      eval[j] = A[idx[j] + (size_t)n * idx[j]];
      for (int i = 0; i < n; i++) evec[i + (size_t)n * j] = V[i + (size_t)n * idx[j]];
   }
}

namespace {
struct Fail { int code; };
}

int davidson_solve(void* stream_v, int64_t n, const MatVec& matvec, double* x_dev, const double* diag_dev, const DavidsonParams& prm,
                   double* eigenvalue, int* n_matvec, char* errbuf, int errlen) {
   cudaStream_t s = (cudaStream_t)stream_v;
   const int MAXV = std::min(prm.max_vec, kMaxVec), KEEP = std::min(prm.keep_vec, MAXV - 1);
   const int64_t stride = (n + 15) / 16 * 16;
   double *slab = nullptr, *scal = nullptr, *h_scal = nullptr;
   auto cleanup = [&]() { cudaFree(slab); cudaFree(scal); if (h_scal) cudaFreeHost(h_scal); };
   auto fail = [&](const char* what, cudaError_t e) {
      snprintf(errbuf, errlen, "davidson: %s: %s", what, cudaGetErrorString(e));
      cleanup();
      return -3;
   };
   // slab: V[MAXV] | HV[MAXV] | t | u | work | E[KEEP]
   const int64_t nvecs = 2 * MAXV + 3 + KEEP;
   cudaError_t e = cudaMalloc(&slab, sizeof(double) * (size_t)stride * nvecs);
   if (e != cudaSuccess) return fail("cudaMalloc(vectors)", e);
   e = cudaMalloc(&scal, sizeof(double) * (64 + kRedScratch));
   if (e != cudaSuccess) return fail("cudaMalloc(scalars)", e);
   e = cudaMallocHost(&h_scal, sizeof(double) * 64);
   if (e != cudaSuccess) return fail("cudaMallocHost", e);
   e = cudaMemsetAsync(scal, 0, sizeof(double) * (64 + kRedScratch), s);
   if (e != cudaSuccess) return fail("memset", e);
   double* V = slab;
   double* HV = slab + stride * MAXV;
   double* t = slab + stride * 2 * MAXV;
   double* u = t + stride;
   double* work = u + stride;
   double* E = work + stride;
   double* scratch = scal + 64;

#define DV_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(#call, e_); } while (0)
#define DV_DEV(call) do { if ((call) != 0) { snprintf(errbuf, errlen, "davidson: kernel launch failed in %s", #call); cleanup(); return -3; } } while (0)
   auto fetch = [&](int count) -> cudaError_t {   // device scalars [0, count) -> pinned host, synchronous
      cudaError_t e2 = cudaMemcpyAsync(h_scal, scal, sizeof(double) * count, cudaMemcpyDeviceToHost, s);
      if (e2 != cudaSuccess) return e2;
      return cudaStreamSynchronize(s);
   };

   DV_CUDA(cudaMemcpyAsync(t, x_dev, sizeof(double) * n, cudaMemcpyDeviceToDevice, s));
   // SafetyCheckGuess (Davidson.cpp:193-204): a zero guess is replaced by rand() numbers
   DV_DEV(dev_multi_dot(t, t, stride, 1, n, scal, scratch, s));
   DV_CUDA(fetch(1));
   if (h_scal[0] == 0.0) {
      std::vector<double> r((size_t)n);
      for (int64_t i = 0; i < n; i++) r[i] = ((double)rand()) / RAND_MAX;
      DV_CUDA(cudaMemcpyAsync(t, r.data(), sizeof(double) * n, cudaMemcpyHostToDevice, s));
      DV_CUDA(cudaStreamSynchronize(s));
   }

   int num_vec = 0, nmult = 0;
   std::vector<double> mxM((size_t)MAXV * MAXV, 0.0), evec((size_t)MAXV * MAXV), eval(MAXV), sub((size_t)MAXV * MAXV);

   // Orthogonalise t against V[0..num_vec), normalise, V[num_vec] = t.  The reference does modified Gram-Schmidt with one ddot_ + one
   // daxpy_ per basis vector (Davidson.cpp:214-222) = 2 num_vec kernel launches here; classical Gram-Schmidt applied TWICE ("twice is
   // enough": the second pass removes what rounding left behind) spans the same space to working precision with four launches — one
   // fused multi-dot and one fused multi-axpy per pass, each streaming the basis once.
   auto add_new_vec = [&]() -> int {
      for (int pass = 0; pass < 2 && num_vec > 0; pass++) {
         if (dev_multi_dot(t, V, stride, num_vec, n, scal, scratch, s)) return -1;
         if (dev_multi_axpy_dev(t, V, stride, num_vec, scal, -1.0, n, s)) return -1;
      }
      if (dev_multi_dot(t, t, stride, 1, n, scal, scratch, s)) return -1;
      if (dev_scale_rsqrt(t, scal, n, s)) return -1;
      if (cudaMemcpyAsync(V + stride * num_vec, t, sizeof(double) * n, cudaMemcpyDeviceToDevice, s) != cudaSuccess) return -1;
      return 0;
   };
   auto multiply = [&](int idx) -> int { nmult++; return matvec(V + stride * idx, HV + stride * idx); };

   DV_DEV(add_new_vec());
   DV_DEV(multiply(num_vec));
   double theta = 0.0;
   while (true) {
      // ---- state 'N': new column of the projected matrix (Davidson.cpp:246-253)
      DV_DEV(dev_multi_dot(V + stride * num_vec, HV, stride, num_vec + 1, n, scal, scratch, s));
      DV_CUDA(fetch(num_vec + 1));
      for (int c = 0; c <= num_vec; c++) mxM[c + (size_t)MAXV * num_vec] = mxM[num_vec + (size_t)MAXV * c] = h_scal[c];
      num_vec++;
      for (int j = 0; j < num_vec; j++)
         for (int i = 0; i < num_vec; i++) sub[i + (size_t)num_vec * j] = mxM[i + (size_t)MAXV * j];
      small_symmetric_eig(num_vec, sub.data(), num_vec, eval.data(), evec.data());
      theta = eval[0];
      Coefs a{};
      for (int j = 0; j < num_vec; j++) a.c[j] = evec[j];
      DV_DEV(dev_ritz_residual(u, t, V, HV, stride, num_vec, a, theta, n, scal, scratch, s));
      DV_CUDA(fetch(1));
      const double rnorm = std::sqrt(h_scal[0]);
      if (!(rnorm > prm.rtol)) break;   // converged (Davidson.cpp:141,160)
      if (nmult > prm.max_matvec) {
         snprintf(errbuf, errlen, "davidson: no convergence after %d matrix-vector products (residual %.3e, rtol %.3e)", nmult, rnorm, prm.rtol);
         cleanup();
         return -4;
      }
      // ---- CalculateNewVec (Davidson.cpp:322-350)
      DV_DEV(dev_precond_dots(work, u, t, diag_dev, theta, prm.cutoff, n, scal, scratch, s));
      DV_DEV(dev_precond_apply(t, u, diag_dev, scal, theta, prm.cutoff, n, s));
      if (num_vec == MAXV) {
         // ---- Deflation (Davidson.cpp:352-412): keep the KEEP lowest Ritz vectors, Loewdin-orthonormalised
         if (KEEP <= 1) {
            DV_DEV(dev_multi_dot(u, u, stride, 1, n, scal, scratch, s));
            DV_DEV(dev_scale_rsqrt(u, scal, n, s));
            DV_CUDA(cudaMemcpyAsync(V, u, sizeof(double) * n, cudaMemcpyDeviceToDevice, s));
         } else {
            DV_CUDA(cudaMemcpyAsync(E, u, sizeof(double) * n, cudaMemcpyDeviceToDevice, s));
            for (int c = 1; c < KEEP; c++) {
               Coefs b{};
               for (int j = 0; j < MAXV; j++) b.c[j] = evec[j + (size_t)num_vec * c];
               DV_DEV(dev_lincomb(E + stride * c, V, stride, MAXV, b, n, s));
            }
            std::vector<double> ov((size_t)KEEP * KEEP), oval(KEEP), ovec((size_t)KEEP * KEEP), low((size_t)KEEP * KEEP, 0.0);
            for (int c = 0; c < KEEP; c++) {
               DV_DEV(dev_multi_dot(E + stride * c, E, stride, KEEP, n, scal, scratch, s));
               DV_CUDA(fetch(KEEP));
               for (int r = 0; r < KEEP; r++) ov[r + (size_t)KEEP * c] = h_scal[r];
            }
            small_symmetric_eig(KEEP, ov.data(), KEEP, oval.data(), ovec.data());
            for (int k = 0; k < KEEP; k++) {
               const double w = std::pow(oval[k], -0.5);
               for (int j = 0; j < KEEP; j++)
                  for (int i = 0; i < KEEP; i++) low[i + (size_t)KEEP * j] += ovec[i + (size_t)KEEP * k] * w * ovec[j + (size_t)KEEP * k];
            }
            for (int iv = 0; iv < KEEP; iv++) {
               Coefs b{};
               for (int j = 0; j < KEEP; j++) b.c[j] = low[j + (size_t)KEEP * iv];
               DV_DEV(dev_lincomb(V + stride * iv, E, stride, KEEP, b, n, s));
            }
         }
         // ---- state 'F': fresh products for the kept vectors, then their projected matrix (Davidson.cpp:166-186,510-520)
         const int kept = std::max(KEEP, 1);
         for (int c = 0; c < kept; c++) DV_DEV(multiply(c));
         for (int c = 0; c < kept; c++) {
            DV_DEV(dev_multi_dot(V + stride * c, HV, stride, kept, n, scal, scratch, s));
            DV_CUDA(fetch(kept));
            for (int r = 0; r < kept; r++) mxM[c + (size_t)MAXV * r] = h_scal[r];
         }
         for (int c = 0; c < kept; c++)
            for (int r = c + 1; r < kept; r++) mxM[r + (size_t)MAXV * c] = mxM[c + (size_t)MAXV * r];
         num_vec = kept;
      }
      DV_DEV(add_new_vec());
      DV_DEV(multiply(num_vec));
   }
   DV_CUDA(cudaMemcpyAsync(x_dev, u, sizeof(double) * n, cudaMemcpyDeviceToDevice, s));
   DV_CUDA(cudaStreamSynchronize(s));
   *eigenvalue = theta;
   *n_matvec = nmult;
   cleanup();
   return 0;
#undef DV_CUDA
#undef DV_DEV
}

}   // namespace b2
