// b2_svd.cu — batched one-sided Jacobi (Hestenes) SVD on the device: the decomposition step of Sobject::Split
// (Sobject.cpp:412-419 calls dgesdd_ once per centre sector on the host).
//
// All centre-sector matrices of one Split are decomposed together.  For every matrix the columns of the taller orientation
// W (R x C, R >= C) are orthogonalised by plane rotations, V (C x C, starts as identity) accumulates them:  A = W V^T with
// orthogonal columns of W at convergence, so sigma_j = |W(:,j)|, u_j = W(:,j)/sigma_j.  The C(C-1)/2 column pairs of a sweep
// are visited in the round-robin ("chess tournament") order: C-1 steps of C/2 DISJOINT pairs, one CTA per pair, one kernel
// launch per step over all matrices of the batch.  Everything a CTA does is a fixed-order reduction followed by an
// element-wise rotation, so the result is deterministic.  HBM/L2-bound: a step streams every matrix once (they sit in L2).
#include <cuda_runtime.h>

#include "b2_pool.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <vector>

#include "b2_svd.h"

namespace b2 {

namespace {

struct SvdDesc {
   long long w_off, v_off;   // offsets (doubles) in the batch buffers
   int R, C, Ce;             // rows, columns, columns padded to even
   int pair_base;            // first CTA of this matrix in the launch
   double tiny;
};

constexpr int JT = 128;

__device__ __forceinline__ double block_sum3(double& a, double& b, double& c, double* sh) {
   const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_down_sync(0xffffffffu, a, o);
      b += __shfl_down_sync(0xffffffffu, b, o);
      c += __shfl_down_sync(0xffffffffu, c, o);
   }
   if (lane == 0) { sh[w] = a; sh[4 + w] = b; sh[8 + w] = c; }
   __syncthreads();
   a = sh[0] + sh[1] + sh[2] + sh[3];
   b = sh[4] + sh[5] + sh[6] + sh[7];
   c = sh[8] + sh[9] + sh[10] + sh[11];
   return a;
}

// one round-robin step: CTA -> (matrix, pair); rotates the pair when it is not yet orthogonal (Hestenes criterion)
__global__ void __launch_bounds__(JT) k_jacobi_step(const SvdDesc* __restrict__ descs, const int* __restrict__ cta2mat, int step, double* __restrict__ Wb,
                                                    double* __restrict__ Vb, int* __restrict__ rotated, const int* __restrict__ active) {
   __shared__ double sh[12];
   const int mat = cta2mat[blockIdx.x];
   if (!active[mat]) return;
   const SvdDesc d = descs[mat];
   const int k = blockIdx.x - d.pair_base, n1 = d.Ce - 1, s = step % n1;
   int p, q;
   if (k == 0) { p = n1; q = s; }
   else { p = (s + k) % n1; q = (s - k + n1) % n1; }
   if (p > q) { const int t = p; p = q; q = t; }
   if (q >= d.C) return;   // padding column of an odd-sized matrix
   double* wp = Wb + d.w_off + (size_t)d.R * p;
   double* wq = Wb + d.w_off + (size_t)d.R * q;
   double alpha = 0.0, beta = 0.0, gamma = 0.0;
   for (int i = threadIdx.x; i < d.R; i += JT) {
      const double x = wp[i], y = wq[i];
      alpha += x * x; beta += y * y; gamma += x * y;
   }
   block_sum3(alpha, beta, gamma, sh);
   if (fabs(gamma) <= 1e-15 * sqrt(alpha * beta) || fabs(gamma) <= d.tiny) return;
   if (threadIdx.x == 0) rotated[mat] = 1;
   const double zeta = (beta - alpha) / (2.0 * gamma);
   const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
   const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
   for (int i = threadIdx.x; i < d.R; i += JT) {
      const double x = wp[i], y = wq[i];
      wp[i] = c * x - sn * y; wq[i] = sn * x + c * y;
   }
   double* vp = Vb + d.v_off + (size_t)d.C * p;
   double* vq = Vb + d.v_off + (size_t)d.C * q;
   for (int i = threadIdx.x; i < d.C; i += JT) {
      const double x = vp[i], y = vq[i];
      vp[i] = c * x - sn * y; vq[i] = sn * x + c * y;
   }
}

struct DevBuf {
   void* p = nullptr;
   ~DevBuf() { if (p) cudaFree(p); }
   cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 8); }
};

}   // namespace

int dev_svd_batch(std::vector<SvdJob>& jobs, void* stream, char* err, int errlen) {
   cudaStream_t s = (cudaStream_t)stream;
   auto fail = [&](const char* what, cudaError_t e) { snprintf(err, errlen, "dev_svd_batch: %s: %s", what, cudaGetErrorString(e)); return -3; };
   const int nj = (int)jobs.size();
   if (nj == 0) return 0;
   std::vector<SvdDesc> descs(nj);
   std::vector<int> cta2mat;
   long long wtot = 0, vtot = 0;
   int cmax = 0;
   for (int j = 0; j < nj; j++) {
      SvdJob& J = jobs[j];
      const bool flip = J.m < J.n;
      SvdDesc& d = descs[j];
      d.R = flip ? J.n : J.m; d.C = flip ? J.m : J.n; d.Ce = d.C + (d.C & 1);
      d.w_off = wtot; d.v_off = vtot;
      wtot += ((long long)d.R * d.C + 15) / 16 * 16;
      vtot += ((long long)d.C * d.C + 15) / 16 * 16;
      d.pair_base = (int)cta2mat.size();
      for (int k = 0; k < d.Ce / 2; k++) cta2mat.push_back(j);
      cmax = std::max(cmax, d.Ce);
   }
   std::vector<double> W((size_t)wtot, 0.0), V((size_t)vtot, 0.0);
   for (int j = 0; j < nj; j++) {
      const SvdJob& J = jobs[j];
      SvdDesc& d = descs[j];
      const bool flip = J.m < J.n;
      double scale = 0.0;
      double* w = W.data() + d.w_off;
      for (int c = 0; c < d.C; c++)
         for (int r = 0; r < d.R; r++) {
            const double x = flip ? J.a[c + (size_t)J.m * r] : J.a[r + (size_t)J.m * c];
            w[r + (size_t)d.R * c] = x;
            scale = std::max(scale, std::fabs(x));
         }
      d.tiny = scale * scale * 1e-300;
      double* v = V.data() + d.v_off;
      for (int c = 0; c < d.C; c++) v[c + (size_t)d.C * c] = 1.0;
   }
   DevBuf dW, dV, dD, dM, dR, dA;
   cudaError_t e;
   if ((e = dW.alloc(sizeof(double) * W.size())) != cudaSuccess) return fail("alloc W", e);
   if ((e = dV.alloc(sizeof(double) * V.size())) != cudaSuccess) return fail("alloc V", e);
   if ((e = dD.alloc(sizeof(SvdDesc) * nj)) != cudaSuccess) return fail("alloc descs", e);
   if ((e = dM.alloc(sizeof(int) * cta2mat.size())) != cudaSuccess) return fail("alloc map", e);
   if ((e = dR.alloc(sizeof(int) * nj)) != cudaSuccess) return fail("alloc flags", e);
   if ((e = dA.alloc(sizeof(int) * nj)) != cudaSuccess) return fail("alloc flags", e);
   cudaMemcpyAsync(dW.p, W.data(), sizeof(double) * W.size(), cudaMemcpyHostToDevice, s);
   cudaMemcpyAsync(dV.p, V.data(), sizeof(double) * V.size(), cudaMemcpyHostToDevice, s);
   cudaMemcpyAsync(dD.p, descs.data(), sizeof(SvdDesc) * nj, cudaMemcpyHostToDevice, s);
   cudaMemcpyAsync(dM.p, cta2mat.data(), sizeof(int) * cta2mat.size(), cudaMemcpyHostToDevice, s);
   std::vector<int> active(nj, 1), rotated(nj, 0);
   for (int j = 0; j < nj; j++) if (descs[j].C < 2) active[j] = 0;
   const int nsteps = std::max(1, cmax - 1);
   if (!cta2mat.empty() && cmax >= 2) {
      for (int sweep = 0; sweep < 60; sweep++) {
         cudaMemcpyAsync(dA.p, active.data(), sizeof(int) * nj, cudaMemcpyHostToDevice, s);
         cudaMemsetAsync(dR.p, 0, sizeof(int) * nj, s);
         for (int st = 0; st < nsteps; st++)
            k_jacobi_step<<<(unsigned)cta2mat.size(), JT, 0, s>>>((const SvdDesc*)dD.p, (const int*)dM.p, st, (double*)dW.p, (double*)dV.p, (int*)dR.p, (const int*)dA.p);
         if ((e = cudaGetLastError()) != cudaSuccess) return fail("k_jacobi_step launch", e);
         cudaMemcpyAsync(rotated.data(), dR.p, sizeof(int) * nj, cudaMemcpyDeviceToHost, s);
         if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return fail("sweep", e);
         bool any = false;
         // a matrix that saw no rotation during nsteps >= C-1 consecutive steps has had every pair checked: converged
         for (int j = 0; j < nj; j++) { if (active[j] && !rotated[j]) active[j] = 0; any = any || active[j]; }
         if (!any) break;
      }
   }
   cudaMemcpyAsync(W.data(), dW.p, sizeof(double) * W.size(), cudaMemcpyDeviceToHost, s);
   cudaMemcpyAsync(V.data(), dV.p, sizeof(double) * V.size(), cudaMemcpyDeviceToHost, s);
   if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return fail("download", e);
   // singular values = column norms, sorted decreasingly; thin factors in the caller's orientation
   for (int j = 0; j < nj; j++) {
      SvdJob& J = jobs[j];
      const SvdDesc& d = descs[j];
      const bool flip = J.m < J.n;
      const int R = d.R, C = d.C, m = J.m, n = J.n, k = C;
      const double* w = W.data() + d.w_off;
      const double* v = V.data() + d.v_off;
      std::vector<double> nrm(C);
      std::vector<int> idx(C);
      for (int c = 0; c < C; c++) {
         double x = 0.0;
         for (int r = 0; r < R; r++) x += w[r + (size_t)R * c] * w[r + (size_t)R * c];
         nrm[c] = std::sqrt(x); idx[c] = c;
      }
      std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return nrm[x] > nrm[y]; });
      for (int jj = 0; jj < k; jj++) {
         const int c = idx[jj];
         J.s[jj] = nrm[c];
         const double inv = nrm[c] > 0.0 ? 1.0 / nrm[c] : 0.0;
         if (!flip) {
            for (int i = 0; i < m; i++) J.u[i + (size_t)m * jj] = w[i + (size_t)R * c] * inv;
            for (int i = 0; i < n; i++) J.vt[jj + (size_t)k * i] = v[i + (size_t)C * c];
         } else {   // a^T = W V^T  =>  a = V W^T
            for (int i = 0; i < m; i++) J.u[i + (size_t)m * jj] = v[i + (size_t)C * c];
            for (int i = 0; i < n; i++) J.vt[jj + (size_t)k * i] = w[i + (size_t)R * c] * inv;
         }
      }
   }
   return 0;
}

}   // namespace b2
