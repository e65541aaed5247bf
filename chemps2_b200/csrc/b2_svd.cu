// b2_svd.cu — batched one-sided Jacobi (Hestenes) SVD on the device: the decomposition step of Sobject::Split
// (Sobject.cpp:412-419 calls dgesdd_ once per centre sector on the host).
//
// All centre-sector matrices of one Split are decomposed together.  For every matrix the columns of the taller orientation
// W (R x C, R >= C) are orthogonalised by plane rotations, V (C x C, starts as identity) accumulates them:  A = W V^T with
// orthogonal columns of W at convergence, so sigma_j = |W(:,j)|, u_j = W(:,j)/sigma_j.
//
// Two kernels walk the column pairs of a sweep in the round-robin ("chess tournament") order, one launch per step over all
// matrices of the batch, every CTA owning a DISJOINT set of columns:
//   k_jacobi_block (default): the columns are grouped in blocks of 8; a CTA takes a PAIR OF BLOCKS (<= 16 columns), forms their
//      16 x 16 Gram matrix, runs one cyclic Jacobi sweep on it in shared memory (15 steps of 8 disjoint rotations) and applies the
//      accumulated 16 x 16 rotation to the 16 columns of W and V.  A sweep over C columns is ceil(C/8) - 1 launches instead of
//      C - 1: the scalar kernel is launch-latency bound (a D = 2000 Split is ~600 launches per sweep, 15-30 sweeps), the block
//      kernel does 8x fewer, fatter steps.  Convergence is decided on the FRESH Gram matrix of every block pair with the Hestenes
//      criterion of the scalar kernel, so both stop at the same accuracy.
//   k_jacobi_step (B2_SVD_BLOCK=0): one CTA per column pair, C - 1 launches per sweep.
// Everything a CTA does is a fixed-order reduction followed by element-wise rotations, so the result is deterministic.
// HBM/L2-bound: a step streams every matrix once (they sit in L2).
#include <cuda_runtime.h>

#include "b2_pool.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <vector>

#include "b2_core.h"
#include "b2_sigma.h"
#include "b2_svd.h"

#include <atomic>
#include <chrono>
#include <cstdlib>

namespace b2 {

namespace {

struct SvdDesc {
   long long w_off, v_off;   // offsets (doubles) in the batch buffers
   int R, C, Ce;             // rows, columns, columns padded to even
   int pair_base;            // first CTA of this matrix in the launch (scalar kernel)
   double tiny;
   int nbe, bpair_base;      // block kernel: column blocks padded to even (>= 2), first CTA of this matrix in the launch
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

constexpr int NB = 8;            // columns per block
constexpr int NP = 2 * NB;       // columns per CTA
constexpr int BT = 128;
constexpr int RCH = 64;          // rows per chunk of the Gram accumulation

// round-robin schedule over n (even) players: the k-th pair of step s, smaller index first
__device__ __forceinline__ void rr_pair(int n, int s, int k, int& p, int& q) {
   const int n1 = n - 1;
   if (k == 0) { p = n1; q = s; }
   else { p = (s + k) % n1; q = (s - k + n1) % n1; }
   if (p > q) { const int t = p; p = q; q = t; }
}

// one round-robin step over column BLOCKS: CTA -> (matrix, pair of blocks); see the header of this file
__global__ void __launch_bounds__(BT) k_jacobi_block(const SvdDesc* __restrict__ descs, const int* __restrict__ cta2mat, int step, double* __restrict__ Wb,
                                                     double* __restrict__ Vb, int* __restrict__ rotated, const int* __restrict__ active) {
   __shared__ double panel[NP][RCH + 1];
   __shared__ double G[NP][NP + 1], Q[NP][NP + 1];
   __shared__ double rc[NB], rs[NB];
   __shared__ int rp[NB], rq[NB];
   __shared__ int cols[NP];
   const int tid = threadIdx.x;
   const int mat = cta2mat[blockIdx.x];
   if (!active[mat]) return;
   const SvdDesc d = descs[mat];
   int bi, bj;
   rr_pair(d.nbe, step % (d.nbe - 1), blockIdx.x - d.bpair_base, bi, bj);
   const int ni = max(0, min(NB, d.C - bi * NB)), nj = max(0, min(NB, d.C - bj * NB));
   const int np = ni + nj;          // a padding block contributes no column
   if (np < 2) return;
   if (tid < NP) cols[tid] = tid < ni ? bi * NB + tid : (tid < np ? bj * NB + (tid - ni) : -1);
   __syncthreads();
   double* W = Wb + d.w_off;
   double* V = Vb + d.v_off;
   // ---- Gram matrix of the <= 16 columns: thread -> entries (a, b) and (a + 8, b), rows accumulated in a fixed order
   const int a = tid >> 4, b = tid & 15;
   double g0 = 0.0, g1 = 0.0;
   for (int r0 = 0; r0 < d.R; r0 += RCH) {
      for (int idx = tid; idx < NP * RCH; idx += BT) {
         const int c = idx / RCH, i = idx - c * RCH;
         panel[c][i] = (c < np && r0 + i < d.R) ? W[(size_t)d.R * cols[c] + r0 + i] : 0.0;
      }
      __syncthreads();
#pragma unroll 8
      for (int i = 0; i < RCH; i++) {
         const double y = panel[b][i];
         g0 += panel[a][i] * y;
         g1 += panel[a + NB][i] * y;
      }
      __syncthreads();
   }
   G[a][b] = g0; G[a + NB][b] = g1;
   Q[a][b] = (a == b) ? 1.0 : 0.0; Q[a + NB][b] = (a + NB == b) ? 1.0 : 0.0;
   __syncthreads();
   // ---- anything left to do?  Hestenes criterion on the fresh inner products, exactly as the scalar kernel applies it
   int need = 0;
   if (a < b && b < np) need |= (fabs(g0) > 1e-15 * sqrt(G[a][a] * G[b][b]) && fabs(g0) > d.tiny);
   if (a + NB < b && b < np) need |= (fabs(g1) > 1e-15 * sqrt(G[a + NB][a + NB] * G[b][b]) && fabs(g1) > d.tiny);
   if (!__syncthreads_or(need)) return;
   if (tid == 0) rotated[mat] = 1;
   // ---- one cyclic Jacobi sweep on G (two-sided), rotations accumulated in Q:  G <- J^T G J,  Q <- Q J
   const int ne = np + (np & 1);
   for (int st = 0; st < ne - 1; st++) {
      if (tid < ne / 2) {
         int p, q;
         rr_pair(ne, st, tid, p, q);
         double c = 1.0, sn = 0.0;
         if (q < np) {
            const double gamma = G[p][q], alpha = G[p][p], beta = G[q][q];
            if (!(fabs(gamma) <= 1e-15 * sqrt(fabs(alpha * beta)) || fabs(gamma) <= d.tiny)) {
               const double zeta = (beta - alpha) / (2.0 * gamma);
               const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
               c = 1.0 / sqrt(1.0 + t * t); sn = c * t;
            }
         }
         rp[tid] = p; rq[tid] = q; rc[tid] = c; rs[tid] = sn;
      }
      __syncthreads();
      const int kk = tid >> 4, r = tid & 15;      // rotation kk, row / column r
      const bool on = kk < ne / 2 && rs[kk] != 0.0;
      const int p = on ? rp[kk] : 0, q = on ? rq[kk] : 0;
      const double c = on ? rc[kk] : 1.0, sn = on ? rs[kk] : 0.0;
      if (on) {   // columns p, q of G and Q
         const double x = G[r][p], y = G[r][q];
         G[r][p] = c * x - sn * y; G[r][q] = sn * x + c * y;
         const double u = Q[r][p], v = Q[r][q];
         Q[r][p] = c * u - sn * v; Q[r][q] = sn * u + c * v;
      }
      __syncthreads();
      if (on) {   // rows p, q of G
         const double x = G[p][r], y = G[q][r];
         G[p][r] = c * x - sn * y; G[q][r] = sn * x + c * y;
      }
      __syncthreads();
   }
   // ---- the 16 columns of W and V times Q (one row per thread and pass, the row is held in registers)
   for (int pass = 0; pass < 2; pass++) {
      double* M = pass == 0 ? W : V;
      const int rows = pass == 0 ? d.R : d.C;
      for (int i = tid; i < rows; i += BT) {
         double x[NP];
#pragma unroll
         for (int c = 0; c < NP; c++) x[c] = c < np ? M[(size_t)rows * cols[c] + i] : 0.0;
#pragma unroll 1
         for (int bb = 0; bb < np; bb++) {   // rolled on purpose: unrolled, the compiler keeps all of Q in registers and spills
            double y = 0.0;
#pragma unroll
            for (int c = 0; c < NP; c++) y += x[c] * Q[c][bb];
            M[(size_t)rows * cols[bb] + i] = y;
         }
      }
   }
}

// V = identity for every matrix of the batch
__global__ void k_svd_identity(const SvdDesc* __restrict__ descs, double* __restrict__ Vb) {
   const SvdDesc d = descs[blockIdx.x];
   double* v = Vb + d.v_off;
   for (int c = threadIdx.x; c < d.C; c += blockDim.x) v[c + (size_t)d.C * c] = 1.0;
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
   static const bool use_block = [] { const char* e = getenv("B2_SVD_BLOCK"); return !e || atoi(e) != 0; }();
   std::vector<SvdDesc> descs(nj);
   std::vector<int> cta2mat;
   long long wtot = 0, vtot = 0;
   int cmax = 0, nbmax = 0;
   for (int j = 0; j < nj; j++) {
      SvdJob& J = jobs[j];
      const bool flip = J.m < J.n;
      SvdDesc& d = descs[j];
      d.R = flip ? J.n : J.m; d.C = flip ? J.m : J.n; d.Ce = d.C + (d.C & 1);
      d.w_off = wtot; d.v_off = vtot;
      wtot += ((long long)d.R * d.C + 15) / 16 * 16;
      vtot += ((long long)d.C * d.C + 15) / 16 * 16;
      const int nb = (d.C + NB - 1) / NB;
      d.nbe = std::max(2, nb + (nb & 1));
      d.pair_base = d.bpair_base = (int)cta2mat.size();
      for (int k = 0; k < (use_block ? d.nbe : d.Ce) / 2; k++) cta2mat.push_back(j);
      cmax = std::max(cmax, d.Ce);
      nbmax = std::max(nbmax, d.nbe);
   }
   // host staging: the taller orientation of every matrix, filled job by job on the host workers
   // (pinned buffers from the library's staging cache: no zero fill, copies at full PCIe / C2C speed)
   struct Stage {
      double* p = nullptr; bool pinned = false; std::vector<double> fallback;
      void get(size_t n) { p = (double*)pinned_acquire(sizeof(double) * std::max<size_t>(n, 1)); pinned = p != nullptr; if (!p) { fallback.resize(std::max<size_t>(n, 1)); p = fallback.data(); } }
      ~Stage() { if (pinned) pinned_release(p); }
      double* data() const { return p; }
   } W, V;
   const bool timing = getenv("B2_TIMING") != nullptr;
   auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
   const double t_0 = now();
   W.get((size_t)wtot); V.get((size_t)vtot);
   {
      std::atomic<int> next{0};
      const int T = std::max(1, std::min(nj, plan_threads(8 * nj)));
      parallel_run(T, [&](int) {
         for (int j; (j = next.fetch_add(1)) < nj;) {
            const SvdJob& J = jobs[j];
            SvdDesc& d = descs[j];
            const bool flip = J.m < J.n;
            double scale = 0.0;
            double* w = W.data() + d.w_off;
            if (!flip) {
               for (size_t e = 0; e < (size_t)d.R * d.C; e++) { w[e] = J.a[e]; scale = std::max(scale, std::fabs(J.a[e])); }
            } else {
               for (int r = 0; r < d.R; r++)       // w(r, c) = a(c, r): rows of w are columns of a
                  for (int c = 0; c < d.C; c++) {
                     const double x = J.a[c + (size_t)J.m * r];
                     w[r + (size_t)d.R * c] = x;
                     scale = std::max(scale, std::fabs(x));
                  }
            }
            for (size_t e = (size_t)d.R * d.C; e < (size_t)(((long long)d.R * d.C + 15) / 16 * 16); e++) w[e] = 0.0;
            d.tiny = scale * scale * 1e-300;
         }
      });
   }
   DevBuf dW, dV, dD, dM, dR, dA;
   cudaError_t e;
   const double t_1 = now();
   if ((e = dW.alloc(sizeof(double) * (size_t)wtot)) != cudaSuccess) return fail("alloc W", e);
   if ((e = dV.alloc(sizeof(double) * (size_t)vtot)) != cudaSuccess) return fail("alloc V", e);
   if ((e = dD.alloc(sizeof(SvdDesc) * nj)) != cudaSuccess) return fail("alloc descs", e);
   if ((e = dM.alloc(sizeof(int) * cta2mat.size())) != cudaSuccess) return fail("alloc map", e);
   if ((e = dR.alloc(sizeof(int) * nj)) != cudaSuccess) return fail("alloc flags", e);
   if ((e = dA.alloc(sizeof(int) * nj)) != cudaSuccess) return fail("alloc flags", e);
   cudaMemcpyAsync(dW.p, W.data(), sizeof(double) * (size_t)wtot, cudaMemcpyHostToDevice, s);
   cudaMemcpyAsync(dD.p, descs.data(), sizeof(SvdDesc) * nj, cudaMemcpyHostToDevice, s);
   cudaMemcpyAsync(dM.p, cta2mat.data(), sizeof(int) * cta2mat.size(), cudaMemcpyHostToDevice, s);
   cudaMemsetAsync(dV.p, 0, sizeof(double) * (size_t)vtot, s);
   k_svd_identity<<<nj, 128, 0, s>>>((const SvdDesc*)dD.p, (double*)dV.p);
   std::vector<int> active(nj, 1), rotated(nj, 0);
   for (int j = 0; j < nj; j++) if (descs[j].C < 2) active[j] = 0;
   const int nsteps = use_block ? std::max(1, nbmax - 1) : std::max(1, cmax - 1);
   int nsweeps = 0;
   if (!cta2mat.empty() && cmax >= 2) {
      for (int sweep = 0; sweep < 60; sweep++) {
         nsweeps++;
         cudaMemcpyAsync(dA.p, active.data(), sizeof(int) * nj, cudaMemcpyHostToDevice, s);
         cudaMemsetAsync(dR.p, 0, sizeof(int) * nj, s);
         for (int st = 0; st < nsteps; st++) {
            if (use_block)
               k_jacobi_block<<<(unsigned)cta2mat.size(), BT, 0, s>>>((const SvdDesc*)dD.p, (const int*)dM.p, st, (double*)dW.p, (double*)dV.p, (int*)dR.p, (const int*)dA.p);
            else
               k_jacobi_step<<<(unsigned)cta2mat.size(), JT, 0, s>>>((const SvdDesc*)dD.p, (const int*)dM.p, st, (double*)dW.p, (double*)dV.p, (int*)dR.p, (const int*)dA.p);
         }
         if ((e = cudaGetLastError()) != cudaSuccess) return fail("jacobi step launch", e);
         cudaMemcpyAsync(rotated.data(), dR.p, sizeof(int) * nj, cudaMemcpyDeviceToHost, s);
         if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return fail("sweep", e);
         bool any = false;
         // a matrix that saw no rotation during nsteps >= (its number of steps per sweep) consecutive steps has had every pair checked: converged
         for (int j = 0; j < nj; j++) { if (active[j] && !rotated[j]) active[j] = 0; any = any || active[j]; }
         if (!any) break;
      }
   }
   const double t_2 = now();
   cudaMemcpyAsync(W.data(), dW.p, sizeof(double) * (size_t)wtot, cudaMemcpyDeviceToHost, s);
   cudaMemcpyAsync(V.data(), dV.p, sizeof(double) * (size_t)vtot, cudaMemcpyDeviceToHost, s);
   if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return fail("download", e);
   const double t_3 = now();
   // singular values = column norms, sorted decreasingly; thin factors in the caller's orientation (job by job on the host workers)
   std::atomic<int> next{0};
   const int T = std::max(1, std::min(nj, plan_threads(8 * nj)));
   parallel_run(T, [&](int) {
      for (int j; (j = next.fetch_add(1)) < nj;) {
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
   });
   if (timing)
      fprintf(stderr, "dev_svd_batch: %d matrices, %.1f MB, widest %d columns: stage %.4f s, %d sweeps x %d launches (+upload) %.4f s, download %.4f s, factors %.4f s\n", nj,
              (wtot + vtot) * 8e-6, cmax, t_1 - t_0, nsweeps, nsteps, t_2 - t_1, t_3 - t_2, now() - t_3);
   return 0;
}

}   // namespace b2
