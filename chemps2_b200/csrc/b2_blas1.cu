// b2_blas1.cu — Davidson vector algebra on the device (replaces the ddot_/daxpy_/dscal_/dlange_ loops of Davidson.cpp:214-412).
// HBM-bound streaming kernels: 128-bit-friendly coalesced grid-stride loops, fused so that every vector is read once per
// step.  Reductions are two-level and fixed-order (per-block partials, then the last block to finish sums them), so results
// are bitwise reproducible run to run.
#include <cuda_runtime.h>

#include <cstdio>

#include "b2_device.h"

namespace b2 {

static thread_local char g_err1[256] = "";
static int fail1(cudaError_t e, const char* what) {
   snprintf(g_err1, sizeof(g_err1), "%s: %s", what, cudaGetErrorString(e));
   return -3;
}
#define LAUNCH_CHECK(what)                                  \
   do {                                                     \
      cudaError_t e_ = cudaGetLastError();                  \
      if (e_ != cudaSuccess) return fail1(e_, what);        \
   } while (0)

constexpr int BT = 256;

// block-level sum of M per-thread values; result valid in thread 0
template <int M> __device__ __forceinline__ void block_sum(double (&v)[M], double* sh /* [M][BT/32] */) {
   const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
   for (int m = 0; m < M; m++) {
      double x = v[m];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
      if (lane == 0) sh[m * (BT / 32) + w] = x;
   }
   __syncthreads();
   if (threadIdx.x == 0) {
#pragma unroll
      for (int m = 0; m < M; m++) {
         double s = 0.0;
         for (int k = 0; k < BT / 32; k++) s += sh[m * (BT / 32) + k];
         v[m] = s;
      }
   }
   __syncthreads();
}

// partial[blockIdx][m] written by thread 0 of every block; the last block sums over blocks in index order
template <int M> __device__ __forceinline__ void grid_finish(double (&v)[M], int m_used, double* partial, unsigned int* counter, double* out) {
   __shared__ bool last;
   if (threadIdx.x == 0) {
      for (int m = 0; m < m_used; m++) partial[(size_t)blockIdx.x * M + m] = v[m];
      __threadfence();
      const unsigned int done = atomicAdd(counter, 1u);
      last = (done == gridDim.x - 1);
   }
   __syncthreads();
   if (last) {
      __threadfence();
      for (int m = threadIdx.x; m < m_used; m += blockDim.x) {
         double s = 0.0;
         for (unsigned int b = 0; b < gridDim.x; b++) s += partial[(size_t)b * M + m];
         out[m] = s;
      }
      if (threadIdx.x == 0) *counter = 0u;
   }
}

// grid-stride loop with four independent iterations in flight per thread: these kernels are pure streaming, and one 8-byte load per
// thread at a time keeps only ~1.2 MB in flight on the whole chip (measured: 1.0 TB/s on a 62 MB vector); four give the memory system
// enough outstanding requests to approach the HBM rate
template <class F> __device__ __forceinline__ void stream4(int64_t n, F&& f) {
   const int64_t step = (int64_t)gridDim.x * blockDim.x;
   int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   for (; e + 3 * step < n; e += 4 * step) { f(e); f(e + step); f(e + 2 * step); f(e + 3 * step); }
   for (; e < n; e += step) f(e);
}

static inline int grid_for(int64_t n) {
   int64_t b = (n + BT - 1) / BT;
   return (int)(b < 1 ? 1 : (b > kRedBlocks ? kRedBlocks : b));
}
static inline unsigned int* counter_of(double* scratch) { return (unsigned int*)(scratch + kRedScratch - 8); }

// ---- multi dot: 8 vectors per pass
constexpr int MD = 8;
__global__ void __launch_bounds__(BT) k_multi_dot(const double* __restrict__ x, const double* __restrict__ ybase, int64_t ystride, int m, int64_t n,
                                                  double* out, double* partial, unsigned int* counter) {
   __shared__ double sh[MD * (BT / 32)];
   double v[MD];
#pragma unroll
   for (int j = 0; j < MD; j++) v[j] = 0.0;
   stream4(n, [&](int64_t e) {
      const double xe = x[e];
#pragma unroll
      for (int j = 0; j < MD; j++)
         if (j < m) v[j] += xe * ybase[(size_t)j * ystride + e];
   });
   block_sum<MD>(v, sh);
   grid_finish<MD>(v, m, partial, counter, out);
}
int dev_multi_dot(const double* x, const double* ybase, int64_t ystride, int m, int64_t n, double* out, double* scratch, void* stream) {
   for (int j0 = 0; j0 < m; j0 += MD) {
      const int mm = (m - j0 < MD) ? m - j0 : MD;
      k_multi_dot<<<grid_for(n), BT, 0, (cudaStream_t)stream>>>(x, ybase + (size_t)j0 * ystride, ystride, mm, n, out + j0, scratch, counter_of(scratch));
      LAUNCH_CHECK("k_multi_dot");
   }
   return 0;
}

__global__ void k_axpy_dev(double* __restrict__ y, const double* __restrict__ x, const double* __restrict__ coef, double sign, int64_t n) {
   const double a = sign * coef[0];
   stream4(n, [&](int64_t e) { y[e] += a * x[e]; });
}
int dev_axpy_dev(double* y, const double* x, const double* coef, double sign, int64_t n, void* stream) {
   k_axpy_dev<<<grid_for(n), BT, 0, (cudaStream_t)stream>>>(y, x, coef, sign, n);
   LAUNCH_CHECK("k_axpy_dev");
   return 0;
}

// y += sign * sum_j coef[j] * X_j (coefficients on the device): the projection step of the classical Gram-Schmidt pass — the whole basis
// is streamed once, y is read and written once
__global__ void __launch_bounds__(BT) k_multi_axpy_dev(double* __restrict__ y, const double* __restrict__ xbase, int64_t xstride, int m, const double* __restrict__ coef,
                                                       double sign, int64_t n) {
   __shared__ double c[kMaxVec];
   if (threadIdx.x < m) c[threadIdx.x] = sign * coef[threadIdx.x];
   __syncthreads();
   stream4(n, [&](int64_t e) {
      double v = y[e];
      for (int j = 0; j < m; j++) v += c[j] * xbase[(size_t)j * xstride + e];
      y[e] = v;
   });
}
int dev_multi_axpy_dev(double* y, const double* xbase, int64_t xstride, int m, const double* coef, double sign, int64_t n, void* stream) {
   if (m <= 0 || n <= 0) return 0;
   k_multi_axpy_dev<<<grid_for(n), BT, 0, (cudaStream_t)stream>>>(y, xbase, xstride, m, coef, sign, n);
   LAUNCH_CHECK("k_multi_axpy_dev");
   return 0;
}

__global__ void k_add_square(double* __restrict__ y, const double* __restrict__ x, int64_t n) {
   stream4(n, [&](int64_t e) { y[e] += x[e] * x[e]; });
}
int dev_add_square(double* y, const double* x, int64_t n, void* stream) {
   if (n <= 0) return 0;
   k_add_square<<<grid_for(n), BT, 0, (cudaStream_t)stream>>>(y, x, n);
   return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

__global__ void k_scale_rsqrt(double* __restrict__ x, const double* __restrict__ ss, int64_t n) {
   const double a = 1.0 / sqrt(ss[0]);
   stream4(n, [&](int64_t e) { x[e] *= a; });
}
int dev_scale_rsqrt(double* x, const double* ss, int64_t n, void* stream) {
   k_scale_rsqrt<<<grid_for(n), BT, 0, (cudaStream_t)stream>>>(x, ss, n);
   LAUNCH_CHECK("k_scale_rsqrt");
   return 0;
}

__global__ void __launch_bounds__(BT) k_ritz_residual(double* __restrict__ u, double* __restrict__ t, const double* __restrict__ V, const double* __restrict__ HV,
                                                      int64_t stride, int m, Coefs a, double theta, int64_t n, double* out, double* partial,
                                                      unsigned int* counter) {
   __shared__ double sh[BT / 32];
   double v[1] = {0.0};
   stream4(n, [&](int64_t e) {
      double ue = 0.0, te = 0.0;
      for (int j = 0; j < m; j++) {
         ue += a.c[j] * V[(size_t)j * stride + e];
         te += a.c[j] * HV[(size_t)j * stride + e];
      }
      te -= theta * ue;
      u[e] = ue; t[e] = te;
      v[0] += te * te;
   });
   block_sum<1>(v, sh);
   grid_finish<1>(v, 1, partial, counter, out);
}
int dev_ritz_residual(double* u, double* t, const double* V, const double* HV, int64_t stride, int m, Coefs a, double theta, int64_t n, double* out,
                      double* scratch, void* stream) {
   k_ritz_residual<<<grid_for(n), BT, 0, (cudaStream_t)stream>>>(u, t, V, HV, stride, m, a, theta, n, out, scratch, counter_of(scratch));
   LAUNCH_CHECK("k_ritz_residual");
   return 0;
}

__device__ __forceinline__ double clamp_diff(double d, double cutoff) { return (fabs(d) > cutoff) ? d : cutoff; }

__global__ void __launch_bounds__(BT) k_precond_dots(double* __restrict__ work, const double* __restrict__ u, const double* __restrict__ t,
                                                     const double* __restrict__ diag, double theta, double cutoff, int64_t n, double* out, double* partial,
                                                     unsigned int* counter) {
   __shared__ double sh[2 * (BT / 32)];
   double v[2] = {0.0, 0.0};
   stream4(n, [&](int64_t e) {
      const double w = u[e] / clamp_diff(diag[e] - theta, cutoff);
      work[e] = w;
      v[0] += w * t[e];
      v[1] += w * u[e];
   });
   block_sum<2>(v, sh);
   grid_finish<2>(v, 2, partial, counter, out);
}
int dev_precond_dots(double* work, const double* u, const double* t, const double* diag, double theta, double cutoff, int64_t n, double* out,
                     double* scratch, void* stream) {
   k_precond_dots<<<grid_for(n), BT, 0, (cudaStream_t)stream>>>(work, u, t, diag, theta, cutoff, n, out, scratch, counter_of(scratch));
   LAUNCH_CHECK("k_precond_dots");
   return 0;
}

__global__ void k_precond_apply(double* __restrict__ t, const double* __restrict__ u, const double* __restrict__ diag, const double* __restrict__ dots,
                                double theta, double cutoff, int64_t n) {
   const double alpha = -dots[0] / dots[1];
   stream4(n, [&](int64_t e) { t[e] = -(t[e] + alpha * u[e]) / clamp_diff(diag[e] - theta, cutoff); });
}
int dev_precond_apply(double* t, const double* u, const double* diag, const double* dots, double theta, double cutoff, int64_t n, void* stream) {
   k_precond_apply<<<grid_for(n), BT, 0, (cudaStream_t)stream>>>(t, u, diag, dots, theta, cutoff, n);
   LAUNCH_CHECK("k_precond_apply");
   return 0;
}

__global__ void k_lincomb(double* __restrict__ out, const double* __restrict__ V, int64_t stride, int m, Coefs a, int64_t n) {
   stream4(n, [&](int64_t e) {
      double s = 0.0;
      for (int j = 0; j < m; j++) s += a.c[j] * V[(size_t)j * stride + e];
      out[e] = s;
   });
}
int dev_lincomb(double* out, const double* V, int64_t stride, int m, Coefs a, int64_t n, void* stream) {
   k_lincomb<<<grid_for(n), BT, 0, (cudaStream_t)stream>>>(out, V, stride, m, a, n);
   LAUNCH_CHECK("k_lincomb");
   return 0;
}

__global__ void k_scale_blocks(double* __restrict__ x, const int64_t* __restrict__ off, const double* __restrict__ scale) {
   const int64_t b = off[blockIdx.x], e1 = off[blockIdx.x + 1];
   const double a = scale[blockIdx.x];
   for (int64_t e = b + (int64_t)blockIdx.y * blockDim.x + threadIdx.x; e < e1; e += (int64_t)gridDim.y * blockDim.x) x[e] *= a;
}
int dev_scale_blocks(double* x, const int64_t* d_off, const double* d_scale, int nblocks, void* stream) {
   if (nblocks <= 0) return 0;
   k_scale_blocks<<<dim3(nblocks, 16), BT, 0, (cudaStream_t)stream>>>(x, d_off, d_scale);   // 16 CTAs per block: the large blocks no longer serialise
   LAUNCH_CHECK("k_scale_blocks");
   return 0;
}

const char* dev_last_error_blas1() { return g_err1; }

}   // namespace b2
