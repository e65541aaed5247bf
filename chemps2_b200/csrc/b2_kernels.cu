// b2_kernels.cu — sm_100a kernels of the sigma build / operator update.
//
// k_tiles<TM,TN,WM,WN>: grouped FP64 contraction.  One CTA owns one output tile of one symmetry block and
//   accumulates EVERY term that lands on it in registers (deterministic, no atomics):
//       C_tile = sum_items alpha * opX(X) * opY(Y)
//   Operand panels are staged through shared memory (k-major, padded so that the DMMA fragment loads are
//   bank-conflict free for 64-bit accesses) and multiplied with FP64 tensor-core MMA
//   (mma.sync.aligned.m8n8k4.f64 -> SASS DMMA).  tcgen05/UMMA has no FP64 path, so warp-level DMMA is the
//   tensor pipe this workload can use on sm_100a.
// k_presum: integral-weighted operator pre-sums (HBM-bound, vectorised, coalesced).
#include <cuda_runtime.h>

#include <cstdio>

#include "b2_device.h"

namespace b2 {

static thread_local char g_dev_err[256] = "";
const char* dev_last_error() { return g_dev_err; }
static int cuda_fail(cudaError_t e, const char* what) {
   snprintf(g_dev_err, sizeof(g_dev_err), "%s: %s", what, cudaGetErrorString(e));
   return -3;
}

__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int KC = 8;   // k-chunk staged per shared-memory panel

template <int TM, int TN, int WM, int WN>
__global__ void __launch_bounds__(WM * WN * 32) k_tiles(const Tile* __restrict__ tiles, const GemmItem* __restrict__ items, DevBases bases) {
   constexpr int NT = WM * WN * 32;
   constexpr int WTM = TM / WM, WTN = TN / WN;   // warp tile
   constexpr int MI = WTM / 8, NI = WTN / 8;     // 8x8 MMA tiles per warp
   constexpr int SX = TM + 4, SY = TN + 4;       // strides == 4 (mod 16) doubles: conflict-free half-warp fragment loads
   __shared__ double Xs[KC * SX];
   __shared__ double Ys[KC * SY];

   const Tile t = tiles[blockIdx.x];
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   const int wm = warp % WM, wn = warp / WM;
   const int g = lane >> 2, q = lane & 3;        // fragment coordinates
   // 8x8 sub-tiles of this warp that lie (partly) inside the tile; the rest is skipped (warp-uniform)
   const int mi_n = min(MI, max(0, (t.mrem - wm * WTM + 7) >> 3));
   const int ni_n = min(NI, max(0, (t.nrem - wn * WTN + 7) >> 3));

   double acc[MI][NI][2];
#pragma unroll
   for (int i = 0; i < MI; i++)
#pragma unroll
      for (int j = 0; j < NI; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

   for (int it = t.item_begin; it < t.item_end; it++) {
      const GemmItem I = items[it];
      const double* __restrict__ X = bases.p[I.xs] + I.xoff;
      if (I.flags & IF_AXPY) {
#pragma unroll
         for (int i = 0; i < MI; i++)
#pragma unroll
            for (int j = 0; j < NI; j++) {
               const int r = wm * WTM + i * 8 + g, c = wn * WTN + j * 8 + 2 * q;
               if (r < t.mrem) {
                  if (c < t.nrem) acc[i][j][0] += I.alpha * X[(size_t)(t.m0 + r) + (size_t)(t.n0 + c) * I.ldx];
                  if (c + 1 < t.nrem) acc[i][j][1] += I.alpha * X[(size_t)(t.m0 + r) + (size_t)(t.n0 + c + 1) * I.ldx];
               }
            }
         continue;
      }
      const double* __restrict__ Y = bases.p[I.ys] + I.yoff;
      const int K = I.k;
      const bool tx = I.flags & IF_TX, ty = I.flags & IF_TY;
      for (int k0 = 0; k0 < K; k0 += KC) {
         // ---- stage the X panel: Xs[k][m] = alpha * opX(X)[m0+m][k0+k]
         if (!tx) {
            for (int idx = tid; idx < TM * KC; idx += NT) {
               const int m = idx % TM, k = idx / TM;
               double v = 0.0;
               if (m < t.mrem && k0 + k < K) v = I.alpha * X[(size_t)(t.m0 + m) + (size_t)(k0 + k) * I.ldx];
               Xs[k * SX + m] = v;
            }
         } else {
            for (int idx = tid; idx < TM * KC; idx += NT) {
               const int k = idx % KC, m = idx / KC;
               double v = 0.0;
               if (m < t.mrem && k0 + k < K) v = I.alpha * X[(size_t)(k0 + k) + (size_t)(t.m0 + m) * I.ldx];
               Xs[k * SX + m] = v;
            }
         }
         // ---- stage the Y panel: Ys[k][n] = opY(Y)[k0+k][n0+n]
         if (!ty) {
            for (int idx = tid; idx < TN * KC; idx += NT) {
               const int k = idx % KC, n = idx / KC;
               double v = 0.0;
               if (n < t.nrem && k0 + k < K) v = Y[(size_t)(k0 + k) + (size_t)(t.n0 + n) * I.ldy];
               Ys[k * SY + n] = v;
            }
         } else {
            for (int idx = tid; idx < TN * KC; idx += NT) {
               const int n = idx % TN, k = idx / TN;
               double v = 0.0;
               if (n < t.nrem && k0 + k < K) v = Y[(size_t)(t.n0 + n) + (size_t)(k0 + k) * I.ldy];
               Ys[k * SY + n] = v;
            }
         }
         __syncthreads();
#pragma unroll
         for (int kk = 0; kk < KC; kk += 4) {
            double a[MI], b[NI];
#pragma unroll
            for (int i = 0; i < MI; i++) a[i] = Xs[(kk + q) * SX + wm * WTM + i * 8 + g];
#pragma unroll
            for (int j = 0; j < NI; j++) b[j] = Ys[(kk + q) * SY + wn * WTN + j * 8 + g];
#pragma unroll
            for (int i = 0; i < MI; i++)
#pragma unroll
               for (int j = 0; j < NI; j++)
                  if (i < mi_n && j < ni_n) dmma8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
         }
         __syncthreads();
      }
   }

   double* __restrict__ C = bases.p[t.cspace] + t.coff;
#pragma unroll
   for (int i = 0; i < MI; i++)
#pragma unroll
      for (int j = 0; j < NI; j++) {
         const int r = wm * WTM + i * 8 + g, c = wn * WTN + j * 8 + 2 * q;
         if (r < t.mrem) {
            double* p0 = C + (size_t)(t.cm0 + r) + (size_t)(t.cn0 + c) * t.ldc;
            if (t.accumulate) {
               if (c < t.nrem) p0[0] += acc[i][j][0];
               if (c + 1 < t.nrem) p0[t.ldc] += acc[i][j][1];
            } else {
               if (c < t.nrem) p0[0] = acc[i][j][0];
               if (c + 1 < t.nrem) p0[t.ldc] = acc[i][j][1];
            }
         }
      }
}

int dev_launch_tiles(int tile_class, const Tile* d_tiles, int ntiles, const GemmItem* d_items, const DevBases& bases, void* stream) {
   if (ntiles <= 0) return 0;
   cudaStream_t s = (cudaStream_t)stream;
   switch (tile_class) {
      case 0: k_tiles<64, 64, 2, 2><<<ntiles, 128, 0, s>>>(d_tiles, d_items, bases); break;
      case 1: k_tiles<32, 32, 2, 2><<<ntiles, 128, 0, s>>>(d_tiles, d_items, bases); break;
      case 2: k_tiles<16, 16, 1, 1><<<ntiles, 32, 0, s>>>(d_tiles, d_items, bases); break;
      case 3: k_tiles<8, 8, 1, 1><<<ntiles, 32, 0, s>>>(d_tiles, d_items, bases); break;
      default: snprintf(g_dev_err, sizeof(g_dev_err), "bad tile class %d", tile_class); return -1;
   }
   cudaError_t e = cudaGetLastError();
   if (e != cudaSuccess) return cuda_fail(e, "k_tiles launch");
   return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// split-K epilogue: sigma tile += partial slots, summed in slot order (deterministic)
__global__ void k_reduce(const ReduceJob* __restrict__ jobs, DevBases bases) {
   const ReduceJob j = jobs[blockIdx.x];
   const double* __restrict__ part = bases.p[SP_PART] + j.part_off;
   double* __restrict__ C = bases.p[SP_VOUT] + j.dst_off;
   const int n = j.mrem * j.nrem;
   for (int e = threadIdx.x; e < n; e += blockDim.x) {
      double v = 0.0;
      for (int p = 0; p < j.nparts; p++) v += part[(size_t)p * j.part_stride + e];
      const int r = e % j.mrem, c = e / j.mrem;
      C[(size_t)(j.m0 + r) + (size_t)(j.n0 + c) * j.ldc] += v;
   }
}
int dev_launch_reduce(const ReduceJob* d_jobs, int njobs, const DevBases& bases, void* stream) {
   if (njobs <= 0) return 0;
   k_reduce<<<njobs, 256, 0, (cudaStream_t)stream>>>(d_jobs, bases);
   cudaError_t e = cudaGetLastError();
   if (e != cudaSuccess) return cuda_fail(e, "k_reduce launch");
   return 0;
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void k_presum(const PresumJob* __restrict__ jobs, const PresumPart* __restrict__ parts, DevBases bases) {
   const PresumJob j = jobs[blockIdx.y];
   double* __restrict__ out = bases.p[SP_PRESUM] + j.dst_off;
   for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < j.size; e += (int64_t)gridDim.x * blockDim.x) {
      double v = 0.0;
      for (int p = j.part_begin; p < j.part_end; p++) v += parts[p].coef * bases.p[parts[p].space][parts[p].src_off + e];
      out[e] = v;
   }
}

int dev_launch_presum(const PresumJob* d_jobs, int njobs, const PresumPart* d_parts, const DevBases& bases, void* stream) {
   if (njobs <= 0) return 0;
   // gridDim.y is limited to 65535: launch in slabs
   for (int j0 = 0; j0 < njobs; j0 += 65535) {
      const int nj = (njobs - j0 < 65535) ? njobs - j0 : 65535;
      dim3 grid(8, nj);
      k_presum<<<grid, 256, 0, (cudaStream_t)stream>>>(d_jobs + j0, d_parts, bases);
   }
   cudaError_t e = cudaGetLastError();
   if (e != cudaSuccess) return cuda_fail(e, "k_presum launch");
   return 0;
}

__global__ void k_zero(double* p, int64_t n) {
   for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) p[e] = 0.0;
}
int dev_fill_zero(double* d_ptr, int64_t n, void* stream) {
   if (n <= 0) return 0;
   k_zero<<<592, 256, 0, (cudaStream_t)stream>>>(d_ptr, n);
   cudaError_t e = cudaGetLastError();
   if (e != cudaSuccess) return cuda_fail(e, "k_zero launch");
   return 0;
}

__global__ void k_fill_hash(double* p, int64_t n, uint64_t seed, uint64_t key, double amp) {
   for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
      uint64_t z = seed + 0x9E3779B97F4A7C15ULL * (key + 1) + 0xD1B54A32D192ED03ULL * ((uint64_t)e + 1);
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
      z = z ^ (z >> 31);
      p[e] = amp * ((double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5);
   }
}
int dev_fill_hash(double* d_ptr, int64_t n, uint64_t seed, uint64_t key, double amp, void* stream) {
   if (n <= 0) return 0;
   const int blocks = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
   k_fill_hash<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_ptr, n, seed, key, amp);
   cudaError_t e = cudaGetLastError();
   if (e != cudaSuccess) return cuda_fail(e, "k_fill_hash launch");
   return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// FP64 peak probes (register resident, no memory traffic): the measured roofline denominator for the DMMA kernels.
__global__ void k_probe_mma(double* out, int iters) {
   double c[8][2];
   for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = 0.0;
   double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
   for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++) dmma8x8x4(c[i][0], c[i][1], a, b);
   }
   double s = 0.0;
   for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_probe_fma(double* out, int iters) {
   double c[16];
   for (int i = 0; i < 16; i++) c[i] = threadIdx.x * 1e-9;
   double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
   for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 16; i++) c[i] = fma(a, c[i], b);
   }
   double s = 0.0;
   for (int i = 0; i < 16; i++) s += c[i];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int dev_probe_fp64(int use_mma, double* tflops_out) {
   const int blocks = 148 * 8, threads = 256, iters = 20000;
   double* d = nullptr;
   cudaError_t e = cudaMalloc(&d, sizeof(double) * blocks * threads);
   if (e != cudaSuccess) return cuda_fail(e, "probe malloc");
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0); cudaEventCreate(&e1);
   double best = 0.0;
   for (int rep = 0; rep < 4; rep++) {
      cudaEventRecord(e0);
      if (use_mma) k_probe_mma<<<blocks, threads>>>(d, iters); else k_probe_fma<<<blocks, threads>>>(d, iters);
      cudaEventRecord(e1);
      e = cudaEventSynchronize(e1);
      if (e != cudaSuccess) { cudaFree(d); return cuda_fail(e, "probe run"); }
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      // per warp and iteration: 8 MMAs x (8*8*4*2) flops ; per thread and iteration: 16 FMAs x 2 flops
      const double flops = use_mma ? (double)blocks * (threads / 32) * iters * 8.0 * 512.0 : (double)blocks * threads * iters * 16.0 * 2.0;
      const double tf = flops / (ms * 1e-3) / 1e12;
      if (rep > 0 && tf > best) best = tf;
   }
   cudaEventDestroy(e0); cudaEventDestroy(e1);
   cudaFree(d);
   *tflops_out = best;
   return 0;
}

}   // namespace b2
