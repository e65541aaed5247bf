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

constexpr int KC = 16;       // k-chunk staged per pipeline stage
constexpr int STAGES = 3;    // cp.async pipeline depth

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc, bool valid) {
   const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
   const int bytes = valid ? 8 : 0;   // src-size 0 => the 8 destination bytes are zero-filled
   asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gsrc), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Shared-memory operand panels keep the operand's own contiguous direction (so global reads stay coalesced whatever
// the transposition flag) and are padded so that the 64-bit DMMA fragment loads of a half-warp hit 16 distinct bank pairs:
//   "m-major"  P[k][m], stride T+4   (T = 64/32/16/8: stride == 4 mod 8 doubles)
//   "k-major"  P[m][k], stride KC+4  (20 doubles)
template <int T> struct Panel {
   static constexpr int SM = T + 4, SK = KC + 4;
   static constexpr int SIZE = (KC * SM > T * SK) ? KC * SM : T * SK;
};

// Per-item staging cursor of one operand panel.  For a T x KC panel and NT threads every thread copies E = T*KC/NT
// elements per chunk; with the thread -> element map below all E elements of a thread share one coordinate, so the
// global pointer, the validity of that coordinate and the shared-memory slot are computed once per item and a chunk
// costs one predicate + one pointer bump + one cp.async per element (the index arithmetic used to dominate issue slots).
//   contig_r  (stored R x K, rows contiguous)  -> "m-major" panel P[k][r]:  r = tid % T fixed,   k = tid / T + e * (NT / T)
//   !contig_r (stored K x R, k contiguous)     -> "k-major" panel P[r][k]:  k = tid % KC fixed,  r = tid / KC + e * (NT / KC)
template <int T, int NT> struct Stager {
   static constexpr int E = (T * KC) / NT;
   static_assert((T * KC) % NT == 0 && E >= 1, "panel must split evenly over the CTA");
   const double* g;      // global pointer of element e = 0 of the current chunk
   long long estep;      // global stride between consecutive e
   long long cstep;      // global stride between consecutive chunks
   int soff, sstep;      // shared-memory slot of e = 0 and stride between consecutive e
   int fix_ok;           // the fixed coordinate is inside the matrix (contig_r: r < rrem; else evaluated per chunk: k0 + k < K)
   int var0, varstep;    // the varying coordinate of e = 0 and its step
   int kfix;             // !contig_r: k of this thread
   bool contig;

   __device__ __forceinline__ void init(const double* G, int ld, bool contig_r, int r0, int rrem, int tid) {
      contig = contig_r;
      if (contig_r) {
         const int r = tid % T, kb = tid / T;
         g = G + (size_t)(r0 + r) + (size_t)kb * ld;
         estep = (long long)(NT / T) * ld; cstep = (long long)KC * ld;
         soff = kb * Panel<T>::SM + r; sstep = (NT / T) * Panel<T>::SM;
         fix_ok = r < rrem; var0 = kb; varstep = NT / T; kfix = 0;
      } else {
         const int k = tid % KC, rb = tid / KC;
         g = G + (size_t)k + (size_t)(r0 + rb) * ld;
         estep = (long long)(NT / KC) * ld; cstep = KC;
         soff = rb * Panel<T>::SK + k; sstep = (NT / KC) * Panel<T>::SK;
         fix_ok = 1; var0 = rb; varstep = NT / KC; kfix = k;
      }
   }
   // copies the chunk that starts at k0 (kleft = K - k0 > 0 columns left); rrem = rows of the tile
   __device__ __forceinline__ void chunk(double* P, int kleft, int rrem) {
      const double* p = g;
      if (contig) {
#pragma unroll
         for (int e = 0; e < E; e++) {
            const bool ok = fix_ok && (var0 + e * varstep < kleft);
            cp_async8(P + soff + e * sstep, ok ? p : g, ok);
            p += estep;
         }
      } else {
         const bool kok = kfix < kleft;
#pragma unroll
         for (int e = 0; e < E; e++) {
            const bool ok = kok && (var0 + e * varstep < rrem);
            cp_async8(P + soff + e * sstep, ok ? p : g, ok);
            p += estep;
         }
      }
      g += cstep;
   }
};

// one k-chunk of MMAs; XM / YM: the X / Y panel is m-major (compile-time so that the fragment addresses fold to immediates)
template <int TM, int TN, int MI, int NI, bool XM, bool YM>
__device__ __forceinline__ void mma_chunk(double (&acc)[MI][NI][2], const double* __restrict__ xs, const double* __restrict__ ys, int rbase, int cbase,
                                          int q, int kvalid, double alpha, int mi_n, int ni_n) {
   const double* xa = XM ? xs + q * Panel<TM>::SM + rbase : xs + rbase * Panel<TM>::SK + q;
   const double* yb = YM ? ys + q * Panel<TN>::SM + cbase : ys + cbase * Panel<TN>::SK + q;
#pragma unroll
   for (int kk = 0; kk < KC; kk += 4) {
      if (kk < kvalid) {
         double a[MI], b[NI];
#pragma unroll
         for (int i = 0; i < MI; i++) a[i] = alpha * (XM ? xa[kk * Panel<TM>::SM + i * 8] : xa[i * 8 * Panel<TM>::SK + kk]);
#pragma unroll
         for (int j = 0; j < NI; j++) b[j] = YM ? yb[kk * Panel<TN>::SM + j * 8] : yb[j * 8 * Panel<TN>::SK + kk];
#pragma unroll
         for (int i = 0; i < MI; i++)
#pragma unroll
            for (int j = 0; j < NI; j++)
               if (i < mi_n && j < ni_n) dmma8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
   }
}

template <int TM, int TN, int WM, int WN>
__global__ void __launch_bounds__(WM * WN * 32) k_tiles(const Tile* __restrict__ tiles, const GemmItem* __restrict__ items, DevBases bases) {
   constexpr int NT = WM * WN * 32;
   constexpr int WTM = TM / WM, WTN = TN / WN;   // warp tile
   constexpr int MI = WTM / 8, NI = WTN / 8;     // 8x8 MMA tiles per warp
   constexpr int XSZ = Panel<TM>::SIZE, YSZ = Panel<TN>::SIZE;
   extern __shared__ double smem[];
   double* Xs = smem;                    // STAGES panels
   double* Ys = smem + STAGES * XSZ;

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

   // ---- block-axpy items come first in every item range (b2_compile.cpp sorts them there)
   int it0 = t.item_begin;
   for (; it0 < t.item_end; it0++) {
      const GemmItem I = items[it0];
      if (!(I.flags & IF_AXPY)) break;
      const double* __restrict__ X = bases.p[I.xs] + I.xoff;
#pragma unroll
      for (int i = 0; i < MI; i++)
#pragma unroll
         for (int j = 0; j < NI; j++) {
            const int r = wm * WTM + i * 8 + g, c = wn * WTN + j * 8 + 2 * q;
            if (r < t.mrem) {
               if (I.flags & IF_TX) {
                  if (c < t.nrem) acc[i][j][0] += I.alpha * X[(size_t)(t.n0 + c) + (size_t)(t.m0 + r) * I.ldx];
                  if (c + 1 < t.nrem) acc[i][j][1] += I.alpha * X[(size_t)(t.n0 + c + 1) + (size_t)(t.m0 + r) * I.ldx];
               } else {
                  if (c < t.nrem) acc[i][j][0] += I.alpha * X[(size_t)(t.m0 + r) + (size_t)(t.n0 + c) * I.ldx];
                  if (c + 1 < t.nrem) acc[i][j][1] += I.alpha * X[(size_t)(t.m0 + r) + (size_t)(t.n0 + c + 1) * I.ldx];
               }
            }
         }
   }

   // ---- GEMM items: one flattened stream of k-chunks over all items, software-pipelined with cp.async
   int p_it = it0, p_left = 0;          // producer cursor: item, columns of K still to stage
   Stager<TM, NT> sx;
   Stager<TN, NT> sy;
   auto producer_load_item = [&]() {
      const GemmItem P = items[p_it];
      sx.init(bases.p[P.xs] + P.xoff, P.ldx, !(P.flags & IF_TX), t.m0, t.mrem, tid);
      sy.init(bases.p[P.ys] + P.yoff, P.ldy, (P.flags & IF_TY) != 0, t.n0, t.nrem, tid);
      p_left = P.k;
   };
   if (p_it < t.item_end) producer_load_item();
   auto issue = [&](int stage) {
      if (p_it < t.item_end) {
         sx.chunk(Xs + stage * XSZ, p_left, t.mrem);
         sy.chunk(Ys + stage * YSZ, p_left, t.nrem);
         p_left -= KC;
         if (p_left <= 0 && ++p_it < t.item_end) producer_load_item();
      }
      cp_async_commit();
   };
#pragma unroll
   for (int s = 0; s < STAGES - 1; s++) issue(s);

   int c_it = it0, c_left = 0, stage = 0, cflags = 0;
   double alpha = 0.0;
   if (c_it < t.item_end) { const GemmItem Cn = items[c_it]; c_left = Cn.k; cflags = Cn.flags; alpha = Cn.alpha; }
   const int rbase = wm * WTM + g, cbase = wn * WTN + g;
   while (c_it < t.item_end) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      issue((stage + STAGES - 1) % STAGES);   // refills the stage consumed in the previous iteration
      const double* xs = Xs + stage * XSZ;
      const double* ys = Ys + stage * YSZ;
      const int kvalid = min(KC, c_left);
      if (!(cflags & IF_TX)) {
         if (cflags & IF_TY) mma_chunk<TM, TN, MI, NI, true, true>(acc, xs, ys, rbase, cbase, q, kvalid, alpha, mi_n, ni_n);
         else mma_chunk<TM, TN, MI, NI, true, false>(acc, xs, ys, rbase, cbase, q, kvalid, alpha, mi_n, ni_n);
      } else {
         if (cflags & IF_TY) mma_chunk<TM, TN, MI, NI, false, true>(acc, xs, ys, rbase, cbase, q, kvalid, alpha, mi_n, ni_n);
         else mma_chunk<TM, TN, MI, NI, false, false>(acc, xs, ys, rbase, cbase, q, kvalid, alpha, mi_n, ni_n);
      }
      c_left -= KC;
      if (c_left <= 0 && ++c_it < t.item_end) { const GemmItem Cn = items[c_it]; c_left = Cn.k; cflags = Cn.flags; alpha = Cn.alpha; }
      stage = (stage + 1) % STAGES;
   }
   cp_async_wait<0>();

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

template <int TM, int TN, int WM, int WN>
static cudaError_t launch_tiles_t(const Tile* d_tiles, int ntiles, const GemmItem* d_items, const DevBases& bases, cudaStream_t s) {
   constexpr size_t smem = sizeof(double) * STAGES * (Panel<TM>::SIZE + Panel<TN>::SIZE);
   static bool configured = false;
   if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(k_tiles<TM, TN, WM, WN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      configured = true;
   }
   k_tiles<TM, TN, WM, WN><<<ntiles, WM * WN * 32, smem, s>>>(d_tiles, d_items, bases);
   return cudaGetLastError();
}

int dev_launch_tiles(int tile_class, const Tile* d_tiles, int ntiles, const GemmItem* d_items, const DevBases& bases, void* stream) {
   if (ntiles <= 0) return 0;
   cudaStream_t s = (cudaStream_t)stream;
   cudaError_t e;
   switch (tile_class) {
      case 0: e = launch_tiles_t<64, 64, 2, 2>(d_tiles, ntiles, d_items, bases, s); break;
      case 1: e = launch_tiles_t<32, 32, 2, 2>(d_tiles, ntiles, d_items, bases, s); break;
      case 2: e = launch_tiles_t<16, 16, 1, 1>(d_tiles, ntiles, d_items, bases, s); break;
      case 3: e = launch_tiles_t<8, 8, 1, 1>(d_tiles, ntiles, d_items, bases, s); break;
      default: snprintf(g_dev_err, sizeof(g_dev_err), "bad tile class %d", tile_class); return -1;
   }
   if (e != cudaSuccess) return cuda_fail(e, "k_tiles launch");
   return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// split-K epilogue: sigma tile += partial slots, summed in slot order (deterministic); one thread per tile element
__global__ void __launch_bounds__(256) k_reduce(const ReduceJob* __restrict__ jobs, DevBases bases) {
   const ReduceJob j = jobs[blockIdx.x];
   const int n = j.mrem * j.nrem;
   const int e = blockIdx.y * 256 + threadIdx.x;
   if (e >= n) return;
   const double* __restrict__ part = bases.p[SP_PART] + j.part_off + e;
   double v = 0.0;
   int p = 0;
   for (; p + 4 <= j.nparts; p += 4) {   // four independent loads in flight, summed in slot order
      const double a0 = part[(size_t)p * j.part_stride], a1 = part[(size_t)(p + 1) * j.part_stride];
      const double a2 = part[(size_t)(p + 2) * j.part_stride], a3 = part[(size_t)(p + 3) * j.part_stride];
      v += a0; v += a1; v += a2; v += a3;
   }
   for (; p < j.nparts; p++) v += part[(size_t)p * j.part_stride];
   double* __restrict__ C = bases.p[j.dst_space] + j.dst_off;
   const int r = e % j.mrem, c = e / j.mrem;
   C[(size_t)(j.m0 + r) + (size_t)(j.n0 + c) * j.ldc] += v;
}
int dev_launch_reduce(const ReduceJob* d_jobs, int njobs, const DevBases& bases, void* stream) {
   if (njobs <= 0) return 0;
   for (int j0 = 0; j0 < njobs; j0 += 32768) {
      const int nj = (njobs - j0 < 32768) ? njobs - j0 : 32768;
      dim3 grid(nj, 16);   // 16 x 256 threads cover the largest (64 x 64) tile
      k_reduce<<<grid, 256, 0, (cudaStream_t)stream>>>(d_jobs + j0, bases);
   }
   cudaError_t e = cudaGetLastError();
   if (e != cudaSuccess) return cuda_fail(e, "k_reduce launch");
   return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Heff diagonal: 32x32 output tile per CTA, rank-1 updates from gathered operator-block diagonals (HBM/L2 gather bound)
__global__ void __launch_bounds__(256) k_diag(const DiagTile* __restrict__ tiles, const DiagItem* __restrict__ items, DevBases bases, double* __restrict__ out) {
   constexpr int T = 32, CH = 16;
   __shared__ double as[CH][T + 1], bs[CH][T + 1];
   const DiagTile t = tiles[blockIdx.x];
   const int tid = threadIdx.x, i = tid % T, j0 = tid / T;   // thread owns rows i, cols j0, j0+8, j0+16, j0+24
   double acc[4] = {0.0, 0.0, 0.0, 0.0};
   for (int c0 = t.item_begin; c0 < t.item_end; c0 += CH) {
      for (int idx = tid; idx < 2 * CH * T; idx += 256) {
         const int which = idx / (CH * T), rem = idx % (CH * T), c = rem / T, e = rem % T;
         double v = 0.0;
         if (c0 + c < t.item_end) {
            const DiagItem I = items[c0 + c];
            if (which == 0) {
               if (e < t.mrem) v = I.f * (I.as ? bases.p[I.as][I.aoff + (size_t)(t.m0 + e) * (I.lda + 1)] : 1.0);
            } else {
               if (e < t.nrem) v = I.bs ? bases.p[I.bs][I.boff + (size_t)(t.n0 + e) * (I.ldb + 1)] : 1.0;
            }
         }
         if (which == 0) as[c][e] = v; else bs[c][e] = v;
      }
      __syncthreads();
#pragma unroll
      for (int c = 0; c < CH; c++) {
         const double a = as[c][i];
#pragma unroll
         for (int r = 0; r < 4; r++) acc[r] += a * bs[c][j0 + 8 * r];
      }
      __syncthreads();
   }
   if (i < t.mrem)
#pragma unroll
      for (int r = 0; r < 4; r++)
         if (j0 + 8 * r < t.nrem) out[t.coff + (size_t)(t.m0 + i) + (size_t)(t.n0 + j0 + 8 * r) * t.ldc] = acc[r];
}
int dev_launch_diag(const DiagTile* d_tiles, int ntiles, const DiagItem* d_items, const DevBases& bases, double* d_out, void* stream) {
   if (ntiles <= 0) return 0;
   k_diag<<<ntiles, 256, 0, (cudaStream_t)stream>>>(d_tiles, d_items, bases, d_out);
   cudaError_t e = cudaGetLastError();
   if (e != cudaSuccess) return cuda_fail(e, "k_diag launch");
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
