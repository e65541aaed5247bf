// b2_kernels.cu — sm_100a kernels of the sigma build / operator update.
//
// k_tiles<TM,TN,WM,WN>: grouped FP64 contraction.  One CTA owns one output tile of one symmetry block and
//   accumulates EVERY term that lands on it in registers (deterministic, no atomics):
//       C_tile = sum_items alpha * opX(X) * opY(Y)
//   Operand panels are staged through shared memory (one odd-stride layout, conflict free for the staging writes of both
//   operand orientations and for the DMMA fragment loads) and multiplied with FP64 tensor-core MMA
//   (mma.sync.aligned.m8n8k4.f64 -> SASS DMMA).  tcgen05/UMMA has no FP64 path, so warp-level DMMA is the
//   tensor pipe this workload can use on sm_100a.
// k_presum: integral-weighted operator pre-sums (HBM-bound, vectorised, coalesced).
#include <cuda_runtime.h>

#include "b2_pool.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <type_traits>

#include "b2_device.h"

namespace b2 {

static thread_local char g_dev_err[256] = "";
const char* dev_last_error() { return g_dev_err; }
static int cuda_fail(cudaError_t e, const char* what) {
   snprintf(g_dev_err, sizeof(g_dev_err), "%s: %s", what, cudaGetErrorString(e));
   return -3;
}

__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
   asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int KC = 16;       // k-chunk staged per pipeline stage

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
   const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
   asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async8_zfill(double* smem_dst, const double* gsrc, bool valid) {
   const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
   const int bytes = valid ? 8 : 0;   // src-size 0 => nothing is read, the 8 destination bytes are zero-filled
   asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gsrc), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Shared-memory operand panels: ONE layout whatever the transposition flag of the operand, P[k][r] with an ODD row
// stride RS = T + 1 doubles.  Together with the k-interleaved MMA steps (step s multiplies k = s, s+4, s+8, s+12, i.e.
// fragment lane q holds k = s + 4q) every access pattern of a half-warp hits 16 distinct 8-byte bank pairs:
//   fragment loads     (4 q) x (4 rows):  (4q*RS + g) mod 16  -> 4*(q*RS mod 4) + g        distinct because RS is odd
//   row-contiguous staging: 16 consecutive r at one k                                       consecutive addresses
//   k-contiguous staging:   16 consecutive k at one r:  k*RS mod 16                         distinct because RS is odd
// and every address is (lane part) + (compile-time part): no swizzle arithmetic in the inner loop.
template <int T> struct Panel {
   static constexpr int RS = T + 1;
   static constexpr int SIZE = KC * RS;
};

// Per-item staging cursor of one T x KC operand panel copied by NT threads, E = T*KC/NT elements per thread and chunk,
// one 8-byte cp.async each.  The thread -> element map keeps the GLOBAL reads in as few 128-byte lines as possible (the
// LSU spends one shared-memory wavefront per line a cp.async touches):
//   rows contiguous (stored R x K):  r = tid % T, k = tid / T + e*(NT/T)     a warp reads 32 consecutive rows of one column
//   k contiguous    (stored K x R):  k = tid % 16, r = tid / 16 + e*(NT/16)  a half-warp reads the 16 k of one row
// Element e sits at a compile-time shared-memory offset from element 0 and one constant pointer bump further in global
// memory.  Rows past the tile edge: the row-contiguous map clamps its (single) row, the k-contiguous map zero-fills them
// (rowmask); k past the end of the item is zero-filled in the tail chunk.
template <int T, int NT> struct Stager {
   static constexpr int E = (T * KC) / NT;
   static_assert((T * KC) % NT == 0 && E >= 2 && E <= 16 && NT % 16 == 0 && NT % T == 0, "panel must split evenly over the CTA");
   static constexpr unsigned FULLMASK = (1u << E) - 1u;
   static constexpr int RS = Panel<T>::RS;
   static constexpr int KSTEP = NT / T;     // row-contiguous map: k advance per element
   static constexpr int RSTEP = NT / 16;    // k-contiguous map: row advance per element
   const double* g;         // global address of element 0 of the current chunk
   int estep;               // pointer bump element -> element (doubles)
   int soff;                // shared-memory slot of element 0
   unsigned rowmask;        // bit e: the row of element e is inside the tile; bit 31: rows contiguous

   __device__ __forceinline__ void init(const double* G, int ld, bool contig_r, int r0, int rrem, int tid) {
      if (contig_r) {
         const int r = min(tid % T, rrem - 1);
         const int kt = tid / T;
         g = G + (size_t)(r0 + r) + (size_t)kt * ld;
         estep = KSTEP * ld;
         soff = kt * RS + tid % T;
         rowmask = FULLMASK | 0x80000000u;
      } else {
         const int kt = tid & 15;
         const int rt = tid >> 4;
         g = G + (size_t)kt + (size_t)(r0 + rt) * ld;
         estep = RSTEP * ld;
         soff = kt * RS + rt;
         rowmask = 0;
#pragma unroll
         for (int e = 0; e < E; e++) rowmask |= (rt + RSTEP * e < rrem) ? (1u << e) : 0u;
      }
   }
   // copies the chunk whose first column is the cursor position; kleft > 0 columns of the item are left
   __device__ __forceinline__ void chunk(double* P, int kleft, const double* safe, int tid) {
      const double* p = g;
      double* s = P + soff;
      const bool contig = (rowmask >> 31) != 0u;
      if (kleft >= KC && (rowmask & FULLMASK) == FULLMASK) {
         if (contig) {
#pragma unroll
            for (int e = 0; e < E; e++) { cp_async8(s + e * KSTEP * RS, p); p += estep; }
         } else {
#pragma unroll
            for (int e = 0; e < E; e++) { cp_async8(s + e * RSTEP, p); p += estep; }
         }
      } else {
         if (contig) {
            const int kt = tid / T;
#pragma unroll
            for (int e = 0; e < E; e++) {
               const bool ok = kt + e * KSTEP < kleft;
               cp_async8_zfill(s + e * KSTEP * RS, ok ? p : safe, ok);
               p += estep;
            }
         } else {
            const bool kok = (tid & 15) < kleft;
#pragma unroll
            for (int e = 0; e < E; e++) {
               const bool ok = kok && ((rowmask >> e) & 1u);
               cp_async8_zfill(s + e * RSTEP, ok ? p : safe, ok);
               p += estep;
            }
         }
      }
      g += contig ? (long long)estep * (KC / KSTEP) : (long long)KC;
   }
};

// one MMA step of a k-chunk: step S multiplies k = S + 4q (q = fragment lane).  MIE x NIE of the warp's MI x NI 8x8 sub-tiles lie (partly)
// inside the tile and are computed — compile-time counts, so a warp on a ragged tile edge issues exactly the DMMAs it needs without
// predicates (sub-tile rows / columns beyond the tile see clamped or zero-filled panel rows and are never stored).  SCALE: alpha != 1.
template <int TM, int TN, int MI, int NI, int MIE, int NIE, bool SCALE, int S>
__device__ __forceinline__ void mma_step(double (&acc)[MI][NI][2], const double* __restrict__ xa, const double* __restrict__ yb, double alpha, int mi_n, int ni_n) {
   constexpr int RSX = Panel<TM>::RS, RSY = Panel<TN>::RS;
   constexpr bool RUNTIME = MIE < 0;                       // generic path: warp-uniform runtime counts mi_n x ni_n, one predicate per sub-tile
   constexpr int ME = RUNTIME ? MI : MIE, NE = RUNTIME ? NI : NIE;
   double a[ME > 0 ? ME : 1], b[NE > 0 ? NE : 1];
#pragma unroll
   for (int i = 0; i < ME; i++) a[i] = SCALE ? alpha * xa[S * RSX + i * 8] : xa[S * RSX + i * 8];
#pragma unroll
   for (int j = 0; j < NE; j++) b[j] = yb[S * RSY + j * 8];
#pragma unroll
   for (int i = 0; i < ME; i++)
#pragma unroll
      for (int j = 0; j < NE; j++)
         if (!RUNTIME || (i < mi_n && j < ni_n)) dmma8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
}

// one k-chunk: the four MMA steps with the cp.async copies of a later chunk issued in between (the copies do not depend on
// the MMAs; spreading them over the chunk keeps the DMMA pipe fed right after the barrier).  FULLK: all KC columns valid,
// else steps >= kvalid are skipped (columns >= kvalid are zero).
template <int TM, int TN, int MI, int NI, int MIE, int NIE, bool SCALE, bool FULLK, class FX, class FY>
__device__ __forceinline__ void mma_chunk(double (&acc)[MI][NI][2], const double* __restrict__ xa, const double* __restrict__ yb, int kvalid, double alpha,
                                          int mi_n, int ni_n, FX&& stage_x, FY&& stage_y) {
   mma_step<TM, TN, MI, NI, MIE, NIE, SCALE, 0>(acc, xa, yb, alpha, mi_n, ni_n);
   stage_x();
   if (FULLK || 1 < kvalid) mma_step<TM, TN, MI, NI, MIE, NIE, SCALE, 1>(acc, xa, yb, alpha, mi_n, ni_n);
   stage_y();
   if (FULLK || 2 < kvalid) mma_step<TM, TN, MI, NI, MIE, NIE, SCALE, 2>(acc, xa, yb, alpha, mi_n, ni_n);
   if (FULLK || 3 < kvalid) mma_step<TM, TN, MI, NI, MIE, NIE, SCALE, 3>(acc, xa, yb, alpha, mi_n, ni_n);
}

// KSUB k-chunks are consumed per barrier (the pipeline unit), NSTG units are in flight; CPS = resident CTAs per SM the launch bounds ask for.
template <int TM, int TN, int WM, int WN, int KSUB = 1, int NSTG = 3, int CPS = 4, bool HYB = false>
__global__ void __launch_bounds__(WM * WN * 32, (TM == 64) ? CPS : 1) k_tiles(const Tile* __restrict__ tiles, const GemmItem* __restrict__ items, DevBases bases) {
   constexpr int NT = WM * WN * 32;
   constexpr int STAGES = KSUB * NSTG;   // panel buffers
   constexpr int WTM = TM / WM, WTN = TN / WN;   // warp tile
   constexpr int MI = WTM / 8, NI = WTN / 8;     // 8x8 MMA tiles per warp
   constexpr int XSZ = Panel<TM>::SIZE, YSZ = Panel<TN>::SIZE;
   extern __shared__ double smem[];
   double* Xs = smem;                    // STAGES panels
   double* Ys = smem + STAGES * XSZ;

   const Tile* __restrict__ tp = tiles + blockIdx.x;
   const int m0 = tp->m0, n0 = tp->n0, mrem = tp->mrem, nrem = tp->nrem, item_end = tp->item_end;
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   // the warp -> sub-tile map alternates with the CTA index: partial tiles always leave the LAST warp row / column short,
   // and warp w of every CTA lives on scheduler w % 4, so a fixed map would systematically underload two of the four
   const int wflip = (int)(blockIdx.x & 1u);
   const int wm = (warp % WM) ^ (wflip & (WM - 1)), wn = (warp / WM) ^ (wflip & (WN - 1));
   const int g = lane >> 2, q = lane & 3;        // fragment coordinates

   double acc[MI][NI][2];
#pragma unroll
   for (int i = 0; i < MI; i++)
#pragma unroll
      for (int j = 0; j < NI; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

   // ---- GEMM items: one flattened stream of k-chunks over all items, software-pipelined with cp.async.  Block-axpy
   // items come first in every item range (b2_compile.cpp sorts them there).
   int it0 = tp->item_begin;
   while (it0 < item_end && (items[it0].flags & IF_AXPY)) it0++;
   const double* safe = reinterpret_cast<const double*>(tiles);   // valid, aligned dummy source of zero-filled copies
   int p_it = it0, p_left = 0;          // producer cursor: item, columns of K still to stage
   Stager<TM, NT> sx;
   Stager<TN, NT> sy;
   auto producer_load_item = [&]() {
      const GemmItem P = items[p_it];
      sx.init(bases.p[P.xs] + P.xoff, P.ldx, !(P.flags & IF_TX), m0, mrem, tid);
      sy.init(bases.p[P.ys] + P.yoff, P.ldy, (P.flags & IF_TY) != 0, n0, nrem, tid);
      p_left = P.k;
   };
   if (p_it < item_end) producer_load_item();
   auto stage_x_at = [&](int ps) {
      if (p_it < item_end) sx.chunk(Xs + ps * XSZ, p_left, safe, tid);
   };
   auto stage_y_at = [&](int ps) {
      if (p_it < item_end) {
         sy.chunk(Ys + ps * YSZ, p_left, safe, tid);
         p_left -= KC;
         if (p_left <= 0 && ++p_it < item_end) producer_load_item();
      }
      cp_async_commit();
   };
#pragma unroll
   for (int s = 0; s < STAGES - KSUB; s++) { stage_x_at(s); stage_y_at(s); }

   // ---- block-axpy items (their global loads overlap the first panel copies)
   for (int it = tp->item_begin; it < it0; it++) {
      const GemmItem I = items[it];
      const double* __restrict__ X = bases.p[I.xs] + I.xoff;
#pragma unroll
      for (int i = 0; i < MI; i++)
#pragma unroll
         for (int j = 0; j < NI; j++) {
            const int r = wm * WTM + i * 8 + g, c = wn * WTN + j * 8 + 2 * q;
            if (r < mrem) {
               if (I.flags & IF_TX) {
                  if (c < nrem) acc[i][j][0] += I.alpha * X[(size_t)(n0 + c) + (size_t)(m0 + r) * I.ldx];
                  if (c + 1 < nrem) acc[i][j][1] += I.alpha * X[(size_t)(n0 + c + 1) + (size_t)(m0 + r) * I.ldx];
               } else {
                  if (c < nrem) acc[i][j][0] += I.alpha * X[(size_t)(m0 + r) + (size_t)(n0 + c) * I.ldx];
                  if (c + 1 < nrem) acc[i][j][1] += I.alpha * X[(size_t)(m0 + r) + (size_t)(n0 + c + 1) * I.ldx];
               }
            }
         }
   }

   int c_it = it0, c_left = 0, stage = 0;
   double alpha = 1.0;
   auto consumer_load_item = [&]() {
      const GemmItem* __restrict__ Cn = items + c_it;
      c_left = Cn->k; alpha = Cn->alpha;
   };
   if (c_it < item_end) consumer_load_item();
   const int xfrag = 4 * q * Panel<TM>::RS + wm * WTM + g, yfrag = 4 * q * Panel<TN>::RS + wn * WTN + g;
   // 8x8 sub-tile rows / columns of this warp that lie (partly) inside the tile (warp-uniform); the main loop is instantiated per count
   const int mi_n = min(MI, max(0, (mrem - wm * WTM + 7) >> 3)), ni_n = min(NI, max(0, (nrem - wn * WTN + 7) >> 3));
   auto main_loop = [&](auto mie_tag, auto nie_tag) {
      constexpr int MIE = decltype(mie_tag)::value, NIE = decltype(nie_tag)::value;
      while (c_it < item_end) {
         cp_async_wait<STAGES - 2 * KSUB>();
         __syncthreads();          // the KSUB chunks from `stage` on have landed; everybody is done with the buffers refilled below
#pragma unroll
         for (int sub = 0; sub < KSUB; sub++) {
         if (KSUB > 1 && c_it >= item_end) {   // the item stream ended inside the unit: keep the group accounting of the producer
            cp_async_commit();
            continue;
         }
         const double* xa = Xs + stage * XSZ + xfrag;
         const double* yb = Ys + stage * YSZ + yfrag;
         const int ps = (stage < KSUB) ? stage + STAGES - KSUB : stage - KSUB;   // a buffer consumed in the previous unit is refilled
         auto stage_x = [&]() { stage_x_at(ps); };
         auto stage_y = [&]() { stage_y_at(ps); };
         if (MIE < 0) {            // generic path: one code variant (scaled, any chunk length) keeps the kernel at 128 registers without spills
            mma_chunk<TM, TN, MI, NI, MIE, NIE, true, false>(acc, xa, yb, c_left, alpha, mi_n, ni_n, stage_x, stage_y);
         } else if (c_left >= KC) {
            // alpha == 1 compared on the bit pattern: an FP64 compare would queue behind the DMMAs
            if (__double2hiint(alpha) == 0x3FF00000 && __double2loint(alpha) == 0)
               mma_chunk<TM, TN, MI, NI, MIE, NIE, false, true>(acc, xa, yb, KC, alpha, mi_n, ni_n, stage_x, stage_y);
            else mma_chunk<TM, TN, MI, NI, MIE, NIE, true, true>(acc, xa, yb, KC, alpha, mi_n, ni_n, stage_x, stage_y);
         } else {
            mma_chunk<TM, TN, MI, NI, MIE, NIE, true, false>(acc, xa, yb, c_left, alpha, mi_n, ni_n, stage_x, stage_y);
         }
         c_left -= KC;
         if (c_left <= 0 && ++c_it < item_end) consumer_load_item();
         stage = (stage + 1 == STAGES) ? 0 : stage + 1;
         }
      }
   };
   // dispatch on the (warp-uniform) sub-tile counts of this warp: the full warp tile and — for the 64 x 64 class — the three ragged shapes
   // that would otherwise waste a quarter of their DMMAs get compile-time loops; everything else takes the predicated generic path
   using I4 = std::integral_constant<int, 4>;
   using I3 = std::integral_constant<int, 3>;
   using IR = std::integral_constant<int, -1>;
   if constexpr (HYB) {
      if (mi_n == MI && ni_n == NI) main_loop(std::integral_constant<int, MI>{}, std::integral_constant<int, NI>{});
      else if (MI == 4 && NI == 4 && mi_n == 3 && ni_n == 4) main_loop(I3{}, I4{});
      else if (MI == 4 && NI == 4 && mi_n == 4 && ni_n == 3) main_loop(I4{}, I3{});
      else if (MI == 4 && NI == 4 && mi_n == 3 && ni_n == 3) main_loop(I3{}, I3{});
      else main_loop(IR{}, IR{});
   } else {   // two paths: a warp with >= 3/4 of its sub-tiles inside computes all of them without predicates, the others take the generic path
      if (mi_n * ni_n * 4 >= MI * NI * 3) main_loop(std::integral_constant<int, MI>{}, std::integral_constant<int, NI>{});
      else main_loop(IR{}, IR{});
   }
   cp_async_wait<0>();

   const Tile t = *tp;
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

template <int TM, int TN, int WM, int WN, int KSUB = 1, int NSTG = 3, int CPS = 4, bool HYB = false>
static cudaError_t launch_tiles_t(const Tile* d_tiles, int ntiles, const GemmItem* d_items, const DevBases& bases, cudaStream_t s) {
   constexpr size_t smem = sizeof(double) * KSUB * NSTG * (Panel<TM>::SIZE + Panel<TN>::SIZE);
   // the opt-in above 48 KiB is a per-DEVICE function attribute: one flag per device ordinal (a process may hold contexts on several)
   static std::atomic<bool> configured[64];
   int dev = 0;
   cudaGetDevice(&dev);
   if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
      cudaError_t e = cudaFuncSetAttribute(k_tiles<TM, TN, WM, WN, KSUB, NSTG, CPS, HYB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_release);
   }
   k_tiles<TM, TN, WM, WN, KSUB, NSTG, CPS, HYB><<<ntiles, WM * WN * 32, smem, s>>>(d_tiles, d_items, bases);
   return cudaGetLastError();
}

int dev_launch_tiles(int tile_class, const Tile* d_tiles, int ntiles, const GemmItem* d_items, const DevBases& bases, void* stream) {
   if (ntiles <= 0) return 0;
   cudaStream_t s = (cudaStream_t)stream;
   cudaError_t e;
   switch (tile_class) {
      // pipeline variants measured in round 2 (profiles/r2_kernel_variants.md): 2 chunks per barrier at 2 or 3 CTAs/SM, 4 stages at 3 CTAs/SM,
      // 8-warp CTAs with 32 x 16 / 16 x 32 warp tiles at 3 or 4 CTAs/SM — all slower than 4 warps x (32 x 32), 3 stages, 4 CTAs/SM
      case 0: {
         static const bool hyb = getenv("B2_KHYBRID") && atoi(getenv("B2_KHYBRID")) != 0;   // experiment switch, see profiles/r2_kernel_variants.md
         e = hyb ? launch_tiles_t<64, 64, 2, 2, 1, 3, 4, true>(d_tiles, ntiles, d_items, bases, s) : launch_tiles_t<64, 64, 2, 2>(d_tiles, ntiles, d_items, bases, s);
         break;
      }
      case 1: e = launch_tiles_t<32, 32, 2, 2>(d_tiles, ntiles, d_items, bases, s); break;
      case 2: e = launch_tiles_t<16, 16, 1, 1>(d_tiles, ntiles, d_items, bases, s); break;
      case 3: e = launch_tiles_t<8, 8, 1, 1>(d_tiles, ntiles, d_items, bases, s); break;
      default: snprintf(g_dev_err, sizeof(g_dev_err), "bad tile class %d", tile_class); return -1;
   }
   if (e != cudaSuccess) return cuda_fail(e, "k_tiles launch");
   return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// k_axpy_tiles: tiles whose items are ALL block axpys — the A/B/C/D mixing pass of the operator update (DMRGoperators.cpp:367-405:
// TensorOperator::daxpy / daxpy_transpose_tensorCD of the two-operator tensors with integral weights).  HBM/L2-bound: one CTA streams
// the <= 64 x 64 destination tile once and adds every source tile; plain sources are read coalesced along the rows, transposed sources
// are read coalesced along THEIR rows into shared memory and added transposed.  The item record of the next source is fetched while
// the current one is added (no dependent-load chain per item); the launch order groups the tiles that share their sources
// (b2_capi_update.cpp), so the sources come from L2.
__global__ void __launch_bounds__(256) k_axpy_tiles(const Tile* __restrict__ tiles, const GemmItem* __restrict__ items, DevBases bases) {
   constexpr int EPT = 16;   // 64 x 64 / 256
   __shared__ double sT[64 * 65];
   const Tile t = tiles[blockIdx.x];
   const int tid = threadIdx.x, mrem = t.mrem, nrem = t.nrem, n = mrem * nrem;
   double acc[EPT];
   int rr[EPT], cc[EPT];
#pragma unroll
   for (int j = 0; j < EPT; j++) {
      const int idx = tid + 256 * j;
      acc[j] = 0.0;
      rr[j] = (idx < n) ? idx % mrem : -1;
      cc[j] = (idx < n) ? idx / mrem : 0;
   }
   GemmItem I = items[t.item_begin];
   for (int it = t.item_begin; it < t.item_end; it++) {
      const GemmItem cur = I;
      if (it + 1 < t.item_end) I = items[it + 1];
      const double* __restrict__ X = bases.p[cur.xs] + cur.xoff;
      if (!(cur.flags & IF_TX)) {
#pragma unroll
         for (int j = 0; j < EPT; j++)
            if (rr[j] >= 0) acc[j] += cur.alpha * X[(size_t)(t.m0 + rr[j]) + (size_t)(t.n0 + cc[j]) * cur.ldx];
      } else {
         __syncthreads();   // the previous transposed source has been consumed
         for (int idx = tid; idx < n; idx += 256) {   // stored block: rows = destination columns (contiguous), columns = destination rows
            const int c = idx % nrem, r = idx / nrem;
            sT[c * 65 + r] = X[(size_t)(t.n0 + c) + (size_t)(t.m0 + r) * cur.ldx];
         }
         __syncthreads();
#pragma unroll
         for (int j = 0; j < EPT; j++)
            if (rr[j] >= 0) acc[j] += cur.alpha * sT[cc[j] * 65 + rr[j]];
      }
   }
   double* __restrict__ C = bases.p[t.cspace] + t.coff;
#pragma unroll
   for (int j = 0; j < EPT; j++)
      if (rr[j] >= 0) {
         double* p = C + (size_t)(t.cm0 + rr[j]) + (size_t)(t.cn0 + cc[j]) * t.ldc;
         if (t.accumulate) *p += acc[j]; else *p = acc[j];
      }
}
int dev_launch_axpy_tiles(const Tile* d_tiles, int ntiles, const GemmItem* d_items, const DevBases& bases, void* stream) {
   if (ntiles <= 0) return 0;
   k_axpy_tiles<<<ntiles, 256, 0, (cudaStream_t)stream>>>(d_tiles, d_items, bases);
   cudaError_t e = cudaGetLastError();
   if (e != cudaSuccess) return cuda_fail(e, "k_axpy_tiles launch");
   return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// k_mix_flat: whole-operator mixing of one layout group as a tall-skinny GEMM,  Dst[e, d] += sum_s Src[e, s] * coef[s][d]  over the
// elements e of the (identical) operator layouts: nd destination operators (A / B / C / D of the outside pairs), ns sources (the
// two-operator tensors with a leg on the new site, plain or transposed copies).  HBM-bound: every destination element is read and
// written once, the sources are read once per 16 destinations (adjacent CTAs share them through L2).  One thread per element,
// 16 destination accumulators in registers, the coefficient tile in shared memory.
constexpr int MIX_DT = 16, MIX_SC = 64;
__global__ void __launch_bounds__(256) k_mix_flat(const int64_t* __restrict__ dst_off, int nd, const int64_t* __restrict__ src_off, const uint8_t* __restrict__ src_space, int ns,
                                                  const double* __restrict__ coef, int64_t size, DevBases bases) {
   __shared__ double c[MIX_SC * MIX_DT];
   __shared__ const double* sp[MIX_SC];
   const int d0 = blockIdx.x * MIX_DT, ndl = min(MIX_DT, nd - d0);
   const int64_t e = (int64_t)blockIdx.y * 256 + threadIdx.x;
   double acc[MIX_DT];
#pragma unroll
   for (int d = 0; d < MIX_DT; d++) acc[d] = 0.0;
   for (int s0 = 0; s0 < ns; s0 += MIX_SC) {
      const int nsl = min(MIX_SC, ns - s0);
      __syncthreads();
      for (int i = threadIdx.x; i < nsl * MIX_DT; i += 256) {
         const int sl = i / MIX_DT, d = i % MIX_DT;
         c[i] = (d < ndl) ? coef[(size_t)(s0 + sl) * nd + d0 + d] : 0.0;
      }
      if (threadIdx.x < nsl) sp[threadIdx.x] = bases.p[src_space[s0 + threadIdx.x]] + src_off[s0 + threadIdx.x];
      __syncthreads();
      if (e < size) {
         int sl = 0;
         for (; sl + 4 <= nsl; sl += 4) {   // four independent source loads in flight per thread
            const double x0 = sp[sl][e], x1 = sp[sl + 1][e], x2 = sp[sl + 2][e], x3 = sp[sl + 3][e];
#pragma unroll
            for (int d = 0; d < MIX_DT; d++)
               acc[d] += c[sl * MIX_DT + d] * x0 + c[(sl + 1) * MIX_DT + d] * x1 + c[(sl + 2) * MIX_DT + d] * x2 + c[(sl + 3) * MIX_DT + d] * x3;
         }
         for (; sl < nsl; sl++) {
            const double x = sp[sl][e];
#pragma unroll
            for (int d = 0; d < MIX_DT; d++) acc[d] += c[sl * MIX_DT + d] * x;
         }
      }
   }
   if (e < size) {
      double* __restrict__ out = bases.p[SP_VOUT];
#pragma unroll
      for (int d = 0; d < MIX_DT; d++)
         if (d < ndl) out[dst_off[d0 + d] + e] += acc[d];
   }
}
int dev_launch_mix_flat(const int64_t* d_dst_off, int nd, const int64_t* d_src_off, const uint8_t* d_src_space, int ns, const double* d_coef, int64_t size, const DevBases& bases,
                        void* stream) {
   if (nd <= 0 || ns <= 0 || size <= 0) return 0;
   const int64_t chunks = (size + 255) / 256;
   for (int64_t y0 = 0; y0 < chunks; y0 += 65535) {   // gridDim.y limit
      const int ny = (int)std::min<int64_t>(65535, chunks - y0);
      dim3 grid((nd + MIX_DT - 1) / MIX_DT, ny);      // x (destination tiles) varies fastest: the CTAs that read the same source elements run together
      DevBases b = bases;
      // element offset of this slab: shift the bases instead of passing another argument
      for (int i = 0; i < SP_COUNT; i++) if (b.p[i]) b.p[i] += y0 * 256;
      k_mix_flat<<<grid, 256, 0, (cudaStream_t)stream>>>(d_dst_off, nd, d_src_off, d_src_space, ns, d_coef, size - y0 * 256, b);
   }
   cudaError_t e = cudaGetLastError();
   if (e != cudaSuccess) return cuda_fail(e, "k_mix_flat launch");
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
      dim3 grid(32, nj);   // 32 CTAs per pre-summed operator: a site with few jobs still fills the chip
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
