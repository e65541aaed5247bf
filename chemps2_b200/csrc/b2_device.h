// b2_device.h — device-side work lists shared between the host plan compiler (b2_heff.cpp) and the kernels
// (b2_kernels.cu).  Plain structs, no CUDA types, so host-only translation units can include it.
#pragma once
#include <cstdint>

namespace b2 {

// address spaces an operand can live in; resolved to base pointers at launch time
// sigma build: LEFT/RIGHT = operator arenas of the two boundaries, VIN/VOUT = S and sigma.
// operator update: LEFT = old operator arena, RIGHT = MPS site tensor, VOUT = new operator arena.
enum Space : uint8_t { SP_NONE = 0, SP_LEFT = 1, SP_RIGHT = 2, SP_PRESUM = 3, SP_WORK = 4, SP_VIN = 5, SP_VOUT = 6, SP_PART = 7, SP_COUNT = 8 };

struct DevBases { double* p[SP_COUNT]; };

enum ItemKind : uint8_t { IT_GEMM = 0, IT_AXPY = 1 };
enum ItemFlags : uint8_t { IF_TX = 1, IF_TY = 2, IF_AXPY = 4 };

// C_tile += alpha * opX(X)[M x K] * opY(Y)[K x N]     (GEMM)
// C_tile += alpha * opX(X)[M x N]                     (IF_AXPY)
// X, Y are column-major with leading dimensions ldx, ldy; IF_TX / IF_TY: the stored matrix enters transposed.
struct GemmItem {
   int64_t xoff, yoff;
   double alpha;
   int32_t ldx, ldy, k;
   uint8_t xs, ys, flags, pad;
};
static_assert(sizeof(GemmItem) == 40, "GemmItem layout");

// One unit of CTA work: the partial sum over items[item_begin, item_end) for rows [m0, m0+mrem) x cols [n0, n0+nrem)
// of a target matrix.  The result goes to the column-major matrix at (cspace, coff, ldc), rows from cm0, cols from cn0:
//   accumulate = 0:  C  = acc   (stage-1 intermediates in the workspace; split-K partial slots)
//   accumulate = 1:  C += acc   (sigma; at most one CTA per launch touches a given sigma tile => deterministic, no atomics)
struct Tile {
   int64_t coff;
   int32_t ldc, m0, n0, mrem, nrem, cm0, cn0;
   int32_t item_begin, item_end;
   uint8_t cspace, accumulate, pad[2];
};
static_assert(sizeof(Tile) == 48, "Tile layout");

// sigma tile += sum_{p < nparts} partial slot p   (fixed order => deterministic)
struct ReduceJob {
   int64_t dst_off, part_off;
   int32_t ldc, m0, n0, mrem, nrem, nparts;
   int64_t part_stride;
   uint8_t dst_space, pad[7];
};

// out[dst_off + e] = sum_{parts} coef * src[e]   for e < size
struct PresumPart { int64_t src_off; double coef; uint8_t space, pad[7]; };
struct PresumJob { int64_t dst_off, size; int32_t part_begin, part_end; };

// Diagonal of H_eff: diag[dst tile](i, j) = sum_items f * a(i) * b(j) with a(i) = A[i*(lda+1)] (or 1), b(j) likewise
struct DiagItem { int64_t aoff, boff; double f; int32_t lda, ldb; uint8_t as, bs, pad[6]; };
struct DiagTile { int64_t coff; int32_t ldc, m0, n0, mrem, nrem, item_begin, item_end, pad; };

// tile classes: CTA tile edge and threads per CTA
constexpr int kNumTileClasses = 4;
constexpr int kTileEdge[kNumTileClasses] = {64, 32, 16, 8};
constexpr int kTileThreads[kNumTileClasses] = {128, 128, 32, 32};

// ---- launchers implemented in b2_kernels.cu (all asynchronous on `stream`, a cudaStream_t passed as void*)
int dev_launch_tiles(int tile_class, const Tile* d_tiles, int ntiles, const GemmItem* d_items, const DevBases& bases, void* stream);
// tiles whose items are all block axpys (IF_AXPY): the mixing pass of the operator update
int dev_launch_axpy_tiles(const Tile* d_tiles, int ntiles, const GemmItem* d_items, const DevBases& bases, void* stream);
// whole-operator mixing of one layout group: out[dst_off[d] + e] += sum_s coef[s * nd + d] * base(src_space[s])[src_off[s] + e], e < size
int dev_launch_mix_flat(const int64_t* d_dst_off, int nd, const int64_t* d_src_off, const uint8_t* d_src_space, int ns, const double* d_coef, int64_t size, const DevBases& bases,
                        void* stream);
int dev_launch_reduce(const ReduceJob* d_jobs, int njobs, const DevBases& bases, void* stream);
int dev_launch_diag(const DiagTile* d_tiles, int ntiles, const DiagItem* d_items, const DevBases& bases, double* d_out, void* stream);
int dev_launch_presum(const PresumJob* d_jobs, int njobs, const PresumPart* d_parts, const DevBases& bases, void* stream);
int dev_fill_zero(double* d_ptr, int64_t n, void* stream);
// p[e] = amp * hash(seed, key, e): deterministic synthetic operator contents (bench / full-size parity vs oracle/ref_driver synth)
int dev_fill_hash(double* d_ptr, int64_t n, uint64_t seed, uint64_t key, double amp, void* stream);
inline double hash_value(uint64_t seed, uint64_t key, uint64_t e) {
   uint64_t z = seed + 0x9E3779B97F4A7C15ULL * (key + 1) + 0xD1B54A32D192ED03ULL * (e + 1);
   z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
   z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
   z = z ^ (z >> 31);
   return (double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5;
}
// ---- Davidson vector algebra (b2_blas1.cu).  All scalars stay on the device unless stated otherwise; every reduction is a
// fixed-order two-level tree (deterministic).  `scratch` holds >= kRedScratch doubles + one counter per call site.
constexpr int kRedBlocks = 1184;                // 8 x 148 SMs: full residency of 256-thread CTAs
constexpr int kMaxVec = 32;                     // Davidson MAX_NUM_VEC (Options.h:70)
constexpr int kRedScratch = kRedBlocks * (kMaxVec + 2) + 64;
struct Coefs { double c[kMaxVec]; };
// out[j] = <x, Y_j>, Y_j = ybase + j*ystride, j < m
int dev_multi_dot(const double* x, const double* ybase, int64_t ystride, int m, int64_t n, double* out, double* scratch, void* stream);
// y += sign * coef[0] * x   (coef on device)
int dev_axpy_dev(double* y, const double* x, const double* coef, double sign, int64_t n, void* stream);
// y += sign * sum_{j<m} coef[j] * X_j, X_j = xbase + j*xstride (coef on device, m <= kMaxVec)
int dev_multi_axpy_dev(double* y, const double* xbase, int64_t xstride, int m, const double* coef, double sign, int64_t n, void* stream);
// y += x .* x   (diagonal of the excited-state projector, HeffDiagonal.cpp:621-640)
int dev_add_square(double* y, const double* x, int64_t n, void* stream);
// x *= 1/sqrt(ss[0])
int dev_scale_rsqrt(double* x, const double* ss, int64_t n, void* stream);
// u = sum_j a.c[j] V_j ; t = sum_j a.c[j] HV_j - theta*u ; out[0] = ||t||^2
int dev_ritz_residual(double* u, double* t, const double* V, const double* HV, int64_t stride, int m, Coefs a, double theta, int64_t n,
                      double* out, double* scratch, void* stream);
// work = u / clamp(diag - theta) ; out[0] = <work,t>, out[1] = <work,u>     (Davidson.cpp:328-338)
int dev_precond_dots(double* work, const double* u, const double* t, const double* diag, double theta, double cutoff, int64_t n, double* out,
                     double* scratch, void* stream);
// t = -(t - (out[0]/out[1]) u) / clamp(diag - theta)                          (Davidson.cpp:339-348)
int dev_precond_apply(double* t, const double* u, const double* diag, const double* dots, double theta, double cutoff, int64_t n, void* stream);
// out = sum_j a.c[j] V_j
int dev_lincomb(double* out, const double* V, int64_t stride, int m, Coefs a, int64_t n, void* stream);
// x[off_k .. off_k+len_k) *= scale_k for every block k (prog2symm / symm2prog, Sobject.cpp:624-650)
int dev_scale_blocks(double* x, const int64_t* d_off, const double* d_scale, int nblocks, void* stream);

// FP64 peak probes: returns achieved TFLOP/s of a register-resident DMMA (m8n8k4) / DFMA loop
int dev_probe_fp64(int use_mma, double* tflops_out);
const char* dev_last_error();

}   // namespace b2
