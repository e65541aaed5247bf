// b2_capi.cpp — the C ABI declared in include/chemps2_b200.h.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/chemps2_b200.h"
#include "b2_core.h"
#include "b2_davidson.h"
#include "b2_device.h"
#include "b2_heff.h"
#include "b2_ops.h"
#include "b2_sigma.h"
#include "b2_sobject.h"
#include "b2_twodm.h"
#include "b2_update.h"

using namespace b2;

static thread_local std::string g_err;
static double wall_seconds() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static int fail(int code, const char* fmt, ...) {
   char buf[512];
   va_list ap;
   va_start(ap, fmt);
   vsnprintf(buf, sizeof(buf), fmt, ap);
   va_end(ap);
   g_err = buf;
   return code;
}
#define CUDA_TRY(call)                                                                         \
   do {                                                                                        \
      cudaError_t e_ = (call);                                                                 \
      if (e_ != cudaSuccess) return fail(B2_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
   } while (0)

struct b2_ctx {
   int device = -1;
   cudaStream_t stream = nullptr;
   bool own_stream = true;
   Problem prob;
   Bookkeeper bk;
   bool have_problem = false, have_bk = false;
   CompileOptions copt;
   int simulate_oom = 0;                // test hook: the next N device allocations of operator sets / plans report B2_ERR_CUDA
   // plans below this many reference FLOPs per apply are scheduled on all host cores: segment-wise scheduling shares stage-1 products
   // only inside a segment (+1-3 % executed FLOPs, measured on the N2/cc-pVDZ D=2000 and tetracene D=3000 shapes) but builds the plan
   // 2-4x faster, which wins as long as a Davidson solve (~15 sigma builds) is shorter than the planning it saves
   double parallel_plan_flops = 1e13;
};

struct b2_opset {
   b2_ctx* ctx = nullptr;
   OpSet set;
   std::vector<double> host;   // host mirror, allocated on first use (upload / download / planning-only contexts)
   double* dev = nullptr;
   double* spill = nullptr;    // pinned host copy while the set is offloaded (b2_opset_offload): the device arena is released
   bool offloaded = false;
   void ensure_host() { if (host.size() != (size_t)set.size) host.assign((size_t)set.size, 0.0); }
   ~b2_opset() { if (spill) cudaFreeHost(spill); if (dev) cudaFree(dev); }
};

struct b2_heff {
   b2_ctx* ctx = nullptr;
   b2_opset *left = nullptr, *right = nullptr;
   SigmaPlan plan;
   CompiledSigma comp;
   // device copies
   GemmItem *d_items1 = nullptr, *d_items2 = nullptr;
   ReduceJob* d_reduces = nullptr;
   DiagItem* d_diag_items = nullptr;
   DiagTile* d_diag_tiles = nullptr;
   int64_t* d_blk_off = nullptr;                 // Sobject block offsets (nkappa + 1)
   double *d_p2s = nullptr, *d_s2p = nullptr;    // sqrt(2SR+1) and its inverse per block (Sobject.cpp:624-650)
   double* d_part = nullptr;
   Tile* d_tiles1[kNumTileClasses] = {nullptr, nullptr, nullptr, nullptr};
   Tile* d_tiles2[kNumTileClasses] = {nullptr, nullptr, nullptr, nullptr};
   PresumJob* d_jobs = nullptr;
   PresumPart* d_parts = nullptr;
   double *d_presum = nullptr, *d_work = nullptr, *d_vin = nullptr, *d_vout = nullptr;
   double *h_vin = nullptr, *h_vout = nullptr;   // pinned staging
   cudaEvent_t ev0 = nullptr, ev1 = nullptr;
   b2_allreduce_fn allreduce = nullptr;          // sums partial sigma / diag vectors over the GPUs (NCCL in the caller)
   void* allreduce_user = nullptr;
   double last_kernel_s = 0.0;
   long long launches = 0;
   int world = 1, rank = 0;
   double list_bytes = 0.0;                      // size of the device work lists (decides whether the sweep driver keeps the plan)
   // excited states (Heff::addDiagramExcitations): n_exc level-shifted lower states, one vector of veclength doubles each
   int n_exc = 0;
   double *d_exc = nullptr, *d_exc_coef = nullptr, *d_exc_scratch = nullptr;
   ~b2_heff() {   // also runs when b2_heff_create bails out half-way (e.g. out of HBM): nothing leaks
      cudaFree(d_diag_items); cudaFree(d_diag_tiles); cudaFree(d_blk_off); cudaFree(d_p2s); cudaFree(d_s2p);
      cudaFree(d_items1); cudaFree(d_items2); cudaFree(d_reduces); cudaFree(d_part);
      for (int c = 0; c < kNumTileClasses; c++) { cudaFree(d_tiles1[c]); cudaFree(d_tiles2[c]); }
      cudaFree(d_jobs); cudaFree(d_parts); cudaFree(d_presum); cudaFree(d_work); cudaFree(d_vin); cudaFree(d_vout);
      cudaFree(d_exc); cudaFree(d_exc_coef); cudaFree(d_exc_scratch);
      if (h_vin) cudaFreeHost(h_vin);
      if (h_vout) cudaFreeHost(h_vout);
      if (ev0) cudaEventDestroy(ev0);
      if (ev1) cudaEventDestroy(ev1);
   }
};

// stage-1 workspace budget of a plan: the configured value, capped by half of the HBM that is free right now
static CompileOptions budgeted(const b2_ctx* ctx) {
   CompileOptions o = ctx->copt;
   if (ctx->device >= 0) {
      size_t free_b = 0, total_b = 0;
      if (cudaSetDevice(ctx->device) == cudaSuccess && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)
         o.work_budget = std::max<int64_t>((int64_t)1 << 22, std::min<int64_t>(o.work_budget, (int64_t)(free_b / 2 / sizeof(double))));
   }
   return o;
}

template <class T> static int upload_vec(T** dptr, const std::vector<T>& v, cudaStream_t s) {
   *dptr = nullptr;
   if (v.empty()) return B2_OK;
   CUDA_TRY(cudaMalloc(dptr, sizeof(T) * v.size()));
   CUDA_TRY(cudaMemcpyAsync(*dptr, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice, s));
   return B2_OK;
}

static DevBases bases_of(const b2_heff* h, const double* vin, double* vout) {
   DevBases b;
   for (int i = 0; i < SP_COUNT; i++) b.p[i] = nullptr;
   b.p[SP_LEFT] = h->left ? h->left->dev : nullptr;
   b.p[SP_RIGHT] = h->right ? h->right->dev : nullptr;
   b.p[SP_PRESUM] = h->d_presum;
   b.p[SP_WORK] = h->d_work;
   b.p[SP_PART] = h->d_part;
   b.p[SP_VIN] = const_cast<double*>(vin);
   b.p[SP_VOUT] = vout;
   return b;
}

extern "C" {

const char* b2_last_error(void) { return g_err.c_str(); }
const char* b2_version(void) { return "chemps2_b200 0.1 (sm_100a)"; }

int b2_ctx_create(int device, b2_ctx** out) {
   if (!out) return fail(B2_ERR_ARG, "b2_ctx_create: out is NULL");
   std::unique_ptr<b2_ctx> c(new b2_ctx);
   c->device = device;
   if (device >= 0) {
      int n = 0;
      cudaError_t e = cudaGetDeviceCount(&n);
      if (e != cudaSuccess || n <= device) return fail(B2_ERR_NO_DEVICE, "b2_ctx_create: CUDA device %d not available (%s)", device, cudaGetErrorString(e));
      CUDA_TRY(cudaSetDevice(device));
      CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
   }
   *out = c.release();
   return B2_OK;
}
void b2_ctx_destroy(b2_ctx* ctx) {
   if (!ctx) return;
   if (ctx->stream && ctx->own_stream) cudaStreamDestroy(ctx->stream);
   delete ctx;
}
int b2_ctx_device(const b2_ctx* ctx) { return ctx ? ctx->device : -1; }
int b2_ctx_set_stream(b2_ctx* ctx, void* cuda_stream) {
   if (!ctx || ctx->device < 0) return fail(B2_ERR_NO_DEVICE, "b2_ctx_set_stream: no CUDA device");
   if (ctx->stream && ctx->own_stream) cudaStreamDestroy(ctx->stream);
   ctx->stream = (cudaStream_t)cuda_stream; ctx->own_stream = false;
   return B2_OK;
}
void* b2_ctx_stream(const b2_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

static int set_problem_common(b2_ctx* ctx, int L, int group, int N, int twoS, int irrep, const int* orb_irrep, double econst) {
   if (!ctx || L < 2 || !orb_irrep) return fail(B2_ERR_ARG, "b2_problem_set: bad arguments");
   const int nirr = num_irreps_of_group(group);
   if (nirr < 0) return fail(B2_ERR_ARG, "b2_problem_set: group %d out of range", group);
   if (irrep < 0 || irrep >= nirr) return fail(B2_ERR_ARG, "b2_problem_set: target irrep %d out of range", irrep);
   if (N < 2) return fail(B2_ERR_ARG, "b2_problem_set: N must be >= 2 (one-body part is folded in with 1/(N-1))");
   for (int i = 0; i < L; i++)
      if (orb_irrep[i] < 0 || orb_irrep[i] >= nirr) return fail(B2_ERR_ARG, "b2_problem_set: orbital irrep out of range");
   Problem& p = ctx->prob;
   p.L = L; p.group = group; p.N = N; p.twoS = twoS; p.irrep = irrep; p.econst = econst;
   p.orb_irrep.assign(orb_irrep, orb_irrep + L);
   ctx->have_problem = true; ctx->have_bk = false;
   return B2_OK;
}
int b2_problem_set(b2_ctx* ctx, int L, int group, int N, int twoS, int irrep, const int* orb_irrep, const double* mx_elem, double econst) {
   if (!mx_elem) return fail(B2_ERR_ARG, "b2_problem_set: mx_elem is NULL");
   int rc = set_problem_common(ctx, L, group, N, twoS, irrep, orb_irrep, econst);
   if (rc) return rc;
   ctx->prob.mx.assign(mx_elem, mx_elem + (size_t)L * L * L * L);
   return B2_OK;
}
int b2_problem_set_integrals(b2_ctx* ctx, int L, int group, int N, int twoS, int irrep, const int* orb_irrep, const double* tmat,
                             const double* vmat, double econst) {
   if (!tmat || !vmat) return fail(B2_ERR_ARG, "b2_problem_set_integrals: NULL integrals");
   int rc = set_problem_common(ctx, L, group, N, twoS, irrep, orb_irrep, econst);
   if (rc) return rc;
   ctx->prob.build(tmat, vmat);
   return B2_OK;
}

int b2_problem_mx(const b2_ctx* ctx, double* mx_out) {
   if (!ctx || !ctx->have_problem || !mx_out) return fail(B2_ERR_STATE, "b2_problem_mx: no problem set");
   std::memcpy(mx_out, ctx->prob.mx.data(), sizeof(double) * ctx->prob.mx.size());
   return B2_OK;
}
double b2_wigner6j(int a, int b, int c, int d, int e, int f) { return wigner6j(a, b, c, d, e, f); }
double b2_wigner9j(int a, int b, int c, int d, int e, int f, int g, int h, int i) { return wigner9j(a, b, c, d, e, f, g, h, i); }

int b2_small_symmetric_eig(int n, const double* a, double* eval, double* evec) {
   if (n < 1 || n > 32 || !a || !eval || !evec) return fail(B2_ERR_ARG, "b2_small_symmetric_eig: bad arguments");
   small_symmetric_eig(n, a, n, eval, evec);
   return B2_OK;
}

int b2_bk_init(b2_ctx* ctx, int D) {
   if (!ctx || !ctx->have_problem) return fail(B2_ERR_STATE, "b2_bk_init: set the problem first");
   if (D < 1) return fail(B2_ERR_ARG, "b2_bk_init: D < 1");
   ctx->bk.init(ctx->prob, D);
   ctx->have_bk = true;
   if (!ctx->bk.is_possible()) return fail(B2_ERR_ARG, "b2_bk_init: target sector not reachable (SyBookkeeper::IsPossible)");
   return B2_OK;
}
int b2_bk_set_dim(b2_ctx* ctx, int boundary, int N, int twoS, int irrep, int dim) {
   if (!ctx || !ctx->have_bk) return fail(B2_ERR_STATE, "b2_bk_set_dim: no bookkeeper");
   ctx->bk.set_dim(boundary, N, twoS, irrep, dim);
   return B2_OK;
}
int b2_bk_dim(const b2_ctx* ctx, int b, int N, int twoS, int irrep) { return (ctx && ctx->have_bk) ? ctx->bk.dim(b, N, twoS, irrep) : 0; }
int b2_bk_fcidim(const b2_ctx* ctx, int b, int N, int twoS, int irrep) { return (ctx && ctx->have_bk) ? ctx->bk.fcidim(b, N, twoS, irrep) : 0; }
int b2_bk_nmin(const b2_ctx* ctx, int b) { return ctx->bk.Nmin[b]; }
int b2_bk_nmax(const b2_ctx* ctx, int b) { return ctx->bk.Nmax[b]; }
int b2_bk_twosmin(const b2_ctx* ctx, int b, int N) { return ctx->bk.tsmin[b][N - ctx->bk.Nmin[b]]; }
int b2_bk_twosmax(const b2_ctx* ctx, int b, int N) { return ctx->bk.tsmax[b][N - ctx->bk.Nmin[b]]; }

int64_t b2_tensor_t_size(const b2_ctx* ctx, int site) { TLayout t; t.build(ctx->bk, site); return t.size; }
int64_t b2_sobject_size(const b2_ctx* ctx, int site) { SLayout s; s.build(ctx->bk, site); return s.size; }
int b2_sobject_nkappa(const b2_ctx* ctx, int site) { SLayout s; s.build(ctx->bk, site); return s.nkappa(); }
int b2_sobject_table(const b2_ctx* ctx, int site, int* labels, int64_t* offsets) {
   SLayout s; s.build(ctx->bk, site);
   for (int k = 0; k < s.nkappa(); k++) {
      int* l = labels + 9 * k;
      l[0] = s.NL[k]; l[1] = s.twoSL[k]; l[2] = s.IL[k]; l[3] = s.N1[k]; l[4] = s.N2[k]; l[5] = s.twoJ[k]; l[6] = s.NR[k]; l[7] = s.twoSR[k]; l[8] = s.IR[k];
      offsets[k] = s.blk[k].off;
   }
   offsets[s.nkappa()] = s.size;
   return B2_OK;
}

// ------------------------------------------------------------------------------------------------ operator sets
int b2_opset_create(b2_ctx* ctx, int boundary, int moving_right, b2_opset** out) {
   if (!ctx || !ctx->have_bk || !out) return fail(B2_ERR_STATE, "b2_opset_create: no bookkeeper");
   if (boundary < 1 || boundary > ctx->bk.L - 1) return fail(B2_ERR_ARG, "b2_opset_create: boundary %d out of range", boundary);
   std::unique_ptr<b2_opset> s(new b2_opset);
   s->ctx = ctx;
   s->set.build_all(ctx->bk, boundary, moving_right != 0);
   if (ctx->device >= 0 && ctx->simulate_oom > 0) { ctx->simulate_oom--; return fail(B2_ERR_CUDA, "b2_opset_create: out of memory (simulated)"); }
   if (ctx->device >= 0 && s->set.size > 0) {
      CUDA_TRY(cudaSetDevice(ctx->device));
      CUDA_TRY(cudaMalloc(&s->dev, sizeof(double) * (size_t)s->set.size));
      CUDA_TRY(cudaMemsetAsync(s->dev, 0, sizeof(double) * (size_t)s->set.size, ctx->stream));
   }
   *out = s.release();
   return B2_OK;
}
static int opset_create_reduced(b2_ctx* ctx, int boundary, bool mr, bool only_L, b2_opset** out) {
   std::unique_ptr<b2_opset> s(new b2_opset);
   s->ctx = ctx;
   s->set.build_reduced(ctx->bk, boundary, mr, only_L);
   if (ctx->device >= 0 && s->set.size > 0) {
      CUDA_TRY(cudaSetDevice(ctx->device));
      CUDA_TRY(cudaMalloc(&s->dev, sizeof(double) * (size_t)s->set.size));
      CUDA_TRY(cudaMemsetAsync(s->dev, 0, sizeof(double) * (size_t)s->set.size, ctx->stream));
   }
   *out = s.release();
   return B2_OK;
}
int b2_opset_create_correlation(b2_ctx* ctx, int boundary, b2_opset** out) {
   if (!ctx || !ctx->have_bk || !out) return fail(B2_ERR_STATE, "b2_opset_create_correlation: no bookkeeper");
   if (boundary < 1 || boundary > ctx->bk.L) return fail(B2_ERR_ARG, "b2_opset_create_correlation: boundary %d out of range", boundary);
   std::unique_ptr<b2_opset> s(new b2_opset);
   s->ctx = ctx;
   s->set.build_correlation(ctx->bk, boundary);
   if (ctx->device >= 0 && s->set.size > 0) {
      CUDA_TRY(cudaSetDevice(ctx->device));
      CUDA_TRY(cudaMalloc(&s->dev, sizeof(double) * (size_t)s->set.size));
      CUDA_TRY(cudaMemsetAsync(s->dev, 0, sizeof(double) * (size_t)s->set.size, ctx->stream));
   }
   *out = s.release();
   return B2_OK;
}

/* Operator life-cycle (replaces DMRG::OperatorsOnDisk / deleteTensors / allocateTensors, DMRGoperators.cpp:33-231,1147-1433: the
 * reference spills the operator tables of the boundaries it is not working on to HDF5 files; here they go to pinned host memory
 * over PCIe/C2C and the HBM arena is released). */
int b2_opset_offload(b2_opset* set) {
   if (!set) return fail(B2_ERR_ARG, "b2_opset_offload: NULL");
   if (set->ctx->device < 0) return fail(B2_ERR_NO_DEVICE, "b2_opset_offload: planning-only context, no CUDA device");
   if (set->offloaded || !set->dev) return B2_OK;
   const size_t bytes = sizeof(double) * (size_t)set->set.size;
   if (!set->spill) CUDA_TRY(cudaMallocHost(&set->spill, bytes));
   CUDA_TRY(cudaMemcpyAsync(set->spill, set->dev, bytes, cudaMemcpyDeviceToHost, set->ctx->stream));
   CUDA_TRY(cudaStreamSynchronize(set->ctx->stream));
   CUDA_TRY(cudaFree(set->dev));
   set->dev = nullptr; set->offloaded = true;
   return B2_OK;
}
int b2_opset_reload(b2_opset* set) {
   if (!set) return fail(B2_ERR_ARG, "b2_opset_reload: NULL");
   if (!set->offloaded) return B2_OK;
   const size_t bytes = sizeof(double) * (size_t)set->set.size;
   CUDA_TRY(cudaSetDevice(set->ctx->device));
   CUDA_TRY(cudaMalloc(&set->dev, bytes));
   CUDA_TRY(cudaMemcpyAsync(set->dev, set->spill, bytes, cudaMemcpyHostToDevice, set->ctx->stream));
   CUDA_TRY(cudaStreamSynchronize(set->ctx->stream));
   cudaFreeHost(set->spill);
   set->spill = nullptr; set->offloaded = false;
   return B2_OK;
}
int b2_opset_resident(const b2_opset* set) { return (set && set->dev && !set->offloaded) ? 1 : 0; }

void b2_opset_destroy(b2_opset* set) { delete set; }
int b2_opset_count(const b2_opset* set) { return set ? (int)set->set.ops.size() : 0; }
int b2_opset_info(const b2_opset* set, int index, int* kind, int* si, int* sj, int64_t* size) {
   if (!set || index < 0 || index >= (int)set->set.ops.size()) return fail(B2_ERR_ARG, "b2_opset_info: bad index");
   const OpTensor& t = set->set.ops[index];
   if (kind) *kind = t.kind;
   if (si) *si = t.i;
   if (sj) *sj = t.j;
   if (size) *size = t.lay->size;
   return B2_OK;
}
int b2_opset_find(const b2_opset* set, int kind, int si, int sj) { return set ? set->set.find(kind, si, sj) : -1; }
int b2_opset_upload(b2_opset* set, int index, const double* packed) {
   if (!set || index < 0 || index >= (int)set->set.ops.size() || !packed) return fail(B2_ERR_ARG, "b2_opset_upload: bad arguments");
   const OpTensor& t = set->set.ops[index];
   if (t.lay->size == 0) return B2_OK;
   set->ensure_host();
   std::memcpy(set->host.data() + t.off, packed, sizeof(double) * (size_t)t.lay->size);
   if (set->offloaded) std::memcpy(set->spill + t.off, packed, sizeof(double) * (size_t)t.lay->size);
   if (set->dev) {
      CUDA_TRY(cudaMemcpyAsync(set->dev + t.off, set->host.data() + t.off, sizeof(double) * (size_t)t.lay->size, cudaMemcpyHostToDevice, set->ctx->stream));
      CUDA_TRY(cudaStreamSynchronize(set->ctx->stream));
   }
   return B2_OK;
}
int b2_opset_download(b2_opset* set, int index, double* packed) {
   if (!set || index < 0 || index >= (int)set->set.ops.size() || !packed) return fail(B2_ERR_ARG, "b2_opset_download: bad arguments");
   const OpTensor& t = set->set.ops[index];
   if (t.lay->size == 0) return B2_OK;
   set->ensure_host();
   if (set->dev) {
      CUDA_TRY(cudaMemcpyAsync(set->host.data() + t.off, set->dev + t.off, sizeof(double) * (size_t)t.lay->size, cudaMemcpyDeviceToHost, set->ctx->stream));
      CUDA_TRY(cudaStreamSynchronize(set->ctx->stream));
   }
   if (set->offloaded) std::memcpy(set->host.data() + t.off, set->spill + t.off, sizeof(double) * (size_t)t.lay->size);
   std::memcpy(packed, set->host.data() + t.off, sizeof(double) * (size_t)t.lay->size);
   return B2_OK;
}
int b2_opset_clear(b2_opset* set) {
   if (!set) return fail(B2_ERR_ARG, "b2_opset_clear: NULL");
   if (!set->host.empty()) std::fill(set->host.begin(), set->host.end(), 0.0);
   if (set->dev) CUDA_TRY(cudaMemsetAsync(set->dev, 0, sizeof(double) * (size_t)set->set.size, set->ctx->stream));
   return B2_OK;
}
const double* b2_opset_host_arena(const b2_opset* set) {
   if (!set) return nullptr;
   const_cast<b2_opset*>(set)->ensure_host();
   return set->host.data();
}
int b2_opset_fill_hash(b2_opset* set, uint64_t seed, double amp) {
   if (!set) return fail(B2_ERR_ARG, "b2_opset_fill_hash: NULL");
   const uint64_t side = set->set.moving_right ? 1 : 2;
   for (const OpTensor& t : set->set.ops) {
      const uint64_t key = (side << 60) | ((uint64_t)t.kind << 40) | ((uint64_t)(t.i + 1) << 20) | (uint64_t)(t.j + 1);
      if (set->dev) {
         if (dev_fill_hash(set->dev + t.off, t.lay->size, seed, key, amp, set->ctx->stream)) return fail(B2_ERR_CUDA, "%s", dev_last_error());
      } else {
         set->ensure_host();
         double* p = set->host.data() + t.off;
         for (int64_t e = 0; e < t.lay->size; e++) p[e] = amp * hash_value(seed, key, (uint64_t)e);
      }
   }
   if (set->dev) CUDA_TRY(cudaStreamSynchronize(set->ctx->stream));
   return B2_OK;
}
int b2_hash_fill(double* out, int64_t n, uint64_t seed, uint64_t key, double amp) {
   if (!out) return fail(B2_ERR_ARG, "b2_hash_fill: NULL");
   for (int64_t e = 0; e < n; e++) out[e] = amp * hash_value(seed, key, (uint64_t)e);
   return B2_OK;
}
int64_t b2_opset_arena_size(const b2_opset* set) { return set ? set->set.size : 0; }

// ------------------------------------------------------------------------------------------------ heff
int b2_heff_create(b2_ctx* ctx, int site, b2_opset* left, b2_opset* right, int world, int rank, b2_heff** out) {
   if (!ctx || !ctx->have_bk || !out) return fail(B2_ERR_STATE, "b2_heff_create: no bookkeeper");
   const int L = ctx->bk.L;
   if (site < 0 || site > L - 2) return fail(B2_ERR_ARG, "b2_heff_create: site %d out of range", site);
   if (site > 0 && (!left || left->set.boundary != site || !left->set.moving_right)) return fail(B2_ERR_ARG, "b2_heff_create: left operator set must sit at boundary %d moving right", site);
   if (site < L - 2 && (!right || right->set.boundary != site + 2 || right->set.moving_right)) return fail(B2_ERR_ARG, "b2_heff_create: right operator set must sit at boundary %d moving left", site + 2);
   if (world < 1 || rank < 0 || rank >= world) return fail(B2_ERR_ARG, "b2_heff_create: bad world/rank");
   if ((site > 0 && left && left->set.reduced) || (site < L - 2 && right && right->set.reduced)) return fail(B2_ERR_STATE, "b2_heff_create: a reduced operator set (2-RDM chain / correlation tensors) cannot drive a sigma build");
   if (ctx->device >= 0 && ctx->simulate_oom > 0) { ctx->simulate_oom--; return fail(B2_ERR_CUDA, "b2_heff_create: out of memory (simulated)"); }
   if ((site > 0 && left && left->offloaded) || (site < L - 2 && right && right->offloaded)) return fail(B2_ERR_STATE, "b2_heff_create: operator set is offloaded (b2_opset_reload first)");
   std::unique_ptr<b2_heff> h(new b2_heff);
   h->ctx = ctx; h->world = world; h->rank = rank;
   h->left = (site > 0) ? left : nullptr;
   h->right = (site < L - 2) ? right : nullptr;
   const double tb0 = wall_seconds();
   build_sigma_plan(h->plan, ctx->bk, ctx->prob, h->left ? &h->left->set : nullptr, h->right ? &h->right->set : nullptr, site, world);
   const double tb1 = wall_seconds();
   CompileOptions copt = budgeted(ctx);
   // plans whose sigma build is a few milliseconds are dominated by the time to BUILD them: compile those on all host cores
   copt.threads = (h->plan.flops_ref < ctx->parallel_plan_flops) ? plan_threads(h->plan.S.nkappa()) : 1;
   compile_sigma(h->comp, h->plan, h->left ? &h->left->set : nullptr, h->right ? &h->right->set : nullptr, rank, world, copt);
   h->list_bytes = h->comp.bytes();
   if (getenv("B2_TIMING")) fprintf(stderr, "b2_heff_create: enumerate %.3f s, schedule %.3f s, %zu terms\n", tb1 - tb0, wall_seconds() - tb1, h->plan.terms.size());
   const double tb2 = wall_seconds();
   if (ctx->device >= 0) {
      CUDA_TRY(cudaSetDevice(ctx->device));
      cudaStream_t s = ctx->stream;
      int rc;
      if ((rc = upload_vec(&h->d_items1, h->comp.items1, s))) return rc;
      if ((rc = upload_vec(&h->d_items2, h->comp.items2, s))) return rc;
      if ((rc = upload_vec(&h->d_reduces, h->comp.reduces, s))) return rc;
      if ((rc = upload_vec(&h->d_diag_items, h->comp.diag_items, s))) return rc;
      if ((rc = upload_vec(&h->d_diag_tiles, h->comp.diag_tiles, s))) return rc;
      {
         const SLayout& S = h->plan.S;
         std::vector<int64_t> off(S.nkappa() + 1);
         std::vector<double> p2s(S.nkappa()), s2p(S.nkappa());
         for (int k = 0; k < S.nkappa(); k++) { off[k] = S.blk[k].off; p2s[k] = std::sqrt(S.twoSR[k] + 1.0); s2p[k] = 1.0 / p2s[k]; }
         off[S.nkappa()] = S.size;
         if ((rc = upload_vec(&h->d_blk_off, off, s))) return rc;
         if ((rc = upload_vec(&h->d_p2s, p2s, s))) return rc;
         if ((rc = upload_vec(&h->d_s2p, s2p, s))) return rc;
         CUDA_TRY(cudaStreamSynchronize(s));   // the staging vectors above go out of scope
      }
      if (h->comp.part_size > 0) CUDA_TRY(cudaMalloc(&h->d_part, sizeof(double) * (size_t)h->comp.part_size));
      for (int c = 0; c < kNumTileClasses; c++) {
         if ((rc = upload_vec(&h->d_tiles1[c], h->comp.tiles1[c], s))) return rc;
         if ((rc = upload_vec(&h->d_tiles2[c], h->comp.tiles2[c], s))) return rc;
      }
      if ((rc = upload_vec(&h->d_jobs, h->comp.presum_jobs, s))) return rc;
      if ((rc = upload_vec(&h->d_parts, h->comp.presum_parts, s))) return rc;
      const size_t n = (size_t)h->plan.S.size;
      if (h->plan.presum_size > 0) CUDA_TRY(cudaMalloc(&h->d_presum, sizeof(double) * (size_t)h->plan.presum_size));
      if (h->comp.work_size > 0) CUDA_TRY(cudaMalloc(&h->d_work, sizeof(double) * (size_t)h->comp.work_size));
      CUDA_TRY(cudaMalloc(&h->d_vin, sizeof(double) * (n ? n : 1)));
      CUDA_TRY(cudaMalloc(&h->d_vout, sizeof(double) * (n ? n : 1)));
      CUDA_TRY(cudaMallocHost(&h->h_vin, sizeof(double) * (n ? n : 1)));
      CUDA_TRY(cudaMallocHost(&h->h_vout, sizeof(double) * (n ? n : 1)));
      CUDA_TRY(cudaEventCreate(&h->ev0));
      CUDA_TRY(cudaEventCreate(&h->ev1));
      // materialise the integral-weighted operator pre-sums once (operators are fixed during the Davidson solve)
      DevBases b = bases_of(h.get(), nullptr, nullptr);
      if (dev_launch_presum(h->d_jobs, (int)h->comp.presum_jobs.size(), h->d_parts, b, s)) return fail(B2_ERR_CUDA, "%s", dev_last_error());
      CUDA_TRY(cudaStreamSynchronize(s));
      if (getenv("B2_TIMING"))
         fprintf(stderr, "b2_heff_create: device setup %.3f s (work lists %.1f MB uploaded, workspace %.2f GB, pre-sums %.1f MB)\n", wall_seconds() - tb2,
                 h->list_bytes / 1e6, h->comp.work_size * 8e-9, h->plan.presum_size * 8e-6);
   }
   *out = h.release();
   return B2_OK;
}

void b2_heff_destroy(b2_heff* h) { delete h; }

// a plan parked in the sweep driver's cache keeps only its device work lists: workspaces, vectors and host copies of the lists go
static void heff_park(b2_heff* h) {
   cudaFree(h->d_work); cudaFree(h->d_part); cudaFree(h->d_vin); cudaFree(h->d_vout); cudaFree(h->d_presum);
   cudaFree(h->d_exc); cudaFree(h->d_exc_coef); cudaFree(h->d_exc_scratch);
   h->d_work = h->d_part = h->d_vin = h->d_vout = h->d_presum = h->d_exc = h->d_exc_coef = h->d_exc_scratch = nullptr;
   h->n_exc = 0;
   if (h->h_vin) { cudaFreeHost(h->h_vin); h->h_vin = nullptr; }
   if (h->h_vout) { cudaFreeHost(h->h_vout); h->h_vout = nullptr; }
   std::vector<SigmaTerm>().swap(h->plan.terms);
   std::vector<GemmItem>().swap(h->comp.items1); std::vector<GemmItem>().swap(h->comp.items2);
   std::vector<ReduceJob>().swap(h->comp.reduces);
   for (int c = 0; c < kNumTileClasses; c++) { std::vector<Tile>().swap(h->comp.tiles1[c]); std::vector<Tile>().swap(h->comp.tiles2[c]); }
   h->left = h->right = nullptr;
}
// brings a parked plan back: new operator sets (same layouts: same dimensions), workspaces, pre-summed operators of the new contents
static int heff_unpark(b2_heff* h, b2_opset* left, b2_opset* right) {
   b2_ctx* ctx = h->ctx;
   cudaStream_t s = ctx->stream;
   h->left = left; h->right = right;
   const size_t n = (size_t)h->plan.S.size;
   if (h->comp.part_size > 0) CUDA_TRY(cudaMalloc(&h->d_part, sizeof(double) * (size_t)h->comp.part_size));
   if (h->plan.presum_size > 0) CUDA_TRY(cudaMalloc(&h->d_presum, sizeof(double) * (size_t)h->plan.presum_size));
   if (h->comp.work_size > 0) CUDA_TRY(cudaMalloc(&h->d_work, sizeof(double) * (size_t)h->comp.work_size));
   CUDA_TRY(cudaMalloc(&h->d_vin, sizeof(double) * (n ? n : 1)));
   CUDA_TRY(cudaMalloc(&h->d_vout, sizeof(double) * (n ? n : 1)));
   CUDA_TRY(cudaMallocHost(&h->h_vin, sizeof(double) * (n ? n : 1)));
   CUDA_TRY(cudaMallocHost(&h->h_vout, sizeof(double) * (n ? n : 1)));
   DevBases b = bases_of(h, nullptr, nullptr);
   if (dev_launch_presum(h->d_jobs, (int)h->comp.presum_jobs.size(), h->d_parts, b, s)) return fail(B2_ERR_CUDA, "%s", dev_last_error());
   CUDA_TRY(cudaStreamSynchronize(s));
   return B2_OK;
}

int64_t b2_heff_veclength(const b2_heff* h) { return h ? h->plan.S.size : 0; }

int b2_heff_apply_device(b2_heff* h, const double* dev_in, double* dev_out) {
   if (!h) return fail(B2_ERR_ARG, "b2_heff_apply_device: NULL");
   if (h->ctx->device < 0) return fail(B2_ERR_NO_DEVICE, "b2_heff_apply: planning-only context, no CUDA device (there is no CPU fallback)");
   cudaStream_t s = h->ctx->stream;
   DevBases b = bases_of(h, dev_in, dev_out);
   CUDA_TRY(cudaEventRecord(h->ev0, s));
   if (dev_fill_zero(dev_out, h->plan.S.size, s)) return fail(B2_ERR_CUDA, "%s", dev_last_error());
   for (const Wave& w : h->comp.waves) {
      for (int c = 0; c < kNumTileClasses; c++)
         if (dev_launch_tiles(c, h->d_tiles1[c] + w.t1_begin[c], w.t1_end[c] - w.t1_begin[c], h->d_items1, b, s)) return fail(B2_ERR_CUDA, "%s", dev_last_error());
      for (int c = 0; c < kNumTileClasses; c++)
         if (dev_launch_tiles(c, h->d_tiles2[c] + w.t2_begin[c], w.t2_end[c] - w.t2_begin[c], h->d_items2, b, s)) return fail(B2_ERR_CUDA, "%s", dev_last_error());
      if (dev_launch_reduce(h->d_reduces + w.red_begin, w.red_end - w.red_begin, b, s)) return fail(B2_ERR_CUDA, "%s", dev_last_error());
      h->launches += 1;
   }
   // level-shift projector of the lower states (HeffDiagrams1.cpp:65-85): sigma += sum_s <V_s|S> V_s.  State s belongs to GPU
   // s % world (MPIchemps2.h owner_specific_excitation), the partial sigma vectors are summed by the caller's all-reduce.
   for (int st = 0; st < h->n_exc; st++) {
      if (st % h->world != h->rank) continue;
      const double* v = h->d_exc + (size_t)st * (size_t)h->plan.S.size;
      if (dev_multi_dot(dev_in, v, h->plan.S.size, 1, h->plan.S.size, h->d_exc_coef + st, h->d_exc_scratch, s) ||
          dev_axpy_dev(dev_out, v, h->d_exc_coef + st, 1.0, h->plan.S.size, s))
         return fail(B2_ERR_CUDA, "%s", dev_last_error());
   }
   CUDA_TRY(cudaEventRecord(h->ev1, s));
   return B2_OK;
}

int b2_heff_set_excitations(b2_heff* h, int n_lower, const double* const* veff_tilde) {
   if (!h || n_lower < 0 || (n_lower > 0 && !veff_tilde)) return fail(B2_ERR_ARG, "b2_heff_set_excitations: bad arguments");
   if (h->ctx->device < 0) return fail(B2_ERR_NO_DEVICE, "b2_heff_set_excitations: planning-only context, no CUDA device (there is no CPU fallback)");
   cudaStream_t s = h->ctx->stream;
   cudaFree(h->d_exc); cudaFree(h->d_exc_coef); cudaFree(h->d_exc_scratch);
   h->d_exc = h->d_exc_coef = h->d_exc_scratch = nullptr;
   h->n_exc = 0;
   if (n_lower == 0) return B2_OK;
   const size_t n = (size_t)h->plan.S.size;
   CUDA_TRY(cudaMalloc(&h->d_exc, sizeof(double) * n * n_lower));
   CUDA_TRY(cudaMalloc(&h->d_exc_coef, sizeof(double) * n_lower));
   CUDA_TRY(cudaMalloc(&h->d_exc_scratch, sizeof(double) * kRedScratch));
   CUDA_TRY(cudaMemsetAsync(h->d_exc_scratch, 0, sizeof(double) * kRedScratch, s));
   for (int st = 0; st < n_lower; st++) {
      if (!veff_tilde[st]) return fail(B2_ERR_ARG, "b2_heff_set_excitations: NULL vector %d", st);
      std::memcpy(h->h_vin, veff_tilde[st], sizeof(double) * n);
      CUDA_TRY(cudaMemcpyAsync(h->d_exc + (size_t)st * n, h->h_vin, sizeof(double) * n, cudaMemcpyHostToDevice, s));
      CUDA_TRY(cudaStreamSynchronize(s));
   }
   h->n_exc = n_lower;
   return B2_OK;
}

int b2_heff_apply(b2_heff* h, const double* vec_in, double* vec_out) {
   if (!h || !vec_in || !vec_out) return fail(B2_ERR_ARG, "b2_heff_apply: NULL argument");
   if (h->ctx->device < 0) return fail(B2_ERR_NO_DEVICE, "b2_heff_apply: planning-only context, no CUDA device (there is no CPU fallback)");
   cudaStream_t s = h->ctx->stream;
   const size_t bytes = sizeof(double) * (size_t)h->plan.S.size;
   std::memcpy(h->h_vin, vec_in, bytes);
   CUDA_TRY(cudaMemcpyAsync(h->d_vin, h->h_vin, bytes, cudaMemcpyHostToDevice, s));
   int rc = b2_heff_apply_device(h, h->d_vin, h->d_vout);
   if (rc) return rc;
   CUDA_TRY(cudaMemcpyAsync(h->h_vout, h->d_vout, bytes, cudaMemcpyDeviceToHost, s));
   CUDA_TRY(cudaStreamSynchronize(s));
   std::memcpy(vec_out, h->h_vout, bytes);
   float ms = 0.f;
   CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
   h->last_kernel_s = ms * 1e-3;
   return B2_OK;
}

double b2_heff_last_kernel_seconds(const b2_heff* h) {
   if (!h || !h->ev0) return 0.0;
   float ms = 0.f;
   if (cudaEventSynchronize(h->ev1) != cudaSuccess) return 0.0;
   if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) != cudaSuccess) return 0.0;
   return ms * 1e-3;
}

int b2_heff_diag_device(b2_heff* h, double* dev_diag) {
   if (!h || !dev_diag) return fail(B2_ERR_ARG, "b2_heff_diag_device: NULL argument");
   if (h->ctx->device < 0) return fail(B2_ERR_NO_DEVICE, "b2_heff_diag: planning-only context, no CUDA device (there is no CPU fallback)");
   cudaStream_t s = h->ctx->stream;
   DevBases b = bases_of(h, nullptr, nullptr);
   if (dev_fill_zero(dev_diag, h->plan.S.size, s)) return fail(B2_ERR_CUDA, "%s", dev_last_error());
   if (dev_launch_diag(h->d_diag_tiles, (int)h->comp.diag_tiles.size(), h->d_diag_items, b, dev_diag, s)) return fail(B2_ERR_CUDA, "%s", dev_last_error());
   for (int st = 0; st < h->n_exc; st++)   // HeffDiagonal.cpp:621-640: diag += V_s .* V_s
      if (st % h->world == h->rank && dev_add_square(dev_diag, h->d_exc + (size_t)st * (size_t)h->plan.S.size, h->plan.S.size, s)) return fail(B2_ERR_CUDA, "%s", dev_last_error());
   return B2_OK;
}

int b2_heff_diag(b2_heff* h, double* diag) {
   if (!h || !diag) return fail(B2_ERR_ARG, "b2_heff_diag: NULL argument");
   if (h->ctx->device < 0) return fail(B2_ERR_NO_DEVICE, "b2_heff_diag: planning-only context, no CUDA device (there is no CPU fallback)");
   int rc = b2_heff_diag_device(h, h->d_vout);
   if (rc) return rc;
   const size_t bytes = sizeof(double) * (size_t)h->plan.S.size;
   CUDA_TRY(cudaMemcpyAsync(h->h_vout, h->d_vout, bytes, cudaMemcpyDeviceToHost, h->ctx->stream));
   CUDA_TRY(cudaStreamSynchronize(h->ctx->stream));
   std::memcpy(diag, h->h_vout, bytes);
   return B2_OK;
}

int b2_heff_solve_device(b2_heff* h, double* dev_s, double rtol, double* eigenvalue, int* n_matvec) {
   if (!h || !dev_s || !eigenvalue) return fail(B2_ERR_ARG, "b2_heff_solve: NULL argument");
   if (h->ctx->device < 0) return fail(B2_ERR_NO_DEVICE, "b2_heff_solve: planning-only context, no CUDA device (there is no CPU fallback)");
   cudaStream_t s = h->ctx->stream;
   const int64_t n = h->plan.S.size;
   const int nk = h->plan.S.nkappa();
   double* d_diag = nullptr;
   CUDA_TRY(cudaMalloc(&d_diag, sizeof(double) * (size_t)(n ? n : 1)));
   int rc = B2_OK, nm = 0;
   char err[256] = "";
   do {
      if (dev_scale_blocks(dev_s, h->d_blk_off, h->d_p2s, nk, s)) { rc = fail(B2_ERR_CUDA, "prog2symm launch failed"); break; }   // Heff.cpp:345
      if ((rc = b2_heff_diag_device(h, d_diag))) break;                                                                            // Heff.cpp:352
      if (h->allreduce && (rc = h->allreduce(h->allreduce_user, d_diag, n, (void*)s))) { rc = fail(B2_ERR_STATE, "all-reduce callback failed"); break; }
      DavidsonParams prm;
      prm.rtol = rtol;
      MatVec mv = [h](const double* in, double* out) -> int {
         int r = b2_heff_apply_device(h, in, out);
         if (r) return r;
         if (h->allreduce) return h->allreduce(h->allreduce_user, out, h->plan.S.size, (void*)h->ctx->stream);
         return 0;
      };
      if (davidson_solve((void*)s, n, mv, dev_s, d_diag, prm, eigenvalue, &nm, err, sizeof(err))) { rc = fail(B2_ERR_CUDA, "%s", err); break; }
      if (dev_scale_blocks(dev_s, h->d_blk_off, h->d_s2p, nk, s)) { rc = fail(B2_ERR_CUDA, "symm2prog launch failed"); break; }   // Heff.cpp:374
      cudaError_t e = cudaStreamSynchronize(s);
      if (e != cudaSuccess) { rc = fail(B2_ERR_CUDA, "b2_heff_solve: %s", cudaGetErrorString(e)); break; }
   } while (0);
   cudaFree(d_diag);
   if (n_matvec) *n_matvec = nm;
   return rc;
}

int b2_heff_solve(b2_heff* h, double* s_host, double rtol, double* eigenvalue, int* n_matvec) {
   if (!h || !s_host || !eigenvalue) return fail(B2_ERR_ARG, "b2_heff_solve: NULL argument");
   if (h->ctx->device < 0) return fail(B2_ERR_NO_DEVICE, "b2_heff_solve: planning-only context, no CUDA device (there is no CPU fallback)");
   cudaStream_t s = h->ctx->stream;
   const size_t bytes = sizeof(double) * (size_t)h->plan.S.size;
   double* d_s = nullptr;
   CUDA_TRY(cudaMalloc(&d_s, bytes ? bytes : 8));
   std::memcpy(h->h_vin, s_host, bytes);
   cudaError_t e = cudaMemcpyAsync(d_s, h->h_vin, bytes, cudaMemcpyHostToDevice, s);
   int rc = (e == cudaSuccess) ? b2_heff_solve_device(h, d_s, rtol, eigenvalue, n_matvec) : fail(B2_ERR_CUDA, "H2D: %s", cudaGetErrorString(e));
   if (!rc) {
      e = cudaMemcpyAsync(h->h_vin, d_s, bytes, cudaMemcpyDeviceToHost, s);
      if (e == cudaSuccess) e = cudaStreamSynchronize(s);
      if (e != cudaSuccess) rc = fail(B2_ERR_CUDA, "D2H: %s", cudaGetErrorString(e));
      else std::memcpy(s_host, h->h_vin, bytes);
   }
   cudaFree(d_s);
   return rc;
}

int b2_heff_set_allreduce(b2_heff* h, b2_allreduce_fn fn, void* user) {
   if (!h) return fail(B2_ERR_ARG, "b2_heff_set_allreduce: NULL");
   h->allreduce = fn; h->allreduce_user = user;
   return B2_OK;
}

int b2_heff_stats(const b2_heff* h, double* o) {
   if (!h || !o) return fail(B2_ERR_ARG, "b2_heff_stats: NULL");
   o[0] = (double)h->plan.terms.size(); o[1] = (double)h->plan.skipped_zero; o[2] = (double)h->plan.presums.size();
   o[3] = h->plan.flops_ref; o[4] = h->comp.flops_exec; o[5] = (double)h->comp.work_size; o[6] = (double)h->comp.n_stage1; o[7] = (double)h->comp.n_tiles;
   o[8] = (double)h->comp.waves.size();
   o[9] = 1.0 + h->comp.launches(); o[10] = (double)h->comp.part_size;
   const double bytes = h->comp.bytes();
   o[11] = bytes;
   return B2_OK;
}

// ---- flat exports for the CPU checker
static void flat_ref(const BRef& r, const SigmaPlan& plan, const OpSet* left, const OpSet* right, int8_t* space, int8_t* trans, int64_t* off, int32_t* rows, int32_t* cols) {
   *space = 0; *trans = 0; *off = 0; *rows = 0; *cols = 0;
   if (r.src == SRC_NONE || r.op < 0 || r.blk < 0) return;
   const OpLayout* lay; int64_t base;
   if (r.src == SRC_LEFT) { lay = left->ops[r.op].lay.get(); base = left->ops[r.op].off; *space = 1; }
   else if (r.src == SRC_RIGHT) { lay = right->ops[r.op].lay.get(); base = right->ops[r.op].off; *space = 2; }
   else { lay = plan.presums[r.op].lay.get(); base = plan.presums[r.op].off; *space = 3; }
   *trans = r.trans; *off = base + lay->blk[r.blk].off; *rows = lay->blk[r.blk].rows; *cols = lay->blk[r.blk].cols;
}
int64_t b2_heff_num_terms(const b2_heff* h) { return h ? (int64_t)h->plan.terms.size() : 0; }
int b2_heff_export_terms(const b2_heff* h, b2_flat_term* out) {
   if (!h || !out) return fail(B2_ERR_ARG, "b2_heff_export_terms: NULL");
   const OpSet* l = h->left ? &h->left->set : nullptr;
   const OpSet* r = h->right ? &h->right->set : nullptr;
   for (size_t i = 0; i < h->plan.terms.size(); i++) {
      const SigmaTerm& t = h->plan.terms[i];
      b2_flat_term& f = out[i];
      f.dst = t.dst; f.src = t.src; f.owner = t.owner; f.factor = t.factor;
      flat_ref(t.l, h->plan, l, r, &f.a_space, &f.a_trans, &f.a_off, &f.a_rows, &f.a_cols);
      flat_ref(t.r, h->plan, l, r, &f.b_space, &f.b_trans, &f.b_off, &f.b_rows, &f.b_cols);
   }
   return B2_OK;
}
int64_t b2_heff_num_presum_parts(const b2_heff* h) { return h ? (int64_t)h->comp.presum_parts.size() : 0; }
int64_t b2_heff_presum_size(const b2_heff* h) { return h ? h->plan.presum_size : 0; }
int b2_heff_export_presums(const b2_heff* h, b2_flat_presum* out) {
   if (!h || !out) return fail(B2_ERR_ARG, "b2_heff_export_presums: NULL");
   size_t n = 0;
   for (const PresumJob& j : h->comp.presum_jobs)
      for (int p = j.part_begin; p < j.part_end; p++) {
         const PresumPart& pp = h->comp.presum_parts[p];
         out[n].dst_off = j.dst_off; out[n].src_off = pp.src_off; out[n].size = j.size; out[n].space = pp.space; out[n].coef = pp.coef;
         n++;
      }
   return B2_OK;
}

int b2_heff_worklists(const b2_heff* h, b2_worklists* o) {
   if (!h || !o) return fail(B2_ERR_ARG, "b2_heff_worklists: NULL");
   const CompiledSigma& c = h->comp;
   o->items1 = c.items1.data(); o->n_items1 = (int64_t)c.items1.size();
   o->items2 = c.items2.data(); o->n_items2 = (int64_t)c.items2.size();
   for (int k = 0; k < kNumTileClasses; k++) {
      o->tiles1[k] = c.tiles1[k].data(); o->n_tiles1[k] = (int64_t)c.tiles1[k].size();
      o->tiles2[k] = c.tiles2[k].data(); o->n_tiles2[k] = (int64_t)c.tiles2[k].size();
   }
   o->reduces = c.reduces.data(); o->n_reduces = (int64_t)c.reduces.size();
   o->waves = c.waves.data(); o->n_waves = (int64_t)c.waves.size();
   o->work_size = c.work_size; o->part_size = c.part_size;
   return B2_OK;
}
int b2_ctx_set_option(b2_ctx* ctx, const char* name, double value) {
   if (!ctx || !name) return fail(B2_ERR_ARG, "b2_ctx_set_option: NULL");
   if (!std::strcmp(name, "work_budget")) { if (value < 1024) return fail(B2_ERR_ARG, "work_budget too small"); ctx->copt.work_budget = (int64_t)value; }
   else if (!std::strcmp(name, "chunk_k")) { if (value < 8) return fail(B2_ERR_ARG, "chunk_k too small"); ctx->copt.chunk_k = (int64_t)value; }
   else if (!std::strcmp(name, "parallel_plan_flops")) ctx->parallel_plan_flops = value;
   else if (!std::strcmp(name, "simulate_oom")) ctx->simulate_oom = (int)value;
   else if (!std::strcmp(name, "parallel_min_terms")) ctx->copt.parallel_min_terms = (int64_t)value;
   else return fail(B2_ERR_ARG, "b2_ctx_set_option: unknown option %s", name);
   return B2_OK;
}

// ------------------------------------------------------------------------------------------------ operator update
struct b2_update {
   b2_ctx* ctx = nullptr;
   b2_opset *old_set = nullptr, *new_set = nullptr;
   UpdatePlan plan;
   CompiledWork pass[2];
   std::vector<PresumJob> presum_jobs;
   std::vector<PresumPart> presum_parts;
   GemmItem *d_items1[2] = {nullptr, nullptr}, *d_items2[2] = {nullptr, nullptr};
   ReduceJob* d_reduces[2] = {nullptr, nullptr};
   Tile* d_tiles1[2][kNumTileClasses] = {};
   Tile* d_tiles2[2][kNumTileClasses] = {};
   PresumJob* d_jobs = nullptr;
   PresumPart* d_parts = nullptr;
   double *d_presum = nullptr, *d_work = nullptr, *d_part = nullptr, *d_t = nullptr, *h_t = nullptr;
   int world = 1, rank = 0;
   double list_bytes[2] = {0.0, 0.0};      // device work-list bytes per pass
   std::vector<int> op_owner;              // GPU that computes new operator i in pass 0
   b2_allreduce_fn allreduce = nullptr;
   void* allreduce_user = nullptr;
   ~b2_update() {
      for (int p = 0; p < 2; p++) {
         cudaFree(d_items1[p]); cudaFree(d_items2[p]); cudaFree(d_reduces[p]);
         for (int c = 0; c < kNumTileClasses; c++) { cudaFree(d_tiles1[p][c]); cudaFree(d_tiles2[p][c]); }
      }
      cudaFree(d_jobs); cudaFree(d_parts); cudaFree(d_presum); cudaFree(d_work); cudaFree(d_part); cudaFree(d_t);
      if (h_t) cudaFreeHost(h_t);
   }
};

// FLOPs the scheduler will spend on one update term (cheaper association order, as compile_terms picks it)
static double term_cost(const Term3& t, const DstBlock& d) {
   const double M = d.rows, N = d.cols;
   const bool hp = t.p.present(), hq = t.q.present(), hr = t.r.present();
   if (hp && hq && hr) {
      const double k1 = t.q.op_rows(), k2 = t.q.op_cols();
      return 2.0 * std::min(M * k1 * k2 + M * k2 * N, k1 * k2 * N + M * k1 * N);
   }
   if (hp && hq) return 2.0 * M * N * t.p.op_cols();
   if (hq && hr) return 2.0 * M * N * t.q.op_cols();
   if (hp && hr) return 2.0 * M * N * t.p.op_cols();
   return 2.0 * M * N;
}

static void fill_worklists(const CompiledWork& c, b2_worklists* o) {
   o->items1 = c.items1.data(); o->n_items1 = (int64_t)c.items1.size();
   o->items2 = c.items2.data(); o->n_items2 = (int64_t)c.items2.size();
   for (int k = 0; k < kNumTileClasses; k++) {
      o->tiles1[k] = c.tiles1[k].data(); o->n_tiles1[k] = (int64_t)c.tiles1[k].size();
      o->tiles2[k] = c.tiles2[k].data(); o->n_tiles2[k] = (int64_t)c.tiles2[k].size();
   }
   o->reduces = c.reduces.data(); o->n_reduces = (int64_t)c.reduces.size();
   o->waves = c.waves.data(); o->n_waves = (int64_t)c.waves.size();
   o->work_size = c.work_size; o->part_size = c.part_size;
}

int b2_update_create(b2_ctx* ctx, int index, int moving_right, b2_opset* old_set, b2_opset* new_set, b2_update** out) {
   return b2_update_create_sharded(ctx, index, moving_right, old_set, new_set, 1, 0, out);
}

int b2_update_create_sharded(b2_ctx* ctx, int index, int moving_right, b2_opset* old_set, b2_opset* new_set, int world, int rank, b2_update** out) {
   if (!ctx || !ctx->have_bk || !new_set || !out) return fail(B2_ERR_STATE, "b2_update_create: bad arguments");
   if (world < 1 || rank < 0 || rank >= world) return fail(B2_ERR_ARG, "b2_update_create: bad world/rank");
   const int L = ctx->bk.L;
   if (index < 0 || index > L - 1) return fail(B2_ERR_ARG, "b2_update_create: site %d out of range", index);
   const bool mr = moving_right != 0;
   const int b_old = mr ? index : index + 1, b_new = mr ? index + 1 : index;
   if (new_set->set.boundary != b_new || new_set->set.moving_right != mr) return fail(B2_ERR_ARG, "b2_update_create: new_set must sit at boundary %d", b_new);
   const bool need_old = mr ? (index > 0) : (index < L - 1);
   if (need_old && (!old_set || old_set->set.boundary != b_old || old_set->set.moving_right != mr)) return fail(B2_ERR_ARG, "b2_update_create: old_set must sit at boundary %d", b_old);
   if ((need_old && old_set->offloaded) || new_set->offloaded) return fail(B2_ERR_STATE, "b2_update_create: operator set is offloaded (b2_opset_reload first)");
   std::unique_ptr<b2_update> u(new b2_update);
   u->ctx = ctx; u->old_set = need_old ? old_set : nullptr; u->new_set = new_set;
   build_update_plan(u->plan, ctx->bk, ctx->prob, u->old_set ? &u->old_set->set : nullptr, new_set->set, index, mr);
   u->world = world; u->rank = rank;
   {  // pass 0 is sharded by NEW operator: greedy longest-processing-time assignment of the operators to the GPUs by the
      // FLOPs of their terms (the reference's static owner maps, MPIchemps2.h:158-231, balance counts, not work; every rank
      // evaluates the same deterministic assignment).  The partial arenas are summed by the all-reduce callback.
      const int nops = (int)u->plan.block_base.size();
      std::vector<double> cost(nops, 0.0);
      auto op_of_block = [&](int blk) { return (int)(std::upper_bound(u->plan.block_base.begin(), u->plan.block_base.end(), blk) - u->plan.block_base.begin()) - 1; };
      for (const Term3& t : u->plan.terms) cost[op_of_block(t.dst)] += term_cost(t, u->plan.dst[t.dst]);
      std::vector<int> order(nops);
      for (int i = 0; i < nops; i++) order[i] = i;
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
      std::vector<double> load(world, 0.0);
      u->op_owner.assign(nops, 0);
      for (int i : order) {
         const int r = (int)(std::min_element(load.begin(), load.end()) - load.begin());
         u->op_owner[i] = r; load[r] += cost[i];
      }
      if (world > 1) {
         std::vector<Term3> mine;
         for (const Term3& t : u->plan.terms) if (u->op_owner[op_of_block(t.dst)] == rank) mine.push_back(t);
         u->plan.terms.swap(mine);
      }
   }
   CompileOptions copt = budgeted(ctx);
   copt.threads = (u->plan.flops_ref < ctx->parallel_plan_flops) ? plan_threads((int)u->plan.dst.size()) : 1;
   compile_terms(u->pass[0], u->plan.terms, u->plan.dst, SP_VOUT, copt);
   compile_terms(u->pass[1], u->plan.mix_terms, u->plan.dst, SP_VOUT, copt);
   for (int p = 0; p < 2; p++) u->list_bytes[p] = u->pass[p].bytes();
   for (const Presum& p : u->plan.presums) {
      PresumJob j{};
      j.dst_off = p.off; j.size = p.lay->size; j.part_begin = (int)u->presum_parts.size();
      for (auto& pr : p.parts) {
         PresumPart pp{};
         pp.src_off = u->old_set->set.ops[pr.second].off; pp.coef = pr.first; pp.space = SP_LEFT;
         u->presum_parts.push_back(pp);
      }
      j.part_end = (int)u->presum_parts.size();
      u->presum_jobs.push_back(j);
   }
   if (ctx->device >= 0) {
      CUDA_TRY(cudaSetDevice(ctx->device));
      cudaStream_t s = ctx->stream;
      int rc;
      int64_t work = 0, part = 0;
      for (int p = 0; p < 2; p++) {
         if ((rc = upload_vec(&u->d_items1[p], u->pass[p].items1, s))) return rc;
         if ((rc = upload_vec(&u->d_items2[p], u->pass[p].items2, s))) return rc;
         if ((rc = upload_vec(&u->d_reduces[p], u->pass[p].reduces, s))) return rc;
         for (int c = 0; c < kNumTileClasses; c++) {
            if ((rc = upload_vec(&u->d_tiles1[p][c], u->pass[p].tiles1[c], s))) return rc;
            if ((rc = upload_vec(&u->d_tiles2[p][c], u->pass[p].tiles2[c], s))) return rc;
         }
         work = std::max(work, u->pass[p].work_size); part = std::max(part, u->pass[p].part_size);
      }
      if ((rc = upload_vec(&u->d_jobs, u->presum_jobs, s))) return rc;
      if ((rc = upload_vec(&u->d_parts, u->presum_parts, s))) return rc;
      if (u->plan.presum_size > 0) CUDA_TRY(cudaMalloc(&u->d_presum, sizeof(double) * (size_t)u->plan.presum_size));
      if (work > 0) CUDA_TRY(cudaMalloc(&u->d_work, sizeof(double) * (size_t)work));
      if (part > 0) CUDA_TRY(cudaMalloc(&u->d_part, sizeof(double) * (size_t)part));
      const size_t nt = (size_t)(u->plan.T.size ? u->plan.T.size : 1);
      CUDA_TRY(cudaMalloc(&u->d_t, sizeof(double) * nt));
      CUDA_TRY(cudaMallocHost(&u->h_t, sizeof(double) * nt));
      CUDA_TRY(cudaStreamSynchronize(s));
   }
   *out = u.release();
   return B2_OK;
}
void b2_update_destroy(b2_update* u) { delete u; }
// update plans kept by the sweep driver between visits of a boundary (same idea as heff_park / heff_unpark)
static void update_park(b2_update* u) {
   cudaFree(u->d_presum); cudaFree(u->d_work); cudaFree(u->d_part); cudaFree(u->d_t);
   u->d_presum = u->d_work = u->d_part = u->d_t = nullptr;
   if (u->h_t) { cudaFreeHost(u->h_t); u->h_t = nullptr; }
   std::vector<Term3>().swap(u->plan.terms); std::vector<Term3>().swap(u->plan.mix_terms);
   for (int p = 0; p < 2; p++) {
      std::vector<GemmItem>().swap(u->pass[p].items1); std::vector<GemmItem>().swap(u->pass[p].items2);
      std::vector<ReduceJob>().swap(u->pass[p].reduces);
      for (int c = 0; c < kNumTileClasses; c++) { std::vector<Tile>().swap(u->pass[p].tiles1[c]); std::vector<Tile>().swap(u->pass[p].tiles2[c]); }
   }
   u->old_set = u->new_set = nullptr;
}
static int update_unpark(b2_update* u, b2_opset* old_set, b2_opset* new_set) {
   u->old_set = old_set; u->new_set = new_set;
   int64_t work = 0, part = 0;
   for (int p = 0; p < 2; p++) { work = std::max(work, u->pass[p].work_size); part = std::max(part, u->pass[p].part_size); }
   if (u->plan.presum_size > 0) CUDA_TRY(cudaMalloc(&u->d_presum, sizeof(double) * (size_t)u->plan.presum_size));
   if (work > 0) CUDA_TRY(cudaMalloc(&u->d_work, sizeof(double) * (size_t)work));
   if (part > 0) CUDA_TRY(cudaMalloc(&u->d_part, sizeof(double) * (size_t)part));
   const size_t nt = (size_t)(u->plan.T.size ? u->plan.T.size : 1);
   CUDA_TRY(cudaMalloc(&u->d_t, sizeof(double) * nt));
   CUDA_TRY(cudaMallocHost(&u->h_t, sizeof(double) * nt));
   return B2_OK;
}
int b2_update_run_device(b2_update* u, const double* t_dev) {
   if (!u || !t_dev) return fail(B2_ERR_ARG, "b2_update_run_device: NULL");
   if (u->ctx->device < 0) return fail(B2_ERR_NO_DEVICE, "b2_update_run: planning-only context, no CUDA device (there is no CPU fallback)");
   cudaStream_t s = u->ctx->stream;
   DevBases b;
   for (int i = 0; i < SP_COUNT; i++) b.p[i] = nullptr;
   b.p[SP_LEFT] = u->old_set ? u->old_set->dev : nullptr;
   b.p[SP_RIGHT] = const_cast<double*>(t_dev);
   b.p[SP_PRESUM] = u->d_presum; b.p[SP_WORK] = u->d_work; b.p[SP_PART] = u->d_part;
   b.p[SP_VOUT] = u->new_set->dev;
   if (dev_launch_presum(u->d_jobs, (int)u->presum_jobs.size(), u->d_parts, b, s)) return fail(B2_ERR_CUDA, "%s", dev_last_error());
   if (dev_fill_zero(u->new_set->dev, u->new_set->set.size, s)) return fail(B2_ERR_CUDA, "%s", dev_last_error());
   if (u->world > 1 && !u->allreduce) return fail(B2_ERR_STATE, "b2_update_run: sharded update (world %d) without an all-reduce callback", u->world);
   for (int p = 0; p < 2; p++) {
      for (const Wave& w : u->pass[p].waves) {
         for (int c = 0; c < kNumTileClasses; c++)
            if (dev_launch_tiles(c, u->d_tiles1[p][c] + w.t1_begin[c], w.t1_end[c] - w.t1_begin[c], u->d_items1[p], b, s)) return fail(B2_ERR_CUDA, "%s", dev_last_error());
         for (int c = 0; c < kNumTileClasses; c++)
            if (dev_launch_tiles(c, u->d_tiles2[p][c] + w.t2_begin[c], w.t2_end[c] - w.t2_begin[c], u->d_items2[p], b, s)) return fail(B2_ERR_CUDA, "%s", dev_last_error());
         if (dev_launch_reduce(u->d_reduces[p] + w.red_begin, w.red_end - w.red_begin, b, s)) return fail(B2_ERR_CUDA, "%s", dev_last_error());
      }
      // every GPU computed the operators it was assigned; summing the (otherwise zero) arenas replicates all of them before
      // the mixing pass, which every GPU then runs in full (block axpys, replaces the MPI exchanges of DMRGoperators.cpp:449-533)
      if (p == 0 && u->world > 1 && u->allreduce(u->allreduce_user, u->new_set->dev, u->new_set->set.size, (void*)s)) return fail(B2_ERR_STATE, "b2_update_run: all-reduce callback failed");
   }
   return B2_OK;
}
int b2_update_run(b2_update* u, const double* t_host) {
   if (!u || !t_host) return fail(B2_ERR_ARG, "b2_update_run: NULL");
   if (u->ctx->device < 0) return fail(B2_ERR_NO_DEVICE, "b2_update_run: planning-only context, no CUDA device (there is no CPU fallback)");
   const size_t bytes = sizeof(double) * (size_t)u->plan.T.size;
   std::memcpy(u->h_t, t_host, bytes);
   CUDA_TRY(cudaMemcpyAsync(u->d_t, u->h_t, bytes, cudaMemcpyHostToDevice, u->ctx->stream));
   int rc = b2_update_run_device(u, u->d_t);
   if (rc) return rc;
   CUDA_TRY(cudaStreamSynchronize(u->ctx->stream));
   return B2_OK;
}
int b2_update_set_allreduce(b2_update* u, b2_allreduce_fn fn, void* user) {
   if (!u) return fail(B2_ERR_ARG, "b2_update_set_allreduce: NULL");
   u->allreduce = fn; u->allreduce_user = user;
   return B2_OK;
}
int b2_update_stats(const b2_update* u, double* o) {
   if (!u || !o) return fail(B2_ERR_ARG, "b2_update_stats: NULL");
   o[0] = (double)u->plan.terms.size(); o[1] = (double)u->plan.mix_terms.size(); o[2] = (double)u->plan.presums.size(); o[3] = u->plan.flops_ref;
   o[4] = u->pass[0].flops_exec + u->pass[1].flops_exec; o[5] = (double)std::max(u->pass[0].work_size, u->pass[1].work_size);
   o[6] = (double)(u->pass[0].waves.size() + u->pass[1].waves.size()); o[7] = 2.0 + u->pass[0].launches() + u->pass[1].launches();
   return B2_OK;
}
int b2_update_worklists(const b2_update* u, int pass, b2_worklists* o) {
   if (!u || !o || pass < 0 || pass > 1) return fail(B2_ERR_ARG, "b2_update_worklists: bad arguments");
   fill_worklists(u->pass[pass], o);
   return B2_OK;
}
int64_t b2_update_num_presum_parts(const b2_update* u) { return u ? (int64_t)u->presum_parts.size() : 0; }
int64_t b2_update_presum_size(const b2_update* u) { return u ? u->plan.presum_size : 0; }
int b2_update_export_presums(const b2_update* u, b2_flat_presum* out) {
   if (!u || !out) return fail(B2_ERR_ARG, "b2_update_export_presums: NULL");
   size_t n = 0;
   for (const PresumJob& j : u->presum_jobs)
      for (int p = j.part_begin; p < j.part_end; p++) {
         const PresumPart& pp = u->presum_parts[p];
         out[n].dst_off = j.dst_off; out[n].src_off = pp.src_off; out[n].size = j.size; out[n].space = pp.space; out[n].coef = pp.coef;
         n++;
      }
   return B2_OK;
}

// ------------------------------------------------------------------------------------------------ sweep driver
// upload a compiled work list, run it once on the context stream, free it (used for small one-shot contractions: Join)
static int run_compiled_once(b2_ctx* ctx, const CompiledWork& w, DevBases b) {
   cudaStream_t s = ctx->stream;
   GemmItem *i1 = nullptr, *i2 = nullptr;
   ReduceJob* red = nullptr;
   Tile *t1[kNumTileClasses] = {}, *t2[kNumTileClasses] = {};
   double *work = nullptr, *part = nullptr;
   int rc = B2_OK;
   auto cleanup = [&]() {
      cudaFree(i1); cudaFree(i2); cudaFree(red); cudaFree(work); cudaFree(part);
      for (int c = 0; c < kNumTileClasses; c++) { cudaFree(t1[c]); cudaFree(t2[c]); }
   };
   do {
      if ((rc = upload_vec(&i1, w.items1, s))) break;
      if ((rc = upload_vec(&i2, w.items2, s))) break;
      if ((rc = upload_vec(&red, w.reduces, s))) break;
      for (int c = 0; c < kNumTileClasses && !rc; c++) { rc = upload_vec(&t1[c], w.tiles1[c], s); if (!rc) rc = upload_vec(&t2[c], w.tiles2[c], s); }
      if (rc) break;
      if (w.work_size > 0 && cudaMalloc(&work, sizeof(double) * (size_t)w.work_size) != cudaSuccess) { rc = fail(B2_ERR_CUDA, "workspace allocation failed"); break; }
      if (w.part_size > 0 && cudaMalloc(&part, sizeof(double) * (size_t)w.part_size) != cudaSuccess) { rc = fail(B2_ERR_CUDA, "workspace allocation failed"); break; }
      b.p[SP_WORK] = work; b.p[SP_PART] = part;
      for (const Wave& wv : w.waves) {
         for (int c = 0; c < kNumTileClasses && !rc; c++)
            if (dev_launch_tiles(c, t1[c] + wv.t1_begin[c], wv.t1_end[c] - wv.t1_begin[c], i1, b, s)) rc = fail(B2_ERR_CUDA, "%s", dev_last_error());
         for (int c = 0; c < kNumTileClasses && !rc; c++)
            if (dev_launch_tiles(c, t2[c] + wv.t2_begin[c], wv.t2_end[c] - wv.t2_begin[c], i2, b, s)) rc = fail(B2_ERR_CUDA, "%s", dev_last_error());
         if (!rc && dev_launch_reduce(red + wv.red_begin, wv.red_end - wv.red_begin, b, s)) rc = fail(B2_ERR_CUDA, "%s", dev_last_error());
         if (rc) break;
      }
      cudaError_t e = cudaStreamSynchronize(s);
      if (!rc && e != cudaSuccess) rc = fail(B2_ERR_CUDA, "run_compiled_once: %s", cudaGetErrorString(e));
   } while (0);
   cleanup();
   return rc;
}

// overlap tensor <current MPS | stored lower state> on one boundary (CheMPS2::TensorO, TensorO.cpp): one block dim_current x dim_stored per
// symmetry sector that is populated in both bookkeepers
struct Overlap {
   struct Blk { int n, ts, ir, rows, cols; int64_t off; };
   std::vector<Blk> blk;
   std::vector<double> data;
   bool valid = false;
   const Blk* find(int n, int ts, int ir) const {
      for (const Blk& b : blk) if (b.n == n && b.ts == ts && b.ir == ir) return &b;
      return nullptr;
   }
};
// a converged lower state kept for the level-shift projector (DMRG::newExcitation, DMRG.cpp:475-505): Exc_MPSs, Exc_BKs, Exc_Eshifts, Exc_Overlaps
struct ExcState {
   double eshift = 0.0;
   Bookkeeper bk;
   std::vector<std::vector<double>> mps;
   std::vector<Overlap> left, right;   // per boundary: built moving right (covers sites < b) / moving left (sites >= b)
};

struct b2_dmrg {
   b2_ctx* ctx = nullptr;
   int L = 0;
   std::vector<ExcState> exc;              // lower states (excited-state calculations)
   std::vector<std::vector<double>> mps;   // TensorT storage per site in the layouts of the current bookkeeper
   std::vector<b2_opset*> left, right;     // operator sets per boundary: moving right (sites < b) / moving left (sites >= b)
   // Sigma plans of earlier visits, one slot per site: at a fixed virtual dimension the sector dimensions stop changing once the sweeps
   // converge, and a plan depends on nothing but those dimensions — re-using it removes the host-side plan building (the largest part
   // of a small/medium-D sweep) from every later visit.  Key = the exact dimension tables of the three boundaries + the sharding.
   struct PlanSlot { std::vector<int> key; b2_heff* h = nullptr; };
   std::vector<PlanSlot> plan_cache;
   struct UpdSlot { std::vector<int> key; b2_update* u = nullptr; };
   std::vector<UpdSlot> upd_cache;        // index = 2 * site + moving_right
   bool use_plan_cache = true;
   long long plan_hits = 0, plan_misses = 0;
   bool swept_once = false;                // false until the first left sweep (which runs with fixed virtual dimensions, DMRG.cpp:270)
   double max_disc_last_sweep = 0.0;       // DMRG::MaxDiscWeightLastSweep (DMRG.cpp:360-362): scales the noise of the next half sweep
   double last_energy = 0.0;               // energy of the last site solved (what DMRG::sweepleft / sweepright return)
   double last_min_energy = 1e8;           // DMRG::LastMinEnergy: lowest energy of the last half sweep
   double total_min_energy = 1e8;          // DMRG::TotalMinEnergy: lowest energy since the last PreSolve
   bool spill = false;                     // keep only the operator sets of the site being optimised in HBM (b2_dmrg_set_spill)
   int world = 1, rank = 0;                // GPUs sharing the sweep: sigma terms and operator updates are sharded, the rest is replicated
   b2_allreduce_fn allreduce = nullptr;
   void* allreduce_user = nullptr;
   double t_solve = 0.0, t_update = 0.0, t_split = 0.0, t_plan = 0.0;   // wall-clock seconds spent per phase (b2_dmrg_timers)
   long long n_matvec = 0;
   double t_join = 0.0, t_release = 0.0, t_tail = 0.0;   // B2_TIMING diagnostics: Join + vector copies, releasing plans / buffers / stale sets, update epilogue
   unsigned long long rng = 0x9E3779B97F4A7C15ULL;
   double next_uniform() {                 // xorshift64*: our own stream (the reference uses rand(), Sobject.cpp:652-659)
      rng ^= rng >> 12; rng ^= rng << 25; rng ^= rng >> 27;
      return (double)((rng * 0x2545F4914F6CDD1DULL) >> 11) * (1.0 / 9007199254740992.0);
   }
};

int b2_dmrg_create(b2_ctx* ctx, b2_dmrg** out) {
   if (!ctx || !ctx->have_bk || !out) return fail(B2_ERR_STATE, "b2_dmrg_create: no bookkeeper");
   if (ctx->device < 0) return fail(B2_ERR_NO_DEVICE, "b2_dmrg_create: planning-only context, no CUDA device (there is no CPU fallback)");
   std::unique_ptr<b2_dmrg> d(new b2_dmrg);
   d->ctx = ctx; d->L = ctx->bk.L;
   d->mps.resize(d->L);
   for (int s = 0; s < d->L; s++) { TLayout t; t.build(ctx->bk, s); d->mps[s].assign((size_t)t.size, 0.0); }
   d->left.assign(d->L + 1, nullptr); d->right.assign(d->L + 1, nullptr);
   *out = d.release();
   return B2_OK;
}
static void dmrg_clear_plan_cache(b2_dmrg* d) {
   for (b2_dmrg::PlanSlot& p : d->plan_cache) { b2_heff_destroy(p.h); p.h = nullptr; p.key.clear(); }
   for (b2_dmrg::UpdSlot& p : d->upd_cache) { b2_update_destroy(p.u); p.u = nullptr; p.key.clear(); }
}
void b2_dmrg_destroy(b2_dmrg* d) {
   if (!d) return;
   dmrg_clear_plan_cache(d);
   for (b2_opset* s : d->left) b2_opset_destroy(s);
   for (b2_opset* s : d->right) b2_opset_destroy(s);
   delete d;
}
int64_t b2_dmrg_mps_size(const b2_dmrg* d, int site) { return (d && site >= 0 && site < d->L) ? (int64_t)d->mps[site].size() : -1; }
int b2_dmrg_set_mps(b2_dmrg* d, int site, const double* t) {
   if (!d || site < 0 || site >= d->L || !t) return fail(B2_ERR_ARG, "b2_dmrg_set_mps: bad arguments");
   TLayout lay; lay.build(d->ctx->bk, site);
   d->mps[site].assign(t, t + lay.size);
   return B2_OK;
}
int b2_dmrg_get_mps(const b2_dmrg* d, int site, double* t) {
   if (!d || site < 0 || site >= d->L || !t) return fail(B2_ERR_ARG, "b2_dmrg_get_mps: bad arguments");
   std::memcpy(t, d->mps[site].data(), sizeof(double) * d->mps[site].size());
   return B2_OK;
}
int b2_dmrg_random_mps(b2_dmrg* d, uint64_t seed) {
   if (!d) return fail(B2_ERR_ARG, "b2_dmrg_random_mps: NULL");
   d->rng = seed * 0x9E3779B97F4A7C15ULL + 0xD1B54A32D192ED03ULL;
   for (int s = 0; s < d->L; s++) {   // DMRG::setupBookkeeperAndMPS (DMRG.cpp:149-169): random() then left_normalize with R discarded
      TLayout lay; lay.build(d->ctx->bk, s);
      d->mps[s].resize((size_t)lay.size);
      for (double& x : d->mps[s]) x = d->next_uniform();
      left_normalize_host(d->ctx->bk, lay, d->mps[s].data());
   }
   return B2_OK;
}
b2_opset* b2_dmrg_opset(b2_dmrg* d, int boundary, int moving_right) {
   if (!d || boundary < 0 || boundary > d->L) return nullptr;
   return moving_right ? d->left[boundary] : d->right[boundary];
}
int b2_dmrg_set_opset(b2_dmrg* d, int boundary, int moving_right, b2_opset* set) {
   if (!d || boundary < 1 || boundary > d->L - 1) return fail(B2_ERR_ARG, "b2_dmrg_set_opset: bad arguments");
   auto& slot = moving_right ? d->left[boundary] : d->right[boundary];
   if (slot && slot != set) b2_opset_destroy(slot);
   slot = set;
   return B2_OK;
}

int b2_dmrg_set_world(b2_dmrg* d, int world, int rank, b2_allreduce_fn fn, void* user) {
   if (!d || world < 1 || rank < 0 || rank >= world || (world > 1 && !fn)) return fail(B2_ERR_ARG, "b2_dmrg_set_world: bad arguments");
   d->world = world; d->rank = rank; d->allreduce = fn; d->allreduce_user = user;
   return B2_OK;
}
int b2_dmrg_set_spill(b2_dmrg* d, int enabled) {
   if (!d) return fail(B2_ERR_ARG, "b2_dmrg_set_spill: NULL");
   d->spill = enabled != 0;
   if (!d->spill)
      for (int b = 0; b <= d->L; b++) {
         int rc;
         if (d->left[b] && (rc = b2_opset_reload(d->left[b]))) return rc;
         if (d->right[b] && (rc = b2_opset_reload(d->right[b]))) return rc;
      }
   return B2_OK;
}
// make the sets `keep_l` (moving right) and `keep_r` (moving left) resident and, in spill mode, offload every other set
static int dmrg_residency(b2_dmrg* d, int keep_l, int keep_r) {
   int rc;
   if (keep_l >= 0 && keep_l <= d->L && d->left[keep_l] && (rc = b2_opset_reload(d->left[keep_l]))) return rc;
   if (keep_r >= 0 && keep_r <= d->L && d->right[keep_r] && (rc = b2_opset_reload(d->right[keep_r]))) return rc;
   if (!d->spill) return B2_OK;
   for (int b = 0; b <= d->L; b++) {
      if (b != keep_l && d->left[b] && (rc = b2_opset_offload(d->left[b]))) return rc;
      if (b != keep_r && d->right[b] && (rc = b2_opset_offload(d->right[b]))) return rc;
   }
   return B2_OK;
}
int b2_dmrg_set_plan_cache(b2_dmrg* d, int enabled) {
   if (!d) return fail(B2_ERR_ARG, "b2_dmrg_set_plan_cache: NULL");
   d->use_plan_cache = enabled != 0;
   if (!d->use_plan_cache) dmrg_clear_plan_cache(d);
   return B2_OK;
}
int b2_dmrg_plan_cache_stats(const b2_dmrg* d, long long* hits, long long* misses) {
   if (!d) return fail(B2_ERR_ARG, "b2_dmrg_plan_cache_stats: NULL");
   if (hits) *hits = d->plan_hits;
   if (misses) *misses = d->plan_misses;
   return B2_OK;
}
int b2_dmrg_timers(b2_dmrg* d, double* out5, int reset) {
   if (!d || !out5) return fail(B2_ERR_ARG, "b2_dmrg_timers: NULL");
   out5[0] = d->t_plan; out5[1] = d->t_solve; out5[2] = d->t_split; out5[3] = d->t_update; out5[4] = (double)d->n_matvec;
   if (reset) { d->t_plan = d->t_solve = d->t_split = d->t_update = 0.0; d->n_matvec = 0; }
   return B2_OK;
}

// DMRG::updateMovingRight(index) / updateMovingLeft(index-1): operators of the boundary next to site `index` from T = MPS[index]
// TensorO::update_ownmem / create (TensorO.cpp:38-196, formulas of TensorOperator::update with two_j = 0, no Jordan-Wigner phase) for every
// stored state: the overlap tensor of the boundary next to site `index` from MPS[index] of both states.
static int dmrg_update_overlaps(b2_dmrg* d, int index, bool mr) {
   if (d->exc.empty()) return B2_OK;
   b2_ctx* ctx = d->ctx;
   const Bookkeeper& bk = ctx->bk;
   const int L = d->L, b_old = mr ? index : index + 1, b_new = mr ? index + 1 : index;
   cudaStream_t s = ctx->stream;
   for (ExcState& e : d->exc) {
      if ((int)e.left.size() != L + 1) { e.left.assign(L + 1, Overlap()); e.right.assign(L + 1, Overlap()); }
      const Overlap& oldo = mr ? e.left[b_old] : e.right[b_old];
      const bool edge = mr ? (index == 0) : (index == L - 1);
      if (!edge && !oldo.valid) return fail(B2_ERR_STATE, "overlap tensor of boundary %d is missing", b_old);
      Overlap fresh;
      int64_t off = 0;
      bk.for_sectors(b_new, [&](int n, int ts, int ir) {
         const int r = bk.dim(b_new, n, ts, ir), c = e.bk.dim(b_new, n, ts, ir);
         if (r > 0 && c > 0) { fresh.blk.push_back({n, ts, ir, r, c, off}); off += ((int64_t)r * c + 15) / 16 * 16; }
      });
      fresh.data.assign((size_t)std::max<int64_t>(off, 1), 0.0);
      TLayout Tc, Te;
      Tc.build(bk, index); Te.build(e.bk, index);
      std::vector<Term3> terms;
      std::vector<DstBlock> dst;
      for (size_t k = 0; k < fresh.blk.size(); k++) {
         const Overlap::Blk& nb = fresh.blk[k];
         dst.push_back(DstBlock{nb.off, nb.rows, nb.cols});
         for (int geval = 0; geval < 4; geval++) {   // site empty / doubly occupied / singly occupied with spin down or up coupling
            int on, ots, oir;
            const int sg = mr ? -1 : +1;
            if (geval == 0) { on = nb.n; ots = nb.ts; oir = nb.ir; }
            else if (geval == 1) { on = nb.n + 2 * sg; ots = nb.ts; oir = nb.ir; }
            else { on = nb.n + sg; ots = nb.ts + (geval == 2 ? -1 : 1); oir = xorp(nb.ir, bk.orb_irrep[index]); }
            if (ots < 0) continue;
            const int kc = mr ? Tc.kappa(bk, on, ots, oir, nb.n, nb.ts, nb.ir) : Tc.kappa(bk, nb.n, nb.ts, nb.ir, on, ots, oir);
            const int ke = mr ? Te.kappa(e.bk, on, ots, oir, nb.n, nb.ts, nb.ir) : Te.kappa(e.bk, nb.n, nb.ts, nb.ir, on, ots, oir);
            if (kc < 0 || ke < 0) continue;
            Term3 t;
            t.dst = (int)k;
            t.f = (!mr && geval >= 2) ? (ots + 1.0) / (nb.ts + 1.0) : 1.0;                       // TensorOperator.cpp:370-372 with two_j = 0
            t.p.space = SP_LEFT; t.p.off = Tc.blk[kc].off; t.p.rows = Tc.blk[kc].rows; t.p.cols = Tc.blk[kc].cols; t.p.trans = mr ? 1 : 0;
            t.r.space = SP_RIGHT; t.r.off = Te.blk[ke].off; t.r.rows = Te.blk[ke].rows; t.r.cols = Te.blk[ke].cols; t.r.trans = mr ? 0 : 1;
            if (edge) {   // TensorO::create: the outer boundary carries the 1 x 1 identity
               if (bk.dim(b_old, on, ots, oir) != e.bk.dim(b_old, on, ots, oir)) continue;
            } else {
               const Overlap::Blk* ob = oldo.find(on, ots, oir);
               if (!ob) continue;
               t.q.space = SP_VIN; t.q.off = ob->off; t.q.rows = ob->rows; t.q.cols = ob->cols;
            }
            terms.push_back(t);
         }
      }
      struct Buf { double* p = nullptr; ~Buf() { cudaFree(p); } } dTc, dTe, dOld, dNew;
      CUDA_TRY(cudaMalloc(&dTc.p, sizeof(double) * (size_t)std::max<int64_t>(Tc.size, 1)));
      CUDA_TRY(cudaMalloc(&dTe.p, sizeof(double) * (size_t)std::max<int64_t>(Te.size, 1)));
      CUDA_TRY(cudaMalloc(&dOld.p, sizeof(double) * std::max<size_t>(oldo.data.size(), 1)));
      CUDA_TRY(cudaMalloc(&dNew.p, sizeof(double) * fresh.data.size()));
      CUDA_TRY(cudaMemcpyAsync(dTc.p, d->mps[index].data(), sizeof(double) * (size_t)Tc.size, cudaMemcpyHostToDevice, s));
      CUDA_TRY(cudaMemcpyAsync(dTe.p, e.mps[index].data(), sizeof(double) * (size_t)Te.size, cudaMemcpyHostToDevice, s));
      if (!oldo.data.empty()) CUDA_TRY(cudaMemcpyAsync(dOld.p, oldo.data.data(), sizeof(double) * oldo.data.size(), cudaMemcpyHostToDevice, s));
      CUDA_TRY(cudaMemsetAsync(dNew.p, 0, sizeof(double) * fresh.data.size(), s));
      CompiledWork w;
      compile_terms(w, terms, dst, SP_VOUT, budgeted(ctx));
      DevBases b;
      for (int i = 0; i < SP_COUNT; i++) b.p[i] = nullptr;
      b.p[SP_LEFT] = dTc.p; b.p[SP_RIGHT] = dTe.p; b.p[SP_VIN] = dOld.p; b.p[SP_VOUT] = dNew.p;
      int rc = run_compiled_once(ctx, w, b);
      if (rc) return rc;
      CUDA_TRY(cudaMemcpyAsync(fresh.data.data(), dNew.p, sizeof(double) * fresh.data.size(), cudaMemcpyDeviceToHost, s));
      CUDA_TRY(cudaStreamSynchronize(s));
      fresh.valid = true;
      (mr ? e.left[b_new] : e.right[b_new]) = std::move(fresh);
   }
   return B2_OK;
}

// DMRG::calcVeffTilde (DMRGtechnics.cpp:540-620) for every stored state, straight into the device slab of the sigma plan:
//   Vtilde[kappa] = sqrt(Eshift) / (2S+1) * sqrt(2SR+1) * O_left[l] * Sup[kappa] * O_right[r]^T ,   Sup = Join of the stored state's two tensors
static int dmrg_attach_excitations(b2_dmrg* d, b2_heff* h, int index) {
   if (d->exc.empty()) return B2_OK;
   b2_ctx* ctx = d->ctx;
   const int L = d->L, nexc = (int)d->exc.size();
   cudaStream_t s = ctx->stream;
   const SLayout& S = h->plan.S;
   const size_t n = (size_t)S.size;
   cudaFree(h->d_exc); cudaFree(h->d_exc_coef); cudaFree(h->d_exc_scratch);
   h->d_exc = h->d_exc_coef = h->d_exc_scratch = nullptr; h->n_exc = 0;
   CUDA_TRY(cudaMalloc(&h->d_exc, sizeof(double) * std::max<size_t>(n, 1) * nexc));
   CUDA_TRY(cudaMalloc(&h->d_exc_coef, sizeof(double) * nexc));
   CUDA_TRY(cudaMalloc(&h->d_exc_scratch, sizeof(double) * kRedScratch));
   CUDA_TRY(cudaMemsetAsync(h->d_exc_scratch, 0, sizeof(double) * kRedScratch, s));
   CUDA_TRY(cudaMemsetAsync(h->d_exc, 0, sizeof(double) * std::max<size_t>(n, 1) * nexc, s));
   for (int st = 0; st < nexc; st++) {
      ExcState& e = d->exc[st];
      const Overlap* ol = index > 0 ? &e.left[index] : nullptr;
      const Overlap* orr = index < L - 2 ? &e.right[index + 2] : nullptr;
      if ((ol && !ol->valid) || (orr && !orr->valid)) return fail(B2_ERR_STATE, "overlap tensors for site %d are missing", index);
      SLayout Se;
      Se.build(e.bk, index);
      TLayout TLe, TRe;
      TLe.build(e.bk, index); TRe.build(e.bk, index + 1);
      struct Buf { double* p = nullptr; ~Buf() { cudaFree(p); } } dTl, dTr, dSup, dOl, dOr;
      CUDA_TRY(cudaMalloc(&dTl.p, sizeof(double) * (size_t)std::max<int64_t>(TLe.size, 1)));
      CUDA_TRY(cudaMalloc(&dTr.p, sizeof(double) * (size_t)std::max<int64_t>(TRe.size, 1)));
      CUDA_TRY(cudaMalloc(&dSup.p, sizeof(double) * (size_t)std::max<int64_t>(Se.size, 1)));
      CUDA_TRY(cudaMemcpyAsync(dTl.p, e.mps[index].data(), sizeof(double) * (size_t)TLe.size, cudaMemcpyHostToDevice, s));
      CUDA_TRY(cudaMemcpyAsync(dTr.p, e.mps[index + 1].data(), sizeof(double) * (size_t)TRe.size, cudaMemcpyHostToDevice, s));
      CUDA_TRY(cudaMemsetAsync(dSup.p, 0, sizeof(double) * (size_t)std::max<int64_t>(Se.size, 1), s));
      {  // Sup = Join of the stored state (Sobject::Join with its own bookkeeper)
         std::vector<Term3> jt; std::vector<DstBlock> jd;
         join_terms(jt, jd, e.bk, Se, TLe, TRe);
         CompiledWork jw;
         compile_terms(jw, jt, jd, SP_VOUT, budgeted(ctx));
         DevBases b;
         for (int i = 0; i < SP_COUNT; i++) b.p[i] = nullptr;
         b.p[SP_LEFT] = dTl.p; b.p[SP_RIGHT] = dTr.p; b.p[SP_VOUT] = dSup.p;
         int rc = run_compiled_once(ctx, jw, b);
         if (rc) return rc;
      }
      if (ol) { CUDA_TRY(cudaMalloc(&dOl.p, sizeof(double) * ol->data.size())); CUDA_TRY(cudaMemcpyAsync(dOl.p, ol->data.data(), sizeof(double) * ol->data.size(), cudaMemcpyHostToDevice, s)); }
      if (orr) { CUDA_TRY(cudaMalloc(&dOr.p, sizeof(double) * orr->data.size())); CUDA_TRY(cudaMemcpyAsync(dOr.p, orr->data.data(), sizeof(double) * orr->data.size(), cudaMemcpyHostToDevice, s)); }
      std::vector<Term3> terms;
      std::vector<DstBlock> dst(S.nkappa());
      const double pref = std::sqrt(e.eshift) / (ctx->prob.twoS + 1.0);
      for (int k = 0; k < S.nkappa(); k++) {
         dst[k] = DstBlock{S.blk[k].off, S.blk[k].rows, S.blk[k].cols};
         const int ke = Se.kappa(e.bk, S.NL[k], S.twoSL[k], S.IL[k], S.N1[k], S.N2[k], S.twoJ[k], S.NR[k], S.twoSR[k], S.IR[k]);
         if (ke < 0) continue;
         Term3 t;
         t.dst = k; t.f = pref * std::sqrt(S.twoSR[k] + 1.0);
         t.q.space = SP_VIN; t.q.off = Se.blk[ke].off; t.q.rows = Se.blk[ke].rows; t.q.cols = Se.blk[ke].cols;
         if (ol) {
            const Overlap::Blk* ob = ol->find(S.NL[k], S.twoSL[k], S.IL[k]);
            if (!ob) continue;
            t.p.space = SP_LEFT; t.p.off = ob->off; t.p.rows = ob->rows; t.p.cols = ob->cols;
         } else if (S.blk[k].rows != Se.blk[ke].rows) continue;
         if (orr) {
            const Overlap::Blk* ob = orr->find(S.NR[k], S.twoSR[k], S.IR[k]);
            if (!ob) continue;
            t.r.space = SP_RIGHT; t.r.off = ob->off; t.r.rows = ob->rows; t.r.cols = ob->cols; t.r.trans = 1;
         } else if (S.blk[k].cols != Se.blk[ke].cols) continue;
         terms.push_back(t);
      }
      CompiledWork w;
      compile_terms(w, terms, dst, SP_VOUT, budgeted(ctx));
      DevBases b;
      for (int i = 0; i < SP_COUNT; i++) b.p[i] = nullptr;
      b.p[SP_LEFT] = dOl.p; b.p[SP_RIGHT] = dOr.p; b.p[SP_VIN] = dSup.p; b.p[SP_VOUT] = h->d_exc + (size_t)st * n;
      int rc = run_compiled_once(ctx, w, b);
      if (rc) return rc;
   }
   h->n_exc = nexc;
   return B2_OK;
}

// DMRG::activateExcitations + newExcitation (DMRG.cpp:464-505): the current MPS becomes lower state number nStates-1 with level shift
// `eshift`; a fresh random MPS (bookkeeper re-initialised for virtual dimension D) takes its place and every operator set is dropped.
int b2_dmrg_new_excitation(b2_dmrg* d, double eshift, int D, uint64_t seed) {
   if (!d || D < 1) return fail(B2_ERR_ARG, "b2_dmrg_new_excitation: bad arguments");
   ExcState e;
   e.eshift = eshift; e.bk = d->ctx->bk; e.mps = d->mps;
   e.left.assign(d->L + 1, Overlap()); e.right.assign(d->L + 1, Overlap());
   d->exc.push_back(std::move(e));
   for (int b = 0; b <= d->L; b++) {
      if (d->left[b]) { b2_opset_destroy(d->left[b]); d->left[b] = nullptr; }
      if (d->right[b]) { b2_opset_destroy(d->right[b]); d->right[b] = nullptr; }
   }
   for (ExcState& x : d->exc) { x.left.assign(d->L + 1, Overlap()); x.right.assign(d->L + 1, Overlap()); }
   d->ctx->bk.init(d->ctx->prob, D);
   dmrg_clear_plan_cache(d);
   d->max_disc_last_sweep = 0.0;
   d->swept_once = false;
   return b2_dmrg_random_mps(d, seed);
}
int b2_dmrg_num_lower_states(const b2_dmrg* d) { return d ? (int)d->exc.size() : 0; }

static int dmrg_update_mode(b2_dmrg* d, int index, int moving_right, int mode);
int b2_dmrg_update(b2_dmrg* d, int index, int moving_right) { return dmrg_update_mode(d, index, moving_right, 0); }
// mode 0: the full operator complement of a sweep; 1: L, S0, S1, F0, F1 (updateMovingLeftSafe2DM); 2: L only
static int dmrg_update_mode(b2_dmrg* d, int index, int moving_right, int mode) {
   if (!d || index < 0 || index >= d->L) return fail(B2_ERR_ARG, "b2_dmrg_update: bad arguments");
   b2_ctx* ctx = d->ctx;
   const bool mr = moving_right != 0;
   const int b_old = mr ? index : index + 1, b_new = mr ? index + 1 : index;
   if (b_new < 1 || b_new > d->L - 1) return fail(B2_ERR_ARG, "b2_dmrg_update: no operators live at boundary %d", b_new);
   b2_opset* old_set = mr ? d->left[b_old] : d->right[b_old];
   const bool need_old = mr ? (index > 0) : (index < d->L - 1);
   if (need_old && old_set) { int rr = b2_opset_reload(old_set); if (rr) return rr; }
   if (need_old && !old_set) return fail(B2_ERR_STATE, "b2_dmrg_update: operators of boundary %d are missing", b_old);
   b2_opset* fresh = nullptr;
   b2_update* u = nullptr;
   const double t0 = wall_seconds();
   int rc = B2_OK;
   std::vector<int> key;
   b2_dmrg::UpdSlot* slot = nullptr;
   if (d->use_plan_cache) {
      if ((int)d->upd_cache.size() != 2 * d->L) d->upd_cache.assign(2 * d->L, b2_dmrg::UpdSlot());
      slot = &d->upd_cache[2 * index + (mr ? 1 : 0)];
      key.push_back(d->world); key.push_back(d->rank); key.push_back(mode);
      for (int b = index; b <= index + 1; b++) key.insert(key.end(), ctx->bk.cur[b].begin(), ctx->bk.cur[b].end());
   }
   for (int attempt = 0; attempt < 2; attempt++) {
      rc = mode == 0 ? b2_opset_create(ctx, b_new, mr, &fresh) : opset_create_reduced(ctx, b_new, mr, mode == 2, &fresh);
      if (!rc && slot && slot->u && slot->key == key) {   // the plan of the previous visit fits: re-bind it to the new arenas
         u = slot->u; slot->u = nullptr;
         rc = update_unpark(u, need_old ? old_set : nullptr, fresh);
         if (!rc) d->plan_hits++;
      } else if (!rc) {
         d->plan_misses++;
         rc = b2_update_create_sharded(ctx, index, mr, need_old ? old_set : nullptr, fresh, d->world, d->rank, &u);
      }
      if (rc != B2_ERR_CUDA || attempt == 1 || d->spill) break;
      // HBM exhausted (O(L) boundaries x O(L^2 D^2) operators): from now on only the sets in use stay resident — the
      // reference's OperatorsOnDisk mode, switched on when it is needed instead of by the user
      cudaGetLastError();
      b2_update_destroy(u); u = nullptr;
      b2_opset_destroy(fresh); fresh = nullptr;
      d->spill = true;
      dmrg_clear_plan_cache(d);
      if ((rc = dmrg_residency(d, mr ? b_old : -1, mr ? -1 : b_old))) return rc;
   }
   if (rc) { b2_update_destroy(u); b2_opset_destroy(fresh); return rc; }
   if (d->world > 1) rc = b2_update_set_allreduce(u, d->allreduce, d->allreduce_user);
   d->t_plan += wall_seconds() - t0;
   const double t1 = wall_seconds();
   if (!rc) rc = b2_update_run(u, d->mps[index].data());
   d->t_update += wall_seconds() - t1;
   const double t2 = wall_seconds();
   double ubytes = 0.0;
   if (u) for (int p = 0; p < 2; p++) ubytes += u->list_bytes[p];
   if (!rc && slot && ubytes <= 1.0e9) {
      b2_update_destroy(slot->u);
      update_park(u);
      slot->u = u; slot->key = key;
   } else b2_update_destroy(u);
   if (rc) { b2_opset_destroy(fresh); return rc; }
   if ((rc = b2_dmrg_set_opset(d, b_new, mr, fresh))) return rc;
   rc = dmrg_update_overlaps(d, index, mr);   // DMRGoperators.cpp:556-567 / :889-900
   d->t_tail += wall_seconds() - t2;
   return rc;
}

// DMRG::solve_site (DMRG.cpp:419-452): Join -> Heff::SolveDAVIDSON -> (noise) -> Split.  *energy includes Econst.
int b2_dmrg_solve_site(b2_dmrg* d, int index, double rtol, double noise, int D, int moving_right, int change, double* energy,
                       double* discarded_weight, int* n_matvec) {
   if (!d || index < 0 || index > d->L - 2 || !energy) return fail(B2_ERR_ARG, "b2_dmrg_solve_site: bad arguments");
   b2_ctx* ctx = d->ctx;
   const int L = d->L;
   cudaStream_t s = ctx->stream;
   { int rr = dmrg_residency(d, index > 0 ? index : -1, index < L - 2 ? index + 2 : -1); if (rr) return rr; }
   b2_opset* lset = index > 0 ? d->left[index] : nullptr;
   b2_opset* rset = index < L - 2 ? d->right[index + 2] : nullptr;
   if ((index > 0 && !lset) || (index < L - 2 && !rset)) return fail(B2_ERR_STATE, "b2_dmrg_solve_site: boundary operators for site %d are missing", index);
   b2_heff* h = nullptr;
   const double tp0 = wall_seconds();
   std::vector<int> key;
   if (d->use_plan_cache) {
      if ((int)d->plan_cache.size() != L) d->plan_cache.assign(L, b2_dmrg::PlanSlot());
      key.push_back(d->world); key.push_back(d->rank);
      for (int b = index; b <= index + 2; b++) key.insert(key.end(), ctx->bk.cur[b].begin(), ctx->bk.cur[b].end());
      b2_dmrg::PlanSlot& slot = d->plan_cache[index];
      if (slot.h && slot.key == key) {
         h = slot.h; slot.h = nullptr;
         int ur = heff_unpark(h, lset, rset);
         if (ur) { b2_heff_destroy(h); h = nullptr; cudaGetLastError(); } else d->plan_hits++;
      }
   }
   int rc = B2_OK;
   if (!h) { d->plan_misses++; rc = b2_heff_create(ctx, index, lset, rset, d->world, d->rank, &h); }
   if (rc == B2_ERR_CUDA && !d->spill) {   // out of HBM: park every operator set that this site does not use and retry
      cudaGetLastError();
      b2_heff_destroy(h); h = nullptr;
      d->spill = true;
      dmrg_clear_plan_cache(d);
      if ((rc = dmrg_residency(d, index > 0 ? index : -1, index < L - 2 ? index + 2 : -1))) return rc;
      rc = b2_heff_create(ctx, index, lset, rset, d->world, d->rank, &h);
   }
   if (rc) return rc;
   if (d->world > 1) b2_heff_set_allreduce(h, d->allreduce, d->allreduce_user);
   if ((rc = dmrg_attach_excitations(d, h, index))) { b2_heff_destroy(h); return rc; }   // DMRG::prepare_excitations (DMRG.cpp:434)
   d->t_plan += wall_seconds() - tp0;
   const double tj0 = wall_seconds();
   const SLayout& S = h->plan.S;
   TLayout TL, TR;
   TL.build(ctx->bk, index); TR.build(ctx->bk, index + 1);
   double *d_tl = nullptr, *d_tr = nullptr, *d_s = nullptr;
   std::vector<double> s_host((size_t)S.size);
   do {
      if (cudaMalloc(&d_tl, sizeof(double) * (size_t)std::max<int64_t>(TL.size, 1)) != cudaSuccess || cudaMalloc(&d_tr, sizeof(double) * (size_t)std::max<int64_t>(TR.size, 1)) != cudaSuccess ||
          cudaMalloc(&d_s, sizeof(double) * (size_t)std::max<int64_t>(S.size, 1)) != cudaSuccess) { rc = fail(B2_ERR_CUDA, "b2_dmrg_solve_site: allocation failed"); break; }
      cudaMemcpyAsync(d_tl, d->mps[index].data(), sizeof(double) * (size_t)TL.size, cudaMemcpyHostToDevice, s);
      cudaMemcpyAsync(d_tr, d->mps[index + 1].data(), sizeof(double) * (size_t)TR.size, cudaMemcpyHostToDevice, s);
      // ---- Join (Sobject.cpp:212-258) on the device
      std::vector<Term3> jt; std::vector<DstBlock> jd;
      join_terms(jt, jd, ctx->bk, S, TL, TR);
      CompiledWork jw;
      compile_terms(jw, jt, jd, SP_VOUT, ctx->copt);
      DevBases b;
      for (int i = 0; i < SP_COUNT; i++) b.p[i] = nullptr;
      b.p[SP_LEFT] = d_tl; b.p[SP_RIGHT] = d_tr; b.p[SP_VOUT] = d_s;
      if (dev_fill_zero(d_s, S.size, s)) { rc = fail(B2_ERR_CUDA, "%s", dev_last_error()); break; }
      if ((rc = run_compiled_once(ctx, jw, b))) break;
      // ---- Heff::SolveDAVIDSON on the device
      double ev = 0.0; int nm = 0;
      const double ts0 = wall_seconds();
      d->t_join += ts0 - tj0;
      if ((rc = b2_heff_solve_device(h, d_s, rtol, &ev, &nm))) break;
      d->t_solve += wall_seconds() - ts0; d->n_matvec += nm;
      *energy = ev + ctx->prob.econst;
      if (n_matvec) *n_matvec = nm;
      if (cudaMemcpyAsync(s_host.data(), d_s, sizeof(double) * (size_t)S.size, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) { rc = fail(B2_ERR_CUDA, "b2_dmrg_solve_site: D2H failed"); break; }
      if (noise > 0.0) for (double& x : s_host) x += (d->next_uniform() - 0.5) * noise;   // Sobject::addNoise
      // ---- Split (host SVD + truncation); the bookkeeper dims of boundary index+1 change here
      SLayout Scopy = S;
      const double tq0 = wall_seconds();
      char svd_err[256] = "";
      SvdBatchFn svd = [&](std::vector<SvdJob>& jobs) { return dev_svd_batch(jobs, (void*)s, svd_err, (int)sizeof(svd_err)); };
      const double dw = split_host(ctx->bk, index, Scopy, s_host.data(), D, moving_right != 0, change != 0, d->mps[index], d->mps[index + 1], svd);
      d->t_split += wall_seconds() - tq0;
      if (dw < 0.0) { rc = fail(B2_ERR_CUDA, "b2_dmrg_solve_site: Split: %s", svd_err); break; }
      if (discarded_weight) *discarded_weight = dw;
   } while (0);
   const double tr0 = wall_seconds();
   cudaFree(d_tl); cudaFree(d_tr); cudaFree(d_s);
   if (!rc && d->use_plan_cache && h->list_bytes <= 1.0e9) {   // keep the plan for the next visit of this site (device work lists only, < 1 GB)
      b2_dmrg::PlanSlot& slot = d->plan_cache[index];
      b2_heff_destroy(slot.h);
      heff_park(h);
      slot.h = h; slot.key = key;
   } else b2_heff_destroy(h);
   if (!rc) {   // operator sets living at the re-dimensioned boundary are stale now
      b2_dmrg_set_opset(d, index + 1, 1, nullptr);
      b2_dmrg_set_opset(d, index + 1, 0, nullptr);
   }
   d->t_release += wall_seconds() - tr0;
   return rc;
}

// MPS checkpoint: the content of DMRG::saveMPS / loadDIM / loadMPS (DMRGmpsio.cpp:30-131: converged flag, every virtual dimension in
// the bookkeeper's enumeration order, the packed TensorT storage of every site) as one flat little-endian binary file — this image has
// no HDF5 library, so the reference's HDF5 container is not reproduced, only its payload (a shim can copy dataset by dataset).
int b2_dmrg_save_mps(const b2_dmrg* d, const char* path, int converged) {
   if (!d || !path) return fail(B2_ERR_ARG, "b2_dmrg_save_mps: NULL");
   FILE* f = std::fopen(path, "wb");
   if (!f) return fail(B2_ERR_ARG, "b2_dmrg_save_mps: cannot open %s", path);
   const Bookkeeper& bk = d->ctx->bk;
   const char magic[8] = {'B', '2', 'M', 'P', 'S', '0', '0', '1'};
   const int32_t hdr[6] = {bk.L, bk.N, bk.twoS, bk.irrep, bk.nirr, converged ? 1 : 0};
   bool ok = std::fwrite(magic, 1, 8, f) == 8 && std::fwrite(hdr, 4, 6, f) == 6;
   for (int b = 0; b <= bk.L && ok; b++)
      bk.for_sectors(b, [&](int n, int ts, int ir) { const int32_t v = bk.dim(b, n, ts, ir); ok = ok && std::fwrite(&v, 4, 1, f) == 1; });
   for (int sdx = 0; sdx < d->L && ok; sdx++) {
      const int64_t n = (int64_t)d->mps[sdx].size();
      ok = std::fwrite(&n, 8, 1, f) == 1 && (n == 0 || std::fwrite(d->mps[sdx].data(), 8, (size_t)n, f) == (size_t)n);
   }
   std::fclose(f);
   return ok ? B2_OK : fail(B2_ERR_STATE, "b2_dmrg_save_mps: write to %s failed", path);
}
int b2_dmrg_load_mps(b2_dmrg* d, const char* path, int* converged) {
   if (!d || !path) return fail(B2_ERR_ARG, "b2_dmrg_load_mps: NULL");
   FILE* f = std::fopen(path, "rb");
   if (!f) return fail(B2_ERR_ARG, "b2_dmrg_load_mps: cannot open %s", path);
   Bookkeeper& bk = d->ctx->bk;
   char magic[8];
   int32_t hdr[6];
   bool ok = std::fread(magic, 1, 8, f) == 8 && std::memcmp(magic, "B2MPS001", 8) == 0 && std::fread(hdr, 4, 6, f) == 6;
   if (ok && (hdr[0] != bk.L || hdr[1] != bk.N || hdr[2] != bk.twoS || hdr[3] != bk.irrep || hdr[4] != bk.nirr)) {
      std::fclose(f);
      return fail(B2_ERR_STATE, "b2_dmrg_load_mps: %s belongs to another problem (L, N, 2S, irrep, group differ)", path);
   }
   for (int b = 0; b <= bk.L && ok; b++)
      bk.for_sectors(b, [&](int n, int ts, int ir) { int32_t v = 0; ok = ok && std::fread(&v, 4, 1, f) == 1; if (ok) bk.set_dim(b, n, ts, ir, v); });
   for (int sdx = 0; sdx < d->L && ok; sdx++) {
      TLayout lay;
      lay.build(bk, sdx);
      int64_t n = -1;
      ok = std::fread(&n, 8, 1, f) == 1 && n == lay.size;
      if (ok) { d->mps[sdx].resize((size_t)n); ok = n == 0 || std::fread(d->mps[sdx].data(), 8, (size_t)n, f) == (size_t)n; }
   }
   std::fclose(f);
   if (!ok) return fail(B2_ERR_STATE, "b2_dmrg_load_mps: %s is truncated or inconsistent with the bookkeeper", path);
   if (converged) *converged = hdr[5];
   for (int b = 0; b <= d->L; b++) {   // the operators of the previous MPS are stale
      if (d->left[b]) { b2_opset_destroy(d->left[b]); d->left[b] = nullptr; }
      if (d->right[b]) { b2_opset_destroy(d->right[b]); d->right[b] = nullptr; }
   }
   for (ExcState& x : d->exc) { x.left.assign(d->L + 1, Overlap()); x.right.assign(d->L + 1, Overlap()); }
   dmrg_clear_plan_cache(d);
   return B2_OK;
}

// DMRG::PreSolve (DMRG.cpp:257-266): the moving-right operators of every boundary from the current MPS
int b2_dmrg_presolve(b2_dmrg* d) {
   if (!d) return fail(B2_ERR_ARG, "b2_dmrg_presolve: NULL");
   for (int i = 0; i < d->L - 2; i++) { int rc = b2_dmrg_update(d, i, 1); if (rc) return rc; }
   d->total_min_energy = 1e8;       // DMRG.cpp:263-264
   d->max_disc_last_sweep = 0.0;
   return B2_OK;
}

// DMRG::Solve (DMRG.cpp:268-355) for a ConvergenceScheme given as arrays: per instruction the virtual dimension, the energy convergence
// threshold, the maximum number of (left + right) sweeps, the noise prefactor and the Davidson residual tolerance.  The very first left
// sweep of a fresh MPS keeps the virtual dimensions fixed (`change` = false), exactly like the reference; returns the lowest energy met.
int b2_dmrg_solve(b2_dmrg* d, int n_instructions, const int* D, const double* energy_conv, const int* max_sweeps, const double* noise_prefactor,
                  const double* davidson_rtol, double* energy_out) {
   if (!d || n_instructions < 1 || !D || !energy_conv || !max_sweeps || !noise_prefactor || !davidson_rtol || !energy_out)
      return fail(B2_ERR_ARG, "b2_dmrg_solve: bad arguments");
   bool have_ops = true;
   for (int b = 1; b <= d->L - 2 && have_ops; b++) have_ops = d->left[b] != nullptr && !d->left[b]->set.reduced;
   int rc;
   if (!have_ops && (rc = b2_dmrg_presolve(d))) return rc;
   double energy = 0.0, lowest = 1e300;
   for (int ins = 0; ins < n_instructions; ins++) {
      int it = 0;
      double prev = energy + 10 * energy_conv[ins];   // at least one left-right sweep per instruction (DMRG.cpp:283)
      while (std::fabs(energy - prev) > energy_conv[ins] && it < max_sweeps[ins]) {
         prev = energy;
         double el, er, dw;
         if ((rc = b2_dmrg_sweep(d, 0, davidson_rtol[ins], noise_prefactor[ins], D[ins], d->swept_once ? 1 : 0, &el, &dw))) return rc;
         d->swept_once = true;
         if ((rc = b2_dmrg_sweep(d, 1, davidson_rtol[ins], noise_prefactor[ins], D[ins], 1, &er, &dw))) return rc;
         energy = d->last_energy;         // the convergence test compares what sweepright returns: the energy of its last site
         lowest = std::min(lowest, std::min(el, er));
         it++;
      }
   }
   *energy_out = lowest;
   return B2_OK;
}

// Move the orthogonality centre of the MPS by one site (TensorT::QR + LeftMultiply = DMRG::left_normalize, TensorT::LQ +
// RightMultiply = DMRG::right_normalize; TensorT.cpp:188-420, DMRGtechnics.cpp).  The orthogonal factor comes from the batched
// device SVD (any orthonormal basis of the same space is a valid gauge: Q = U resp. V^T, the other factor S V^T resp. U S), the
// neighbour absorbs that factor through the grouped contraction kernels.
//   to_left  = 0: MPS[site] becomes left-normalised, the factor goes into MPS[site+1]      (left_normalize)
//   to_left != 0: MPS[site] becomes right-normalised (U-convention weights sqrt((2SR+1)/(2SL+1)), TensorT.cpp:289-299),
//                 the factor goes into MPS[site-1]                                         (right_normalize)
static int dmrg_gauge_move(b2_dmrg* d, int site, bool to_left) {
   b2_ctx* ctx = d->ctx;
   const Bookkeeper& bk = ctx->bk;
   const int L = d->L;
   TLayout T;
   T.build(bk, site);
   std::vector<double>& t = d->mps[site];
   const int b_fix = to_left ? site : site + 1;        // boundary whose sectors index the decompositions
   struct Sector { int n, ts, ir, dim, tot; std::vector<int> blocks; std::vector<int> start; };
   std::vector<Sector> secs;
   bk.for_sectors(b_fix, [&](int n, int ts, int ir) {
      const int dm = bk.dim(b_fix, n, ts, ir);
      if (dm <= 0) return;
      Sector sc{n, ts, ir, dm, 0, {}, {}};
      for (int k = 0; k < T.nkappa(); k++) {
         const bool match = to_left ? (T.NL[k] == n && T.twoSL[k] == ts && T.IL[k] == ir) : (T.NR[k] == n && T.twoSR[k] == ts && T.IR[k] == ir);
         if (!match) continue;
         sc.blocks.push_back(k); sc.start.push_back(sc.tot);
         sc.tot += to_left ? T.blk[k].cols : T.blk[k].rows;
      }
      if (sc.tot > 0) secs.push_back(sc);
   });
   // stacked matrices: to_left: dim x tot (blocks side by side, weighted); else tot x dim (blocks on top of each other)
   std::vector<std::vector<double>> mem(secs.size()), sv(secs.size()), U(secs.size()), VT(secs.size());
   std::vector<SvdJob> jobs(secs.size());
   for (size_t i = 0; i < secs.size(); i++) {
      const Sector& sc = secs[i];
      const int m = to_left ? sc.dim : sc.tot, n = to_left ? sc.tot : sc.dim, kk = std::min(m, n);
      mem[i].assign((size_t)m * n, 0.0);
      for (size_t bi = 0; bi < sc.blocks.size(); bi++) {
         const int k = sc.blocks[bi];
         const Block& B = T.blk[k];
         const double f = to_left ? std::sqrt((T.twoSR[k] + 1.0) / (sc.ts + 1.0)) : 1.0;
         for (int c = 0; c < B.cols; c++)
            for (int r = 0; r < B.rows; r++) {
               const double x = f * t[B.off + r + (size_t)B.rows * c];
               if (to_left) mem[i][r + (size_t)m * (sc.start[bi] + c)] = x;
               else mem[i][sc.start[bi] + r + (size_t)m * c] = x;
            }
      }
      sv[i].resize(kk); U[i].resize((size_t)m * kk); VT[i].resize((size_t)kk * n);
      jobs[i].m = m; jobs[i].n = n; jobs[i].a = mem[i].data(); jobs[i].s = sv[i].data(); jobs[i].u = U[i].data(); jobs[i].vt = VT[i].data();
   }
   char err[256] = "";
   if (dev_svd_batch(jobs, (void*)ctx->stream, err, (int)sizeof(err))) return fail(B2_ERR_CUDA, "gauge move: %s", err);
   // ---- the orthonormal factor goes back into MPS[site]; the square factor F (dim x dim per sector) is collected for the neighbour
   std::vector<int64_t> foff(secs.size());
   int64_t ftot = 0;
   for (size_t i = 0; i < secs.size(); i++) { foff[i] = ftot; ftot += ((int64_t)secs[i].dim * secs[i].dim + 15) / 16 * 16; }
   std::vector<double> F((size_t)std::max<int64_t>(ftot, 1), 0.0);
   for (size_t i = 0; i < secs.size(); i++) {
      const Sector& sc = secs[i];
      const int m = to_left ? sc.dim : sc.tot, n = to_left ? sc.tot : sc.dim, kk = std::min(m, n), dm = sc.dim;
      double* Fi = F.data() + foff[i];
      if (to_left) {   // mem = (U S) V^T : F = U S (dim x kk, zero-padded to dim x dim), Q = V^T (kk x tot, zero rows below)
         for (int j = 0; j < kk; j++)
            for (int r = 0; r < dm; r++) Fi[r + (size_t)dm * j] = U[i][r + (size_t)m * j] * sv[i][j];
      } else {         // mem = U (S V^T) : Q = U (tot x kk, zero columns beyond), F = S V^T (kk x dim, zero rows below)
         for (int c = 0; c < dm; c++)
            for (int j = 0; j < kk; j++) Fi[j + (size_t)dm * c] = sv[i][j] * VT[i][j + (size_t)kk * c];
      }
      for (size_t bi = 0; bi < sc.blocks.size(); bi++) {
         const int k = sc.blocks[bi];
         const Block& B = T.blk[k];
         const double f = to_left ? std::sqrt((sc.ts + 1.0) / (T.twoSR[k] + 1.0)) : 1.0;
         for (int c = 0; c < B.cols; c++)
            for (int r = 0; r < B.rows; r++) {
               double x;
               if (to_left) x = (r < kk) ? f * VT[i][r + (size_t)kk * (sc.start[bi] + c)] : 0.0;
               else x = (c < kk) ? U[i][sc.start[bi] + r + (size_t)m * c] : 0.0;
               t[B.off + r + (size_t)B.rows * c] = x;
            }
      }
   }
   // ---- neighbour:  T_prev[. -> sector] <- T_prev F   resp.   T_next[sector -> .] <- F T_next      (device GEMMs)
   const int nb = to_left ? site - 1 : site + 1;
   if (nb < 0 || nb >= L) return B2_OK;
   TLayout N;
   N.build(bk, nb);
   std::vector<Term3> terms;
   std::vector<DstBlock> dst;
   for (int k = 0; k < N.nkappa(); k++) {
      dst.push_back(DstBlock{N.blk[k].off, N.blk[k].rows, N.blk[k].cols});
      const int sn = to_left ? N.NR[k] : N.NL[k], sts = to_left ? N.twoSR[k] : N.twoSL[k], sir = to_left ? N.IR[k] : N.IL[k];
      for (size_t i = 0; i < secs.size(); i++) {
         if (secs[i].n != sn || secs[i].ts != sts || secs[i].ir != sir) continue;
         Term3 x;
         x.dst = k; x.f = 1.0;
         MatRef tb, fb;
         tb.space = SP_LEFT; tb.off = N.blk[k].off; tb.rows = N.blk[k].rows; tb.cols = N.blk[k].cols;
         fb.space = SP_RIGHT; fb.off = foff[i]; fb.rows = secs[i].dim; fb.cols = secs[i].dim;
         if (to_left) { x.q = tb; x.r = fb; } else { x.p = fb; x.q = tb; }
         terms.push_back(x);
      }
   }
   cudaStream_t s = ctx->stream;
   struct Buf { double* p = nullptr; ~Buf() { cudaFree(p); } } dOld, dNew, dF;
   const size_t nbytes = sizeof(double) * (size_t)std::max<int64_t>(N.size, 1);
   CUDA_TRY(cudaMalloc(&dOld.p, nbytes));
   CUDA_TRY(cudaMalloc(&dNew.p, nbytes));
   CUDA_TRY(cudaMalloc(&dF.p, sizeof(double) * F.size()));
   CUDA_TRY(cudaMemcpyAsync(dOld.p, d->mps[nb].data(), sizeof(double) * (size_t)N.size, cudaMemcpyHostToDevice, s));
   CUDA_TRY(cudaMemcpyAsync(dF.p, F.data(), sizeof(double) * F.size(), cudaMemcpyHostToDevice, s));
   CUDA_TRY(cudaMemsetAsync(dNew.p, 0, nbytes, s));
   CompiledWork w;
   compile_terms(w, terms, dst, SP_VOUT, budgeted(ctx));
   DevBases b;
   for (int i = 0; i < SP_COUNT; i++) b.p[i] = nullptr;
   b.p[SP_LEFT] = dOld.p; b.p[SP_RIGHT] = dF.p; b.p[SP_VOUT] = dNew.p;
   int rc = run_compiled_once(ctx, w, b);
   if (rc) return rc;
   CUDA_TRY(cudaMemcpyAsync(d->mps[nb].data(), dNew.p, sizeof(double) * (size_t)N.size, cudaMemcpyDeviceToHost, s));
   CUDA_TRY(cudaStreamSynchronize(s));
   return B2_OK;
}

// DMRG::calc_rdms_and_correlations, 2-RDM part (DMRGtechnics.cpp:40-113): whole MPS into left-canonical form, moving-right operators
// of every boundary, then site by site from the right: TwoDM::FillSite, right-normalise, moving-left operators one boundary further.
int b2_dmrg_calc_2rdm(b2_dmrg* d, double* two_rdm_A, double* two_rdm_B) {
   if (!d || !two_rdm_A || !two_rdm_B) return fail(B2_ERR_ARG, "b2_dmrg_calc_2rdm: NULL");
   const int L = d->L;
   const size_t n4 = (size_t)L * L * L * L;
   std::fill(two_rdm_A, two_rdm_A + n4, 0.0);
   std::fill(two_rdm_B, two_rdm_B + n4, 0.0);
   int rc;
   for (int s = 0; s < L; s++) {
      if ((rc = dmrg_gauge_move(d, s, false))) return rc;          // the last one discards the norm (left_normalize(MPS[L-1], NULL))
      if (s < L - 1 && (rc = dmrg_update_mode(d, s, 1, 2))) return rc;   // L operators of boundary s+1 from the left-normalised MPS[s]
   }
   for (int site = L - 1; site >= 0; site--) {
      b2_opset* lset = site > 0 ? d->left[site] : nullptr;
      b2_opset* rset = site < L - 1 ? d->right[site + 1] : nullptr;
      if (lset && (rc = b2_opset_reload(lset))) return rc;
      if (rset && (rc = b2_opset_reload(rset))) return rc;
      if ((rc = b2_twodm_fill_site(d->ctx, site, d->mps[site].data(), lset, rset, two_rdm_A, two_rdm_B))) return rc;
      if (site > 0) {
         if ((rc = dmrg_gauge_move(d, site, true))) return rc;
         if ((rc = dmrg_update_mode(d, site, 0, 1))) return rc;    // updateMovingLeftSafe2DM(site-1): L, S0, S1, F0, F1 of boundary `site`
      }
   }
   if (d->ctx->prob.twoS != 0) {                                    // TwoDM::correct_higher_multiplicities (TwoDM.cpp:630-640)
      const double alpha = 1.0 / (d->ctx->prob.twoS + 1.0);
      for (size_t i = 0; i < n4; i++) { two_rdm_A[i] *= alpha; two_rdm_B[i] *= alpha; }
   }
   return B2_OK;
}

// The Correlations part of DMRG::calc_rdms_and_correlations (DMRGtechnics.cpp:150-175): spin / density / spin-flip / singlet-diradical
// correlation functions from the 2-RDM (Correlations::FillSpinDensSpinflip, Correlations.cpp:69-103) and the two-orbital mutual
// information from the G/Y/Z/K/M tensors, site by site from the left.
int b2_dmrg_calc_correlations(b2_dmrg* d, const double* A, const double* B, double* Cspin, double* Cdens, double* Cspinflip, double* Cdirad,
                              double* MutInfo) {
   if (!d || !A || !B || !Cspin || !Cdens || !Cspinflip || !Cdirad || !MutInfo) return fail(B2_ERR_ARG, "b2_dmrg_calc_correlations: NULL");
   b2_ctx* ctx = d->ctx;
   const int L = d->L, N = ctx->prob.N;
   auto irr = [&](int o) { return ctx->bk.orb_irrep[o]; };
   auto getA = [&](int i, int j, int k, int l) { return (xorp(irr(i), irr(j)) == xorp(irr(k), irr(l))) ? A[i + L * (j + L * (k + L * (size_t)l))] : 0.0; };
   auto getB = [&](int i, int j, int k, int l) { return (xorp(irr(i), irr(j)) == xorp(irr(k), irr(l))) ? B[i + L * (j + L * (k + L * (size_t)l))] : 0.0; };
   std::vector<double> n1(L);
   for (int i = 0; i < L; i++) { double v = 0.0; for (int o = 0; o < L; o++) v += getA(i, o, i, o); n1[i] = v / (N - 1.0); }
   for (int r = 0; r < L; r++)
      for (int c = 0; c < L; c++) {
         Cspin[r + L * c] = getB(r, c, r, c) + (r == c ? n1[r] : 0.0);
         Cdens[r + L * c] = getA(r, c, r, c) - n1[r] * n1[c] + (r == c ? n1[r] : 0.0);
         Cspinflip[r + L * c] = 0.5 * (getB(r, c, c, r) - getA(r, c, c, r)) + (r == c ? n1[r] : 0.0);
         Cdirad[r + L * c] = -0.5 * (n1[r] - getA(r, r, r, r)) * (n1[c] - getA(c, c, c, c));
         MutInfo[r + L * c] = 0.0;
      }
   int rc;
   for (int site = L - 1; site >= 1; site--)
      if ((rc = dmrg_gauge_move(d, site, true))) return rc;        // right-canonical, orthogonality centre on site 0
   b2_opset* old_set = nullptr;
   for (int site = 1; site < L; site++) {
      if ((rc = dmrg_gauge_move(d, site - 1, false))) { b2_opset_destroy(old_set); return rc; }   // left_normalize(MPS[site-1], MPS[site])
      b2_opset* fresh = nullptr;
      b2_update* u = nullptr;
      rc = b2_opset_create_correlation(ctx, site, &fresh);                                        // update_correlations_tensors(site)
      if (!rc) rc = b2_update_create(ctx, site - 1, 1, old_set, fresh, &u);
      if (!rc) rc = b2_update_run(u, d->mps[site - 1].data());
      b2_update_destroy(u);
      if (!rc) rc = b2_corr_fill_site(ctx, site, d->mps[site].data(), fresh, A, B, Cdirad, MutInfo);
      b2_opset_destroy(old_set);
      old_set = fresh;
      if (rc) { b2_opset_destroy(old_set); return rc; }
   }
   b2_opset_destroy(old_set);
   return B2_OK;
}

// DMRG::sweepleft / sweepright (DMRG.cpp:357-417): returns the lowest site energy of the half sweep
int b2_dmrg_sweep(b2_dmrg* d, int to_right, double rtol, double noise, int D, int change, double* min_energy, double* max_discarded) {
   if (!d || !min_energy) return fail(B2_ERR_ARG, "b2_dmrg_sweep: bad arguments");
   const int L = d->L;
   double emin = 1e300, dmax = 0.0;
   int rc;
   const double tw0 = wall_seconds();
   const double base[7] = {d->t_plan, d->t_join, d->t_solve, d->t_split, d->t_release, d->t_update, d->t_tail};
   // DMRG.cpp:360,391: the noise added before Split is |noise prefactor| x (largest discarded weight of the previous half sweep)
   noise = std::fabs(noise) * d->max_disc_last_sweep;
   if (!to_right) {
      for (int index = L - 2; index > 0; index--) {
         double e, dw;
         if ((rc = b2_dmrg_solve_site(d, index, rtol, noise, D, 0, change, &e, &dw, nullptr))) return rc;
         emin = std::min(emin, e); dmax = std::max(dmax, dw); d->last_energy = e;
         if ((rc = b2_dmrg_update(d, index + 1, 0))) return rc;
      }
   } else {
      for (int index = 0; index < L - 2; index++) {
         double e, dw;
         if ((rc = b2_dmrg_solve_site(d, index, rtol, noise, D, 1, change, &e, &dw, nullptr))) return rc;
         emin = std::min(emin, e); dmax = std::max(dmax, dw); d->last_energy = e;
         if ((rc = b2_dmrg_update(d, index, 1))) return rc;
      }
   }
   if (getenv("B2_TIMING"))
      fprintf(stderr, "b2_dmrg_sweep %s D=%d: wall %.3f s = plan %.3f + join %.3f + solve %.3f + split %.3f + release %.3f + update %.3f + update epilogue %.3f + rest\n",
              to_right ? "->" : "<-", D, wall_seconds() - tw0, d->t_plan - base[0], d->t_join - base[1], d->t_solve - base[2], d->t_split - base[3],
              d->t_release - base[4], d->t_update - base[5], d->t_tail - base[6]);
   d->max_disc_last_sweep = dmax;
   d->last_min_energy = emin;
   d->total_min_energy = std::min(d->total_min_energy, emin);
   *min_energy = emin;
   if (max_discarded) *max_discarded = dmax;
   return B2_OK;
}
int b2_dmrg_sweep_info(const b2_dmrg* d, double* out4) {
   if (!d || !out4) return fail(B2_ERR_ARG, "b2_dmrg_sweep_info: NULL");
   out4[0] = d->last_energy; out4[1] = d->last_min_energy; out4[2] = d->max_disc_last_sweep; out4[3] = d->total_min_energy;
   return B2_OK;
}

/* TwoDM::FillSite (TwoDM.cpp:445-628): contribution of one site to the spin-summed 2-RDM arrays A and B.  See b2_twodm.cpp. */
struct b2_twodm {
   b2_ctx* ctx = nullptr;
   b2_opset *left = nullptr, *right = nullptr;
   TwoDMPlan plan;
   CompiledWork build;    // pass 1: the effective operators
};

int b2_twodm_create(b2_ctx* ctx, int site, b2_opset* left, b2_opset* right, b2_twodm** out) {
   if (!ctx || !ctx->have_bk || !out) return fail(B2_ERR_ARG, "b2_twodm_create: bad arguments");
   const int L = ctx->bk.L;
   if (site < 0 || site >= L) return fail(B2_ERR_ARG, "b2_twodm_create: site %d out of range", site);
   if (site > 0 && (!left || left->set.boundary != site || !left->set.moving_right)) return fail(B2_ERR_ARG, "b2_twodm_create: left set must sit at boundary %d moving right", site);
   if (site < L - 1 && (!right || right->set.boundary != site + 1 || right->set.moving_right)) return fail(B2_ERR_ARG, "b2_twodm_create: right set must sit at boundary %d moving left", site + 1);
   std::unique_ptr<b2_twodm> p(new b2_twodm);
   p->ctx = ctx;
   p->left = site > 0 ? left : nullptr;
   p->right = site < L - 1 ? right : nullptr;
   build_twodm_plan(p->plan, ctx->bk, site, p->left ? &p->left->set : nullptr, p->right ? &p->right->set : nullptr);
   CompileOptions copt = budgeted(ctx);
   copt.threads = plan_threads((int)p->plan.dst.size());
   compile_terms(p->build, p->plan.terms, p->plan.dst, SP_VOUT, copt);
   *out = p.release();
   return B2_OK;
}
void b2_twodm_destroy(b2_twodm* p) { delete p; }
int b2_twodm_worklists(const b2_twodm* p, b2_worklists* o) {
   if (!p || !o) return fail(B2_ERR_ARG, "b2_twodm_worklists: NULL");
   fill_worklists(p->build, o);
   return B2_OK;
}
int64_t b2_twodm_m_size(const b2_twodm* p) { return p ? p->plan.m_size : 0; }
int b2_twodm_num_groups(const b2_twodm* p) { return p ? (int)p->plan.groups.size() : 0; }
int b2_twodm_group_info(const b2_twodm* p, int g, int* left_side, int64_t* off, int64_t* stride, int64_t* op_size, int* n_members, int* n_partners,
                        int* partners, int cap) {
   if (!p || g < 0 || g >= (int)p->plan.groups.size()) return fail(B2_ERR_ARG, "b2_twodm_group_info: bad arguments");
   const TwoDMPlan::Group& grp = p->plan.groups[g];
   if (left_side) *left_side = grp.left_side;
   if (off) *off = grp.off;
   if (stride) *stride = grp.stride;
   if (op_size) *op_size = grp.members.empty() ? 0 : p->plan.mops[grp.members[0]].lay->size;
   if (n_members) *n_members = (int)grp.members.size();
   if (n_partners) *n_partners = (int)grp.partners.size();
   if (partners) for (int i = 0; i < std::min<int>(cap, (int)grp.partners.size()); i++) partners[i] = grp.partners[i];
   return B2_OK;
}
int b2_twodm_d1_scale(const b2_twodm* p, double* per_block, int cap) {
   if (!p || !per_block) return fail(B2_ERR_ARG, "b2_twodm_d1_scale: NULL");
   for (int k = 0; k < std::min<int>(cap, (int)p->plan.d1_scale.size()); k++) per_block[k] = p->plan.d1_scale[k];
   return (int)p->plan.d1_scale.size();
}
int b2_twodm_scatter(const b2_twodm* p, const double* const* gram, double d1, double* two_rdm_A, double* two_rdm_B) {
   if (!p || !gram || !two_rdm_A || !two_rdm_B) return fail(B2_ERR_ARG, "b2_twodm_scatter: NULL");
   std::vector<std::vector<double>> g(p->plan.groups.size());
   for (size_t i = 0; i < g.size(); i++) {
      const size_t n = p->plan.groups[i].members.size() * p->plan.groups[i].partners.size();
      if (n) g[i].assign(gram[i], gram[i] + n);
   }
   twodm_scatter(p->plan, p->ctx->bk, p->left ? &p->left->set : nullptr, p->right ? &p->right->set : nullptr, d1, g, two_rdm_A, two_rdm_B);
   return B2_OK;
}

// executes a TwoDMPlan on the device: effective operators, diagram-1 weight sum, Gram matrices with the stored operators
static int twodm_execute(b2_ctx* ctx, const TwoDMPlan& plan, const CompiledWork& build, const double* t_host, b2_opset* left, b2_opset* right, double* d1_out,
                         std::vector<std::vector<double>>& gram) {
   if (ctx->device < 0) return fail(B2_ERR_NO_DEVICE, "2-RDM / correlations: planning-only context, no CUDA device (there is no CPU fallback)");
   if ((left && left->offloaded) || (right && right->offloaded)) return fail(B2_ERR_STATE, "2-RDM / correlations: operator set is offloaded (b2_opset_reload first)");
   CUDA_TRY(cudaSetDevice(ctx->device));
   cudaStream_t s = ctx->stream;
   const int64_t tsize = plan.T.size;
   struct Buf { double* p = nullptr; ~Buf() { cudaFree(p); } } dT, dTs, dM, dY, dG, dScal, dScale;
   struct IBuf { int64_t* p = nullptr; ~IBuf() { cudaFree(p); } } dOff;
   CUDA_TRY(cudaMalloc(&dT.p, sizeof(double) * (size_t)std::max<int64_t>(tsize, 1)));
   CUDA_TRY(cudaMalloc(&dTs.p, sizeof(double) * (size_t)std::max<int64_t>(tsize, 1)));
   CUDA_TRY(cudaMalloc(&dM.p, sizeof(double) * (size_t)std::max<int64_t>(plan.m_size, 1)));
   CUDA_TRY(cudaMemcpyAsync(dT.p, t_host, sizeof(double) * (size_t)tsize, cudaMemcpyHostToDevice, s));
   CUDA_TRY(cudaMemcpyAsync(dTs.p, dT.p, sizeof(double) * (size_t)tsize, cudaMemcpyDeviceToDevice, s));
   CUDA_TRY(cudaMemsetAsync(dM.p, 0, sizeof(double) * (size_t)std::max<int64_t>(plan.m_size, 1), s));
   // ---- effective operators
   {
      DevBases b;
      for (int i = 0; i < SP_COUNT; i++) b.p[i] = nullptr;
      b.p[SP_LEFT] = left ? left->dev : nullptr; b.p[SP_RIGHT] = dT.p; b.p[SP_VOUT] = dM.p;
      int rc = run_compiled_once(ctx, build, b);
      if (rc) return rc;
   }
   // ---- diagram 1: < T , (2SL+1)-scaled doubly-occupied blocks of T >
   double d1 = 0.0;
   {
      const int nk = plan.T.nkappa();
      std::vector<int64_t> off(nk + 1);
      for (int k = 0; k < nk; k++) off[k] = plan.T.blk[k].off;
      off[nk] = tsize;
      CUDA_TRY(cudaMalloc(&dOff.p, sizeof(int64_t) * (nk + 1)));
      CUDA_TRY(cudaMalloc(&dScale.p, sizeof(double) * std::max(nk, 1)));
      CUDA_TRY(cudaMalloc(&dScal.p, sizeof(double) * (kRedScratch + 8)));
      CUDA_TRY(cudaMemcpyAsync(dOff.p, off.data(), sizeof(int64_t) * (nk + 1), cudaMemcpyHostToDevice, s));
      CUDA_TRY(cudaMemcpyAsync(dScale.p, plan.d1_scale.data(), sizeof(double) * nk, cudaMemcpyHostToDevice, s));
      CUDA_TRY(cudaMemsetAsync(dScal.p, 0, sizeof(double) * (kRedScratch + 8), s));
      if (nk > 0 && dev_scale_blocks(dTs.p, dOff.p, dScale.p, nk, s)) return fail(B2_ERR_CUDA, "%s", dev_last_error());
      if (tsize > 0 && dev_multi_dot(dT.p, dTs.p, tsize, 1, tsize, dScal.p + kRedScratch, dScal.p, s)) return fail(B2_ERR_CUDA, "%s", dev_last_error());
      CUDA_TRY(cudaMemcpyAsync(&d1, dScal.p + kRedScratch, sizeof(double), cudaMemcpyDeviceToHost, s));
      CUDA_TRY(cudaStreamSynchronize(s));
   }
   // ---- Gram matrices  G[member, partner] = < M_member , stored operator >: the partners of a group are gathered into a dense
   // [stride x count] matrix, then one K-concatenated GEMM per group through the grouped contraction kernels
   gram.assign(plan.groups.size(), std::vector<double>());
   {
      std::vector<int64_t> yoff(plan.groups.size(), 0), goff(plan.groups.size(), 0);
      int64_t ytot = 0, gtot = 0;
      for (size_t gi = 0; gi < plan.groups.size(); gi++) {
         const TwoDMPlan::Group& grp = plan.groups[gi];
         yoff[gi] = ytot; goff[gi] = gtot;
         ytot += grp.stride * (int64_t)grp.partners.size();
         gtot += ((int64_t)grp.members.size() * (int64_t)grp.partners.size() + 15) / 16 * 16;
      }
      CUDA_TRY(cudaMalloc(&dY.p, sizeof(double) * (size_t)std::max<int64_t>(ytot, 1)));
      CUDA_TRY(cudaMalloc(&dG.p, sizeof(double) * (size_t)std::max<int64_t>(gtot, 1)));
      CUDA_TRY(cudaMemsetAsync(dY.p, 0, sizeof(double) * (size_t)std::max<int64_t>(ytot, 1), s));
      CUDA_TRY(cudaMemsetAsync(dG.p, 0, sizeof(double) * (size_t)std::max<int64_t>(gtot, 1), s));
      std::vector<Term3> terms;
      std::vector<DstBlock> dst;
      for (size_t gi = 0; gi < plan.groups.size(); gi++) {
         const TwoDMPlan::Group& grp = plan.groups[gi];
         const b2_opset* set = grp.left_side ? left : right;
         if (!set || grp.partners.empty() || grp.members.empty() || grp.stride == 0) continue;
         for (size_t c = 0; c < grp.partners.size(); c++) {
            const OpTensor& t = set->set.ops[grp.partners[c]];
            if (t.lay->size > 0)
               CUDA_TRY(cudaMemcpyAsync(dY.p + yoff[gi] + (int64_t)c * grp.stride, set->dev + t.off, sizeof(double) * (size_t)t.lay->size, cudaMemcpyDeviceToDevice, s));
         }
         Term3 x;
         x.dst = (int)dst.size(); x.f = 1.0;
         x.p.space = SP_LEFT; x.p.off = grp.off; x.p.rows = (int32_t)grp.stride; x.p.cols = (int32_t)grp.members.size(); x.p.trans = 1;
         x.q.space = SP_RIGHT; x.q.off = yoff[gi]; x.q.rows = (int32_t)grp.stride; x.q.cols = (int32_t)grp.partners.size(); x.q.trans = 0;
         terms.push_back(x);
         dst.push_back(DstBlock{goff[gi], (int32_t)grp.members.size(), (int32_t)grp.partners.size()});
      }
      if (!terms.empty()) {
         CompiledWork w;
         CompileOptions copt = budgeted(ctx);
         compile_terms(w, terms, dst, SP_VOUT, copt);
         DevBases b;
         for (int i = 0; i < SP_COUNT; i++) b.p[i] = nullptr;
         b.p[SP_LEFT] = dM.p; b.p[SP_RIGHT] = dY.p; b.p[SP_VOUT] = dG.p;
         int rc = run_compiled_once(ctx, w, b);
         if (rc) return rc;
      }
      std::vector<double> gh((size_t)std::max<int64_t>(gtot, 1));
      CUDA_TRY(cudaMemcpyAsync(gh.data(), dG.p, sizeof(double) * (size_t)gtot, cudaMemcpyDeviceToHost, s));
      CUDA_TRY(cudaStreamSynchronize(s));
      for (size_t gi = 0; gi < plan.groups.size(); gi++) {
         const size_t n = plan.groups[gi].members.size() * plan.groups[gi].partners.size();
         gram[gi].assign(gh.begin() + goff[gi], gh.begin() + goff[gi] + n);
      }
   }
   if (d1_out) *d1_out = d1;
   return B2_OK;
}

int b2_twodm_run(b2_twodm* tp, const double* t_host, double* two_rdm_A, double* two_rdm_B) {
   if (!tp || !t_host || !two_rdm_A || !two_rdm_B) return fail(B2_ERR_ARG, "b2_twodm_run: bad arguments");
   double d1 = 0.0;
   std::vector<std::vector<double>> gram;
   int rc = twodm_execute(tp->ctx, tp->plan, tp->build, t_host, tp->left, tp->right, &d1, gram);
   if (rc) return rc;
   twodm_scatter(tp->plan, tp->ctx->bk, tp->left ? &tp->left->set : nullptr, tp->right ? &tp->right->set : nullptr, d1, gram, two_rdm_A, two_rdm_B);
   return B2_OK;
}

/* Correlations::FillSite (Correlations.cpp:212-351) for site `site` (>= 1): T = MPS[site] (orthogonality centre), corr = correlation
 * operator set of boundary `site`; A, B = the finished 2-RDM arrays (TwoDM, after correct_higher_multiplicities).  Fills row/column
 * `site` of MutInfo and adds the two-orbital part to Cdirad, exactly like the reference. */
int b2_corr_fill_site(b2_ctx* ctx, int site, const double* t_host, b2_opset* corr, const double* A, const double* B, double* Cdirad, double* MutInfo) {
   if (!ctx || !ctx->have_bk || !t_host || !corr || !A || !B || !Cdirad || !MutInfo) return fail(B2_ERR_ARG, "b2_corr_fill_site: bad arguments");
   const int L = ctx->bk.L;
   if (site < 1 || site >= L || corr->set.boundary != site) return fail(B2_ERR_ARG, "b2_corr_fill_site: the correlation set must sit at boundary %d", site);
   TwoDMPlan plan;
   build_corr_plan(plan, ctx->bk, site, corr->set);
   CompiledWork build;
   CompileOptions copt = budgeted(ctx);
   compile_terms(build, plan.terms, plan.dst, SP_VOUT, copt);
   std::vector<std::vector<double>> gram;
   int rc = twodm_execute(ctx, plan, build, t_host, corr, nullptr, nullptr, gram);
   if (rc) return rc;
   const Problem& pr = ctx->prob;
   auto irr = [&](int o) { return ctx->bk.orb_irrep[o]; };
   auto getA = [&](int i, int j, int k, int l) { return (xorp(irr(i), irr(j)) == xorp(irr(k), irr(l))) ? A[i + L * (j + L * (k + L * (size_t)l))] : 0.0; };
   auto getB = [&](int i, int j, int k, int l) { return (xorp(irr(i), irr(j)) == xorp(irr(k), irr(l))) ? B[i + L * (j + L * (k + L * (size_t)l))] : 0.0; };
   auto rdm1 = [&](int i, int j) {   // TwoDM::get1RDM_DMRG (TwoDM.cpp:128-142)
      if (irr(i) != irr(j)) return 0.0;
      double v = 0.0;
      for (int o = 0; o < L; o++) v += getA(i, o, j, o);
      return v / (pr.N - 1.0);
   };
   auto entropy1 = [&](int i) {      // Correlations::SingleOrbitalEntropy_DMRG (Correlations.cpp:165-177)
      const double v4 = 0.5 * getA(i, i, i, i), v23 = 0.5 * (rdm1(i, i) - getA(i, i, i, i)), v1 = 1.0 - v4 - 2 * v23;
      double e = 0.0;
      if (v1 > 1e-100) e -= v1 * std::log(v1);
      if (v23 > 1e-100) e -= 2 * v23 * std::log(v23);
      if (v4 > 1e-100) e -= v4 * std::log(v4);
      return e;
   };
   const double ps = 1.0 / (pr.twoS + 1.0), s5 = std::sqrt(0.5);
   const OpSet& cs = corr->set;
   auto v = [&](int tag, int kind, int p) { return corr_value(plan, cs, gram, tag, kind, p); };
   for (int p = 0; p < site; p++) {
      const bool eq = irr(p) == irr(site);
      const double diag1 = v(CORR_D3, K_G, p) * ps * 0.5 * s5;
      const double diag2 = 0.125 * (getB(p, site, site, p) - getA(p, site, site, p));
      const double val1 = v(CORR_D1, K_Y, p) * ps, val2 = v(CORR_D2, K_Z, p) * ps, val3 = diag1 + diag2;
      const double val4 = v(CORR_D1, K_G, p) * ps * s5, val5 = v(CORR_D3, K_Y, p) * ps * 0.5, val6 = eq ? v(CORR_D4, K_K, p) * ps * 0.5 : 0.0;
      const double val7 = v(CORR_D2, K_G, p) * ps * s5, val8 = v(CORR_D3, K_Z, p) * ps * 0.5, val9 = eq ? v(CORR_D5, K_M, p) * ps * 0.5 : 0.0;
      const double alpha = v(CORR_D2, K_Y, p) * ps, gamma = v(CORR_D1, K_Z, p) * ps, beta = diag1 - diag2, lambda = 2 * diag2;
      const double delta = eq ? -v(CORR_D5, K_K, p) * ps * 0.5 : 0.0, epsilon = eq ? v(CORR_D4, K_M, p) * ps * 0.5 : 0.0;
      const double kappa = 0.5 * getA(p, p, site, site);
      double R[256] = {0.0}, ev[16], evec[256];
      auto at = [&](int r, int c) -> double& { return R[r + 16 * c]; };
      at(0, 0) = val1; at(15, 15) = val2; at(5, 5) = at(10, 10) = val3;
      at(1, 1) = at(3, 3) = val4; at(2, 2) = at(4, 4) = val5;
      at(1, 2) = at(2, 1) = at(3, 4) = at(4, 3) = val6;
      at(11, 11) = at(13, 13) = val7; at(12, 12) = at(14, 14) = val8;
      at(11, 12) = at(12, 11) = at(13, 14) = at(14, 13) = val9;
      at(6, 6) = alpha; at(7, 7) = at(8, 8) = beta; at(9, 9) = gamma;
      at(6, 7) = at(7, 6) = delta; at(6, 8) = at(8, 6) = -delta;
      at(7, 9) = at(9, 7) = epsilon; at(8, 9) = at(9, 8) = -epsilon;
      at(6, 9) = at(9, 6) = kappa; at(7, 8) = at(8, 7) = lambda;
      if (b2_small_symmetric_eig(16, R, ev, evec)) return fail(B2_ERR_STATE, "b2_corr_fill_site: eigenvalue problem failed");
      double ent = 0.0;
      for (int c = 0; c < 16; c++) if (ev[c] > 1e-100) ent -= ev[c] * std::log(ev[c]);
      const double mi = 0.5 * (entropy1(p) + entropy1(site) - ent);
      MutInfo[p + L * site] = MutInfo[site + L * p] = mi;
      Cdirad[p + L * site] += 2 * beta;
      Cdirad[site + L * p] += 2 * beta;
   }
   return B2_OK;
}

int b2_twodm_fill_site(b2_ctx* ctx, int site, const double* t_host, b2_opset* left, b2_opset* right, double* two_rdm_A, double* two_rdm_B) {
   b2_twodm* p = nullptr;
   int rc = b2_twodm_create(ctx, site, left, right, &p);
   if (!rc) rc = b2_twodm_run(p, t_host, two_rdm_A, two_rdm_B);
   b2_twodm_destroy(p);
   return rc;
}

/* thin SVDs of a batch of host matrices on the GPU (what Sobject::Split needs from dgesdd_, Sobject.cpp:412-419) */
int b2_svd_batch(b2_ctx* ctx, int count, const int* m, const int* n, const double* const* a, double* const* sv, double* const* u, double* const* vt) {
   if (!ctx || count < 0 || (count > 0 && (!m || !n || !a || !sv || !u || !vt))) return fail(B2_ERR_ARG, "b2_svd_batch: bad arguments");
   if (ctx->device < 0) return fail(B2_ERR_NO_DEVICE, "b2_svd_batch: planning-only context, no CUDA device (there is no CPU fallback)");
   CUDA_TRY(cudaSetDevice(ctx->device));
   std::vector<SvdJob> jobs(count);
   for (int i = 0; i < count; i++) {
      if (m[i] < 1 || n[i] < 1) return fail(B2_ERR_ARG, "b2_svd_batch: empty matrix %d", i);
      jobs[i].m = m[i]; jobs[i].n = n[i]; jobs[i].a = a[i]; jobs[i].s = sv[i]; jobs[i].u = u[i]; jobs[i].vt = vt[i];
   }
   char err[256] = "";
   if (dev_svd_batch(jobs, (void*)ctx->stream, err, (int)sizeof(err))) return fail(B2_ERR_CUDA, "%s", err);
   return B2_OK;
}

/* FP64 peak probe (roofline denominator): mode 1 = DMMA m8n8k4, mode 0 = DFMA */
int b2_probe_fp64(b2_ctx* ctx, int use_mma, double* tflops) {
   if (!ctx || ctx->device < 0) return fail(B2_ERR_NO_DEVICE, "b2_probe_fp64: no CUDA device");
   CUDA_TRY(cudaSetDevice(ctx->device));
   if (dev_probe_fp64(use_mma, tflops)) return fail(B2_ERR_CUDA, "%s", dev_last_error());
   return B2_OK;
}

}   // extern "C"
