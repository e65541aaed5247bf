// b2_capi.cpp — the C ABI declared in include/chemps2_b200.h: errors, context, problem, bookkeeper, operator sets, sigma plans.
// (operator updates: b2_capi_update.cpp; sweep driver: b2_capi_dmrg.cpp; 2-RDM / correlations: b2_capi_twodm.cpp)
#include "b2_capi_internal.h"

static thread_local std::string g_err;
int b2capi::fail(int code, const char* fmt, ...) {
   char buf[512];
   va_list ap;
   va_start(ap, fmt);
   vsnprintf(buf, sizeof(buf), fmt, ap);
   va_end(ap);
   g_err = buf;
   return code;
}

CompileOptions b2capi::budgeted(const b2_ctx* ctx) {
   CompileOptions o = ctx->copt;
   if (ctx->device >= 0) {
      size_t free_b = 0, total_b = 0;
      if (cudaSetDevice(ctx->device) == cudaSuccess && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
         free_b += pool_cached_bytes();   // blocks parked in the library's own cache are released on demand
         o.work_budget = std::max<int64_t>((int64_t)1 << 22, std::min<int64_t>(o.work_budget, (int64_t)(free_b / 2 / sizeof(double))));
      }
   }
   return o;
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
      pool_register_stream(c->stream);
   }
   *out = c.release();
   return B2_OK;
}
void b2_ctx_destroy(b2_ctx* ctx) {
   if (!ctx) return;
   if (ctx->stream) pool_unregister_stream(ctx->stream);
   if (ctx->stream && ctx->own_stream) cudaStreamDestroy(ctx->stream);
   delete ctx;
}
int b2_ctx_device(const b2_ctx* ctx) { return ctx ? ctx->device : -1; }
int b2_ctx_set_stream(b2_ctx* ctx, void* cuda_stream) {
   if (!ctx || ctx->device < 0) return fail(B2_ERR_NO_DEVICE, "b2_ctx_set_stream: no CUDA device");
   if (ctx->stream) pool_unregister_stream(ctx->stream);
   if (ctx->stream && ctx->own_stream) cudaStreamDestroy(ctx->stream);
   ctx->stream = (cudaStream_t)cuda_stream; ctx->own_stream = false;
   CUDA_TRY(cudaSetDevice(ctx->device));
   pool_register_stream(ctx->stream);
   return B2_OK;
}
void* b2_ctx_stream(const b2_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

static int set_problem_common(b2_ctx* ctx, int L, int group, int N, int twoS, int irrep, const int* orb_irrep, double econst) {
   if (!ctx || L < 2 || !orb_irrep) return fail(B2_ERR_ARG, "b2_problem_set: bad arguments");
   const int nirr = num_irreps_of_group(group);
   if (nirr < 0) return fail(B2_ERR_ARG, "b2_problem_set: group %d out of range", group);
   if (irrep < 0 || irrep >= nirr) return fail(B2_ERR_ARG, "b2_problem_set: target irrep %d out of range", irrep);
   if (N < 2) return fail(B2_ERR_ARG, "b2_problem_set: N must be >= 2 (one-body part is folded in with 1/(N-1))");
   // Problem::checkConsistency (Problem.cpp:386-430)
   if (twoS < 0) return fail(B2_ERR_ARG, "b2_problem_set: TwoS = %d", twoS);
   if (N > 2 * L) return fail(B2_ERR_ARG, "b2_problem_set: N > 2*L ; N = %d and L = %d", N, L);
   if ((N % 2) != (twoS % 2)) return fail(B2_ERR_ARG, "b2_problem_set: N %% 2 != TwoS %% 2 ; N = %d and TwoS = %d", N, twoS);
   if (twoS > L - std::abs(N - L)) return fail(B2_ERR_ARG, "b2_problem_set: TwoS > L - |N-L| ; N = %d and TwoS = %d and L = %d", N, twoS, L);
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

int b2_problem_update_mx(b2_ctx* ctx, const double* mx_elem) {
   if (!ctx || !ctx->have_problem || !mx_elem) return fail(B2_ERR_STATE, "b2_problem_update_mx: no problem set");
   const size_t L = (size_t)ctx->prob.L;
   ctx->prob.mx.assign(mx_elem, mx_elem + L * L * L * L);
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
   if (dim < 0) return fail(B2_ERR_ARG, "b2_bk_set_dim: negative dimension");
   ctx->bk.set_dim(boundary, N, twoS, irrep, dim);     // sectors outside the table or with FCI dimension 0 are ignored (SyBookkeeper.cpp:163-169)
   return B2_OK;
}
int b2_bk_dim(const b2_ctx* ctx, int b, int N, int twoS, int irrep) { return (ctx && ctx->have_bk) ? ctx->bk.dim(b, N, twoS, irrep) : 0; }
int b2_bk_fcidim(const b2_ctx* ctx, int b, int N, int twoS, int irrep) { return (ctx && ctx->have_bk) ? ctx->bk.fcidim(b, N, twoS, irrep) : 0; }
// range queries follow the reference's sentinels instead of crashing: nothing set / boundary or N outside the table -> an empty range
static bool bk_boundary_ok(const b2_ctx* ctx, int b) { return ctx && ctx->have_bk && b >= 0 && b <= ctx->bk.L; }
static bool bk_n_ok(const b2_ctx* ctx, int b, int N) { return bk_boundary_ok(ctx, b) && N >= ctx->bk.Nmin[b] && N <= ctx->bk.Nmax[b]; }
int b2_bk_nmin(const b2_ctx* ctx, int b) { return bk_boundary_ok(ctx, b) ? ctx->bk.Nmin[b] : 0; }
int b2_bk_nmax(const b2_ctx* ctx, int b) { return bk_boundary_ok(ctx, b) ? ctx->bk.Nmax[b] : -1; }
int b2_bk_twosmin(const b2_ctx* ctx, int b, int N) { return bk_n_ok(ctx, b, N) ? ctx->bk.tsmin[b][N - ctx->bk.Nmin[b]] : 0; }
int b2_bk_twosmax(const b2_ctx* ctx, int b, int N) { return bk_n_ok(ctx, b, N) ? ctx->bk.tsmax[b][N - ctx->bk.Nmin[b]] : -1; }

static bool site_ok(const b2_ctx* ctx, int site, int last) { return ctx && ctx->have_bk && site >= 0 && site <= last; }
int64_t b2_tensor_t_size(const b2_ctx* ctx, int site) {
   if (!site_ok(ctx, site, ctx ? ctx->bk.L - 1 : 0)) { fail(B2_ERR_ARG, "b2_tensor_t_size: no bookkeeper or site %d out of range", site); return -1; }
   TLayout t; t.build(ctx->bk, site); return t.size;
}
int64_t b2_sobject_size(const b2_ctx* ctx, int site) {
   if (!site_ok(ctx, site, ctx ? ctx->bk.L - 2 : 0)) { fail(B2_ERR_ARG, "b2_sobject_size: no bookkeeper or site %d out of range", site); return -1; }
   SLayout s; s.build(ctx->bk, site); return s.size;
}
int b2_sobject_nkappa(const b2_ctx* ctx, int site) {
   if (!site_ok(ctx, site, ctx ? ctx->bk.L - 2 : 0)) { fail(B2_ERR_ARG, "b2_sobject_nkappa: no bookkeeper or site %d out of range", site); return -1; }
   SLayout s; s.build(ctx->bk, site); return s.nkappa();
}
int b2_sobject_table(const b2_ctx* ctx, int site, int* labels, int64_t* offsets) {
   if (!site_ok(ctx, site, ctx ? ctx->bk.L - 2 : 0) || !labels || !offsets) return fail(B2_ERR_ARG, "b2_sobject_table: no bookkeeper, NULL output or site %d out of range", site);
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
int b2capi::opset_create_reduced(b2_ctx* ctx, int boundary, bool mr, bool only_L, b2_opset** out) {
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
/* Second tier: the arena goes to a FILE (NVMe scratch; the reference's OperatorsOnDisk writes CheMPS2_Operators_*.h5 the same way,
 * DMRGoperators.cpp:1213-1433) and neither HBM nor host memory is held while the set is parked.  The copy runs through a bounded pinned
 * staging buffer (64 MiB pieces), so parking a set never needs a host allocation of its size. */
int b2_opset_offload_file(b2_opset* set, const char* path) {
   if (!set || !path) return fail(B2_ERR_ARG, "b2_opset_offload_file: NULL");
   if (set->ctx->device < 0) return fail(B2_ERR_NO_DEVICE, "b2_opset_offload_file: planning-only context, no CUDA device");
   if (!set->spill_file.empty()) return B2_OK;
   if (set->offloaded) { int rc = b2_opset_reload(set); if (rc) return rc; }
   if (!set->dev) return B2_OK;
   FILE* f = std::fopen(path, "wb");
   if (!f) return fail(B2_ERR_ARG, "b2_opset_offload_file: cannot open %s", path);
   const size_t total = (size_t)set->set.size, piece = (size_t)8 << 20;   // doubles per piece
   double* stage = nullptr;
   if (cudaMallocHost(&stage, sizeof(double) * std::min(total, piece)) != cudaSuccess) { std::fclose(f); return fail(B2_ERR_CUDA, "b2_opset_offload_file: pinned staging allocation failed"); }
   bool ok = true;
   for (size_t o = 0; o < total && ok; o += piece) {
      const size_t n = std::min(piece, total - o);
      ok = cudaMemcpyAsync(stage, set->dev + o, sizeof(double) * n, cudaMemcpyDeviceToHost, set->ctx->stream) == cudaSuccess &&
           cudaStreamSynchronize(set->ctx->stream) == cudaSuccess && std::fwrite(stage, sizeof(double), n, f) == n;
   }
   cudaFreeHost(stage);
   ok = (std::fclose(f) == 0) && ok;
   if (!ok) { std::remove(path); return fail(B2_ERR_STATE, "b2_opset_offload_file: writing %s failed", path); }
   CUDA_TRY(cudaFree(set->dev));
   set->dev = nullptr; set->offloaded = true; set->spill_file = path;
   return B2_OK;
}
static int reload_from_file(b2_opset* set) {
   FILE* f = std::fopen(set->spill_file.c_str(), "rb");
   if (!f) return fail(B2_ERR_STATE, "b2_opset_reload: cannot open %s", set->spill_file.c_str());
   const size_t total = (size_t)set->set.size, piece = (size_t)8 << 20;
   CUDA_TRY(cudaSetDevice(set->ctx->device));
   if (cudaMalloc(&set->dev, sizeof(double) * total) != cudaSuccess) { std::fclose(f); set->dev = nullptr; return fail(B2_ERR_CUDA, "b2_opset_reload: device allocation failed"); }
   double* stage = nullptr;
   if (cudaMallocHost(&stage, sizeof(double) * std::min(total, piece)) != cudaSuccess) { std::fclose(f); return fail(B2_ERR_CUDA, "b2_opset_reload: pinned staging allocation failed"); }
   bool ok = true;
   for (size_t o = 0; o < total && ok; o += piece) {
      const size_t n = std::min(piece, total - o);
      ok = std::fread(stage, sizeof(double), n, f) == n && cudaMemcpyAsync(set->dev + o, stage, sizeof(double) * n, cudaMemcpyHostToDevice, set->ctx->stream) == cudaSuccess &&
           cudaStreamSynchronize(set->ctx->stream) == cudaSuccess;
   }
   cudaFreeHost(stage);
   std::fclose(f);
   if (!ok) return fail(B2_ERR_STATE, "b2_opset_reload: reading %s failed", set->spill_file.c_str());
   std::remove(set->spill_file.c_str());
   set->spill_file.clear(); set->offloaded = false;
   return B2_OK;
}
int b2_opset_reload(b2_opset* set) {
   if (!set) return fail(B2_ERR_ARG, "b2_opset_reload: NULL");
   if (!set->offloaded) return B2_OK;
   if (!set->spill_file.empty()) return reload_from_file(set);
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
   if (set->offloaded && !set->spill_file.empty()) { int rr = b2_opset_reload(set); if (rr) return rr; }
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
   if (set->offloaded && !set->spill_file.empty()) { int rr = b2_opset_reload(set); if (rr) return rr; }   // parked in a file: bring it back first
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
   if (world > 1) set_plan_local_ranks(world);   // the ranks of one box build their plans concurrently
   h->left = (site > 0) ? left : nullptr;
   h->right = (site < L - 2) ? right : nullptr;
   heff_build_host(h.get(), site, budgeted(ctx));
   int rc = heff_setup_device(h.get());
   if (rc) return rc;
   *out = h.release();
   return B2_OK;
}

// Host half of b2_heff_create: term enumeration + scheduling.  Reads the bookkeeper, the integrals and the LAYOUTS of the two
// operator sets, never their contents, and makes no CUDA call — the sweep driver runs it on a helper thread for the next site while
// the operator update of the current one is planned and executed (b2_capi_dmrg.cpp, dmrg_prefetch_start).
void b2capi::heff_build_host(b2_heff* h, int site, const CompileOptions& budget) {
   b2_ctx* ctx = h->ctx;
   const double tb0 = wall_seconds();
   build_sigma_plan(h->plan, ctx->bk, ctx->prob, h->left ? &h->left->set : nullptr, h->right ? &h->right->set : nullptr, site, h->world);
   const double tb1 = wall_seconds();
   CompileOptions copt = budget;
   // plans whose sigma build is a few milliseconds are dominated by the time to BUILD them: compile those on all host cores
   copt.threads = (h->plan.flops_ref < ctx->parallel_plan_flops) ? plan_threads(h->plan.S.nkappa()) : 1;
   compile_sigma(h->comp, h->plan, h->left ? &h->left->set : nullptr, h->right ? &h->right->set : nullptr, h->rank, h->world, copt);
   h->list_bytes = h->comp.bytes();
   if (getenv("B2_TIMING")) fprintf(stderr, "b2_heff_create: enumerate %.3f s, schedule %.3f s, %zu terms\n", tb1 - tb0, wall_seconds() - tb1, h->plan.terms.size());
}

// Device half of b2_heff_create: work lists to the device, workspaces, pre-summed operators (reads the operator contents)
int b2capi::heff_setup_device(b2_heff* h) {
   b2_ctx* ctx = h->ctx;
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
      CUDA_TRY(cudaEventCreate(&h->ev0));
      CUDA_TRY(cudaEventCreate(&h->ev1));
      // materialise the integral-weighted operator pre-sums once (operators are fixed during the Davidson solve)
      DevBases b = bases_of(h, nullptr, nullptr);
      if (dev_launch_presum(h->d_jobs, (int)h->comp.presum_jobs.size(), h->d_parts, b, s)) return fail(B2_ERR_CUDA, "%s", dev_last_error());
      CUDA_TRY(cudaStreamSynchronize(s));
      if (getenv("B2_TIMING"))
         fprintf(stderr, "b2_heff_create: device setup %.3f s (work lists %.1f MB uploaded, workspace %.2f GB, pre-sums %.1f MB)\n", wall_seconds() - tb2,
                 h->list_bytes / 1e6, h->comp.work_size * 8e-9, h->plan.presum_size * 8e-6);
   }
   return B2_OK;
}

void b2_heff_destroy(b2_heff* h) { delete h; }

// a plan parked in the sweep driver's cache keeps only its device work lists: workspaces, vectors and host copies of the lists go
void b2capi::heff_park(b2_heff* h) {
   cudaFree(h->d_work); cudaFree(h->d_part); cudaFree(h->d_vin); cudaFree(h->d_vout); cudaFree(h->d_presum);
   cudaFree(h->d_exc); cudaFree(h->d_exc_coef); cudaFree(h->d_exc_scratch);
   h->d_work = h->d_part = h->d_vin = h->d_vout = h->d_presum = h->d_exc = h->d_exc_coef = h->d_exc_scratch = nullptr;
   h->n_exc = 0;
   if (h->h_vin) { cudaFreeHost(h->h_vin); h->h_vin = nullptr; }
   if (h->h_vout) { cudaFreeHost(h->h_vout); h->h_vout = nullptr; }
   BigVec<SigmaTerm>().swap(h->plan.terms);
   ListVec<GemmItem>().swap(h->comp.items1); ListVec<GemmItem>().swap(h->comp.items2);
   ListVec<ReduceJob>().swap(h->comp.reduces);
   for (int c = 0; c < kNumTileClasses; c++) { ListVec<Tile>().swap(h->comp.tiles1[c]); ListVec<Tile>().swap(h->comp.tiles2[c]); }
   h->left = h->right = nullptr;
}
// brings a parked plan back: new operator sets (same layouts: same dimensions), workspaces, pre-summed operators of the new contents
int b2capi::heff_unpark(b2_heff* h, b2_opset* left, b2_opset* right) {
   b2_ctx* ctx = h->ctx;
   cudaStream_t s = ctx->stream;
   h->left = left; h->right = right;
   const size_t n = (size_t)h->plan.S.size;
   if (h->comp.part_size > 0) CUDA_TRY(cudaMalloc(&h->d_part, sizeof(double) * (size_t)h->comp.part_size));
   if (h->plan.presum_size > 0) CUDA_TRY(cudaMalloc(&h->d_presum, sizeof(double) * (size_t)h->plan.presum_size));
   if (h->comp.work_size > 0) CUDA_TRY(cudaMalloc(&h->d_work, sizeof(double) * (size_t)h->comp.work_size));
   CUDA_TRY(cudaMalloc(&h->d_vin, sizeof(double) * (n ? n : 1)));
   CUDA_TRY(cudaMalloc(&h->d_vout, sizeof(double) * (n ? n : 1)));
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

// Pinned staging vectors of the host-buffer entry points (b2_heff_apply / _diag / _solve / _set_excitations), allocated on their first
// use: the sweep driver works on device vectors only and does not pay for pinning 2 x veclength doubles at every site.
static int ensure_staging(b2_heff* h) {
   const size_t n = (size_t)h->plan.S.size;
   if (!h->h_vin) CUDA_TRY(cudaMallocHost(&h->h_vin, sizeof(double) * (n ? n : 1)));
   if (!h->h_vout) CUDA_TRY(cudaMallocHost(&h->h_vout, sizeof(double) * (n ? n : 1)));
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
   { int rs = ensure_staging(h); if (rs) return rs; }
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
   { int rs = ensure_staging(h); if (rs) return rs; }
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
   int rc = ensure_staging(h);
   if (!rc) rc = b2_heff_diag_device(h, h->d_vout);
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
      prm.max_matvec = h->ctx->davidson_max_matvec;
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
   { int rs = ensure_staging(h); if (rs) return rs; }
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

int b2_heff_diag_lists(const b2_heff* h, const void** items, int64_t* n_items, const void** tiles, int64_t* n_tiles) {
   if (!h || !items || !n_items || !tiles || !n_tiles) return fail(B2_ERR_ARG, "b2_heff_diag_lists: NULL");
   *items = h->comp.diag_items.data(); *n_items = (int64_t)h->comp.diag_items.size();
   *tiles = h->comp.diag_tiles.data(); *n_tiles = (int64_t)h->comp.diag_tiles.size();
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
   else if (!std::strcmp(name, "davidson_max_matvec")) { if (value < 1) return fail(B2_ERR_ARG, "davidson_max_matvec too small"); ctx->davidson_max_matvec = (int)value; }
   else return fail(B2_ERR_ARG, "b2_ctx_set_option: unknown option %s", name);
   return B2_OK;
}

