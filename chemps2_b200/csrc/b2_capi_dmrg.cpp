// b2_capi_dmrg.cpp — C ABI of the sweep driver (DMRG::PreSolve / Solve / sweepleft / sweepright / solve_site, DMRG.cpp:257-452), excited
// states, MPS checkpoints, and the 2-RDM / correlation chains of DMRG::calc_rdms_and_correlations (DMRGtechnics.cpp:40-215).
#include "b2_capi_internal.h"

#include <thread>

int b2capi::run_compiled_once(b2_ctx* ctx, const CompiledWork& w, DevBases b) {
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
   // of a small/medium-D sweep) from every later visit.  Key = the exact dimension tables of the boundaries the plan reads + the sharding.
   struct PlanSlot { std::vector<int> key; b2_heff* h = nullptr; };
   std::vector<PlanSlot> plan_cache;
   struct UpdSlot { std::vector<int> key; b2_update* u = nullptr; };
   std::vector<UpdSlot> upd_cache;        // index = 2 * site + moving_right
   bool use_plan_cache = true;
   long long plan_hits = 0, plan_misses = 0;
   // Inside b2_dmrg_sweep the HOST half of the next site's sigma plan (enumeration + scheduling: no CUDA call, reads only the bookkeeper,
   // the integrals and operator-set layouts, all frozen until the next Split) is built on a helper thread while the calling thread plans
   // and runs the operator update that precedes it; solve_site takes the finished plan over and only does the device half.
   struct Prefetch {
      std::thread th;
      bool active = false;
      int site = -1, world = 1, rank = 0;
      b2_opset *left = nullptr, *right = nullptr;
      std::vector<int> key;                // the plan-cache key of the site: sharding + dimension tables of boundaries site and site + 2
      std::unique_ptr<b2_heff> h;
      bool ok = false;
   } prefetch;
   bool in_sweep = false;
   bool use_prefetch = true;               // B2_PLAN_PREFETCH=0 switches it off
   bool prefetch_multi = false;            // B2_PLAN_PREFETCH=2: also in sharded (multi-GPU) sweeps
   long long plan_prefetched = 0;
   bool right_canonical = false;           // b2_dmrg_calc_2rdm leaves the MPS right-canonical (centre on site 0): PreSolve must restore the gauge first
   double max_disc_last_sweep = 0.0;       // DMRG::MaxDiscWeightLastSweep (DMRG.cpp:360-362): scales the noise of the next half sweep
   double last_energy = 0.0;               // energy of the last site solved (what DMRG::sweepleft / sweepright return)
   double last_min_energy = 1e8;           // DMRG::LastMinEnergy: lowest energy of the last half sweep
   double total_min_energy = 1e8;          // DMRG::TotalMinEnergy: lowest energy since the last PreSolve
   bool spill = false;                     // keep only the operator sets of the site being optimised in HBM (b2_dmrg_set_spill)
   std::string spill_dir;                  // spill mode parks the sets in files of this directory (NVMe) instead of pinned host memory
   int world = 1, rank = 0;                // GPUs sharing the sweep: sigma terms and operator updates are sharded, the rest is replicated
   b2_allreduce_fn allreduce = nullptr;
   void* allreduce_user = nullptr;
   double t_solve = 0.0, t_update = 0.0, t_split = 0.0, t_plan = 0.0;   // wall-clock seconds spent per phase (b2_dmrg_timers)
   long long n_matvec = 0;
   double t_join = 0.0, t_release = 0.0, t_tail = 0.0;   // B2_TIMING diagnostics: Join + vector copies, releasing plans / buffers / stale sets, update epilogue
   // rand() of the reference's process: TensorT::random and Sobject::addNoise draw from ONE stream seeded by the caller's srand()
   // (b2_dmrg_random_mps = srand(seed) + the draws of DMRG::setupBookkeeperAndMPS; every noisy solve_site continues the stream)
   GlibcRand rng{1};
};

int b2_dmrg_create(b2_ctx* ctx, b2_dmrg** out) {
   if (!ctx || !ctx->have_bk || !out) return fail(B2_ERR_STATE, "b2_dmrg_create: no bookkeeper");
   if (ctx->device < 0) return fail(B2_ERR_NO_DEVICE, "b2_dmrg_create: planning-only context, no CUDA device (there is no CPU fallback)");
   std::unique_ptr<b2_dmrg> d(new b2_dmrg);
   d->ctx = ctx; d->L = ctx->bk.L;
   d->mps.resize(d->L);
   for (int s = 0; s < d->L; s++) { TLayout t; t.build(ctx->bk, s); d->mps[s].assign((size_t)t.size, 0.0); }
   d->left.assign(d->L + 1, nullptr); d->right.assign(d->L + 1, nullptr);
   if (const char* e = getenv("B2_PLAN_PREFETCH")) { d->use_prefetch = atoi(e) != 0; d->prefetch_multi = atoi(e) >= 2; }
   *out = d.release();
   return B2_OK;
}
static void dmrg_clear_plan_cache(b2_dmrg* d) {
   for (b2_dmrg::PlanSlot& p : d->plan_cache) { b2_heff_destroy(p.h); p.h = nullptr; p.key.clear(); }
   for (b2_dmrg::UpdSlot& p : d->upd_cache) { b2_update_destroy(p.u); p.u = nullptr; p.key.clear(); }
}
// key of the sigma plan of `site`: sharding + the exact dimension tables of the boundaries the plan reads.  The two-site object and both
// operator sets live on boundaries site and site + 2 only (Sobject.cpp:36-78 never reads the contracted boundary site + 1, which EVERY
// visit re-dimensions): keying on it too would turn most re-visits into misses
static std::vector<int> dmrg_plan_key(const b2_dmrg* d, int site) {
   std::vector<int> key;
   const Bookkeeper& bk = d->ctx->bk;
   key.push_back(d->world); key.push_back(d->rank);
   for (int b = site; b <= site + 2; b += 2) key.insert(key.end(), bk.cur[b].begin(), bk.cur[b].end());
   return key;
}
// waits for the helper thread and drops what it built
static void dmrg_prefetch_cancel(b2_dmrg* d) {
   b2_dmrg::Prefetch& pf = d->prefetch;
   if (!pf.active) return;
   if (pf.th.joinable()) pf.th.join();
   pf.h.reset();
   pf.active = false;
}
// starts the host half of the sigma plan of `site` (operator sets lset / rset, which must stay alive until the prefetch is taken or cancelled)
static void dmrg_prefetch_start(b2_dmrg* d, int site, b2_opset* lset, b2_opset* rset) {
   dmrg_prefetch_cancel(d);
   const int L = d->L;
   if (!d->in_sweep || !d->use_prefetch || d->spill || site < 0 || site > L - 2) return;
   // several GPUs: validated on one GPU only so far (the round's GPU budget ended before a multi-rank run), so it stays off unless asked for
   if (d->world > 1 && !d->prefetch_multi) return;
   if (site > 0 && (!lset || lset->set.reduced || lset->offloaded || lset->set.boundary != site || !lset->set.moving_right)) return;
   if (site < L - 2 && (!rset || rset->set.reduced || rset->offloaded || rset->set.boundary != site + 2 || rset->set.moving_right)) return;
   if (d->ctx->simulate_oom > 0) return;   // the test hook counts b2_heff_create calls
   std::vector<int> key = dmrg_plan_key(d, site);
   if (d->use_plan_cache && (int)d->plan_cache.size() == L && d->plan_cache[site].h && d->plan_cache[site].key == key) return;   // the parked plan will be re-used
   b2_dmrg::Prefetch& pf = d->prefetch;
   pf.site = site; pf.world = d->world; pf.rank = d->rank;
   pf.left = site > 0 ? lset : nullptr; pf.right = site < L - 2 ? rset : nullptr;
   pf.key.swap(key);
   pf.ok = false;
   pf.h.reset(new b2_heff);
   b2_heff* h = pf.h.get();
   h->ctx = d->ctx; h->world = d->world; h->rank = d->rank; h->left = pf.left; h->right = pf.right;
   const CompileOptions budget = budgeted(d->ctx);   // cudaMemGetInfo on the calling thread: the helper makes no CUDA call
   bool* ok = &pf.ok;
   try {
      pf.th = std::thread([h, site, budget, ok] {
         try { heff_build_host(h, site, budget); *ok = true; } catch (...) { *ok = false; }
      });
      pf.active = true;
   } catch (...) { pf.h.reset(); pf.active = false; }
}
// the prefetched plan (host half only) if it was built for exactly this site, these operator sets and these dimensions; else nullptr
static b2_heff* dmrg_prefetch_take(b2_dmrg* d, int site, b2_opset* lset, b2_opset* rset, const std::vector<int>& key) {
   b2_dmrg::Prefetch& pf = d->prefetch;
   if (!pf.active) return nullptr;
   if (pf.th.joinable()) pf.th.join();
   pf.active = false;
   std::unique_ptr<b2_heff> h(std::move(pf.h));
   if (!pf.ok || pf.site != site || pf.left != lset || pf.right != rset || pf.world != d->world || pf.rank != d->rank || pf.key != key) return nullptr;
   return h.release();
}
void b2_dmrg_destroy(b2_dmrg* d) {
   if (!d) return;
   dmrg_prefetch_cancel(d);
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
   d->right_canonical = false;   // the caller's tensors: assumed left-normalised on sites 0 .. L-3 like the reference's PreSolve assumes
   return B2_OK;
}
int b2_dmrg_get_mps(const b2_dmrg* d, int site, double* t) {
   if (!d || site < 0 || site >= d->L || !t) return fail(B2_ERR_ARG, "b2_dmrg_get_mps: bad arguments");
   std::memcpy(t, d->mps[site].data(), sizeof(double) * d->mps[site].size());
   return B2_OK;
}
int b2_dmrg_random_mps(b2_dmrg* d, uint64_t seed) {
   if (!d) return fail(B2_ERR_ARG, "b2_dmrg_random_mps: NULL");
   d->rng.reseed((unsigned int)seed);   // srand(seed) of the reference's caller (tests, Initialize::Init)
   for (int s = 0; s < d->L; s++) {   // DMRG::setupBookkeeperAndMPS (DMRG.cpp:149-169): random() then left_normalize with R discarded
      TLayout lay; lay.build(d->ctx->bk, s);
      d->mps[s].resize((size_t)lay.size);
      for (double& x : d->mps[s]) x = (2 * ((double)d->rng.next()) / GlibcRand::RANDMAX) - 1.0;   // TensorT.cpp:170, value in [-1, 1[
      left_normalize_host(d->ctx->bk, lay, d->mps[s].data());
   }
   d->right_canonical = false;
   return B2_OK;
}
int b2_dmrg_srand(b2_dmrg* d, uint64_t seed) {
   if (!d) return fail(B2_ERR_ARG, "b2_dmrg_srand: NULL");
   d->rng.reseed((unsigned int)seed);
   return B2_OK;
}
int b2_rand_stream(uint64_t seed, int n, int* out) {
   if (n < 0 || (n > 0 && !out)) return fail(B2_ERR_ARG, "b2_rand_stream: bad arguments");
   GlibcRand g((unsigned int)seed);
   for (int i = 0; i < n; i++) out[i] = g.next();
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
   set_plan_local_ranks(world);
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
   auto park = [&](b2_opset* set, int b, int mr) -> int {
      if (d->spill_dir.empty()) return b2_opset_offload(set);
      char name[64];
      std::snprintf(name, sizeof(name), "/b2_operators_%p_%d_%d.bin", (void*)d, b, mr);
      return b2_opset_offload_file(set, (d->spill_dir + name).c_str());
   };
   for (int b = 0; b <= d->L; b++) {
      if (b != keep_l && d->left[b] && (rc = park(d->left[b], b, 1))) return rc;
      if (b != keep_r && d->right[b] && (rc = park(d->right[b], b, 0))) return rc;
   }
   return B2_OK;
}
int b2_dmrg_set_spill_dir(b2_dmrg* d, const char* dir) {
   if (!d) return fail(B2_ERR_ARG, "b2_dmrg_set_spill_dir: NULL");
   d->spill_dir = dir ? dir : "";
   return B2_OK;
}
int b2_dmrg_set_plan_cache(b2_dmrg* d, int enabled) {
   if (!d) return fail(B2_ERR_ARG, "b2_dmrg_set_plan_cache: NULL");
   d->use_plan_cache = enabled != 0;
   if (!d->use_plan_cache) dmrg_clear_plan_cache(d);
   return B2_OK;
}
int b2_dmrg_set_plan_prefetch(b2_dmrg* d, int enabled) {
   if (!d) return fail(B2_ERR_ARG, "b2_dmrg_set_plan_prefetch: NULL");
   d->use_prefetch = enabled != 0;
   d->prefetch_multi = enabled >= 2;
   return B2_OK;
}
long long b2_dmrg_plan_prefetched(const b2_dmrg* d) { return d ? d->plan_prefetched : 0; }
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

// TensorO::update_ownmem / create (TensorO.cpp:38-196, formulas of TensorOperator::update with two_j = 0, no Jordan-Wigner phase) for every
// stored state: the overlap tensor of the boundary next to site `index` from MPS[index] of both states
// (DMRG::updateMovingRight / updateMovingLeft, DMRGoperators.cpp:556-567, 889-900).
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
   d->total_min_energy = 1e8;
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
      if (!rc && mode == 0 && attempt == 0 && d->in_sweep) {
         // the site that b2_dmrg_sweep solves next needs nothing but the LAYOUT of the fresh set (its arena is filled below): build the host
         // half of its sigma plan on a helper thread while this thread plans and runs the update.  Moving right: pair (index + 1, index + 2)
         // with the fresh set on its left; moving left: pair (index - 2, index - 1) with the fresh set on its right.  The first site of the
         // NEXT half sweep is left alone (the caller may do anything between two sweeps).
         const int next = mr ? index + 1 : index - 2;
         if (mr ? next <= d->L - 3 : next >= 1)
            dmrg_prefetch_start(d, next, mr ? fresh : d->left[next], mr ? (next < d->L - 2 ? d->right[next + 2] : nullptr) : fresh);
      }
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
      dmrg_prefetch_cancel(d);   // it reads the layout of `fresh`
      b2_update_destroy(u); u = nullptr;
      b2_opset_destroy(fresh); fresh = nullptr;
      d->spill = true;
      dmrg_clear_plan_cache(d);
      if ((rc = dmrg_residency(d, mr ? b_old : -1, mr ? -1 : b_old))) return rc;
   }
   if (rc) { dmrg_prefetch_cancel(d); b2_update_destroy(u); b2_opset_destroy(fresh); return rc; }
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
   if (rc) { dmrg_prefetch_cancel(d); b2_opset_destroy(fresh); return rc; }
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
   const std::vector<int> key = dmrg_plan_key(d, index);
   std::unique_ptr<b2_heff> pre(dmrg_prefetch_take(d, index, lset, rset, key));   // joins the helper thread (if any) before anything is changed
   if (d->use_plan_cache) {
      if ((int)d->plan_cache.size() != L) d->plan_cache.assign(L, b2_dmrg::PlanSlot());
      b2_dmrg::PlanSlot& slot = d->plan_cache[index];
      if (slot.h && slot.key == key) {
         h = slot.h; slot.h = nullptr;
         int ur = heff_unpark(h, lset, rset);
         if (ur) { b2_heff_destroy(h); h = nullptr; cudaGetLastError(); } else d->plan_hits++;
      }
   }
   int rc = B2_OK;
   if (!h && pre) {   // host half built during the preceding operator update: only the device half is left
      d->plan_misses++; d->plan_prefetched++;
      h = pre.release();
      rc = heff_setup_device(h);
   } else if (!h) { d->plan_misses++; rc = b2_heff_create(ctx, index, lset, rset, d->world, d->rank, &h); }
   pre.reset();
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
      if (noise > 0.0) for (double& x : s_host) x += (((double)d->rng.next()) / GlibcRand::RANDMAX - 0.5) * noise;   // Sobject::addNoise (Sobject.cpp:652-659)
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

// MPS checkpoint in the reference's schema (DMRG::saveMPS / loadDIM / loadMPS, DMRGmpsio.cpp:30-131): the objects
//    /Convergence/Converged_yn (int32)      /VirtDim_<boundary>_<N>_<2S>_<irrep>/Value (int32)      /MPS_<site>/Values (float64, TensorT::gStorage())
// with exactly the reference's names, types and contents.  This image has no HDF5 library, so the objects travel in the flat "B2H5v1"
// container that the HDF5 bridge of the reference builds (env_shims/hdf5.h) reads and writes too: a checkpoint written here is loaded by
// the unmodified reference (DMRG constructor with makechkpt = true finds CheMPS2_MPS0.h5) and vice versa.  A build with libhdf5 replaces
// the two container helpers below by H5Dwrite / H5Dread on the same object names.
namespace {
struct H5Object { std::string path; int32_t elem; std::vector<char> bytes; };
bool b2h5_write(const char* path, const std::vector<H5Object>& objs) {
   FILE* f = std::fopen(path, "wb");
   if (!f) return false;
   bool ok = std::fwrite("B2H5v1\0\0", 1, 8, f) == 8;
   for (const H5Object& o : objs) {
      const int32_t len = (int32_t)o.path.size();
      const int64_t nbytes = (int64_t)o.bytes.size();
      ok = ok && std::fwrite(&len, 4, 1, f) == 1 && std::fwrite(o.path.data(), 1, (size_t)len, f) == (size_t)len && std::fwrite(&o.elem, 4, 1, f) == 1 &&
           std::fwrite(&nbytes, 8, 1, f) == 1 && (nbytes == 0 || std::fwrite(o.bytes.data(), 1, (size_t)nbytes, f) == (size_t)nbytes);
   }
   std::fclose(f);
   return ok;
}
bool b2h5_read(const char* path, std::map<std::string, H5Object>& objs) {
   FILE* f = std::fopen(path, "rb");
   if (!f) return false;
   char magic[8];
   bool ok = std::fread(magic, 1, 8, f) == 8 && std::memcmp(magic, "B2H5v1\0\0", 8) == 0;
   while (ok) {
      int32_t len = 0;
      if (std::fread(&len, 4, 1, f) != 1) break;
      H5Object o;
      int64_t nbytes = 0;
      o.path.assign((size_t)std::max(len, 0), ' ');
      ok = len > 0 && len < 4096 && std::fread(&o.path[0], 1, (size_t)len, f) == (size_t)len && std::fread(&o.elem, 4, 1, f) == 1 && std::fread(&nbytes, 8, 1, f) == 1 && nbytes >= 0;
      if (!ok) break;
      o.bytes.resize((size_t)nbytes);
      ok = nbytes == 0 || std::fread(o.bytes.data(), 1, (size_t)nbytes, f) == (size_t)nbytes;
      if (ok) objs[o.path] = std::move(o);
   }
   std::fclose(f);
   return ok;
}
H5Object int_object(const std::string& path, int32_t v) { H5Object o; o.path = path; o.elem = 4; o.bytes.resize(4); std::memcpy(o.bytes.data(), &v, 4); return o; }
std::string virtdim_name(int b, int n, int ts, int ir) { char buf[96]; std::snprintf(buf, sizeof(buf), "/VirtDim_%d_%d_%d_%d/Value", b, n, ts, ir); return buf; }
}   // namespace

int b2_dmrg_save_mps(const b2_dmrg* d, const char* path, int converged) {
   if (!d || !path) return fail(B2_ERR_ARG, "b2_dmrg_save_mps: NULL");
   const Bookkeeper& bk = d->ctx->bk;
   std::vector<H5Object> objs;
   objs.push_back(int_object("/Convergence/Converged_yn", converged ? 1 : 0));
   for (int b = 0; b <= bk.L; b++)
      bk.for_sectors(b, [&](int n, int ts, int ir) { objs.push_back(int_object(virtdim_name(b, n, ts, ir), bk.dim(b, n, ts, ir))); });
   for (int sdx = 0; sdx < d->L; sdx++) {
      H5Object o;
      o.path = "/MPS_" + std::to_string(sdx) + "/Values"; o.elem = 8;
      o.bytes.resize(sizeof(double) * d->mps[sdx].size());
      if (!o.bytes.empty()) std::memcpy(o.bytes.data(), d->mps[sdx].data(), o.bytes.size());
      objs.push_back(std::move(o));
   }
   return b2h5_write(path, objs) ? B2_OK : fail(B2_ERR_STATE, "b2_dmrg_save_mps: write to %s failed", path);
}
int b2_dmrg_load_mps(b2_dmrg* d, const char* path, int* converged) {
   if (!d || !path) return fail(B2_ERR_ARG, "b2_dmrg_load_mps: NULL");
   std::map<std::string, H5Object> objs;
   if (!b2h5_read(path, objs)) return fail(B2_ERR_ARG, "b2_dmrg_load_mps: cannot read %s (not a B2H5v1 checkpoint container)", path);
   Bookkeeper& bk = d->ctx->bk;
   auto get_int = [&](const std::string& name, int32_t& v) { auto it = objs.find(name); if (it == objs.end() || it->second.bytes.size() != 4) return false; std::memcpy(&v, it->second.bytes.data(), 4); return true; };
   int32_t conv = 0;
   bool ok = get_int("/Convergence/Converged_yn", conv);
   // DMRG::loadDIM: every sector of the bookkeeper's enumeration must be present (a file of another problem lacks some / has others)
   for (int b = 0; b <= bk.L && ok; b++)
      bk.for_sectors(b, [&](int n, int ts, int ir) { int32_t v = 0; ok = ok && get_int(virtdim_name(b, n, ts, ir), v); if (ok) bk.set_dim(b, n, ts, ir, v); });
   if (!ok) return fail(B2_ERR_STATE, "b2_dmrg_load_mps: %s belongs to another problem or is incomplete (virtual dimensions)", path);
   for (int sdx = 0; sdx < d->L && ok; sdx++) {   // DMRG::loadMPS
      TLayout lay;
      lay.build(bk, sdx);
      auto it = objs.find("/MPS_" + std::to_string(sdx) + "/Values");
      ok = it != objs.end() && it->second.bytes.size() == sizeof(double) * (size_t)lay.size;
      if (ok) { d->mps[sdx].resize((size_t)lay.size); if (lay.size) std::memcpy(d->mps[sdx].data(), it->second.bytes.data(), it->second.bytes.size()); }
   }
   if (!ok) return fail(B2_ERR_STATE, "b2_dmrg_load_mps: %s is truncated or inconsistent with the bookkeeper", path);
   if (converged) *converged = conv;
   for (int b = 0; b <= d->L; b++) {   // the operators of the previous MPS are stale
      if (d->left[b]) { b2_opset_destroy(d->left[b]); d->left[b] = nullptr; }
      if (d->right[b]) { b2_opset_destroy(d->right[b]); d->right[b] = nullptr; }
   }
   for (ExcState& x : d->exc) { x.left.assign(d->L + 1, Overlap()); x.right.assign(d->L + 1, Overlap()); }
   dmrg_clear_plan_cache(d);
   d->right_canonical = false;
   d->total_min_energy = 1e8;   // like a freshly constructed DMRG object that found a checkpoint: the next Solve starts with one fixed-dimension sweep
   return B2_OK;
}

static int dmrg_gauge_move(b2_dmrg* d, int site, bool to_left);

// DMRG::PreSolve (DMRG.cpp:257-266): the moving-right operators of every boundary from the current MPS
int b2_dmrg_presolve(b2_dmrg* d) {
   if (!d) return fail(B2_ERR_ARG, "b2_dmrg_presolve: NULL");
   dmrg_clear_plan_cache(d);   // the plans of earlier visits carry the integrals of that time (b2_problem_update_mx)
   if (d->right_canonical) {   // after b2_dmrg_calc_2rdm: the moving-right operators need left-normalised tensors on sites 0 .. L-3
      for (int s = 0; s < d->L - 2; s++) { int rc = dmrg_gauge_move(d, s, false); if (rc) return rc; }
      d->right_canonical = false;
   }
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
   // DMRG.cpp:270: the first left sweep after a PreSolve (TotalMinEnergy still 1e8) keeps the virtual dimensions fixed
   bool change = d->total_min_energy < 1e8;
   double energy = 0.0;
   for (int ins = 0; ins < n_instructions; ins++) {
      int it = 0;
      double prev = energy + 10 * energy_conv[ins];   // at least one left-right sweep per instruction (DMRG.cpp:283)
      while (std::fabs(energy - prev) > energy_conv[ins] && it < max_sweeps[ins]) {
         prev = energy;
         double el, er, dw;
         if ((rc = b2_dmrg_sweep(d, 0, davidson_rtol[ins], noise_prefactor[ins], D[ins], change ? 1 : 0, &el, &dw))) return rc;
         change = true;
         if ((rc = b2_dmrg_sweep(d, 1, davidson_rtol[ins], noise_prefactor[ins], D[ins], 1, &er, &dw))) return rc;
         energy = d->last_energy;         // the convergence test compares what sweepright returns: the energy of its last site
         it++;
      }
   }
   *energy_out = d->total_min_energy;   // DMRG.cpp:353: lowest energy since the last PreSolve
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
   d->right_canonical = true;
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
   d->right_canonical = false;   // sites 0 .. L-2 are left-normalised again, the centre sits on the last site
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
   struct SweepScope {   // plan prefetching is confined to this call: whatever way it ends, no helper thread outlives it
      b2_dmrg* d;
      explicit SweepScope(b2_dmrg* d_) : d(d_) { d->in_sweep = true; }
      ~SweepScope() { dmrg_prefetch_cancel(d); d->in_sweep = false; }
   } scope(d);
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

