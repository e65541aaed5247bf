// b2_capi_update.cpp — C ABI of the renormalized-operator update (DMRG::updateMovingRight / updateMovingLeft, DMRGoperators.cpp:243-907).
#include "b2_capi_internal.h"

#include <map>


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

void b2capi::fill_worklists(const CompiledWork& c, b2_worklists* o) {
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
   const double tb0 = wall_seconds();
   build_update_plan(u->plan, ctx->bk, ctx->prob, u->old_set ? &u->old_set->set : nullptr, new_set->set, index, mr);
   u->world = world; u->rank = rank;
   const double tb1 = wall_seconds();
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
   const double tb2 = wall_seconds();
   CompileOptions copt = budgeted(ctx);
   copt.threads = (u->plan.flops_ref < ctx->parallel_plan_flops) ? plan_threads((int)u->plan.dst.size()) : 1;
   const double tb3 = wall_seconds();
   compile_terms(u->pass[0], u->plan.terms, u->plan.dst, SP_VOUT, copt);
   const double tb4 = wall_seconds();
   compile_terms(u->pass[1], u->plan.mix_terms, u->plan.mix_dst, SP_PRESUM, copt);   // transposed copies for daxpy_transpose_tensorCD
   const double tb5 = wall_seconds();
   {  // The mixing pass is memory bound and every source tile is shared by up to O(L^2) destination operators (A(s1,s2) of all outside
      // pairs add the same S0(o,i) blocks with different integrals): launch the tiles that read the same sources next to each other, so that
      // the sources are served by L2 and HBM sees every destination tile once.  Key = (first source tile, tile position).
      CompiledWork& mixw = u->pass[1];
      u->mix_all_axpy = true;
      for (const GemmItem& g : mixw.items2) if (!(g.flags & IF_AXPY)) { u->mix_all_axpy = false; break; }
      if (u->mix_all_axpy)
         for (const Wave& w : mixw.waves)
            for (int c = 0; c < kNumTileClasses; c++)
               std::sort(mixw.tiles2[c].begin() + w.t2_begin[c], mixw.tiles2[c].begin() + w.t2_end[c], [&](const Tile& a, const Tile& b) {
                  const GemmItem &ia = mixw.items2[a.item_begin], &ib = mixw.items2[b.item_begin];
                  if (ia.xoff != ib.xoff) return ia.xoff < ib.xoff;
                  if (a.n0 != b.n0) return a.n0 < b.n0;
                  if (a.m0 != b.m0) return a.m0 < b.m0;
                  return a.coff < b.coff;
               });
   }
   {  // whole-operator mixing grouped by destination layout (all destinations of a group have the same size and element order)
      const UpdatePlan& pl = u->plan;
      const OpSet& ns = new_set->set;
      if (!pl.mix_temps.empty()) { u->temp_begin = pl.mix_temps.front().off; u->temp_size = pl.presum_size - u->temp_begin; }
      std::map<const OpLayout*, std::vector<const UpdatePlan::MixFlat*>> by_lay;
      std::vector<const OpLayout*> lay_order;
      for (const UpdatePlan::MixFlat& m : pl.mix_flat) {
         const OpLayout* lay = ns.ops[m.dst_op].lay.get();
         if (by_lay.find(lay) == by_lay.end()) lay_order.push_back(lay);
         by_lay[lay].push_back(&m);
      }
      for (const OpLayout* lay : lay_order) {
         const auto& list = by_lay[lay];
         b2_update::MixGroup g{lay->size, 0, 0, (int64_t)u->mix_dst_off.size(), (int64_t)u->mix_src_off.size(), (int64_t)u->mix_coef.size()};
         std::map<int, int> dcol;                          // destination operator -> column
         std::map<std::pair<int, int64_t>, int> srow;      // (space, offset) -> row
         for (const UpdatePlan::MixFlat* m : list) {
            if (dcol.find(m->dst_op) == dcol.end()) { const int c = (int)dcol.size(); dcol[m->dst_op] = c; u->mix_dst_off.push_back(ns.ops[m->dst_op].off); }
            const int space = m->temp >= 0 ? SP_PRESUM : SP_VOUT;
            const int64_t off = m->temp >= 0 ? pl.mix_temps[m->temp].off : ns.ops[m->src_op].off;
            if (srow.find({space, off}) == srow.end()) { const int r = (int)srow.size(); srow[{space, off}] = r; u->mix_src_off.push_back(off); u->mix_src_space.push_back((uint8_t)space); }
         }
         g.nd = (int)dcol.size(); g.ns = (int)srow.size();
         u->mix_coef.resize((size_t)g.coef_begin + (size_t)g.ns * g.nd, 0.0);
         for (const UpdatePlan::MixFlat* m : list) {
            const int space = m->temp >= 0 ? SP_PRESUM : SP_VOUT;
            const int64_t off = m->temp >= 0 ? pl.mix_temps[m->temp].off : ns.ops[m->src_op].off;
            u->mix_coef[(size_t)g.coef_begin + (size_t)srow[{space, off}] * g.nd + dcol[m->dst_op]] += m->coef;
         }
         u->mix_groups.push_back(g);
      }
   }
   for (int p = 0; p < 2; p++) u->list_bytes[p] = u->pass[p].bytes();
   if (getenv("B2_TIMING"))
      fprintf(stderr, "b2_update_create: enumerate %.3f s, owners %.3f s, schedule %.3f s (budget %.3f, contraction pass %.3f, transposed copies %.3f, mixing lists %.3f), %zu + %zu terms\n",
              tb1 - tb0, tb2 - tb1, wall_seconds() - tb2, tb3 - tb2, tb4 - tb3, tb5 - tb4, wall_seconds() - tb5, u->plan.terms.size(), u->plan.mix_terms.size());
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
      if ((rc = upload_vec(&u->d_mix_dst_off, u->mix_dst_off, s))) return rc;
      if ((rc = upload_vec(&u->d_mix_src_off, u->mix_src_off, s))) return rc;
      if ((rc = upload_vec(&u->d_mix_src_space, u->mix_src_space, s))) return rc;
      if ((rc = upload_vec(&u->d_mix_coef, u->mix_coef, s))) return rc;
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
void b2capi::update_park(b2_update* u) {
   cudaFree(u->d_presum); cudaFree(u->d_work); cudaFree(u->d_part); cudaFree(u->d_t);
   u->d_presum = u->d_work = u->d_part = u->d_t = nullptr;
   if (u->h_t) { cudaFreeHost(u->h_t); u->h_t = nullptr; }
   std::vector<Term3>().swap(u->plan.terms); std::vector<Term3>().swap(u->plan.mix_terms);
   for (int p = 0; p < 2; p++) {
      ListVec<GemmItem>().swap(u->pass[p].items1); ListVec<GemmItem>().swap(u->pass[p].items2);
      ListVec<ReduceJob>().swap(u->pass[p].reduces);
      for (int c = 0; c < kNumTileClasses; c++) { ListVec<Tile>().swap(u->pass[p].tiles1[c]); ListVec<Tile>().swap(u->pass[p].tiles2[c]); }
   }
   u->old_set = u->new_set = nullptr;
}
int b2capi::update_unpark(b2_update* u, b2_opset* old_set, b2_opset* new_set) {
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
   const bool timing = getenv("B2_TIMING") != nullptr;
   cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
   if (timing) { for (auto& e : ev) cudaEventCreate(&e); cudaEventRecord(ev[0], s); }
   for (int p = 0; p < 2; p++) {
      if (p == 1 && u->temp_size > 0 && dev_fill_zero(u->d_presum + u->temp_begin, u->temp_size, s)) return fail(B2_ERR_CUDA, "%s", dev_last_error());
      for (const Wave& w : u->pass[p].waves) {
         for (int c = 0; c < kNumTileClasses; c++)
            if (dev_launch_tiles(c, u->d_tiles1[p][c] + w.t1_begin[c], w.t1_end[c] - w.t1_begin[c], u->d_items1[p], b, s)) return fail(B2_ERR_CUDA, "%s", dev_last_error());
         for (int c = 0; c < kNumTileClasses; c++) {
            const int nt2 = w.t2_end[c] - w.t2_begin[c];
            const int rc2 = (p == 1 && u->mix_all_axpy) ? dev_launch_axpy_tiles(u->d_tiles2[p][c] + w.t2_begin[c], nt2, u->d_items2[p], b, s)
                                                        : dev_launch_tiles(c, u->d_tiles2[p][c] + w.t2_begin[c], nt2, u->d_items2[p], b, s);
            if (rc2) return fail(B2_ERR_CUDA, "%s", dev_last_error());
         }
         if (dev_launch_reduce(u->d_reduces[p] + w.red_begin, w.red_end - w.red_begin, b, s)) return fail(B2_ERR_CUDA, "%s", dev_last_error());
      }
      // every GPU computed the operators it was assigned; summing the (otherwise zero) arenas replicates all of them before
      // the mixing pass, which every GPU then runs in full (block axpys, replaces the MPI exchanges of DMRGoperators.cpp:449-533)
      if (p == 0 && u->world > 1 && u->allreduce(u->allreduce_user, u->new_set->dev, u->new_set->set.size, (void*)s)) return fail(B2_ERR_STATE, "b2_update_run: all-reduce callback failed");
      if (p == 1)   // A/B/C/D += integral-weighted two-operator tensors: one tall-skinny GEMM-like launch per layout
         for (const b2_update::MixGroup& g : u->mix_groups)
            if (dev_launch_mix_flat(u->d_mix_dst_off + g.dst_begin, g.nd, u->d_mix_src_off + g.src_begin, u->d_mix_src_space + g.src_begin, g.ns, u->d_mix_coef + g.coef_begin,
                                    g.size, b, s)) return fail(B2_ERR_CUDA, "%s", dev_last_error());
      if (timing) cudaEventRecord(ev[p + 1], s);
   }
   if (timing) {
      cudaEventSynchronize(ev[2]);
      float t0 = 0.f, t1 = 0.f;
      cudaEventElapsedTime(&t0, ev[0], ev[1]); cudaEventElapsedTime(&t1, ev[1], ev[2]);
      fprintf(stderr, "b2_update_run: contraction pass %.3f ms (%.3f TFLOP executed), mixing pass %.3f ms (%zu transposed block copies + %zu whole-operator axpys in %zu GEMM-like launches)\n", t0,
              u->pass[0].flops_exec / 1e12, t1, u->plan.mix_terms.size(), u->plan.mix_flat.size(), u->mix_groups.size());
      for (auto& e : ev) cudaEventDestroy(e);
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
   o[0] = (double)u->plan.terms.size(); o[1] = (double)(u->plan.mix_terms.size() + u->plan.mix_flat.size()); o[2] = (double)u->plan.presums.size(); o[3] = u->plan.flops_ref;
   o[4] = u->pass[0].flops_exec + u->pass[1].flops_exec; o[5] = (double)std::max(u->pass[0].work_size, u->pass[1].work_size);
   o[6] = (double)(u->pass[0].waves.size() + u->pass[1].waves.size()); o[7] = 2.0 + u->pass[0].launches() + u->pass[1].launches();
   return B2_OK;
}
int b2_update_worklists(const b2_update* u, int pass, b2_worklists* o) {
   if (!u || !o || pass < 0 || pass > 1) return fail(B2_ERR_ARG, "b2_update_worklists: bad arguments");
   fill_worklists(u->pass[pass], o);
   return B2_OK;
}
int64_t b2_update_num_mix_flat(const b2_update* u) { return u ? (int64_t)u->plan.mix_flat.size() : 0; }
int b2_update_export_mix_flat(const b2_update* u, b2_flat_presum* out) {
   if (!u || !out) return fail(B2_ERR_ARG, "b2_update_export_mix_flat: NULL");
   size_t n = 0;
   for (const UpdatePlan::MixFlat& m : u->plan.mix_flat) {
      const OpTensor& d = u->new_set->set.ops[m.dst_op];
      out[n].dst_off = d.off; out[n].size = d.lay->size; out[n].coef = m.coef;
      out[n].space = m.temp >= 0 ? SP_PRESUM : SP_VOUT;
      out[n].src_off = m.temp >= 0 ? u->plan.mix_temps[m.temp].off : u->new_set->set.ops[m.src_op].off;
      n++;
   }
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

