// b2_capi_twodm.cpp — C ABI of the 2-RDM and correlation contractions (TwoDM::FillSite, Correlations::FillSite), the batched SVD and the FP64 probe.
#include "b2_capi_internal.h"

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

/* Sobject::Join (Sobject.cpp:212-258) as its own entry point: the same terms the sweep driver builds inside b2_dmrg_solve_site */
struct b2_join {
   b2_ctx* ctx = nullptr;
   SLayout S;
   TLayout TL, TR;
   CompiledWork work;
};
int b2_join_create(b2_ctx* ctx, int site, b2_join** out) {
   if (!ctx || !ctx->have_bk || !out) return fail(B2_ERR_STATE, "b2_join_create: no bookkeeper");
   if (site < 0 || site > ctx->bk.L - 2) return fail(B2_ERR_ARG, "b2_join_create: site %d out of range", site);
   std::unique_ptr<b2_join> j(new b2_join);
   j->ctx = ctx;
   j->S.build(ctx->bk, site); j->TL.build(ctx->bk, site); j->TR.build(ctx->bk, site + 1);
   std::vector<Term3> terms; std::vector<DstBlock> dst;
   join_terms(terms, dst, ctx->bk, j->S, j->TL, j->TR);
   compile_terms(j->work, terms, dst, SP_VOUT, budgeted(ctx));
   *out = j.release();
   return B2_OK;
}
void b2_join_destroy(b2_join* j) { delete j; }
int b2_join_worklists(const b2_join* j, b2_worklists* o) {
   if (!j || !o) return fail(B2_ERR_ARG, "b2_join_worklists: NULL");
   fill_worklists(j->work, o);
   return B2_OK;
}
int b2_join_run(b2_join* j, const double* t_left, const double* t_right, double* s_out) {
   if (!j || !t_left || !t_right || !s_out) return fail(B2_ERR_ARG, "b2_join_run: NULL argument");
   b2_ctx* ctx = j->ctx;
   if (ctx->device < 0) return fail(B2_ERR_NO_DEVICE, "b2_join_run: planning-only context, no CUDA device (there is no CPU fallback)");
   CUDA_TRY(cudaSetDevice(ctx->device));
   cudaStream_t s = ctx->stream;
   struct Buf { double* p = nullptr; ~Buf() { cudaFree(p); } } dTl, dTr, dS;
   const size_t nl = (size_t)std::max<int64_t>(j->TL.size, 1), nr = (size_t)std::max<int64_t>(j->TR.size, 1), ns = (size_t)std::max<int64_t>(j->S.size, 1);
   CUDA_TRY(cudaMalloc(&dTl.p, sizeof(double) * nl));
   CUDA_TRY(cudaMalloc(&dTr.p, sizeof(double) * nr));
   CUDA_TRY(cudaMalloc(&dS.p, sizeof(double) * ns));
   CUDA_TRY(cudaMemcpyAsync(dTl.p, t_left, sizeof(double) * (size_t)j->TL.size, cudaMemcpyHostToDevice, s));
   CUDA_TRY(cudaMemcpyAsync(dTr.p, t_right, sizeof(double) * (size_t)j->TR.size, cudaMemcpyHostToDevice, s));
   CUDA_TRY(cudaMemsetAsync(dS.p, 0, sizeof(double) * ns, s));
   DevBases b;
   for (int i = 0; i < SP_COUNT; i++) b.p[i] = nullptr;
   b.p[SP_LEFT] = dTl.p; b.p[SP_RIGHT] = dTr.p; b.p[SP_VOUT] = dS.p;
   int rc = run_compiled_once(ctx, j->work, b);
   if (rc) return rc;
   CUDA_TRY(cudaMemcpyAsync(s_out, dS.p, sizeof(double) * (size_t)j->S.size, cudaMemcpyDeviceToHost, s));
   CUDA_TRY(cudaStreamSynchronize(s));
   return B2_OK;
}

/* Sobject::Split (Sobject.cpp:260-622) as its own entry point: the routine the sweep driver calls inside b2_dmrg_solve_site */
struct b2_split {
   std::vector<double> t[2];
};
int b2_sobject_split(b2_ctx* ctx, int site, const double* s_storage, int D, int moving_right, int change, b2_svd_fn svd, void* user,
                     b2_split** out, double* discarded_weight) {
   if (!ctx || !ctx->have_bk || !s_storage || !out) return fail(B2_ERR_ARG, "b2_sobject_split: bad arguments");
   if (site < 0 || site > ctx->bk.L - 2 || D < 1) return fail(B2_ERR_ARG, "b2_sobject_split: site %d or D %d out of range", site, D);
   if (!svd && ctx->device < 0) return fail(B2_ERR_NO_DEVICE, "b2_sobject_split: planning-only context and no caller SVD (there is no CPU fallback)");
   SLayout S;
   S.build(ctx->bk, site);
   char svd_err[256] = "";
   SvdBatchFn fn;
   if (svd) {
      fn = [&](std::vector<SvdJob>& jobs) {
         const int n = (int)jobs.size();
         std::vector<int> m(n), nn(n);
         std::vector<const double*> a(n);
         std::vector<double*> sv(n), u(n), vt(n);
         for (int i = 0; i < n; i++) { m[i] = jobs[i].m; nn[i] = jobs[i].n; a[i] = jobs[i].a; sv[i] = jobs[i].s; u[i] = jobs[i].u; vt[i] = jobs[i].vt; }
         const int rc = svd(user, n, m.data(), nn.data(), a.data(), sv.data(), u.data(), vt.data());
         if (rc) snprintf(svd_err, sizeof(svd_err), "the caller's SVD routine returned %d", rc);
         return rc;
      };
   } else {
      if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(B2_ERR_CUDA, "b2_sobject_split: cudaSetDevice failed");
      fn = [&](std::vector<SvdJob>& jobs) { return dev_svd_batch(jobs, (void*)ctx->stream, svd_err, (int)sizeof(svd_err)); };
   }
   std::unique_ptr<b2_split> r(new b2_split);
   const double dw = split_host(ctx->bk, site, S, s_storage, D, moving_right != 0, change != 0, r->t[0], r->t[1], fn);
   if (dw < 0.0) return fail(svd ? B2_ERR_ARG : B2_ERR_CUDA, "b2_sobject_split: %s", svd_err);
   if (discarded_weight) *discarded_weight = dw;
   *out = r.release();
   return B2_OK;
}
int64_t b2_split_size(const b2_split* r, int right) { return r ? (int64_t)r->t[right ? 1 : 0].size() : -1; }
int b2_split_get(const b2_split* r, int right, double* t_out) {
   if (!r || !t_out) return fail(B2_ERR_ARG, "b2_split_get: NULL");
   const std::vector<double>& t = r->t[right ? 1 : 0];
   std::memcpy(t_out, t.data(), sizeof(double) * t.size());
   return B2_OK;
}
void b2_split_destroy(b2_split* r) { delete r; }

/* FP64 peak probe (roofline denominator): mode 1 = DMMA m8n8k4, mode 0 = DFMA */
int b2_probe_fp64(b2_ctx* ctx, int use_mma, double* tflops) {
   if (!ctx || ctx->device < 0) return fail(B2_ERR_NO_DEVICE, "b2_probe_fp64: no CUDA device");
   CUDA_TRY(cudaSetDevice(ctx->device));
   if (dev_probe_fp64(use_mma, tflops)) return fail(B2_ERR_CUDA, "%s", dev_last_error());
   return B2_OK;
}

