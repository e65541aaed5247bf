// b2_heff.cpp — SigmaPlan -> device work lists.  Host only.
//
// Every term  sigma[dst] += f * op(A) * S[src] * op(B)  becomes a three-factor contraction for the generic scheduler
// (b2_compile.cpp): P = left operator block, Q = S[src], R = right operator block.  The terms that map a block onto
// itself also define the diagonal of H_eff.
#include "b2_heff.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <new>
#include <thread>

namespace b2 {

namespace {

MatRef resolve(const BRef& r, const SigmaPlan& plan, const OpSet* left, const OpSet* right) {
   MatRef a;
   if (r.src == SRC_NONE || r.op < 0 || r.blk < 0) return a;
   const Block* b; int64_t base;
   if (r.src == SRC_LEFT) { const OpTensor& t = left->ops[r.op]; b = &t.lay->blk[r.blk]; base = t.off; a.space = SP_LEFT; }
   else if (r.src == SRC_RIGHT) { const OpTensor& t = right->ops[r.op]; b = &t.lay->blk[r.blk]; base = t.off; a.space = SP_RIGHT; }
   else { const Presum& p = plan.presums[r.op]; b = &p.lay->blk[r.blk]; base = p.off; a.space = SP_PRESUM; }
   a.off = base + b->off; a.rows = b->rows; a.cols = b->cols; a.trans = (uint8_t)r.trans;
   return a;
}

}   // namespace

void compile_sigma(CompiledSigma& out, const SigmaPlan& plan, const OpSet* left, const OpSet* right, int rank, int world,
                   const CompileOptions& opt) {
   out = CompiledSigma();
   const SLayout& S = plan.S;
   const int nk = S.nkappa();
   auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
   const double tc0 = now();

   // ---- presums
   for (const Presum& p : plan.presums) {
      PresumJob j{};
      j.dst_off = p.off; j.size = p.lay->size; j.part_begin = (int)out.presum_parts.size();
      const OpSet* set = (p.side == SRC_LEFT) ? left : right;
      for (auto& pr : p.parts) {
         PresumPart pp{};
         pp.src_off = set->ops[pr.second].off; pp.coef = pr.first; pp.space = (p.side == SRC_LEFT) ? SP_LEFT : SP_RIGHT;
         out.presum_parts.push_back(pp);
      }
      j.part_end = (int)out.presum_parts.size();
      out.presum_jobs.push_back(j);
   }

   // ---- terms per target block (the plan is generated block by block, so this is a counting sort that keeps order)
   std::vector<int> cnt(nk + 1, 0);
   for (const SigmaTerm& t : plan.terms)
      if (world <= 1 || t.owner == rank) cnt[t.dst + 1]++;
   for (int k = 0; k < nk; k++) cnt[k + 1] += cnt[k];
   std::vector<int> order(cnt[nk]);
   {
      std::vector<int> pos(cnt.begin(), cnt.end() - 1);
      for (int t = 0; t < (int)plan.terms.size(); t++)
         if (world <= 1 || plan.terms[t].owner == rank) order[pos[plan.terms[t].dst]++] = t;
   }
   out.n_terms_used = (long long)order.size();

   const double tc1 = now();
   std::vector<DstBlock> dst(nk);
   for (int k = 0; k < nk; k++) dst[k] = DstBlock{S.blk[k].off, S.blk[k].rows, S.blk[k].cols};
   // ~100 bytes per term, a million terms: raw storage, constructed by the thread that resolves the range (parallel first touch)
   struct TermStore {
      size_t bytes; Term3* p;
      explicit TermStore(size_t n) : bytes(sizeof(Term3) * std::max<size_t>(n, 1)), p((Term3*)host_block_acquire(bytes)) {}
      ~TermStore() { host_block_release(p, bytes); }
   } term_store(order.size());
   Term3* terms = term_store.p;
   auto resolve_range = [&](size_t b, size_t e) {
      for (size_t i = b; i < e; i++) {
         const SigmaTerm& t = plan.terms[order[i]];
         const Block& sb = S.blk[t.src];
         Term3& x = *::new ((void*)(terms + i)) Term3();
         x.dst = t.dst; x.f = t.factor;
         x.p = resolve(t.l, plan, left, right);
         x.r = resolve(t.r, plan, left, right);
         x.q.space = SP_VIN; x.q.off = sb.off; x.q.rows = sb.rows; x.q.cols = sb.cols; x.q.trans = 0;
      }
   };
   {
      const int T = std::max(1, std::min<int>(opt.threads, (int)(order.size() / 20000)));
      parallel_run(T, [&](int t) { resolve_range(order.size() * t / T, order.size() * (t + 1) / T); });
   }
   const double tc2 = now();
   // ---- diagonal of H_eff (Heff::fillHeffDiag, Heff.cpp:250-315 + HeffDiagonal.cpp): exactly the terms that map a block
   // onto itself — families 1A-1D, 2d3, 2b3/2c3/2e3/2f3 and 2a3 — restricted to the operator-block diagonals:
   //    diag[k](i,j) = sum_t f_t * op(A_t)(i,i) * op(B_t)(j,j)
   for (int k = 0; k < nk; k++) {
      const int ib = (int)out.diag_items.size();
      for (int oi = cnt[k]; oi < cnt[k + 1]; oi++) {
         const SigmaTerm& t = plan.terms[order[oi]];
         if (t.src != t.dst) continue;
         const Term3& x = terms[oi];
         DiagItem d{};
         d.f = t.factor; d.as = x.p.space; d.aoff = x.p.off; d.lda = x.p.rows; d.bs = x.r.space; d.boff = x.r.off; d.ldb = x.r.rows;
         out.diag_items.push_back(d);
      }
      const int ie = (int)out.diag_items.size();
      if (ie == ib) continue;
      const Block& db = S.blk[k];
      for (int n0 = 0; n0 < db.cols; n0 += 32)
         for (int m0 = 0; m0 < db.rows; m0 += 32) {
            DiagTile t{};
            t.coff = db.off; t.ldc = db.rows; t.m0 = m0; t.n0 = n0; t.mrem = std::min(32, db.rows - m0); t.nrem = std::min(32, db.cols - n0);
            t.item_begin = ib; t.item_end = ie;
            out.diag_tiles.push_back(t);
         }
   }

   const double tc3 = now();
   CompiledWork& base = out;
   CompiledWork work;
   compile_terms(work, terms, order.size(), dst, SP_VOUT, opt);
   base = std::move(work);
   if (getenv("B2_TIMING")) fprintf(stderr, "compile_sigma: order %.3f s, resolve %.3f s, diagonal %.3f s, schedule %.3f s\n", tc1 - tc0, tc2 - tc1, tc3 - tc2, now() - tc3);
}

}   // namespace b2
