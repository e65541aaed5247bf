// b2_heff.cpp — SigmaPlan -> device work lists.  Host only.
//
// Every two-sided term  sigma[dst] += f * op(A) * S[src] * op(B)  is split into a stage-1 product W (kept in an
// HBM/L2 workspace) and a stage-2 product that accumulates into the sigma tile.  The multiplication order is chosen
// per term to minimise FLOPs, and identical stage-1 products (same operator block, same source block) are computed
// once and shared by every term that needs them (spin-1 operators reach up to three target sectors from one source).
// All terms of one sigma tile are accumulated by one CTA => K-concatenated GEMM, deterministic, no atomics.
#include "b2_heff.h"

#include <algorithm>
#include <map>
#include <tuple>

namespace b2 {

namespace {

struct BlockAddr { uint8_t space; int64_t off; int rows, cols; };

BlockAddr resolve(const BRef& r, const SigmaPlan& plan, const OpSet* left, const OpSet* right) {
   BlockAddr a{SP_NONE, 0, 0, 0};
   if (r.src == SRC_NONE || r.op < 0 || r.blk < 0) return a;
   if (r.src == SRC_LEFT) {
      const OpTensor& t = left->ops[r.op];
      const Block& b = t.lay->blk[r.blk];
      a = {SP_LEFT, t.off + b.off, b.rows, b.cols};
   } else if (r.src == SRC_RIGHT) {
      const OpTensor& t = right->ops[r.op];
      const Block& b = t.lay->blk[r.blk];
      a = {SP_RIGHT, t.off + b.off, b.rows, b.cols};
   } else {
      const Presum& p = plan.presums[r.op];
      const Block& b = p.lay->blk[r.blk];
      a = {SP_PRESUM, p.off + b.off, b.rows, b.cols};
   }
   return a;
}

int tile_class_for(int m, int n) {
   const int d = std::min(m, n);
   if (d > 32) return 0;
   if (d > 16) return 1;
   if (d > 8) return 2;
   return 3;
}

void add_tiles(std::vector<Tile>* per_class, uint8_t cspace, int64_t coff, int M, int N, int item_begin, int item_end) {
   const int cls = tile_class_for(M, N);
   const int e = kTileEdge[cls];
   for (int n0 = 0; n0 < N; n0 += e)
      for (int m0 = 0; m0 < M; m0 += e) {
         Tile t{};
         t.coff = coff; t.ldc = M; t.m0 = m0; t.n0 = n0;
         t.mrem = std::min(e, M - m0); t.nrem = std::min(e, N - n0);
         t.item_begin = item_begin; t.item_end = item_end; t.cspace = cspace;
         per_class[cls].push_back(t);
      }
}

}   // namespace

void compile_sigma(CompiledSigma& out, const SigmaPlan& plan, const OpSet* left, const OpSet* right, int rank, int world) {
   out = CompiledSigma();
   const SLayout& S = plan.S;
   const int nk = S.nkappa();

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

   // ---- group terms per target block
   std::vector<std::vector<int>> by_dst(nk);
   for (int t = 0; t < (int)plan.terms.size(); t++) {
      if (world > 1 && plan.terms[t].owner != rank) continue;
      by_dst[plan.terms[t].dst].push_back(t);
      out.n_terms_used++;
   }

   // ---- stage 1 products, de-duplicated
   // key: (first? 0 left-first / 1 right-first, operator space, operator offset, trans, source block)
   typedef std::tuple<int, int, int64_t, int, int> WKey;
   struct WInfo { int64_t off; int rows, cols; };
   std::map<WKey, WInfo> wmap;
   std::vector<GemmItem>& items = out.items;

   auto get_w = [&](bool left_first, const BlockAddr& opb, int trans, int src, int dimL, int dimR) -> WInfo {
      WKey key(left_first ? 0 : 1, (int)opb.space, opb.off, trans, src);
      auto it = wmap.find(key);
      if (it != wmap.end()) return it->second;
      const Block& sb = S.blk[src];
      WInfo w{};
      GemmItem g{};
      g.kind = IT_GEMM; g.alpha = 1.0;
      if (left_first) {   // W[dimL x dRs] = op(A)[dimL x dLs] * S[src][dLs x dRs]
         w.rows = dimL; w.cols = sb.cols;
         g.xs = opb.space; g.xoff = opb.off; g.tx = (uint8_t)trans; g.ldx = opb.rows;
         g.ys = SP_VIN; g.yoff = sb.off; g.ty = 0; g.ldy = sb.rows; g.k = sb.rows;
      } else {            // W[dLs x dimR] = S[src][dLs x dRs] * op(B)[dRs x dimR]
         w.rows = sb.rows; w.cols = dimR;
         g.xs = SP_VIN; g.xoff = sb.off; g.tx = 0; g.ldx = sb.rows;
         g.ys = opb.space; g.yoff = opb.off; g.ty = (uint8_t)trans; g.ldy = opb.rows; g.k = sb.cols;
      }
      w.off = out.work_size;
      out.work_size += ((int64_t)w.rows * w.cols + 15) / 16 * 16;
      const int ib = (int)items.size();
      items.push_back(g);
      add_tiles(out.tiles1, SP_WORK, w.off, w.rows, w.cols, ib, ib + 1);
      out.flops_exec += 2.0 * w.rows * w.cols * g.k;
      out.n_stage1++;
      wmap[key] = w;
      return w;
   };

   // ---- stage 2 item lists, one contiguous range per target block
   for (int k = 0; k < nk; k++) {
      const Block& db = S.blk[k];
      const int dimL = db.rows, dimR = db.cols;
      const int ib = (int)items.size();
      for (int ti : by_dst[k]) {
         const SigmaTerm& t = plan.terms[ti];
         const Block& sb = S.blk[t.src];
         const BlockAddr A = resolve(t.l, plan, left, right), B = resolve(t.r, plan, left, right);
         GemmItem g{};
         g.kind = IT_GEMM; g.alpha = t.factor;
         if (A.space != SP_NONE && B.space != SP_NONE) {
            const double dLs = sb.rows, dRs = sb.cols;
            const double f_left = dimL * dLs * dRs + (double)dimL * dRs * dimR;    // (A*S)*B
            const double f_right = dLs * dRs * dimR + (double)dimL * dLs * dimR;   // A*(S*B)
            if (f_left <= f_right) {
               const WInfo w = get_w(true, A, t.l.trans, t.src, dimL, dimR);
               g.xs = SP_WORK; g.xoff = w.off; g.tx = 0; g.ldx = w.rows;
               g.ys = B.space; g.yoff = B.off; g.ty = (uint8_t)t.r.trans; g.ldy = B.rows; g.k = sb.cols;
            } else {
               const WInfo w = get_w(false, B, t.r.trans, t.src, dimL, dimR);
               g.xs = A.space; g.xoff = A.off; g.tx = (uint8_t)t.l.trans; g.ldx = A.rows;
               g.ys = SP_WORK; g.yoff = w.off; g.ty = 0; g.ldy = w.rows; g.k = sb.rows;
            }
         } else if (A.space != SP_NONE) {   // op(A) * S[src]
            g.xs = A.space; g.xoff = A.off; g.tx = (uint8_t)t.l.trans; g.ldx = A.rows;
            g.ys = SP_VIN; g.yoff = sb.off; g.ty = 0; g.ldy = sb.rows; g.k = sb.rows;
         } else if (B.space != SP_NONE) {   // S[src] * op(B)
            g.xs = SP_VIN; g.xoff = sb.off; g.tx = 0; g.ldx = sb.rows;
            g.ys = B.space; g.yoff = B.off; g.ty = (uint8_t)t.r.trans; g.ldy = B.rows; g.k = sb.cols;
         } else {                            // f * S[src]
            g.kind = IT_AXPY;
            g.xs = SP_VIN; g.xoff = sb.off; g.ldx = sb.rows; g.k = 0;
         }
         out.flops_exec += (g.kind == IT_GEMM) ? 2.0 * dimL * dimR * g.k : 2.0 * dimL * dimR;
         items.push_back(g);
      }
      add_tiles(out.tiles2, SP_VOUT, db.off, dimL, dimR, ib, (int)items.size());
   }

   // heaviest tiles first (static load balance across the 148 SMs)
   auto weight = [&](const Tile& t) {
      long long w = 0;
      for (int i = t.item_begin; i < t.item_end; i++) w += items[i].k + 4;
      return w;
   };
   for (int c = 0; c < kNumTileClasses; c++) {
      for (std::vector<Tile>* v : {&out.tiles1[c], &out.tiles2[c]}) {
         std::vector<std::pair<long long, int>> order(v->size());
         for (size_t i = 0; i < v->size(); i++) order[i] = {-weight((*v)[i]), (int)i};
         std::sort(order.begin(), order.end());
         std::vector<Tile> sorted(v->size());
         for (size_t i = 0; i < v->size(); i++) sorted[i] = (*v)[order[i].second];
         v->swap(sorted);
         out.n_tiles += (long long)v->size();
      }
   }
}

}   // namespace b2
