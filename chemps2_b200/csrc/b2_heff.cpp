// b2_heff.cpp — SigmaPlan -> device work lists.  Host only.
//
// Every two-sided term  sigma[dst] += f * op(A) * S[src] * op(B)  is split into a stage-1 product W (kept in a
// workspace that is sized to stay L2/HBM friendly) and a stage-2 product that accumulates into the sigma tile.  The
// multiplication order is chosen per term to minimise FLOPs, and identical stage-1 products (same operator block, same
// source block) are computed once and shared by every term that needs them (spin-1 operators reach up to three target
// sectors from one source).
//
// Scheduling:
//   * the term list (sorted by target block) is cut into WAVES so that the stage-1 workspace of one wave stays below
//     CompileOptions::work_budget — the workspace is reused wave after wave instead of growing with the term count;
//   * inside a wave the terms of one sigma tile are cut into split-K CHUNKS of ~chunk_k accumulated inner dimension, one
//     CTA each, so that a tile with thousands of terms spreads over many SMs.  A tile with one chunk adds straight into
//     sigma; otherwise the chunks write partial slots and a reduce job sums them in a fixed order.
// Everything is deterministic: no atomics anywhere.
#include "b2_heff.h"

#include <algorithm>
#include <cstring>
#include <unordered_map>

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

struct WKey {
   int64_t off; int32_t src; uint8_t space, trans, left_first;
   bool operator==(const WKey& o) const { return off == o.off && src == o.src && space == o.space && trans == o.trans && left_first == o.left_first; }
};
struct WKeyHash {
   size_t operator()(const WKey& k) const {
      uint64_t h = (uint64_t)k.off * 0x9E3779B97F4A7C15ULL ^ ((uint64_t)k.src << 20) ^ ((uint64_t)k.space << 8) ^ ((uint64_t)k.trans << 4) ^ k.left_first;
      return (size_t)(h ^ (h >> 29));
   }
};
struct WInfo { int64_t off; int rows, cols; };

}   // namespace

void compile_sigma(CompiledSigma& out, const SigmaPlan& plan, const OpSet* left, const OpSet* right, int rank, int world,
                   const CompileOptions& opt) {
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

   std::unordered_map<WKey, WInfo, WKeyHash> wmap;
   int64_t wave_work = 0, wave_part = 0;
   Wave wave{};
   auto open_wave = [&]() {
      for (int c = 0; c < kNumTileClasses; c++) {
         wave.t1_begin[c] = (int)out.tiles1[c].size();
         wave.t2_begin[c] = (int)out.tiles2[c].size();
      }
      wave.red_begin = (int)out.reduces.size();
      wave_work = 0; wave_part = 0;
      wmap.clear();
   };
   auto weight_sort = [&](std::vector<Tile>& v, int b, int e, const std::vector<GemmItem>& items) {
      // heaviest CTAs first (static load balance across the SMs)
      std::vector<std::pair<long long, int>> ord(e - b);
      for (int i = b; i < e; i++) {
         long long w = 0;
         for (int it = v[i].item_begin; it < v[i].item_end; it++) w += items[it].k + 4;
         ord[i - b] = {-w, i};
      }
      std::sort(ord.begin(), ord.end());
      std::vector<Tile> sorted(e - b);
      for (int i = 0; i < e - b; i++) sorted[i] = v[ord[i].second];
      std::copy(sorted.begin(), sorted.end(), v.begin() + b);
   };
   auto close_wave = [&]() {
      bool any = (int)out.reduces.size() > wave.red_begin;
      for (int c = 0; c < kNumTileClasses; c++) {
         wave.t1_end[c] = (int)out.tiles1[c].size();
         wave.t2_end[c] = (int)out.tiles2[c].size();
         any = any || wave.t1_end[c] > wave.t1_begin[c] || wave.t2_end[c] > wave.t2_begin[c];
         weight_sort(out.tiles1[c], wave.t1_begin[c], wave.t1_end[c], out.items1);
         weight_sort(out.tiles2[c], wave.t2_begin[c], wave.t2_end[c], out.items2);
      }
      wave.red_end = (int)out.reduces.size();
      out.work_size = std::max(out.work_size, wave_work);
      out.part_size = std::max(out.part_size, wave_part);
      if (any) out.waves.push_back(wave);
   };

   auto get_w = [&](bool left_first, const BlockAddr& opb, int trans, int src, int dimL, int dimR) -> WInfo {
      WKey key{opb.off, src, opb.space, (uint8_t)trans, (uint8_t)(left_first ? 1 : 0)};
      auto it = wmap.find(key);
      if (it != wmap.end()) return it->second;
      const Block& sb = S.blk[src];
      WInfo w{};
      GemmItem g{};
      g.alpha = 1.0;
      if (left_first) {   // W[dimL x dRs] = op(A)[dimL x dLs] * S[src][dLs x dRs]
         w.rows = dimL; w.cols = sb.cols;
         g.xs = opb.space; g.xoff = opb.off; g.flags = trans ? IF_TX : 0; g.ldx = opb.rows;
         g.ys = SP_VIN; g.yoff = sb.off; g.ldy = sb.rows; g.k = sb.rows;
      } else {            // W[dLs x dimR] = S[src][dLs x dRs] * op(B)[dRs x dimR]
         w.rows = sb.rows; w.cols = dimR;
         g.xs = SP_VIN; g.xoff = sb.off; g.ldx = sb.rows;
         g.ys = opb.space; g.yoff = opb.off; g.flags = trans ? IF_TY : 0; g.ldy = opb.rows; g.k = sb.cols;
      }
      w.off = wave_work;
      wave_work += ((int64_t)w.rows * w.cols + 15) / 16 * 16;
      const int ib = (int)out.items1.size();
      out.items1.push_back(g);
      const int cls = tile_class_for(w.rows, w.cols), e = kTileEdge[cls];
      for (int n0 = 0; n0 < w.cols; n0 += e)
         for (int m0 = 0; m0 < w.rows; m0 += e) {
            Tile t{};
            t.coff = w.off; t.ldc = w.rows; t.m0 = t.cm0 = m0; t.n0 = t.cn0 = n0;
            t.mrem = std::min(e, w.rows - m0); t.nrem = std::min(e, w.cols - n0);
            t.item_begin = ib; t.item_end = ib + 1; t.cspace = SP_WORK; t.accumulate = 0;
            out.tiles1[cls].push_back(t);
         }
      out.flops_exec += 2.0 * w.rows * w.cols * g.k;
      out.n_stage1++;
      wmap.emplace(key, w);
      return w;
   };

   // emit the stage-2 CTAs of items2[ib, ie) for target block k (all inside the current wave)
   auto emit_block = [&](int k, int ib, int ie) {
      if (ie <= ib) return;
      const Block& db = S.blk[k];
      const int M = db.rows, N = db.cols;
      // split-K chunk boundaries
      std::vector<int> cuts{ib};
      int64_t acc = 0;
      for (int i = ib; i < ie; i++) {
         acc += out.items2[i].k + 4;
         if (acc >= opt.chunk_k && i + 1 < ie) { cuts.push_back(i + 1); acc = 0; }
      }
      cuts.push_back(ie);
      const int nchunks = (int)cuts.size() - 1;
      const int cls = tile_class_for(M, N), e = kTileEdge[cls];
      for (int n0 = 0; n0 < N; n0 += e)
         for (int m0 = 0; m0 < M; m0 += e) {
            const int mrem = std::min(e, M - m0), nrem = std::min(e, N - n0);
            if (nchunks == 1) {
               Tile t{};
               t.coff = db.off; t.ldc = M; t.m0 = t.cm0 = m0; t.n0 = t.cn0 = n0; t.mrem = mrem; t.nrem = nrem;
               t.item_begin = ib; t.item_end = ie; t.cspace = SP_VOUT; t.accumulate = 1;
               out.tiles2[cls].push_back(t);
            } else {
               const int64_t stride = ((int64_t)mrem * nrem + 15) / 16 * 16;
               ReduceJob r{};
               r.dst_off = db.off; r.ldc = M; r.m0 = m0; r.n0 = n0; r.mrem = mrem; r.nrem = nrem;
               r.part_off = wave_part; r.nparts = nchunks; r.part_stride = stride;
               out.reduces.push_back(r);
               for (int c = 0; c < nchunks; c++) {
                  Tile t{};
                  t.coff = wave_part + c * stride; t.ldc = mrem; t.m0 = m0; t.n0 = n0; t.cm0 = 0; t.cn0 = 0; t.mrem = mrem; t.nrem = nrem;
                  t.item_begin = cuts[c]; t.item_end = cuts[c + 1]; t.cspace = SP_PART; t.accumulate = 0;
                  out.tiles2[cls].push_back(t);
               }
               wave_part += stride * nchunks;
            }
         }
   };

   // block-axpy terms first inside every target block: the kernel consumes them before it starts its GEMM pipeline
   for (int k = 0; k < nk; k++)
      std::stable_partition(order.begin() + cnt[k], order.begin() + cnt[k + 1],
                            [&](int ti) { return plan.terms[ti].l.src == SRC_NONE && plan.terms[ti].r.src == SRC_NONE; });

   open_wave();
   for (int k = 0; k < nk; k++) {
      const Block& db = S.blk[k];
      const int dimL = db.rows, dimR = db.cols;
      int ib = (int)out.items2.size();
      for (int oi = cnt[k]; oi < cnt[k + 1]; oi++) {
         const SigmaTerm& t = plan.terms[order[oi]];
         const Block& sb = S.blk[t.src];
         const BlockAddr A = resolve(t.l, plan, left, right), B = resolve(t.r, plan, left, right);
         GemmItem g{};
         g.alpha = t.factor;
         if (A.space != SP_NONE && B.space != SP_NONE) {
            const double dLs = sb.rows, dRs = sb.cols;
            const double f_left = dimL * dLs * dRs + (double)dimL * dRs * dimR;    // (A*S)*B
            const double f_right = dLs * dRs * dimR + (double)dimL * dLs * dimR;   // A*(S*B)
            if (f_left <= f_right) {
               const WInfo w = get_w(true, A, t.l.trans, t.src, dimL, dimR);
               g.xs = SP_WORK; g.xoff = w.off; g.ldx = w.rows;
               g.ys = B.space; g.yoff = B.off; g.flags = t.r.trans ? IF_TY : 0; g.ldy = B.rows; g.k = sb.cols;
            } else {
               const WInfo w = get_w(false, B, t.r.trans, t.src, dimL, dimR);
               g.xs = A.space; g.xoff = A.off; g.flags = t.l.trans ? IF_TX : 0; g.ldx = A.rows;
               g.ys = SP_WORK; g.yoff = w.off; g.ldy = w.rows; g.k = sb.rows;
            }
         } else if (A.space != SP_NONE) {   // op(A) * S[src]
            g.xs = A.space; g.xoff = A.off; g.flags = t.l.trans ? IF_TX : 0; g.ldx = A.rows;
            g.ys = SP_VIN; g.yoff = sb.off; g.ldy = sb.rows; g.k = sb.rows;
         } else if (B.space != SP_NONE) {   // S[src] * op(B)
            g.xs = SP_VIN; g.xoff = sb.off; g.ldx = sb.rows;
            g.ys = B.space; g.yoff = B.off; g.flags = t.r.trans ? IF_TY : 0; g.ldy = B.rows; g.k = sb.cols;
         } else {                            // f * S[src]
            g.flags = IF_AXPY;
            g.xs = SP_VIN; g.xoff = sb.off; g.ldx = sb.rows; g.k = 0;
         }
         out.flops_exec += (g.flags & IF_AXPY) ? 2.0 * dimL * dimR : 2.0 * dimL * dimR * g.k;
         out.items2.push_back(g);
         if (wave_work >= opt.work_budget) {   // workspace full: flush what this block has so far and start a new wave
            emit_block(k, ib, (int)out.items2.size());
            close_wave();
            open_wave();
            ib = (int)out.items2.size();
         }
      }
      emit_block(k, ib, (int)out.items2.size());
   }
   close_wave();

   // ---- diagonal of H_eff (Heff::fillHeffDiag, Heff.cpp:250-315 + HeffDiagonal.cpp): exactly the terms that map a block
   // onto itself — families 1A-1D, 2d3, 2b3/2c3/2e3/2f3 and 2a3 — restricted to the operator-block diagonals:
   //    diag[k](i,j) = sum_t f_t * op(A_t)(i,i) * op(B_t)(j,j)
   for (int k = 0; k < nk; k++) {
      const int ib = (int)out.diag_items.size();
      for (int oi = cnt[k]; oi < cnt[k + 1]; oi++) {
         const SigmaTerm& t = plan.terms[order[oi]];
         if (t.src != t.dst) continue;
         const BlockAddr A = resolve(t.l, plan, left, right), B = resolve(t.r, plan, left, right);
         DiagItem d{};
         d.f = t.factor; d.as = A.space; d.aoff = A.off; d.lda = A.rows; d.bs = B.space; d.boff = B.off; d.ldb = B.rows;
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
   for (int c = 0; c < kNumTileClasses; c++) out.n_tiles += (long long)out.tiles1[c].size() + (long long)out.tiles2[c].size();
}

}   // namespace b2
