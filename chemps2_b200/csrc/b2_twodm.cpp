// b2_twodm.cpp — the site contribution to the spin-summed 2-RDM: TwoDM::FillSite (TwoDM.cpp:445-628) and its 24 diagram
// functions doD1 .. doD24 (TwoDM.cpp:642-1592), restructured for the GPU.
//
// Every diagram of the reference is a sum over symmetry sectors of
//        f * < T_up , [L_g^T] T_down op(R_jk) >                     (T = MPS tensor of the site, L_g a left-block operator,
// computed with two dgemm_ and one ddot_ PER (g, j, k) triple.         R_jk a right-block L / S0 / S1 / F0 / F1 operator)
// Here the trace is regrouped as an inner product of two operators that live on the SAME boundary,
//        < T_down^T L_g T_up , R_jk >     resp.   < T_up^T L_g^T T_down , R_jk >   when R enters transposed,
// so that per site only O(L) "effective operators" M_g are contracted (three-factor terms of the same form as the operator
// update, executed by the grouped DMMA kernels) and ALL diagram values follow from one Gram matrix per operator type,
//        G[M, R] = < M , R >       (a K-concatenated GEMM over the packed operator storage, again the same kernels).
// The host only enumerates sectors / Wigner factors and scatters the Gram entries into the 2-RDM like FillSite does.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>

#include "b2_twodm.h"

namespace b2 {

namespace {

struct Sec { int n, ts, ir; };

enum Tag { T1 = 0, T2, T3, T45, T6, T7, T8, T9, T10, T11, T12, T1314, T1516, T1718, T1920, T2122, T2324, NTAGS };

int kind_of_tag(int tag) {
   switch (tag) {
      case T2: case T7: case T8: case T9: case T10: case T11: case T12: return K_L;
      case T3: case T1314: return K_S0;
      case T1516: return K_S1;
      case T45: case T1718: case T2122: return K_F0;
      default: return K_F1;   // T6, T1920, T2324
   }
}

}   // namespace

void build_twodm_plan(TwoDMPlan& plan, const Bookkeeper& bk, int site, const OpSet* left, const OpSet* right) {
   plan = TwoDMPlan();
   plan.site = site;
   plan.T.build(bk, site);
   const int th = site, L = bk.L, I_th = bk.orb_irrep[site];
   const TLayout& T = plan.T;

   // ---- effective operators, grouped by (side, kind, irrep) so that every group is a dense [size x count] matrix
   auto add_mop = [&](int tag, int g, int irrep, bool left_side) {
      TwoDMPlan::MOp m;
      m.tag = tag; m.g = g; m.kind = kind_of_tag(tag); m.irrep = irrep; m.left_side = left_side;
      auto lay = std::make_shared<OpLayout>();
      lay->build(bk, left_side ? th : th + 1, kind_two_j(m.kind), kind_nelec(m.kind), irrep);
      m.lay = lay;
      plan.mops.push_back(m);
      return (int)plan.mops.size() - 1;
   };
   std::vector<int> g_list;
   if (left) for (int g = 0; g < th; g++) if (left->find(K_L, g, g) >= 0) g_list.push_back(g);
   int m2 = -1, m3 = -1, m45 = -1, m6 = -1, m7 = -1;
   if (right) { m2 = add_mop(T2, -1, I_th, false); m3 = add_mop(T3, -1, 0, false); m45 = add_mop(T45, -1, 0, false); m6 = add_mop(T6, -1, 0, false); }
   if (left) m7 = add_mop(T7, -1, I_th, true);
   std::vector<std::vector<int>> mg(g_list.size(), std::vector<int>(NTAGS, -1));
   if (right)
      for (size_t a = 0; a < g_list.size(); a++) {
         const int Ig = bk.orb_irrep[g_list[a]], Igt = xorp(Ig, I_th);
         for (int tag : {T8, T9, T10, T11, T12}) mg[a][tag] = add_mop(tag, g_list[a], Ig, false);
         for (int tag : {T1314, T1516, T1718, T1920, T2122, T2324}) mg[a][tag] = add_mop(tag, g_list[a], Igt, false);
      }
   // group layout: groups in first-appearance order, members contiguous with a common stride
   std::map<std::tuple<int, int, int>, int> gid;
   for (size_t i = 0; i < plan.mops.size(); i++) {
      TwoDMPlan::MOp& m = plan.mops[i];
      const auto key = std::make_tuple((int)m.left_side, m.kind, m.irrep);
      auto it = gid.find(key);
      if (it == gid.end()) {
         TwoDMPlan::Group grp;
         grp.left_side = m.left_side; grp.kind = m.kind; grp.irrep = m.irrep;
         grp.stride = (m.lay->size + 15) / 16 * 16;
         plan.groups.push_back(grp);
         it = gid.emplace(key, (int)plan.groups.size() - 1).first;
      }
      m.group = it->second;
      m.col = (int)plan.groups[it->second].members.size();
      plan.groups[it->second].members.push_back((int)i);
   }
   int64_t off = 0;
   for (TwoDMPlan::Group& grp : plan.groups) {
      grp.off = off;
      for (size_t c = 0; c < grp.members.size(); c++) plan.mops[grp.members[c]].off = off + (int64_t)c * grp.stride;
      off += grp.stride * (int64_t)grp.members.size();
   }
   plan.m_size = off;
   // destination blocks
   plan.block_base.resize(plan.mops.size());
   for (size_t i = 0; i < plan.mops.size(); i++) {
      plan.block_base[i] = (int)plan.dst.size();
      for (const Block& b : plan.mops[i].lay->blk) plan.dst.push_back(DstBlock{plan.mops[i].off + b.off, b.rows, b.cols});
   }

   // ---- helpers
   auto tref = [&](const Sec& l, const Sec& r, bool trans) {
      MatRef m;
      const int k = T.kappa(bk, l.n, l.ts, l.ir, r.n, r.ts, r.ir);
      if (k < 0) return m;
      m.space = SP_RIGHT; m.off = T.blk[k].off; m.rows = T.blk[k].rows; m.cols = T.blk[k].cols; m.trans = trans;
      return m;
   };
   auto lref = [&](int g, const Sec& a, const Sec& b, bool trans) {   // block a -> b of the left L operator of site g
      MatRef m;
      const int op = left ? left->find(K_L, g, g) : -1;
      if (op < 0) return m;
      const OpTensor& t = left->ops[op];
      const int k = t.lay->kappa(bk, a.n, a.ts, a.ir, b.n, b.ts, b.ir);
      if (k < 0) return m;
      m.space = SP_LEFT; m.off = t.off + t.lay->blk[k].off; m.rows = t.lay->blk[k].rows; m.cols = t.lay->blk[k].cols; m.trans = trans;
      return m;
   };
   // M[mop][block a -> b] += f * op(p) [op(q)] op(r)
   auto emit = [&](int mop, const Sec& a, const Sec& b, const MatRef& p, const MatRef* q, const MatRef& r, double f) {
      if (mop < 0 || f == 0.0 || !p.present() || !r.present() || (q && !q->present())) return;
      const int k = plan.mops[mop].lay->kappa(bk, a.n, a.ts, a.ir, b.n, b.ts, b.ir);
      if (k < 0) return;
      Term3 t;
      t.dst = plan.block_base[mop] + k; t.f = f; t.p = p; t.r = r;
      if (q) t.q = *q;
      plan.terms.push_back(t);
   };
   const double s5 = std::sqrt(0.5);

   bk.for_sectors(th, [&](int NL, int TwoSL, int IL) {
      if (bk.dim(th, NL, TwoSL, IL) <= 0) return;
      const Sec lu{NL, TwoSL, IL};
      const int IRup = xorp(IL, I_th);
      // ------------------------------------------------------------------ diagrams without a left operator (doD2 .. doD6)
      if (right) {
         for (int TwoSR = TwoSL - 1; TwoSR <= TwoSL + 1; TwoSR += 2) {   // doD2 (:670-713): Lright stored (ru -> rd), enters transposed
            if (TwoSR < 0) continue;
            const Sec ru{NL + 1, TwoSR, IRup}, rd{NL + 2, TwoSL, IL};
            emit(m2, ru, rd, tref(lu, ru, true), nullptr, tref(lu, rd, false), phase(TwoSL + 1 - TwoSR) * 0.5 * std::sqrt((TwoSL + 1) * (TwoSR + 1.0)));
         }
         {  // doD3 (:715-753)
            const Sec rd{NL, TwoSL, IL}, ru{NL + 2, TwoSL, IL};
            emit(m3, rd, ru, tref(lu, rd, true), nullptr, tref(lu, ru, false), s5 * (TwoSL + 1));
         }
         {  // doD4 (:755-791)
            const Sec r{NL + 2, TwoSL, IL};
            emit(m45, r, r, tref(lu, r, true), nullptr, tref(lu, r, false), s5 * (TwoSL + 1));
         }
         for (int TwoSR = TwoSL - 1; TwoSR <= TwoSL + 1; TwoSR += 2) {   // doD5 (:793-832)
            if (TwoSR < 0) continue;
            const Sec r{NL + 1, TwoSR, IRup};
            emit(m45, r, r, tref(lu, r, true), nullptr, tref(lu, r, false), 0.5 * s5 * (TwoSR + 1));
         }
         for (int TwoSRup = TwoSL - 1; TwoSRup <= TwoSL + 1; TwoSRup += 2)   // doD6 (:834-879)
            for (int TwoSRdown = TwoSL - 1; TwoSRdown <= TwoSL + 1; TwoSRdown += 2) {
               if (TwoSRup < 0 || TwoSRdown < 0) continue;
               const Sec ru{NL + 1, TwoSRup, IRup}, rd{NL + 1, TwoSRdown, IRup};
               const double f = std::sqrt((TwoSRup + 1) / 3.0) * (TwoSRdown + 1) * phase(TwoSL + TwoSRdown - 1) * wigner6j(1, 1, 2, TwoSRup, TwoSRdown, TwoSL);
               emit(m6, rd, ru, tref(lu, rd, true), nullptr, tref(lu, ru, false), f);
            }
      }
      // ------------------------------------------------------------------ doD7 (:881-925): only a left operator; N7 lives on the LEFT boundary
      if (left)
         for (int TwoSLdown = TwoSL - 1; TwoSLdown <= TwoSL + 1; TwoSLdown += 2) {
            if (TwoSLdown < 0) continue;
            const Sec ld{NL - 1, TwoSLdown, IRup}, r{NL + 1, TwoSLdown, IRup};
            emit(m7, ld, lu, tref(ld, r, false), nullptr, tref(lu, r, true), 0.5 * std::sqrt((TwoSLdown + 1) * (TwoSL + 1.0)) * phase(TwoSL - TwoSLdown + 3));
         }
      // ------------------------------------------------------------------ diagrams with a left operator L_g (doD8 .. doD24)
      if (!left || !right) return;
      const int TwoSLup = TwoSL;
      for (size_t a = 0; a < g_list.size(); a++) {
         const int g = g_list[a], Ig = bk.orb_irrep[g];
         const int Idown = xorp(IL, Ig);              // irrep of the lower left sector
         const int IRdown = xorp(Idown, I_th);
         for (int TwoSLdown = TwoSLup - 1; TwoSLdown <= TwoSLup + 1; TwoSLdown += 2) {
            if (TwoSLdown < 0) continue;
            const Sec ld{NL - 1, TwoSLdown, Idown};
            if (bk.dim(th, ld.n, ld.ts, ld.ir) <= 0) continue;
            const MatRef Lg = lref(g, ld, lu, false), LgT = lref(g, ld, lu, true);
            if (!Lg.present()) continue;
            {  // doD8 (:927-979)
               const Sec ru{NL + 2, TwoSLup, IL}, rd{NL + 1, TwoSLdown, Idown};
               emit(mg[a][T8], rd, ru, tref(ld, rd, true), &Lg, tref(lu, ru, false), -0.5 * (TwoSLup + 1));
            }
            // doD9, doD10, doD11 (:981-1051)
            for (int TwoSRup = TwoSLup - 1; TwoSRup <= TwoSLup + 1; TwoSRup += 2)
               for (int TwoSRdown = TwoSRup - 1; TwoSRdown <= TwoSRup + 1; TwoSRdown += 2) {
                  if (TwoSRup < 0 || TwoSRdown < 0 || std::abs(TwoSLdown - TwoSRdown) > 1) continue;
                  const Sec ru{NL + 1, TwoSRup, IRup}, rd{NL, TwoSRdown, IRdown};
                  const MatRef p = tref(ld, rd, true), r = tref(lu, ru, false);
                  const double common = (TwoSRup + 1) * std::sqrt((TwoSRdown + 1) * (TwoSLup + 1.0));
                  const double f9 = phase(TwoSLup + TwoSRdown + 2) * common * wigner6j(TwoSRup, 1, TwoSLup, TwoSLdown, 1, TwoSRdown);
                  const double f10 = 2 * common * wigner6j(TwoSRup, TwoSLdown, 2, 1, 1, TwoSLup) * wigner6j(TwoSRup, TwoSLdown, 2, 1, 1, TwoSRdown);
                  const double f11 = (TwoSRdown == TwoSLup) ? (TwoSRup + 1) : 0.0;
                  emit(mg[a][T9], rd, ru, p, &Lg, r, f9);
                  emit(mg[a][T10], rd, ru, p, &Lg, r, f10);
                  emit(mg[a][T11], rd, ru, p, &Lg, r, f11);
               }
            {  // doD12 (:1053-1106): Lright stored (ru -> rd), enters transposed
               const Sec ru{NL, TwoSLup, IL}, rd{NL + 1, TwoSLdown, Idown};
               emit(mg[a][T12], ru, rd, tref(lu, ru, true), &LgT, tref(ld, rd, false), phase(TwoSLdown + 1 - TwoSLup) * 0.5 * std::sqrt((TwoSLup + 1) * (TwoSLdown + 1.0)));
            }
            {  // doD13 (:1108-1162)
               const Sec ru{NL + 2, TwoSLup, IL}, rd{NL, TwoSLup, IRdown};
               emit(mg[a][T1314], rd, ru, tref(ld, rd, true), &Lg, tref(lu, ru, false), -0.5 * s5 * (TwoSLup + 1));
            }
            {  // doD14 (:1164-1218)
               const Sec ru{NL + 1, TwoSLdown, IRup}, rd{NL - 1, TwoSLdown, Idown};
               emit(mg[a][T1314], rd, ru, tref(ld, rd, true), &Lg, tref(lu, ru, false), phase(TwoSLdown + 1 - TwoSLup) * 0.5 * std::sqrt(0.5 * (TwoSLup + 1) * (TwoSLdown + 1)));
            }
            for (int TwoSRdown = TwoSLdown - 1; TwoSRdown <= TwoSLdown + 1; TwoSRdown += 2) {   // doD15 (:1220-1277)
               if (TwoSRdown < 0) continue;
               const Sec ru{NL + 2, TwoSLup, IL}, rd{NL, TwoSRdown, IRdown};
               const double f = phase(TwoSLdown + TwoSLup + 1) * (TwoSLup + 1) * std::sqrt((TwoSRdown + 1) / 3.0) * wigner6j(1, 1, 2, TwoSLup, TwoSRdown, TwoSLdown);
               emit(mg[a][T1516], rd, ru, tref(ld, rd, true), &Lg, tref(lu, ru, false), f);
            }
            for (int TwoSRup = TwoSLup - 1; TwoSRup <= TwoSLup + 1; TwoSRup += 2) {   // doD16 (:1279-1336)
               if (TwoSRup < 0) continue;
               const Sec ru{NL + 1, TwoSRup, IRup}, rd{NL - 1, TwoSLdown, Idown};
               const double f = phase(TwoSRup + TwoSLdown + 2) * (TwoSRup + 1) * std::sqrt((TwoSLup + 1) / 3.0) * wigner6j(1, 1, 2, TwoSRup, TwoSLdown, TwoSLup);
               emit(mg[a][T1516], rd, ru, tref(ld, rd, true), &Lg, tref(lu, ru, false), f);
            }
            {  // doD17 / doD21 (:1338-1393)
               const Sec ru{NL, TwoSLup, IL}, rd{NL, TwoSLup, IRdown};
               const double f = s5 * 0.5 * (TwoSLup + 1);
               emit(mg[a][T1718], rd, ru, tref(ld, rd, true), &Lg, tref(lu, ru, false), f);
               emit(mg[a][T2122], ru, rd, tref(lu, ru, true), &LgT, tref(ld, rd, false), f);
            }
            {  // doD18 / doD22 (:1395-1452)
               const Sec ru{NL + 1, TwoSLdown, IRup}, rd{NL + 1, TwoSLdown, Idown};
               const double f = phase(TwoSLdown + 1 - TwoSLup) * 0.5 * std::sqrt(0.5 * (TwoSLup + 1) * (TwoSLdown + 1));
               emit(mg[a][T1718], rd, ru, tref(ld, rd, true), &Lg, tref(lu, ru, false), f);
               emit(mg[a][T2122], ru, rd, tref(lu, ru, true), &LgT, tref(ld, rd, false), f);
            }
            for (int TwoSRdown = TwoSLdown - 1; TwoSRdown <= TwoSLdown + 1; TwoSRdown += 2) {   // doD19 / doD23 (:1454-1520)
               if (TwoSRdown < 0) continue;
               const Sec ru{NL, TwoSLup, IL}, rd{NL, TwoSRdown, IRdown};
               const double w = wigner6j(1, 1, 2, TwoSLup, TwoSRdown, TwoSLdown);
               const double f19 = phase(TwoSLdown + TwoSRdown - 1) * (TwoSRdown + 1) * std::sqrt((TwoSLup + 1) / 3.0) * w;
               const double f23 = phase(TwoSLdown + TwoSLup - 1) * (TwoSLup + 1) * std::sqrt((TwoSRdown + 1) / 3.0) * w;
               emit(mg[a][T1920], rd, ru, tref(ld, rd, true), &Lg, tref(lu, ru, false), f19);
               emit(mg[a][T2324], ru, rd, tref(lu, ru, true), &LgT, tref(ld, rd, false), f23);
            }
            for (int TwoSRup = TwoSLup - 1; TwoSRup <= TwoSLup + 1; TwoSRup += 2) {   // doD20 / doD24 (:1522-1590)
               if (TwoSRup < 0) continue;
               const Sec ru{NL + 1, TwoSRup, IRup}, rd{NL + 1, TwoSLdown, Idown};
               const double w = wigner6j(1, 1, 2, TwoSRup, TwoSLdown, TwoSLup);
               const double f20 = phase(2 * TwoSLup) * std::sqrt((TwoSLup + 1) * (TwoSRup + 1) * (TwoSLdown + 1) / 3.0) * w;
               const double f24 = phase(2 * TwoSLup + TwoSRup - TwoSLdown) * (TwoSRup + 1) * std::sqrt((TwoSLup + 1) / 3.0) * w;
               emit(mg[a][T1920], rd, ru, tref(ld, rd, true), &Lg, tref(lu, ru, false), f20);
               emit(mg[a][T2324], ru, rd, tref(lu, ru, true), &LgT, tref(ld, rd, false), f24);
            }
         }
      }
   });

   // ---- doD1 (:642-668): sum over blocks l -> (NL+2, 2SL, IL) of (2SL+1) |T block|^2, as < T , scaled copy of T >
   for (int k = 0; k < T.nkappa(); k++) {
      const bool dbl = (T.NR[k] == T.NL[k] + 2);
      plan.d1_scale.push_back(dbl ? (double)(T.twoSL[k] + 1) : 0.0);
   }

   // ---- the stored operators every group is paired with
   for (TwoDMPlan::Group& grp : plan.groups) {
      const OpSet* set = grp.left_side ? left : right;
      if (!set) continue;
      for (size_t i = 0; i < set->ops.size(); i++) {
         const OpTensor& t = set->ops[i];
         const bool kind_ok = grp.partner_kinds.empty() ? (t.kind == grp.kind) : (std::find(grp.partner_kinds.begin(), grp.partner_kinds.end(), t.kind) != grp.partner_kinds.end());
         if (!kind_ok || t.irrep != grp.irrep) continue;
         if (grp.left_side && !(t.i < th)) continue;
         grp.partners.push_back((int)i);
      }
   }
   (void)L;
}

void build_corr_plan(TwoDMPlan& plan, const Bookkeeper& bk, int site, const OpSet& corr) {
   plan = TwoDMPlan();
   plan.site = site;
   plan.T.build(bk, site);
   const int th = site, I_th = bk.orb_irrep[site];
   const TLayout& T = plan.T;
   // N1, N2, N3 pair with G / Y / Z (two_j 0, n_elec 0, irrep 0); N4, N5 with K / M of the sites that carry the irrep of this site
   const int tags[5] = {CORR_D1, CORR_D2, CORR_D3, CORR_D4, CORR_D5};
   for (int a = 0; a < 5; a++) {
      TwoDMPlan::MOp m;
      m.tag = tags[a]; m.g = -1; m.left_side = true;
      m.kind = a < 3 ? K_G : K_K;
      m.irrep = a < 3 ? 0 : I_th;
      auto lay = std::make_shared<OpLayout>();
      lay->build(bk, th, kind_two_j(m.kind), kind_nelec(m.kind), m.irrep);
      m.lay = lay;
      plan.mops.push_back(m);
   }
   for (int gi = 0; gi < 2; gi++) {
      TwoDMPlan::Group grp;
      grp.left_side = true; grp.kind = gi == 0 ? K_G : K_K; grp.irrep = gi == 0 ? 0 : I_th;
      grp.partner_kinds = gi == 0 ? std::vector<int>{K_G, K_Y, K_Z} : std::vector<int>{K_K, K_M};
      const int first = gi == 0 ? 0 : 3, count = gi == 0 ? 3 : 2;
      grp.stride = (plan.mops[first].lay->size + 15) / 16 * 16;
      grp.off = plan.m_size;
      for (int c = 0; c < count; c++) {
         plan.mops[first + c].group = gi; plan.mops[first + c].col = c; plan.mops[first + c].off = grp.off + (int64_t)c * grp.stride;
         grp.members.push_back(first + c);
      }
      plan.m_size += grp.stride * count;
      for (size_t i = 0; i < corr.ops.size(); i++) {
         const OpTensor& t = corr.ops[i];
         if (std::find(grp.partner_kinds.begin(), grp.partner_kinds.end(), t.kind) == grp.partner_kinds.end() || t.irrep != grp.irrep) continue;
         grp.partners.push_back((int)i);
      }
      plan.groups.push_back(grp);
   }
   plan.block_base.resize(plan.mops.size());
   for (size_t i = 0; i < plan.mops.size(); i++) {
      plan.block_base[i] = (int)plan.dst.size();
      for (const Block& b : plan.mops[i].lay->blk) plan.dst.push_back(DstBlock{plan.mops[i].off + b.off, b.rows, b.cols});
   }
   auto tref = [&](int nl, int tsl, int il, int nr, int tsr, int ir, bool trans) {
      MatRef m;
      const int k = T.kappa(bk, nl, tsl, il, nr, tsr, ir);
      if (k < 0) return m;
      m.space = SP_RIGHT; m.off = T.blk[k].off; m.rows = T.blk[k].rows; m.cols = T.blk[k].cols; m.trans = trans;
      return m;
   };
   // N[mop][block a -> b] += f * Ta * Tb^T       (value = < N , stored tensor block a -> b >)
   auto emit = [&](int mop, int an, int ats, int air, int bn, int bts, int bir, const MatRef& ta, const MatRef& tbT, double f) {
      if (!ta.present() || !tbT.present() || f == 0.0) return;
      const int k = plan.mops[mop].lay->kappa(bk, an, ats, air, bn, bts, bir);
      if (k < 0) return;
      Term3 t;
      t.dst = plan.block_base[mop] + k; t.f = f; t.p = ta; t.r = tbT;
      plan.terms.push_back(t);
   };
   bk.for_sectors(th + 1, [&](int NR, int TwoSR, int IR) {
      if (bk.dim(th + 1, NR, TwoSR, IR) <= 0) return;
      const int Ix = xorp(IR, I_th);
      // diagram1 (Correlations.cpp:353-387): site empty;  < T , Y T >  =  < Y , T T^T >
      emit(0, NR, TwoSR, IR, NR, TwoSR, IR, tref(NR, TwoSR, IR, NR, TwoSR, IR, false), tref(NR, TwoSR, IR, NR, TwoSR, IR, true), TwoSR + 1.0);
      // diagram2 (:389-423): site doubly occupied
      emit(1, NR - 2, TwoSR, IR, NR - 2, TwoSR, IR, tref(NR - 2, TwoSR, IR, NR, TwoSR, IR, false), tref(NR - 2, TwoSR, IR, NR, TwoSR, IR, true), TwoSR + 1.0);
      for (int TwoSL = TwoSR - 1; TwoSL <= TwoSR + 1; TwoSL += 2) {
         if (TwoSL < 0) continue;
         // diagram3 (:425-467): site singly occupied
         emit(2, NR - 1, TwoSL, Ix, NR - 1, TwoSL, Ix, tref(NR - 1, TwoSL, Ix, NR, TwoSR, IR, false), tref(NR - 1, TwoSL, Ix, NR, TwoSR, IR, true), TwoSR + 1.0);
         // diagram4 (:469-513): K block (ld -> lu) with lu = (NR, TwoSR, IR) [site empty], ld = (NR-1, TwoSL, Ix) [site single]:  < Tdown , K Tup >
         emit(3, NR - 1, TwoSL, Ix, NR, TwoSR, IR, tref(NR - 1, TwoSL, Ix, NR, TwoSR, IR, false), tref(NR, TwoSR, IR, NR, TwoSR, IR, true), TwoSR + 1.0);
         // diagram5 (:515-560): M block (ld -> lu) with ld = (NR-2, TwoSR, IR) [double], lu = (NR-1, TwoSL, Ix) [single]
         const int fase = ((((TwoSL + 1 - TwoSR) / 2) % 2) != 0) ? -1 : 1;
         emit(4, NR - 2, TwoSR, IR, NR - 1, TwoSL, Ix, tref(NR - 2, TwoSR, IR, NR, TwoSR, IR, false), tref(NR - 1, TwoSL, Ix, NR, TwoSR, IR, true),
              fase * std::sqrt((TwoSL + 1.0) * (TwoSR + 1)));
      }
   });
   for (int k = 0; k < T.nkappa(); k++) plan.d1_scale.push_back(0.0);
}

double corr_value(const TwoDMPlan& plan, const OpSet& corr, const std::vector<std::vector<double>>& gram, int tag, int kind, int p) {
   const int op = corr.find(kind, p, p);
   if (op < 0) return 0.0;
   for (const TwoDMPlan::MOp& mo : plan.mops) {
      if (mo.tag != tag) continue;
      const TwoDMPlan::Group& grp = plan.groups[mo.group];
      for (size_t c = 0; c < grp.partners.size(); c++)
         if (grp.partners[c] == op) return gram[mo.group][mo.col + grp.members.size() * c];
   }
   return 0.0;
}

// scatter the Gram entries into the 2-RDM exactly like TwoDM::FillSite (TwoDM.cpp:445-628)
void twodm_scatter(const TwoDMPlan& plan, const Bookkeeper& bk, const OpSet* left, const OpSet* right, double d1,
                   const std::vector<std::vector<double>>& gram, double* A, double* B) {
   const int L = bk.L, th = plan.site;
   auto setA = [&](int c1, int c2, int c3, int c4, double v) {   // set_2rdm_A_DMRG (:76-85)
      A[c1 + L * (c2 + L * (c3 + L * (size_t)c4))] = v; A[c2 + L * (c1 + L * (c4 + L * (size_t)c3))] = v;
      A[c3 + L * (c4 + L * (c1 + L * (size_t)c2))] = v; A[c4 + L * (c3 + L * (c2 + L * (size_t)c1))] = v;
   };
   auto setB = [&](int c1, int c2, int c3, int c4, double v) {
      B[c1 + L * (c2 + L * (c3 + L * (size_t)c4))] = v; B[c2 + L * (c1 + L * (c4 + L * (size_t)c3))] = v;
      B[c3 + L * (c4 + L * (c1 + L * (size_t)c2))] = v; B[c4 + L * (c3 + L * (c2 + L * (size_t)c1))] = v;
   };
   // value of < M(tag, g) , partner operator (kind, i, j) >
   auto val = [&](int tag, int g, int kind, int i, int j) -> double {
      for (size_t m = 0; m < plan.mops.size(); m++) {
         const TwoDMPlan::MOp& mo = plan.mops[m];
         if (mo.tag != tag || mo.g != g) continue;
         const TwoDMPlan::Group& grp = plan.groups[mo.group];
         const OpSet* set = grp.left_side ? left : right;
         const int op = set->find(kind, i, j);
         for (size_t c = 0; c < grp.partners.size(); c++)
            if (grp.partners[c] == op) return gram[mo.group][mo.col + grp.members.size() * c];
         return 0.0;
      }
      return 0.0;
   };
   auto irr = [&](int o) { return bk.orb_irrep[o]; };
   setA(th, th, th, th, 2 * d1); setB(th, th, th, th, -2 * d1);
   if (right) {
      for (int j = th + 1; j < L; j++)
         if (irr(j) == irr(th)) { const double d2 = val(T2, -1, K_L, j, j); setA(th, j, th, th, 2 * d2); setB(th, j, th, th, -2 * d2); }
      for (int j = th + 1; j < L; j++)
         for (int k = j; k < L; k++) {
            if (irr(j) != irr(k)) continue;
            const double d3 = val(T3, -1, K_S0, j, k), d45 = val(T45, -1, K_F0, j, k), d6 = val(T6, -1, K_F1, j, k);
            setA(th, th, j, k, 2 * d3); setB(th, th, j, k, -2 * d3);
            setA(th, j, k, th, -2 * d45 - 3 * d6); setB(th, j, k, th, -2 * d45 + d6);
            setA(th, j, th, k, 4 * d45); setB(th, j, th, k, 2 * d6);
         }
   }
   if (left)
      for (int g = 0; g < th; g++)
         if (irr(g) == irr(th)) { const double d7 = val(T7, -1, K_L, g, g); setA(g, th, th, th, 2 * d7); setB(g, th, th, th, -2 * d7); }
   if (!left || !right) return;
   for (int g = 0; g < th; g++)
      for (int j = th + 1; j < L; j++) {
         if (irr(g) != irr(j)) continue;
         const double d8 = val(T8, g, K_L, j, j), d9 = val(T9, g, K_L, j, j), d10 = val(T10, g, K_L, j, j), d11 = val(T11, g, K_L, j, j), d12 = val(T12, g, K_L, j, j);
         setA(g, th, j, th, -4 * d8 - d9); setA(g, th, th, j, 2 * d8 + d11);
         setB(g, th, j, th, d9 - 2 * d10); setB(g, th, th, j, 2 * d8 + 2 * d10 - d11);
         setA(g, j, th, th, 2 * d12); setB(g, j, th, th, -2 * d12);
      }
   for (int g = 0; g < th; g++)
      for (int j = th + 1; j < L; j++)
         for (int k = j; k < L; k++) {
            if (xorp(irr(g), irr(th)) != xorp(irr(j), irr(k))) continue;
            const double s0 = val(T1314, g, K_S0, j, k), s1 = (k > j) ? val(T1516, g, K_S1, j, k) : 0.0;
            setA(g, th, j, k, 2 * s0 + 3 * s1); setA(g, th, k, j, 2 * s0 - 3 * s1);
            setB(g, th, j, k, -2 * s0 + s1); setB(g, th, k, j, -2 * s0 - s1);
            const double f0 = val(T1718, g, K_F0, j, k), f1 = val(T1920, g, K_F1, j, k);
            setA(g, j, k, th, -2 * f0 - 3 * f1); setA(g, j, th, k, 4 * f0);
            setB(g, j, k, th, -2 * f0 + f1); setB(g, j, th, k, 2 * f1);
            const double p0 = val(T2122, g, K_F0, j, k), p1 = val(T2324, g, K_F1, j, k);
            setA(g, k, j, th, -2 * p0 - 3 * p1); setA(g, k, th, j, 4 * p0);
            setB(g, k, j, th, -2 * p0 + p1); setB(g, k, th, j, 2 * p1);
         }
}

}   // namespace b2
