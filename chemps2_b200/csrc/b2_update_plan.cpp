// b2_update_plan.cpp — enumerate the renormalized-operator update of one sweep step as three-factor terms (see b2_update.h).
//
// Every function cites the reference routine whose arithmetic it restates.  Sector bookkeeping: a new operator block
// connects an "up" sector U to a "down" sector Dn of the NEW boundary; each contribution picks a pair of sectors
// (Ux, Dx) of the OLD boundary reachable through the site tensor T and, optionally, an old operator block Ux -> Dx:
//    moving right:  new[U -> Dn] += f * T[Ux -> U]^T * mid[Ux -> Dx] * T[Dx -> Dn]
//    moving left :  new[U -> Dn] += f * T[U -> Ux]   * mid[Ux -> Dx] * T[Dn -> Dx]^T
#include <cassert>
#include <cmath>
#include <cstdlib>

#include "b2_update.h"

#include <atomic>

namespace b2 {

namespace {

struct Sec { int n, ts, ir; };

struct UGen {
   UpdatePlan& plan;               // receives terms, mixing terms, pre-sums and the FLOP count (the whole plan, or one thread's fragment)
   const UpdatePlan& shared;       // T layout, destination blocks, block_base: read only
   const Bookkeeper& bk;
   const Problem& prob;
   const OpSet* old_set;
   const OpSet& new_set;
   const int ix, L;
   const bool mr;
   const int b_old, b_new;    // boundaries
   const int site_irr;        // irrep of the site T lives on

   UGen(UpdatePlan& p, const UpdatePlan& sh, const Bookkeeper& b, const Problem& pr, const OpSet* o, const OpSet& n, int index, bool moving_right)
       : plan(p), shared(sh), bk(b), prob(pr), old_set(o), new_set(n), ix(index), L(b.L), mr(moving_right), b_old(moving_right ? index : index + 1),
         b_new(moving_right ? index + 1 : index), site_irr(b.orb_irrep[index]) {}

   double V(int a, int b, int c, int d) const { return prob.V(a, b, c, d); }
   int irr(int orb) const { return bk.orb_irrep[orb]; }
   int dim_old(const Sec& s) const { return bk.dim(b_old, s.n, s.ts, s.ir); }

   // T block between an old-boundary sector and a new-boundary sector
   MatRef tref(const Sec& so, const Sec& sn, bool is_up) const {
      MatRef m;
      const int k = mr ? shared.T.kappa(bk, so.n, so.ts, so.ir, sn.n, sn.ts, sn.ir) : shared.T.kappa(bk, sn.n, sn.ts, sn.ir, so.n, so.ts, so.ir);
      if (k < 0) return m;
      m.space = SP_RIGHT; m.off = shared.T.blk[k].off; m.rows = shared.T.blk[k].rows; m.cols = shared.T.blk[k].cols;
      m.trans = is_up ? (mr ? 1 : 0) : (mr ? 0 : 1);
      return m;
   }
   // block a -> b of an old operator (index in old_set); trans: stored as b -> a and entering transposed
   MatRef oref(int op, const Sec& a, const Sec& b, bool trans = false) const {
      MatRef m;
      if (op < 0 || !old_set) return m;
      const OpTensor& t = old_set->ops[op];
      const int k = trans ? t.lay->kappa(bk, b.n, b.ts, b.ir, a.n, a.ts, a.ir) : t.lay->kappa(bk, a.n, a.ts, a.ir, b.n, b.ts, b.ir);
      if (k < 0) return m;
      m.space = SP_LEFT; m.off = t.off + t.lay->blk[k].off; m.rows = t.lay->blk[k].rows; m.cols = t.lay->blk[k].cols; m.trans = trans;
      return m;
   }
   MatRef pref(int presum, const Sec& a, const Sec& b, bool trans = false) const {
      MatRef m;
      if (presum < 0) return m;
      const Presum& p = plan.presums[presum];
      const int k = trans ? p.lay->kappa(bk, b.n, b.ts, b.ir, a.n, a.ts, a.ir) : p.lay->kappa(bk, a.n, a.ts, a.ir, b.n, b.ts, b.ir);
      if (k < 0) return m;
      m.space = SP_PRESUM; m.off = p.off + p.lay->blk[k].off; m.rows = p.lay->blk[k].rows; m.cols = p.lay->blk[k].cols; m.trans = trans;
      return m;
   }
   int presum(const std::vector<std::pair<double, int>>& parts) {
      Presum p;
      p.side = SRC_LEFT;
      for (auto& pr : parts)
         if (pr.second >= 0 && pr.first != 0.0) p.parts.push_back(pr);
      if (p.parts.empty()) return -1;
      p.lay = old_set->ops[p.parts[0].second].lay;
      p.off = plan.presum_size;
      plan.presum_size += (p.lay->size + 15) / 16 * 16;
      plan.presums.push_back(p);
      return (int)plan.presums.size() - 1;
   }

   // ---- emit  new[op][k] += f * op(T[ux<->U]) * mid * op(T[dx<->Dn]);  mid_kind: 0 identity (ux == dx), 1 given MatRef
   void emit(int new_op, int k, const Sec& U, const Sec& Dn, const Sec& ux, const Sec& dx, const MatRef* mid, double f, bool count = true) {
      if (f == 0.0) return;
      if (dim_old(ux) <= 0 || dim_old(dx) <= 0) return;
      MatRef tu = tref(ux, U, true), td = tref(dx, Dn, false);
      if (!tu.present() || !td.present()) return;
      if (mid && !mid->present()) return;
      Term3 t;
      t.dst = shared.block_base[new_op] + k;
      t.f = f; t.p = tu; t.r = td;
      if (mid) t.q = *mid;
      plan.terms.push_back(t);
      const double m = shared.dst[t.dst].rows, n = shared.dst[t.dst].cols, du = dim_old(ux), dd = dim_old(dx);
      if (count) plan.flops_ref += mid ? 2.0 * (m * dd * du + m * n * dd) : 2.0 * m * n * du;
   }

   // sectors of block k of new operator `op`
   void block_secs(const OpTensor& t, int k, Sec& U, Sec& Dn) const {
      const OpLayout& l = *t.lay;
      U = Sec{l.Nup[k], l.twoSup[k], l.Iup[k]};
      Dn = Sec{l.Nup[k] + l.n_elec, l.twoSdown[k], xorp(l.Iup[k], l.irrep)};
   }
   // the old-boundary sector reached from new-boundary sector s when the site holds `occ` electrons coupled with spin change dts
   Sec step(const Sec& s, int occ, int dts) const {
      const int sgn = mr ? -1 : +1;   // moving right the old boundary is to the left (fewer electrons)
      return Sec{s.n + sgn * occ, s.ts + dts, (occ == 1) ? xorp(s.ir, site_irr) : s.ir};
   }

   // ============================================================================ TensorOperator::update (TensorOperator.cpp:163-405)
   void generic_update(int new_op, int old_op) {
      if (old_op < 0 || !old_set) return;
      const OpTensor& t = new_set.ops[new_op];
      const OpLayout& l = *t.lay;
      const int two_j = l.two_j;
      const bool jw = kind_jw(t.kind), prime_last = t.prime_last;
      for (int k = 0; k < l.nkappa(); k++) {
         Sec U, Dn;
         block_secs(t, k, U, Dn);
         for (int geval = 0; geval < 6; geval++) {
            Sec ux, dx;
            switch (geval) {
               case 0: ux = step(U, 0, 0); dx = step(Dn, 0, 0); break;
               case 1: ux = step(U, 2, 0); dx = step(Dn, 2, 0); break;
               case 2: ux = step(U, 1, -1); dx = step(Dn, 1, -1); break;
               case 3: ux = step(U, 1, -1); dx = step(Dn, 1, +1); break;
               case 4: ux = step(U, 1, +1); dx = step(Dn, 1, -1); break;
               default: ux = step(U, 1, +1); dx = step(Dn, 1, +1); break;
            }
            if (ux.ts < 0 || dx.ts < 0 || std::abs(ux.ts - dx.ts) > two_j) continue;
            double alpha = 1.0;
            if (geval >= 2) {
               if (mr) {   // :256-270 — "left" = old sectors (ux, dx), "right" = new sectors (U, Dn)
                  if (two_j == 0) alpha = jw ? -1.0 : 1.0;
                  else if (prime_last)
                     alpha = phase(U.ts + dx.ts + two_j + (jw ? 3 : 1)) * std::sqrt((dx.ts + 1.0) * (U.ts + 1.0)) * wigner6j(ux.ts, dx.ts, two_j, Dn.ts, U.ts, 1);
                  else
                     alpha = phase(Dn.ts + ux.ts + two_j + (jw ? 3 : 1)) * std::sqrt((ux.ts + 1.0) * (Dn.ts + 1.0)) * wigner6j(dx.ts, ux.ts, two_j, U.ts, Dn.ts, 1);
               } else {    // :370-385 — "left" = new sectors (U, Dn), "right" = old sectors (ux, dx)
                  if (two_j == 0) alpha = (jw ? -1.0 : 1.0) * ((ux.ts + 1.0) / (U.ts + 1));
                  else if (prime_last)
                     alpha = phase(ux.ts + Dn.ts + two_j + (jw ? 3 : 1)) * (dx.ts + 1) * std::sqrt((ux.ts + 1.0) / (Dn.ts + 1)) * wigner6j(ux.ts, dx.ts, two_j, Dn.ts, U.ts, 1);
                  else
                     alpha = phase(dx.ts + U.ts + two_j + (jw ? 3 : 1)) * (ux.ts + 1) * std::sqrt((dx.ts + 1.0) / (U.ts + 1)) * wigner6j(dx.ts, ux.ts, two_j, U.ts, Dn.ts, 1);
               }
            }
            MatRef mid = oref(old_op, ux, dx);
            emit(new_op, k, U, Dn, ux, dx, &mid, alpha);
         }
      }
   }

   // ============================================================================ TensorL::create (TensorL.cpp:41-206)
   void create_L(int new_op) {
      const OpTensor& t = new_set.ops[new_op];
      const OpLayout& l = *t.lay;
      for (int k = 0; k < l.nkappa(); k++) {
         Sec U, Dn;
         block_secs(t, k, U, Dn);
         if (mr) {
            emit(new_op, k, U, Dn, U, U, nullptr, 1.0);                                                               // geval 0 (:84-107)
            const Sec s{U.n - 1, Dn.ts, Dn.ir};                                                                      // geval 1
            emit(new_op, k, U, Dn, s, s, nullptr, phase(Dn.ts - U.ts + 1) * std::sqrt((U.ts + 1.0) / (Dn.ts + 1)));
         } else {
            emit(new_op, k, U, Dn, Dn, Dn, nullptr, 1.0);                                                             // geval 0 (:154-177)
            const Sec s{U.n + 2, U.ts, U.ir};                                                                        // geval 1
            emit(new_op, k, U, Dn, s, s, nullptr, phase(U.ts - Dn.ts + 1) * std::sqrt((U.ts + 1.0) / (Dn.ts + 1)));
         }
      }
   }

   // ============================================================================ two-operator tensors with both operators on the new site
   // TensorS0::makenew(T) (TensorS0.cpp:53-98), TensorF0::makenew(T) (TensorF0.cpp:53-137), TensorF1::makenew(T) (TensorF1.cpp:54-117)
   void makenew_site(int new_op) {
      const OpTensor& t = new_set.ops[new_op];
      const OpLayout& l = *t.lay;
      const double s2 = std::sqrt(2.0);
      for (int k = 0; k < l.nkappa(); k++) {
         Sec U, Dn;
         block_secs(t, k, U, Dn);
         if (t.kind == K_S0) {
            const Sec s = mr ? U : Dn;   // right: old sector = up sector (site empty -> doubly occupied); left: old = down sector
            emit(new_op, k, U, Dn, s, s, nullptr, s2);
         } else if (t.kind == K_F0) {
            for (int geval = 0; geval < 3; geval++) {
               const Sec s = (geval == 0) ? step(U, 2, 0) : step(U, 1, geval == 1 ? -1 : +1);
               if (s.ts < 0) continue;
               double alpha = (geval == 0) ? s2 : 0.5 * s2;
               if (!mr && geval >= 1) alpha = s2 * 0.5 * (s.ts + 1.0) / (U.ts + 1.0);
               emit(new_op, k, U, Dn, s, s, nullptr, alpha);
            }
         } else if (t.kind == K_F1) {
            for (int geval = 0; geval < 2; geval++) {
               const Sec s = step(U, 1, geval == 0 ? -1 : +1);
               if (s.ts < 0 || std::abs(Dn.ts - s.ts) >= 2) continue;
               double alpha;
               if (mr) alpha = phase(s.ts + Dn.ts + 3) * std::sqrt(3.0 * (U.ts + 1)) * wigner6j(1, 1, 2, U.ts, Dn.ts, s.ts);
               else alpha = phase(Dn.ts + s.ts + 1) * std::sqrt(3.0 / (U.ts + 1.0)) * (s.ts + 1) * wigner6j(1, 1, 2, U.ts, Dn.ts, s.ts);
               emit(new_op, k, U, Dn, s, s, nullptr, alpha);
            }
         }
      }
   }

   // ============================================================================ two-operator tensors: one operator inside (old L), one on the new site
   // TensorS0::makenew(L,T) (TensorS0.cpp:100-245), TensorS1 (TensorS1.cpp:46-194), TensorF0 (TensorF0.cpp:139-284), TensorF1 (TensorF1.cpp:119-268)
   void makenew_L(int new_op, int old_L) {
      if (old_L < 0 || !old_set) return;
      const OpTensor& t = new_set.ops[new_op];
      const OpLayout& l = *t.lay;
      const int Lirr = old_set->ops[old_L].irrep;
      const bool pairing = (t.kind == K_S0 || t.kind == K_S1);   // a+ a+ (n_elec 2) vs a+ a (n_elec 0)
      const bool spin1 = (t.kind == K_S1 || t.kind == K_F1);
      for (int k = 0; k < l.nkappa(); k++) {
         Sec U, Dn;
         block_secs(t, k, U, Dn);
         for (int geval = 0; geval < 4; geval++) {
            // (ux -> dx) is a block of the old L operator: dx has one electron more than ux
            Sec ux, dx;
            const int sg = (geval % 2 == 0) ? -1 : +1;
            if (mr) {
               if (pairing) {
                  if (geval <= 1) { ux = U; dx = Sec{U.n + 1, Dn.ts + sg, xorp(U.ir, Lirr)}; }                     // site creates the 2nd electron in the down state
                  else { ux = Sec{U.n - 1, U.ts + sg, xorp(U.ir, site_irr)}; dx = Sec{U.n, Dn.ts, Dn.ir}; }
               } else {
                  if (geval <= 1) { ux = Sec{U.n - 1, U.ts + sg, xorp(U.ir, site_irr)}; dx = Sec{U.n, Dn.ts, Dn.ir}; }
                  else { ux = Sec{U.n - 2, U.ts, U.ir}; dx = Sec{U.n - 1, Dn.ts + sg, xorp(U.ir, Lirr)}; }
               }
            } else {
               if (pairing) {
                  if (geval <= 1) { ux = Sec{U.n + 1, U.ts + sg, xorp(U.ir, site_irr)}; dx = Sec{U.n + 2, Dn.ts, Dn.ir}; }
                  else { ux = Sec{U.n + 2, U.ts, U.ir}; dx = Sec{U.n + 3, Dn.ts + sg, xorp(U.ir, Lirr)}; }
               } else {
                  if (geval <= 1) { ux = U; dx = Sec{U.n + 1, Dn.ts + sg, xorp(U.ir, Lirr)}; }
                  else { ux = Sec{U.n + 1, U.ts + sg, xorp(U.ir, site_irr)}; dx = Sec{U.n + 2, Dn.ts, Dn.ir}; }
               }
            }
            if (ux.ts < 0 || dx.ts < 0) continue;
            if (spin1 && std::abs(ux.ts - dx.ts) >= 2) continue;
            double alpha = 0.0;
            const int su = U.ts, sd = Dn.ts;
            if (t.kind == K_S0) {
               if (mr) alpha = (geval <= 1) ? phase(su - dx.ts + 1) * std::sqrt(0.5 * (dx.ts + 1.0) / (su + 1.0)) : -std::sqrt(0.5);
               else alpha = (geval <= 1) ? phase(su - ux.ts + 1) * std::sqrt(0.5 * (ux.ts + 1.0) / (su + 1.0)) : -std::sqrt(0.5) * (dx.ts + 1.0) / (su + 1.0);
            } else if (t.kind == K_S1) {
               if (mr) alpha = (geval <= 1) ? phase(su + sd + 2) * std::sqrt(3.0 * (dx.ts + 1)) * wigner6j(1, 1, 2, su, sd, dx.ts)
                                            : phase(ux.ts + sd + 1) * std::sqrt(3.0 * (su + 1)) * wigner6j(1, 1, 2, su, sd, ux.ts);
               else alpha = (geval <= 1) ? phase(su + sd + 2) * std::sqrt(3.0 * (ux.ts + 1)) * wigner6j(1, 1, 2, su, sd, ux.ts)
                                         : phase(su + dx.ts + 1) * std::sqrt(3.0 / (sd + 1.0)) * (dx.ts + 1) * wigner6j(1, 1, 2, su, sd, dx.ts);
            } else if (t.kind == K_F0) {
               if (mr) alpha = (geval <= 1) ? std::sqrt(0.5) : phase(su + 1 - dx.ts) * std::sqrt(0.5 * (dx.ts + 1.0) / (su + 1.0));
               else alpha = (geval <= 1) ? std::sqrt(0.5) * (dx.ts + 1.0) / (su + 1.0) : phase(su - ux.ts + 1) * std::sqrt(0.5 * (ux.ts + 1.0) / (su + 1.0));
            } else {   // K_F1
               if (mr) alpha = (geval <= 1) ? phase(ux.ts + sd + 3) * std::sqrt(3.0 * (su + 1)) * wigner6j(1, 1, 2, su, sd, ux.ts)
                                            : phase(su + sd + 2) * std::sqrt(3.0 * (dx.ts + 1)) * wigner6j(1, 1, 2, su, sd, dx.ts);
               else alpha = (geval <= 1) ? phase(sd + dx.ts + 1) * std::sqrt(3.0 / (su + 1.0)) * (dx.ts + 1) * wigner6j(1, 1, 2, su, sd, dx.ts)
                                         : ((su % 2) != 0 ? -1.0 : 1.0) * std::sqrt(3.0 * (ux.ts + 1.0) * (sd + 1.0) / (su + 1.0)) * wigner6j(1, 1, 2, su, sd, ux.ts);
            }
            MatRef mid = oref(old_L, ux, dx);
            emit(new_op, k, U, Dn, ux, dx, &mid, alpha);
         }
      }
   }

   // ============================================================================ TensorGYZ::construct (TensorGYZ.cpp:42-98), TensorKM::construct (TensorKM.cpp:42-95)
   // site-local helper operators of the two-orbital correlation functions, built from T alone (moving right only):
   //   new[U -> Dn] = sum alpha * T[s -> U]^T * T[s -> Dn]
   void construct_corr(int new_op) {
      const OpTensor& t = new_set.ops[new_op];
      const OpLayout& l = *t.lay;
      for (int k = 0; k < l.nkappa(); k++) {
         Sec U, Dn;
         block_secs(t, k, U, Dn);
         switch (t.kind) {
            case K_Y: emit(new_op, k, U, Dn, U, U, nullptr, 1.0); break;                                        // site empty
            case K_Z: { const Sec s{U.n - 2, U.ts, U.ir}; emit(new_op, k, U, Dn, s, s, nullptr, 1.0); break; }  // site doubly occupied
            case K_G:                                                                                           // site singly occupied
               for (int dts = -1; dts <= 1; dts += 2) {
                  const Sec s{U.n - 1, U.ts + dts, xorp(U.ir, site_irr)};
                  if (s.ts >= 0) emit(new_op, k, U, Dn, s, s, nullptr, std::sqrt(0.5));
               }
               break;
            case K_K: emit(new_op, k, U, Dn, U, U, nullptr, 1.0); break;                                        // <empty| ... |single>
            case K_M: {                                                                                         // <single| ... |double>
               const Sec s{U.n - 1, Dn.ts, Dn.ir};
               const int fase = ((((Dn.ts - U.ts + 1) / 2) % 2) != 0) ? -1 : 1;
               emit(new_op, k, U, Dn, s, s, nullptr, fase * std::sqrt((U.ts + 1.0) / (Dn.ts + 1)));
               break;
            }
            default: break;
         }
      }
   }

#include "b2_update_plan_qx.inc"

   // ============================================================================ orchestration (DMRGoperators.cpp:243-907)
   void run_one(int n) {
      const int i = ix;
      auto oldf = [&](int kind, int a, int b) { return old_set ? old_set->find(kind, a, b) : -1; };
      {
         const OpTensor& t = new_set.ops[n];
         switch (t.kind) {
            case K_L:
               if (t.i == i) create_L(n); else generic_update(n, oldf(K_L, t.i, t.i));
               break;
            case K_S0: case K_S1: case K_F0: case K_F1: {
               const bool has_site = (t.i == i || t.j == i);
               if (t.i == i && t.j == i) makenew_site(n);
               else if (has_site) { const int other = (t.i == i) ? t.j : t.i; makenew_L(n, oldf(K_L, other, other)); }
               else generic_update(n, oldf(t.kind, t.i, t.j));
               break;
            }
            case K_A: case K_B: case K_C: case K_D:
               generic_update(n, oldf(t.kind, t.i, t.j));   // absent at the chain end: the tensor starts from zero (:346-360); the mixing
               break;                                       // with the two-operator tensors follows in build_update_plan (sequential, shared temps)
            case K_Q: update_Q(n); break;
            case K_X: update_X(n); break;
            case K_G: case K_Y: case K_Z: case K_K: case K_M:   // DMRG::update_correlations_tensors (DMRGoperators3RDM.cpp:415-479)
               if (t.i == i) construct_corr(n); else generic_update(n, oldf(t.kind, t.i, t.i));
               break;
         }
      }
   }

   // A/B/C/D += integral-weighted two-operator tensors that have one leg on the new site (DMRGoperators.cpp:367-405 right, :700-738 left)
   void mix_complementary(int n) {
      const OpTensor& t = new_set.ops[n];
      const int s1 = t.i, s2 = t.j, i = ix;
      const bool diag = (s1 == s2);
      const int irr_prod = xorp(irr(s1), irr(s2));
      struct Part { int src; double coef; bool tr; };
      std::vector<Part> parts;
      const int lo = mr ? 0 : i, hi = mr ? i : L - 1;   // inside sites: partner `o` of the new site i
      for (int o = lo; o <= hi; o++) {
         const int a = std::min(o, i), b = std::max(o, i);
         if (xorp(irr(a), irr(b)) != irr_prod) continue;
         const bool same = (o == i);
         // A, B: right gMxElement(a, b, s1, s2) (:375-383), left gMxElement(s1, s2, a, b) (:704-712); <pq|rs> = <qp|sr>
         auto M = [&](int p, int q, int r, int s) { return mr ? V(p, q, r, s) : V(r, s, p, q); };
         if (t.kind == K_A) {
            double alpha = M(a, b, s1, s2);
            if (diag && same) alpha *= 0.5;
            if (!diag && !same) alpha += M(a, b, s2, s1);
            parts.push_back({new_set.find(K_S0, a, b), alpha, false});
         } else if (t.kind == K_B) {
            if (!same && !diag) parts.push_back({new_set.find(K_S1, a, b), M(a, b, s1, s2) - M(a, b, s2, s1), false});
         } else {
            // C, D: right gMxElement(o, s1, i, s2) ... (:388-403); left gMxElement(s1, i, s2, o) ... (:721-736)
            double c_plain, c_plain_x, c_tr, c_tr_x;
            if (mr) { c_plain = V(a, s1, b, s2); c_plain_x = V(a, s1, s2, b); c_tr = V(a, s2, b, s1); c_tr_x = V(a, s2, s1, b); }
            else { c_plain = V(s1, a, s2, b); c_plain_x = V(s1, a, b, s2); c_tr = V(s1, b, s2, a); c_tr_x = V(s1, b, a, s2); }
            const int src = new_set.find(t.kind == K_C ? K_F0 : K_F1, a, b);
            parts.push_back({src, t.kind == K_C ? 2 * c_plain - c_plain_x : -c_plain_x, false});
            if (!same) parts.push_back({src, t.kind == K_C ? 2 * c_tr - c_tr_x : -c_tr_x, true});
         }
      }
      const OpLayout& l = *t.lay;
      for (const Part& p : parts) {
         if (p.src < 0 || p.coef == 0.0) continue;
         const OpTensor& so = new_set.ops[p.src];
         if (!p.tr) {   // TensorOperator::daxpy (:407-414): identical layouts, element by element
            assert(so.lay.get() == t.lay.get());
            plan.mix_flat.push_back(UpdatePlan::MixFlat{n, p.src, -1, p.coef});
            plan.flops_ref += 2.0 * (double)l.size;
            continue;
         }
         // daxpy_transpose_tensorCD (:416-455): transposed blocks with a block-dependent spin factor -> a transposed copy in this layout
         int temp = -1;
         for (size_t i = 0; i < plan.mix_temps.size(); i++)
            if (plan.mix_temps[i].src_op == p.src && plan.mix_temps[i].lay == t.lay.get()) { temp = (int)i; break; }
         if (temp < 0) {
            UpdatePlan::MixTemp mt{p.src, t.lay.get(), plan.presum_size, l.size};
            plan.presum_size += (l.size + 15) / 16 * 16;
            temp = (int)plan.mix_temps.size();
            plan.mix_temps.push_back(mt);
            for (int k = 0; k < l.nkappa(); k++) {
               Sec U, Dn;
               block_secs(t, k, U, Dn);
               const int sk = so.lay->kappa(bk, Dn.n, Dn.ts, Dn.ir, U.n, U.ts, U.ir);
               if (sk < 0) continue;
               Term3 x;
               x.dst = (int)plan.mix_dst.size();
               plan.mix_dst.push_back(DstBlock{mt.off + l.blk[k].off, l.blk[k].rows, l.blk[k].cols});
               x.f = (U.ts != Dn.ts) ? phase(U.ts - Dn.ts) * std::sqrt(mr ? ((U.ts + 1.0) / (Dn.ts + 1)) : ((Dn.ts + 1.0) / (U.ts + 1))) : 1.0;
               x.q.space = SP_VOUT; x.q.off = so.off + so.lay->blk[sk].off; x.q.rows = so.lay->blk[sk].rows; x.q.cols = so.lay->blk[sk].cols; x.q.trans = 1;
               plan.mix_terms.push_back(x);
            }
         }
         plan.mix_flat.push_back(UpdatePlan::MixFlat{n, p.src, temp, p.coef});
         for (int k = 0; k < l.nkappa(); k++) {   // the reference's daxpy count: the blocks whose transposed partner exists
            Sec U, Dn;
            block_secs(t, k, U, Dn);
            if (so.lay->kappa(bk, Dn.n, Dn.ts, Dn.ir, U.n, U.ts, U.ir) >= 0) plan.flops_ref += 2.0 * l.blk[k].rows * l.blk[k].cols;
         }
      }
   }
};

}   // namespace

void build_update_plan(UpdatePlan& plan, const Bookkeeper& bk, const Problem& prob, const OpSet* old_set, const OpSet& new_set, int index,
                       bool moving_right) {
   plan = UpdatePlan();
   plan.index = index; plan.moving_right = moving_right;
   plan.T.build(bk, index);
   plan.block_base.resize(new_set.ops.size());
   for (size_t n = 0; n < new_set.ops.size(); n++) {
      const OpTensor& t = new_set.ops[n];
      plan.block_base[n] = (int)plan.dst.size();
      for (const Block& b : t.lay->blk) plan.dst.push_back(DstBlock{t.off + b.off, b.rows, b.cols});
   }
   const int nops = (int)new_set.ops.size();
   const int nthreads = std::min(plan_threads((int)plan.dst.size()), std::max(1, nops / 4));
   if (nthreads <= 1) {
      UGen g(plan, plan, bk, prob, old_set, new_set, index, moving_right);
      for (int n = 0; n < nops; n++) g.run_one(n);
      for (int n = 0; n < nops; n++)
         if (new_set.ops[n].kind >= K_A && new_set.ops[n].kind <= K_D) g.mix_complementary(n);
      return;
   }
   // New operators are independent: every host thread enumerates the operators it grabs into a private fragment; the fragments are
   // stitched together in operator order (pre-sum arena offsets shifted per operator), so the plan has exactly the terms, the term
   // order and the pre-sum layout of the sequential enumeration.
   struct Span { size_t t0, t1, m0, m1, p0, p1; int64_t a0, a1; };   // terms, mixing terms, pre-sums, pre-sum arena range of one operator
   struct Frag { UpdatePlan plan; std::vector<std::pair<int, Span>> spans; };
   std::vector<Frag> frags(nthreads);
   std::atomic<int> next{0};
   parallel_run(nthreads, [&](int t) {
      Frag& f = frags[t];
      UGen g(f.plan, plan, bk, prob, old_set, new_set, index, moving_right);
      for (;;) {
         const int n = next.fetch_add(1);
         if (n >= nops) break;
         Span sp{f.plan.terms.size(), 0, f.plan.mix_terms.size(), 0, f.plan.presums.size(), 0, f.plan.presum_size, 0};
         g.run_one(n);
         sp.t1 = f.plan.terms.size(); sp.m1 = f.plan.mix_terms.size(); sp.p1 = f.plan.presums.size(); sp.a1 = f.plan.presum_size;
         f.spans.push_back({n, sp});
      }
   });
   std::vector<std::pair<int, int>> where(nops, {-1, -1});
   size_t nterms = 0, nmix = 0, npre = 0;
   for (int t = 0; t < nthreads; t++) {
      for (size_t i = 0; i < frags[t].spans.size(); i++) where[frags[t].spans[i].first] = {t, (int)i};
      nterms += frags[t].plan.terms.size(); nmix += frags[t].plan.mix_terms.size(); npre += frags[t].plan.presums.size();
      plan.flops_ref += frags[t].plan.flops_ref;
   }
   plan.terms.reserve(nterms); plan.mix_terms.reserve(nmix); plan.presums.reserve(npre);
   for (int n = 0; n < nops; n++) {
      const Frag& f = frags[where[n].first];
      const Span& sp = f.spans[where[n].second].second;
      const int64_t shift = plan.presum_size - sp.a0;
      for (size_t i = sp.p0; i < sp.p1; i++) { Presum p = f.plan.presums[i]; p.off += shift; plan.presums.push_back(std::move(p)); }
      plan.presum_size += sp.a1 - sp.a0;
      for (size_t i = sp.t0; i < sp.t1; i++) {
         Term3 x = f.plan.terms[i];
         if (x.p.space == SP_PRESUM) x.p.off += shift;
         if (x.q.space == SP_PRESUM) x.q.off += shift;
         if (x.r.space == SP_PRESUM) x.r.off += shift;
         plan.terms.push_back(x);
      }
   }
   // the mixing lists are small (O(L^2) whole-operator axpys + O(L) transposed copies): enumerated here, after the pre-sum arena is final
   UGen g(plan, plan, bk, prob, old_set, new_set, index, moving_right);
   for (int n = 0; n < nops; n++)
      if (new_set.ops[n].kind >= K_A && new_set.ops[n].kind <= K_D) g.mix_complementary(n);
}

}   // namespace b2
