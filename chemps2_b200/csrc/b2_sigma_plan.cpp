// b2_sigma_plan.cpp — enumerate every term of sigma = H_eff * S for one site pair (see b2_sigma.h).
//
// Each function below restates one addDiagram* family of the reference as a list of SigmaTerms; the spin-recoupling
// prefactors (phases, sqrt(2j+1) factors, Wigner 6j/9j symbols) are evaluated here, once per site, and baked into
// the term.  Reference locations are cited per function.  Notation follows the reference: the target block kappa has
// labels (NL,TwoSL,IL | N1,N2,TwoJ | NR,TwoSR,IR); "d"-suffixed labels belong to the source block.
#include <cassert>
#include <cmath>
#include <cstdio>
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <map>
#include <thread>

#include <chrono>

#include "b2_sigma.h"

namespace b2 {

// host threads used to build plans: B2_PLAN_THREADS, else the hardware concurrency (at most 32) divided by the number of processes that
// share this host (one process per GPU: every rank builds its plans at the same time — 8 ranks x 16 threads on 16 cores thrash);
// small plans stay sequential
static std::atomic<int> g_local_ranks{1};
void set_plan_local_ranks(int n) { g_local_ranks.store(std::max(1, n)); }
int plan_threads(int nblocks) {
   static const int configured = [] {
      const char* e = getenv("B2_PLAN_THREADS");
      int n = e ? atoi(e) : (int)std::thread::hardware_concurrency();
      return std::max(1, std::min(n, 32));
   }();
   const int share = std::max(1, configured / g_local_ranks.load());
   return std::max(1, std::min(share, nblocks / 8));
}

namespace {

struct OpH { int8_t src; int op; };   // handle to an operator (left set / right set / presum)

struct Gen {
   SigmaPlan& plan;
   const Bookkeeper& bk;
   const Problem& prob;
   const OpSet* left;
   const OpSet* right;
   const int ix, L, world;
   const SLayout& S;
   std::map<std::string, int> presum_index;
   // labels of the current target block
   int k = 0, NL = 0, TwoSL = 0, IL = 0, N1 = 0, N2 = 0, TwoJ = 0, NR = 0, TwoSR = 0, IR = 0, dimL = 0, dimR = 0;

   Gen(SigmaPlan& p, const Bookkeeper& b, const Problem& pr, const OpSet* l, const OpSet* r, int site, int w)
       : plan(p), bk(b), prob(pr), left(l), right(r), ix(site), L(b.L), world(w), S(p.S) {}

   int irr(int orb) const { return bk.orb_irrep[orb]; }
   double V(int a, int b, int c, int d) const { return prob.V(a, b, c, d); }

   // ---- ownership GROUPS: the arguments of the reference's owner maps before the modulo (MPIchemps2.h:158-231: every operator pair /
   // Q site / specific diagram has its own number; the reference takes it modulo mpi_size()).  A term carries its group number here;
   // balance_owners() below assigns the groups to the GPUs by their FLOPs instead of by the modulo (static round-robin ignores that the
   // pair sums of different operator pairs cost very different amounts: max/mean 1.06 at 8 GPUs).
   int o_master() const { return 0; }
   int o_absigma(int i, int j) const { return 1 + i + (j * (j + 1)) / 2; }
   int o_cdf(int i, int j) const { return 1 + (L * (L + 1)) / 2 + i + (j * (j + 1)) / 2; }
   int o_q(int i) const { return 1 + L * (L + 1) + i; }
   int o_spec(int macro) const { return macro + L * (L + 2); }

   OpH Lo(int kind, int i, int j) const { return OpH{SRC_LEFT, left ? left->find(kind, i, j) : -1}; }
   OpH Ro(int kind, int i, int j) const { return OpH{SRC_RIGHT, right ? right->find(kind, i, j) : -1}; }

   const OpLayout* layout_of(OpH h, int& side) const {
      if (h.op < 0) return nullptr;
      if (h.src == SRC_LEFT) { side = SRC_LEFT; return left->ops[h.op].lay.get(); }
      if (h.src == SRC_RIGHT) { side = SRC_RIGHT; return right->ops[h.op].lay.get(); }
      side = plan.presums[h.op].side;
      return plan.presums[h.op].lay.get();
   }

   // Pre-summed operator  base + sum_l coef_l * part_l  (all parts share one layout).  `key` de-duplicates.
   OpH presum(const std::string& key, int side, int base_op, const std::vector<std::pair<double, int>>& parts) {
      auto it = presum_index.find(key);
      if (it != presum_index.end()) return OpH{SRC_PRESUM, it->second};
      const OpSet* set = (side == SRC_LEFT) ? left : right;
      Presum p;
      p.side = side;
      if (base_op >= 0) p.parts.push_back({1.0, base_op});
      for (auto& pr : parts)
         if (pr.second >= 0) p.parts.push_back(pr);
      if (p.parts.empty()) return OpH{SRC_PRESUM, -1};
      p.lay = set->ops[p.parts[0].second].lay;
      p.off = plan.presum_size;
      plan.presum_size += (p.lay->size + 15) / 16 * 16;
      plan.presums.push_back(p);
      presum_index[key] = (int)plan.presums.size() - 1;
      return OpH{SRC_PRESUM, (int)plan.presums.size() - 1};
   }

   // block of a left operator between the target's left sector and (nls,tsls,ils); trans=true: stored src->dst
   BRef lref(OpH h, bool trans, int nls, int tsls, int ils) const {
      BRef r;
      int side = 0;
      const OpLayout* lay = layout_of(h, side);
      if (!lay) return r;
      r.src = h.src; r.op = h.op; r.trans = trans;
      r.blk = trans ? lay->kappa(bk, nls, tsls, ils, NL, TwoSL, IL) : lay->kappa(bk, NL, TwoSL, IL, nls, tsls, ils);
      return r;
   }
   // block of a right operator; trans=false: stored src->dst (enters as is), trans=true: stored dst->src
   BRef rref(OpH h, bool trans, int nrs, int tsrs, int irs) const {
      BRef r;
      int side = 0;
      const OpLayout* lay = layout_of(h, side);
      if (!lay) return r;
      r.src = h.src; r.op = h.op; r.trans = trans;
      r.blk = trans ? lay->kappa(bk, NR, TwoSR, IR, nrs, tsrs, irs) : lay->kappa(bk, nrs, tsrs, irs, NR, TwoSR, IR);
      return r;
   }

   void push(int src, const BRef& l, const BRef& r, double f, int owner) {
      if (f == 0.0) { plan.skipped_zero++; return; }
      SigmaTerm t;
      t.dst = k; t.src = src; t.l = l; t.r = r; t.factor = f; t.owner = owner;
      plan.terms.push_back(t);
   }
   // sigma[k] += f * op(A) * S[src]            (left operator only; right labels unchanged)
   void L1(OpH a, bool trans, int nls, int tsls, int ils, int n1, int n2, int tj, double f, int owner) {
      if (a.op < 0) return;
      const int src = S.kappa(bk, nls, tsls, ils, n1, n2, tj, NR, TwoSR, IR);
      if (src < 0) return;
      BRef l = lref(a, trans, nls, tsls, ils);
      if (l.blk < 0) return;
      plan.flops_ref += 2.0 * dimL * dimR * S.blk[src].rows;
      push(src, l, BRef(), f, owner);
   }
   // sigma[k] += f * S[src] * op(B)            (right operator only; left labels unchanged)
   void R1(OpH b, bool trans, int nrs, int tsrs, int irs, int n1, int n2, int tj, double f, int owner) {
      if (b.op < 0) return;
      const int src = S.kappa(bk, NL, TwoSL, IL, n1, n2, tj, nrs, tsrs, irs);
      if (src < 0) return;
      BRef r = rref(b, trans, nrs, tsrs, irs);
      if (r.blk < 0) return;
      plan.flops_ref += 2.0 * dimL * dimR * S.blk[src].cols;
      push(src, BRef(), r, f, owner);
   }
   // sigma[k] += f * op(A) * S[src] * op(B)
   void LR(OpH a, bool lt, int nls, int tsls, int ils, OpH b, bool rt, int nrs, int tsrs, int irs, int n1, int n2, int tj,
           double f, int owner, bool right_first = false) {
      if (a.op < 0 || b.op < 0) return;
      const int src = S.kappa(bk, nls, tsls, ils, n1, n2, tj, nrs, tsrs, irs);
      if (src < 0) return;
      BRef l = lref(a, lt, nls, tsls, ils), r = rref(b, rt, nrs, tsrs, irs);
      if (l.blk < 0 || r.blk < 0) return;
      const double dLs = S.blk[src].rows, dRs = S.blk[src].cols;
      // the reference multiplies left-first (temp = A*S, then temp*B) except where noted (right_first)
      plan.flops_ref += right_first ? 2.0 * (dLs * dimR * dRs + dimL * dimR * dLs) : 2.0 * (dimL * dRs * dLs + dimL * dimR * dRs);
      push(src, l, r, f, owner);
   }
   // sigma[k] += f * S[src] with identical outer labels
   void Z(int n1, int n2, int tj, double f, int owner) {
      const int src = S.kappa(bk, NL, TwoSL, IL, n1, n2, tj, NR, TwoSR, IR);
      if (src < 0) return;
      plan.flops_ref += 2.0 * dimL * dimR;
      push(src, BRef(), BRef(), f, owner);
   }

   // valid coupled spins of the site pair
   static int tj_lo(int n1, int n2) { return (n1 + n2) % 2; }
   static int tj_hi(int n1, int n2) { return (n1 == 1 && n2 == 1) ? 2 : (n1 + n2) % 2; }

   void set_block(int kappa) {
      k = kappa;
      NL = S.NL[k]; TwoSL = S.twoSL[k]; IL = S.IL[k]; N1 = S.N1[k]; N2 = S.N2[k]; TwoJ = S.twoJ[k];
      NR = S.NR[k]; TwoSR = S.twoSR[k]; IR = S.IR[k];
      dimL = S.blk[k].rows; dimR = S.blk[k].cols;
   }

   // ============================================================================================ family 1
   // HeffDiagrams1.cpp:26-63
   void d1() {
      if (left) L1(Lo(K_X, -1, -1), false, NL, TwoSL, IL, N1, N2, TwoJ, 1.0, 0);              // 1A: X_L * S
      if (right) R1(Ro(K_X, -1, -1), true, NR, TwoSR, IR, N1, N2, TwoJ, 1.0, 0);             // 1B: S * X_R^T
      if (N1 == 2) Z(N1, N2, TwoJ, V(ix, ix, ix, ix), o_master());                           // 1C
      if (N2 == 2) Z(N1, N2, TwoJ, V(ix + 1, ix + 1, ix + 1, ix + 1), o_master());           // 1D
   }

   // ============================================================================================ family 2 (site-local)
   // HeffDiagrams2.cpp:996-1058
   void d2d() {
      const int i = ix, j = ix + 1;
      if (N1 == 2 && N2 == 0) Z(0, 2, 0, V(i, i, j, j), o_master());
      if (N1 == 0 && N2 == 2) Z(2, 0, 0, V(i, i, j, j), o_master());
      if (N1 == 2 && N2 == 2) Z(2, 2, 0, 4 * V(i, j, i, j) - 2 * V(i, j, j, i), o_master());
      if (N1 == 1 && N2 == 1) Z(1, 1, TwoJ, V(i, j, i, j) + ((TwoJ == 0) ? 1 : -1) * V(i, j, j, i), o_master());
      if (N1 == 2 && N2 == 1) Z(2, 1, 1, 2 * V(i, j, i, j) - V(i, j, j, i), o_master());
      if (N1 == 1 && N2 == 2) Z(1, 2, 1, 2 * V(i, j, i, j) - V(i, j, j, i), o_master());
   }
   // HeffDiagrams2.cpp:850-994 (left A on one site), :1060-1184 (right A)
   void d2bcef() {
      const double s2 = std::sqrt(2.0);
      if (left) {
         OpH Ai = Lo(K_A, ix, ix), Aj = Lo(K_A, ix + 1, ix + 1);
         if (N1 == 0) L1(Ai, true, NL - 2, TwoSL, IL, 2, N2, TwoJ, s2, o_absigma(ix, ix));               // 2b1
         if (N1 == 2) L1(Ai, false, NL + 2, TwoSL, IL, 0, N2, TwoJ, s2, o_absigma(ix, ix));              // 2b2
         if (N2 == 0) L1(Aj, true, NL - 2, TwoSL, IL, N1, 2, TwoJ, s2, o_absigma(ix + 1, ix + 1));       // 2c1
         if (N2 == 2) L1(Aj, false, NL + 2, TwoSL, IL, N1, 0, TwoJ, s2, o_absigma(ix + 1, ix + 1));      // 2c2
      }
      if (right) {
         OpH Ai = Ro(K_A, ix, ix), Aj = Ro(K_A, ix + 1, ix + 1);
         if (N1 == 2) R1(Ai, false, NR - 2, TwoSR, IR, 0, N2, TwoJ, s2, o_absigma(ix, ix));              // 2e1
         if (N1 == 0) R1(Ai, true, NR + 2, TwoSR, IR, 2, N2, TwoJ, s2, o_absigma(ix, ix));               // 2e2
         if (N2 == 2) R1(Aj, false, NR - 2, TwoSR, IR, N1, 0, TwoJ, s2, o_absigma(ix + 1, ix + 1));      // 2f1
         if (N2 == 0) R1(Aj, true, NR + 2, TwoSR, IR, N1, 2, TwoJ, s2, o_absigma(ix + 1, ix + 1));       // 2f2
      }
   }
   // HeffDiagrams2.cpp:1186-1294 (C, spin 0) and :1296-1511 (D, spin 1)
   void d2bcef3() {
      const double s2 = std::sqrt(2.0);
      const int TwoS1 = (N1 == 1) ? 1 : 0, TwoS2 = (N2 == 1) ? 1 : 0;
      if (left) {
         if (N1 != 0) L1(Lo(K_C, ix, ix), true, NL, TwoSL, IL, N1, N2, TwoJ, ((N1 == 2) ? 1.0 : 0.5) * s2, o_cdf(ix, ix));              // 2b3 spin0
         if (N2 != 0) L1(Lo(K_C, ix + 1, ix + 1), true, NL, TwoSL, IL, N1, N2, TwoJ, ((N2 == 2) ? 1.0 : 0.5) * s2, o_cdf(ix + 1, ix + 1)); // 2c3 spin0
         for (int TwoSLd = TwoSL - 2; TwoSLd <= TwoSL + 2; TwoSLd += 2) {
            if (TwoSLd < 0) continue;
            if (N1 == 1)   // 2b3 spin1
               for (int tjd = tj_lo(N1, N2); tjd <= tj_hi(N1, N2); tjd += 2) {
                  if (std::abs(TwoSLd - TwoSR) > tjd) continue;
                  const double f = phase(TwoSLd + TwoSR + TwoJ + TwoS2 + tjd - 1) * std::sqrt(3.0 * (TwoJ + 1) * (tjd + 1) * (TwoSL + 1)) *
                                   wigner6j(tjd, TwoJ, 2, 1, 1, TwoS2) * wigner6j(tjd, TwoJ, 2, TwoSL, TwoSLd, TwoSR);
                  L1(Lo(K_D, ix, ix), true, NL, TwoSLd, IL, N1, N2, tjd, f, o_cdf(ix, ix));
               }
            if (N2 == 1)   // 2c3 spin1
               for (int tjd = tj_lo(N1, N2); tjd <= tj_hi(N1, N2); tjd += 2) {
                  if (std::abs(TwoSLd - TwoSR) > tjd) continue;
                  const double f = phase(TwoSLd + TwoSR + 2 * TwoJ + TwoS1 - 1) * std::sqrt(3.0 * (TwoJ + 1) * (tjd + 1) * (TwoSL + 1)) *
                                   wigner6j(tjd, TwoJ, 2, 1, 1, TwoS1) * wigner6j(tjd, TwoJ, 2, TwoSL, TwoSLd, TwoSR);
                  L1(Lo(K_D, ix + 1, ix + 1), true, NL, TwoSLd, IL, N1, N2, tjd, f, o_cdf(ix + 1, ix + 1));
               }
         }
      }
      if (right) {
         if (N1 != 0) R1(Ro(K_C, ix, ix), false, NR, TwoSR, IR, N1, N2, TwoJ, ((N1 == 2) ? 1.0 : 0.5) * s2, o_cdf(ix, ix));              // 2e3 spin0
         if (N2 != 0) R1(Ro(K_C, ix + 1, ix + 1), false, NR, TwoSR, IR, N1, N2, TwoJ, ((N2 == 2) ? 1.0 : 0.5) * s2, o_cdf(ix + 1, ix + 1)); // 2f3 spin0
         for (int TwoSRd = TwoSR - 2; TwoSRd <= TwoSR + 2; TwoSRd += 2) {
            if (TwoSRd < 0) continue;
            if (N1 == 1)   // 2e3 spin1
               for (int tjd = tj_lo(N1, N2); tjd <= tj_hi(N1, N2); tjd += 2) {
                  if (std::abs(TwoSL - TwoSRd) > tjd) continue;
                  const double f = phase(TwoSRd + TwoSL + 2 * TwoJ + TwoS2 + 1) * std::sqrt(3.0 * (TwoJ + 1) * (tjd + 1) * (TwoSRd + 1)) *
                                   wigner6j(tjd, TwoJ, 2, 1, 1, TwoS2) * wigner6j(tjd, TwoJ, 2, TwoSR, TwoSRd, TwoSL);
                  R1(Ro(K_D, ix, ix), false, NR, TwoSRd, IR, N1, N2, tjd, f, o_cdf(ix, ix));
               }
            if (N2 == 1)   // 2f3 spin1
               for (int tjd = tj_lo(N1, N2); tjd <= tj_hi(N1, N2); tjd += 2) {
                  if (std::abs(TwoSL - TwoSRd) > tjd) continue;
                  const double f = phase(TwoSRd + TwoSL + TwoJ + TwoS1 + tjd + 1) * std::sqrt(3.0 * (TwoJ + 1) * (tjd + 1) * (TwoSRd + 1)) *
                                   wigner6j(tjd, TwoJ, 2, 1, 1, TwoS1) * wigner6j(tjd, TwoJ, 2, TwoSR, TwoSRd, TwoSL);
                  R1(Ro(K_D, ix + 1, ix + 1), false, NR, TwoSRd, IR, N1, N2, tjd, f, o_cdf(ix + 1, ix + 1));
               }
         }
      }
   }

   // ============================================================================================ family 2a (pairs x complementary)
   // HeffDiagrams2.cpp:28-848.  The pair sum runs over the shorter side (leftSum, :50).
   void d2a() {
      const bool leftSum = (ix < L * 0.5);
      const int lo = leftSum ? 0 : ix + 2, hi = leftSum ? ix : L;   // pair sites in [lo, hi)
      // two-op on the summed side, complementary on the other side
      auto two = [&](int kind, int i, int j) { return leftSum ? Lo(kind, i, j) : Ro(kind, i, j); };
      auto cmp = [&](int kind, int i, int j) { return leftSum ? Ro(kind, i, j) : Lo(kind, i, j); };
      auto emit = [&](int kl, bool lt, int kr, bool rt, int i, int j, int dN, int tsld, int tsrd, double f, int owner) {
         // left operator kind kl / right operator kind kr as they sit on the left / right boundary
         OpH a = leftSum ? two(kl, i, j) : cmp(kl, i, j);
         OpH b = leftSum ? cmp(kr, i, j) : two(kr, i, j);
         if (a.op < 0 || b.op < 0) return;
         const int pirr = xorp(irr(i), irr(j));
         LR(a, lt, NL + dN, tsld, xorp(IL, pirr), b, rt, NR + dN, tsrd, xorp(IR, pirr), N1, N2, TwoJ, f, owner);
      };
      const int kS0 = K_S0, kS1 = K_S1, kF0 = K_F0, kF1 = K_F1;
      for (int i = lo; i < hi; i++)
         for (int j = i; j < hi; j++) {
            // 2a1 spin0 (:28-130): S0^T/A^T . S . A/S0        2a2 spin0 (:132-228): S0/A . S . A^T/S0^T
            emit(leftSum ? kS0 : K_A, true, leftSum ? K_A : kS0, false, i, j, -2, TwoSL, TwoSR, 1.0, o_absigma(i, j));
            emit(leftSum ? kS0 : K_A, false, leftSum ? K_A : kS0, true, i, j, +2, TwoSL, TwoSR, 1.0, o_absigma(i, j));
            // 2a3 spin0 (:476-649), second ordering includes i == j
            emit(leftSum ? kF0 : K_C, true, leftSum ? K_C : kF0, false, i, j, 0, TwoSL, TwoSR, 1.0, o_cdf(i, j));
            if (j > i) emit(leftSum ? kF0 : K_C, false, leftSum ? K_C : kF0, true, i, j, 0, TwoSL, TwoSR, 1.0, o_cdf(i, j));
         }
      for (int tsld = TwoSL - 2; tsld <= TwoSL + 2; tsld += 2)
         for (int tsrd = TwoSR - 2; tsrd <= TwoSR + 2; tsrd += 2) {
            if (tsld < 0 || tsrd < 0 || std::abs(tsld - tsrd) > TwoJ) continue;
            const double w = wigner6j(tsld, tsrd, TwoJ, TwoSR, TwoSL, 2);
            const double f_2a1 = phase(tsrd + TwoSL + TwoJ + 2) * std::sqrt((TwoSR + 1) * (TwoSL + 1.0)) * w;      // :261-262
            const double f_2a2 = phase(tsld + TwoSR + TwoJ + 2) * std::sqrt((tsrd + 1) * (tsld + 1.0)) * w;        // :384-385
            const double f_2a3a = phase(tsld + tsrd + TwoJ + 2) * std::sqrt((TwoSR + 1) * (tsld + 1.0)) * w;       // :682-683
            const double f_2a3b = phase(TwoSL + TwoSR + TwoJ + 2) * std::sqrt((tsrd + 1) * (TwoSL + 1.0)) * w;     // :720-721
            for (int i = lo; i < hi; i++)
               for (int j = i; j < hi; j++) {
                  if (j > i) {
                     emit(leftSum ? kS1 : K_B, true, leftSum ? K_B : kS1, false, i, j, -2, tsld, tsrd, f_2a1, o_absigma(i, j));
                     emit(leftSum ? kS1 : K_B, false, leftSum ? K_B : kS1, true, i, j, +2, tsld, tsrd, f_2a2, o_absigma(i, j));
                     emit(leftSum ? kF1 : K_D, false, leftSum ? K_D : kF1, true, i, j, 0, tsld, tsrd, f_2a3a, o_cdf(i, j));
                  }
                  emit(leftSum ? kF1 : K_D, true, leftSum ? K_D : kF1, false, i, j, 0, tsld, tsrd, f_2a3b, o_cdf(i, j));
               }
         }
   }

   // ============================================================================================ family 3
   // 3A+3D / 3B+3I (HeffDiagrams3.cpp:28-292): Q_left(site) (+ on-site L pre-sum) acting with a site creator/annihilator.
   // 3K+3F / 3L+3G (:476-740): mirror on the right.
   void d3_onesided() {
      const int TwoS1 = (N1 == 1) ? 1 : 0, TwoS2 = (N2 == 1) ? 1 : 0;
      if (left) {
         for (int which = 0; which < 2; which++) {   // 0: site ix (3A/3D), 1: site ix+1 (3B/3I)
            const int s = ix + which;
            OpH Q = Lo(K_Q, s, s);
            if (Q.op < 0) continue;
            std::vector<std::pair<double, int>> parts;
            for (int l = 0; l < ix; l++)
               if (irr(l) == irr(s)) parts.push_back({V(l, s, s, s), left->find(K_L, l, l)});   // :69-75, :202-208
            OpH Qp = presum(which == 0 ? "3A" : "3B", SRC_LEFT, Q.op, parts);
            const int ILd = xorp(IL, irr(s));
            const int Ns = which == 0 ? N1 : N2;            // occupation of the active site
            const int TwoSo = which == 0 ? TwoS2 : TwoS1;   // spin of the spectator site
            const int own = o_q(s);
            for (int tsld = TwoSL - 1; tsld <= TwoSL + 1; tsld += 2) {
               if (tsld < 0) continue;
               auto n1n2 = [&](int ns, int& n1, int& n2) { if (which == 0) { n1 = ns; n2 = N2; } else { n1 = N1; n2 = ns; } };
               int n1, n2;
               if (Ns == 2) {          // 3A1A+3D1 (:47-82) / 3B1A+3I2 (:180-215): source has the site singly occupied
                  n1n2(1, n1, n2);
                  for (int tjd = tj_lo(n1, n2); tjd <= tj_hi(n1, n2); tjd += 2) {
                     if (std::abs(tsld - TwoSR) > tjd) continue;
                     const int ph = which == 0 ? phase(TwoSL + TwoSR + 2 + TwoS2) : phase(TwoSL + TwoSR + 3 - tjd);
                     const double f = std::sqrt((tjd + 1) * (tsld + 1.0)) * ph * wigner6j(tjd, TwoSo, 1, TwoSL, tsld, TwoSR);
                     L1(Qp, false, NL + 1, tsld, ILd, n1, n2, tjd, f, own);
                  }
               }
               if (Ns == 1) {          // 3A1B (:85-100) / 3B1B (:218-233): source has the site empty
                  n1n2(0, n1, n2);
                  if (std::abs(tsld - TwoSR) <= TwoSo) {
                     const int ph = which == 0 ? phase(TwoSL + TwoSR + 1 + TwoS2) : phase(TwoSL + TwoSR + 2 - TwoJ);
                     const double f = std::sqrt((tsld + 1) * (TwoJ + 1.0)) * ph * wigner6j(TwoSo, TwoJ, 1, TwoSL, tsld, TwoSR);
                     L1(Q, false, NL + 1, tsld, ILd, n1, n2, TwoSo, f, own);
                  }
               }
               if (Ns == 0) {          // 3A2A (:102-126) / 3B2A (:235-259)
                  n1n2(1, n1, n2);
                  for (int tjd = tj_lo(n1, n2); tjd <= tj_hi(n1, n2); tjd += 2) {
                     if (std::abs(tsld - TwoSR) > tjd) continue;
                     const int ph = which == 0 ? phase(tsld + TwoSR + 1 + TwoS2) : phase(tsld + TwoSR + 2 - tjd);
                     const double f = ph * std::sqrt((TwoSL + 1) * (tjd + 1.0)) * wigner6j(tjd, TwoSo, 1, TwoSL, tsld, TwoSR);
                     L1(Q, true, NL - 1, tsld, ILd, n1, n2, tjd, f, own);
                  }
               }
               if (Ns == 1) {          // 3A2B+3D2 (:128-157) / 3B2B+3I1 (:261-290)
                  n1n2(2, n1, n2);
                  if (std::abs(tsld - TwoSR) <= TwoSo) {
                     const int ph = which == 0 ? phase(tsld + TwoSR + 2 + TwoS2) : phase(tsld + TwoSR + 3 - TwoJ);
                     const double f = ph * std::sqrt((TwoSL + 1) * (TwoJ + 1.0)) * wigner6j(TwoSo, TwoJ, 1, TwoSL, tsld, TwoSR);
                     L1(Qp, true, NL - 1, tsld, ILd, n1, n2, TwoSo, f, own);
                  }
               }
            }
         }
      }
      if (right) {
         for (int which = 0; which < 2; which++) {   // 0: site ix (3K/3F), 1: site ix+1 (3L/3G)
            const int s = ix + which;
            OpH Q = Ro(K_Q, s, s);
            if (Q.op < 0) continue;
            std::vector<std::pair<double, int>> parts;
            for (int l = ix + 2; l < L; l++)
               if (irr(l) == irr(s)) parts.push_back({V(s, s, s, l), right->find(K_L, l, l)});   // :534-540, :667-673
            OpH Qp = presum(which == 0 ? "3K" : "3L", SRC_RIGHT, Q.op, parts);
            const int IRd = xorp(IR, irr(s));
            const int Ns = which == 0 ? N1 : N2;
            const int TwoSo = which == 0 ? TwoS2 : TwoS1;
            const int own = o_q(s);
            for (int tsrd = TwoSR - 1; tsrd <= TwoSR + 1; tsrd += 2) {
               if (tsrd < 0) continue;
               auto n1n2 = [&](int ns, int& n1, int& n2) { if (which == 0) { n1 = ns; n2 = N2; } else { n1 = N1; n2 = ns; } };
               int n1, n2;
               if (Ns == 1) {          // 3K1A (:495-510) / 3L1A (:628-643)
                  n1n2(0, n1, n2);
                  if (std::abs(TwoSL - tsrd) <= TwoSo) {
                     const int ph = which == 0 ? phase(TwoSL + TwoSR + TwoJ + 2 * TwoS2) : phase(TwoSL + TwoSR + TwoS1 + 1);
                     const double f = std::sqrt((TwoJ + 1) * (TwoSR + 1.0)) * ph * wigner6j(TwoSo, TwoJ, 1, TwoSR, tsrd, TwoSL);
                     R1(Q, false, NR - 1, tsrd, IRd, n1, n2, TwoSo, f, own);
                  }
               }
               if (Ns == 2) {          // 3K1B+3F1 (:512-548) / 3L1B+3G1 (:645-681)
                  n1n2(1, n1, n2);
                  for (int tjd = tj_lo(n1, n2); tjd <= tj_hi(n1, n2); tjd += 2) {
                     if (std::abs(TwoSL - tsrd) > tjd) continue;
                     const int ph = which == 0 ? phase(TwoSL + TwoSR + tjd + 1 + 2 * TwoS2) : phase(TwoSL + TwoSR + TwoS1 + 2);
                     const double f = std::sqrt((tjd + 1) * (TwoSR + 1.0)) * ph * wigner6j(tjd, TwoSo, 1, TwoSR, tsrd, TwoSL);
                     R1(Qp, false, NR - 1, tsrd, IRd, n1, n2, tjd, f, own);
                  }
               }
               if (Ns == 0) {          // 3K2A (:550-574) / 3L2A (:683-707)
                  n1n2(1, n1, n2);
                  for (int tjd = tj_lo(n1, n2); tjd <= tj_hi(n1, n2); tjd += 2) {
                     if (std::abs(TwoSL - tsrd) > tjd) continue;
                     const int ph = which == 0 ? phase(TwoSL + tsrd + tjd + 2 * TwoS2) : phase(TwoSL + tsrd + TwoS1 + 1);
                     const double f = std::sqrt((tjd + 1) * (tsrd + 1.0)) * ph * wigner6j(tjd, TwoSo, 1, TwoSR, tsrd, TwoSL);
                     R1(Q, true, NR + 1, tsrd, IRd, n1, n2, tjd, f, own);
                  }
               }
               if (Ns == 1) {          // 3K2B+3F2 (:576-605) / 3L2B+3G2 (:709-738)
                  n1n2(2, n1, n2);
                  if (std::abs(TwoSL - tsrd) <= TwoSo) {
                     const int ph = which == 0 ? phase(TwoSL + tsrd + TwoJ + 1 + 2 * TwoS2) : phase(TwoSL + tsrd + TwoS1 + 2);
                     const double f = std::sqrt((TwoJ + 1) * (tsrd + 1.0)) * ph * wigner6j(TwoSo, TwoJ, 1, TwoSR, tsrd, TwoSL);
                     R1(Qp, true, NR + 1, tsrd, IRd, n1, n2, TwoSo, f, own);
                  }
               }
            }
         }
      }
   }
   // 3C (HeffDiagrams3.cpp:294-397): Q_left(l) x L_right(l), l > ix+1.   3J (:742-848): L_left(l) x Q_right(l), l < ix.
   void d3CJ() {
      const int extra = ((N1 == 1) ? 2 : 0) + ((N2 == 1) ? 2 : 0);
      for (int tsld = TwoSL - 1; tsld <= TwoSL + 1; tsld += 2)
         for (int tsrd = TwoSR - 1; tsrd <= TwoSR + 1; tsrd += 2) {
            if (tsld < 0 || tsrd < 0 || std::abs(tsld - tsrd) > TwoJ) continue;
            const double w = wigner6j(TwoSL, TwoSR, TwoJ, tsrd, tsld, 1);
            const double f_up = phase(tsld + TwoSR + TwoJ + 1 + extra) * std::sqrt((tsld + 1) * (tsrd + 1.0)) * w;   // :320-321, :768-769
            const double f_dn = phase(TwoSL + tsrd + TwoJ + 1 + extra) * std::sqrt((TwoSL + 1) * (TwoSR + 1.0)) * w; // :361-362, :810-811
            for (int l = ix + 2; l < L; l++) {   // 3C
               const int ild = xorp(IL, irr(l)), ird = xorp(IR, irr(l));
               LR(Lo(K_Q, l, l), false, NL + 1, tsld, ild, Ro(K_L, l, l), true, NR + 1, tsrd, ird, N1, N2, TwoJ, f_up, o_q(l));
               LR(Lo(K_Q, l, l), true, NL - 1, tsld, ild, Ro(K_L, l, l), false, NR - 1, tsrd, ird, N1, N2, TwoJ, f_dn, o_q(l));
            }
            for (int l = 0; l < ix; l++) {       // 3J
               const int ild = xorp(IL, irr(l)), ird = xorp(IR, irr(l));
               LR(Lo(K_L, l, l), false, NL + 1, tsld, ild, Ro(K_Q, l, l), true, NR + 1, tsrd, ird, N1, N2, TwoJ, f_up, o_q(l));
               LR(Lo(K_L, l, l), true, NL - 1, tsld, ild, Ro(K_Q, l, l), false, NR - 1, tsrd, ird, N1, N2, TwoJ, f_dn, o_q(l));
            }
         }
   }
   // 3E+3H (HeffDiagrams3.cpp:399-474): site-site terms with three indices on one site
   void d3EH() {
      const int i = ix, j = ix + 1;
      if (irr(i) != irr(j)) return;
      const double s2 = std::sqrt(2.0);
      const double a = V(i, i, i, j), b = V(i, j, j, j);
      if (N1 == 2 && N2 == 0) Z(1, 1, 0, s2 * a, o_master());
      if (N1 == 2 && N2 == 1) Z(1, 2, 1, -(a + b), o_master());
      if (N1 == 1 && N2 == 1 && TwoJ == 0) { Z(2, 0, 0, s2 * a, o_master()); Z(0, 2, 0, s2 * b, o_master()); }
      if (N1 == 1 && N2 == 2) Z(2, 1, 1, -(a + b), o_master());
      if (N1 == 0 && N2 == 2) Z(1, 1, 0, s2 * b, o_master());
   }

#include "b2_sigma_plan_f4.inc"
#include "b2_sigma_plan_f5.inc"
};

}   // namespace

namespace {
// Groups -> GPUs: longest-processing-time-first on the FLOPs the scheduler will execute per term (cheaper association order, like
// compile_terms picks it).  Deterministic: every rank builds the same plan and evaluates the same assignment.  world == 1: all on GPU 0.
void balance_owners(SigmaPlan& plan, const OpSet*, const OpSet*, int world) {
   if (world <= 1) { for (SigmaTerm& t : plan.terms) t.owner = 0; return; }
   int ngroups = 0;
   for (const SigmaTerm& t : plan.terms) ngroups = std::max(ngroups, t.owner + 1);
   std::vector<double> cost((size_t)ngroups, 0.0);
   for (const SigmaTerm& t : plan.terms) {
      const double M = plan.S.blk[t.dst].rows, N = plan.S.blk[t.dst].cols, k1 = plan.S.blk[t.src].rows, k2 = plan.S.blk[t.src].cols;
      const bool hl = t.l.src != SRC_NONE && t.l.blk >= 0, hr = t.r.src != SRC_NONE && t.r.blk >= 0;
      double c;
      if (hl && hr) c = 2.0 * std::min(M * k1 * k2 + M * k2 * N, k1 * k2 * N + M * k1 * N);
      else if (hl) c = 2.0 * M * N * k1;
      else if (hr) c = 2.0 * M * N * k2;
      else c = 2.0 * M * N;
      cost[t.owner] += c;
   }
   std::vector<int> order;
   for (int g = 0; g < ngroups; g++) if (cost[g] > 0.0) order.push_back(g);
   std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
   std::vector<double> load((size_t)world, 0.0);
   std::vector<int> rank_of((size_t)ngroups, 0);
   for (int g : order) {
      const int r = (int)(std::min_element(load.begin(), load.end()) - load.begin());
      rank_of[g] = r; load[r] += cost[g];
   }
   for (SigmaTerm& t : plan.terms) t.owner = rank_of[t.owner];
}
}   // namespace

void build_sigma_plan(SigmaPlan& plan, const Bookkeeper& bk, const Problem& prob, const OpSet* left, const OpSet* right,
                      int site, int world) {
   plan.site = site;
   plan.at_left = (site == 0);
   plan.at_right = (site == bk.L - 2);
   plan.S.build(bk, site);
   plan.terms.clear(); plan.presums.clear(); plan.presum_size = 0; plan.skipped_zero = 0; plan.flops_ref = 0.0;
   if (plan.at_left) left = nullptr;
   if (plan.at_right) right = nullptr;
   const int nk = plan.S.nkappa();
   auto enumerate_block = [&](Gen& g, int k) {
      g.set_block(k);
      g.d1(); g.d2d(); g.d3EH();
      g.d2bcef(); g.d2bcef3(); g.d3_onesided();
      g.d4_onesided();
      if (left && right) { g.d2a(); g.d3CJ(); g.d4_twosided(); g.d5(); }
   };
   const int nthreads = plan_threads(nk);
   if (nthreads <= 1) {
      Gen g(plan, bk, prob, left, right, site, world < 1 ? 1 : world);
      for (int k = 0; k < nk; k++) enumerate_block(g, k);
      balance_owners(plan, left, right, world);
      return;
   }
   // Target blocks are independent: every host thread enumerates the blocks it grabs into a private fragment (terms, pre-sums,
   // FLOP count); the fragments are stitched together in block order, pre-sums de-duplicated by their keys and re-indexed, so the
   // resulting plan has exactly the terms (and the per-block term order) of the sequential enumeration.
   struct Frag {
      SigmaPlan plan;
      std::vector<int> blk_k, blk_begin;             // blocks this thread enumerated and where their terms start
      std::map<std::string, int> keys;
   };
   std::vector<Frag> frags(nthreads);
   std::atomic<int> next{0};
   auto worker = [&](int t) {
      Frag& f = frags[t];
      f.plan.site = site; f.plan.at_left = plan.at_left; f.plan.at_right = plan.at_right; f.plan.S = plan.S;
      Gen g(f.plan, bk, prob, left, right, site, world < 1 ? 1 : world);
      for (;;) {
         const int k0 = next.fetch_add(4);
         if (k0 >= nk) break;
         for (int k = k0; k < std::min(nk, k0 + 4); k++) {
            f.blk_k.push_back(k); f.blk_begin.push_back((int)f.plan.terms.size());
            enumerate_block(g, k);
         }
      }
      f.blk_begin.push_back((int)f.plan.terms.size());
      f.keys.swap(g.presum_index);
   };
   const auto tp0 = std::chrono::steady_clock::now();
   parallel_run(nthreads, worker);
   const auto tp1 = std::chrono::steady_clock::now();
   // ---- pre-sums: global registry in thread order
   std::map<std::string, int> global;
   std::vector<std::vector<int>> remap(nthreads);
   size_t nterms = 0;
   for (int t = 0; t < nthreads; t++) {
      Frag& f = frags[t];
      remap[t].assign(f.plan.presums.size(), -1);
      for (auto& kv : f.keys) {
         auto it = global.find(kv.first);
         if (it == global.end()) {
            Presum p = f.plan.presums[kv.second];
            p.off = plan.presum_size;
            plan.presum_size += (p.lay->size + 15) / 16 * 16;
            plan.presums.push_back(p);
            it = global.emplace(kv.first, (int)plan.presums.size() - 1).first;
         }
         remap[t][kv.second] = it->second;
      }
      plan.skipped_zero += f.plan.skipped_zero;
      plan.flops_ref += f.plan.flops_ref;
      nterms += f.plan.terms.size();
   }
   // ---- terms in block order
   std::vector<std::pair<int, int>> where(nk, {-1, -1});   // block -> (thread, position in its block list)
   for (int t = 0; t < nthreads; t++)
      for (size_t i = 0; i < frags[t].blk_k.size(); i++) where[frags[t].blk_k[i]] = {t, (int)i};
   std::vector<size_t> first(nk + 1, 0);                    // where the terms of block k start in the stitched list
   for (int k = 0; k < nk; k++) {
      const Frag& f = frags[where[k].first];
      first[k + 1] = first[k] + (size_t)(f.blk_begin[where[k].second + 1] - f.blk_begin[where[k].second]);
   }
   plan.terms.resize(nterms);
   parallel_run(nthreads, [&](int piece) {                   // contiguous ranges of blocks, balanced in terms
      const size_t lo = nterms * piece / nthreads, hi = nterms * (piece + 1) / nthreads;
      const int kb = (int)(std::lower_bound(first.begin(), first.end() - 1, lo) - first.begin());
      const int ke = (int)(std::lower_bound(first.begin(), first.end() - 1, hi) - first.begin());
      for (int k = kb; k < (piece == nthreads - 1 ? nk : ke); k++) {
         const int t = where[k].first, i = where[k].second;
         const Frag& f = frags[t];
         size_t out = first[k];
         for (int e = f.blk_begin[i]; e < f.blk_begin[i + 1]; e++) {
            SigmaTerm x = f.plan.terms[e];
            if (x.l.src == SRC_PRESUM && x.l.op >= 0) x.l.op = remap[t][x.l.op];
            if (x.r.src == SRC_PRESUM && x.r.op >= 0) x.r.op = remap[t][x.r.op];
            plan.terms[out++] = x;
         }
      }
   });
   balance_owners(plan, left, right, world);
   if (getenv("B2_TIMING"))
      fprintf(stderr, "build_sigma_plan: %d blocks on %d threads, enumerate %.3f s, stitch %.3f s\n", nk, nthreads,
              std::chrono::duration<double>(tp1 - tp0).count(), std::chrono::duration<double>(std::chrono::steady_clock::now() - tp1).count());
}

}   // namespace b2
