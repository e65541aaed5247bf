// b2_sobject.cpp — two-site object algebra: Join as contraction terms for the device kernels, Split on the host.
//
//   join_terms : Sobject::Join (Sobject.cpp:212-258)  S[kappa] = sum_jM f6j * T_i[L -> M] * T_{i+1}[M -> R]   -> Term3 list
//   split_host : Sobject::Split (Sobject.cpp:260-622) per-centre-sector SVD, global truncation to D, discarded weight,
//                new virtual dimensions written into the bookkeeper, new site tensors (lambda on the side the sweep moves to).
//                The SVD is a one-sided Jacobi written here (the reference calls LAPACK dgesdd_, :412-419); this step is
//                host-side in the reference too and is listed as "next" for the device in SURVEY.md 8(f).
#include "b2_sobject.h"
#include "b2_sigma.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <cmath>
#include <cstring>

namespace b2 {

void join_terms(std::vector<Term3>& terms, std::vector<DstBlock>& dst, const Bookkeeper& bk, const SLayout& S, const TLayout& TL, const TLayout& TR) {
   terms.clear();
   dst.resize(S.nkappa());
   const int ix = S.site;
   for (int k = 0; k < S.nkappa(); k++) {
      dst[k] = DstBlock{S.blk[k].off, S.blk[k].rows, S.blk[k].cols};
      const int NL = S.NL[k], TwoSL = S.twoSL[k], IL = S.IL[k], NR = S.NR[k], TwoSR = S.twoSR[k], IR = S.IR[k];
      const int TwoJ = S.twoJ[k], N1 = S.N1[k], N2 = S.N2[k];
      const int TwoS1 = (N1 == 1) ? 1 : 0, TwoS2 = (N2 == 1) ? 1 : 0;
      const int fase = phase(TwoSL + TwoSR + TwoS1 + TwoS2);
      const int NM = NL + N1;
      const int IM = (TwoS1 == 1) ? xorp(IL, bk.orb_irrep[ix]) : IL;
      const int lo = std::max(std::abs(TwoSL - TwoS1), std::abs(TwoSR - TwoS2)), hi = std::min(TwoSL + TwoS1, TwoSR + TwoS2);
      for (int TwoJM = lo; TwoJM <= hi; TwoJM += 2) {
         if (bk.dim(ix + 1, NM, TwoJM, IM) <= 0) continue;
         const int kl = TL.kappa(bk, NL, TwoSL, IL, NM, TwoJM, IM), kr = TR.kappa(bk, NM, TwoJM, IM, NR, TwoSR, IR);
         if (kl < 0 || kr < 0) continue;
         Term3 t;
         t.dst = k;
         t.f = fase * std::sqrt(1.0 * (TwoJ + 1) * (TwoJM + 1)) * wigner6j(TwoSL, TwoSR, TwoJ, TwoS2, TwoS1, TwoJM);
         t.p.space = SP_LEFT; t.p.off = TL.blk[kl].off; t.p.rows = TL.blk[kl].rows; t.p.cols = TL.blk[kl].cols;
         t.r.space = SP_RIGHT; t.r.off = TR.blk[kr].off; t.r.rows = TR.blk[kr].rows; t.r.cols = TR.blk[kr].cols;
         if (t.f != 0.0) terms.push_back(t);
      }
   }
}

namespace {
struct Center { int NM, TwoJM, IM; };
struct Piece { int n, ts, ir, dim, start; };

std::vector<Piece> left_pieces(const Bookkeeper& bk, int ix, const Center& c) {   // Sobject.cpp:321-334
   std::vector<Piece> v;
   int tot = 0;
   for (int NL = c.NM - 2; NL <= c.NM; NL++) {
      const int TwoS1 = (NL + 1 == c.NM) ? 1 : 0;
      for (int TwoSL = c.TwoJM - TwoS1; TwoSL <= c.TwoJM + TwoS1; TwoSL += 2) {
         if (TwoSL < 0) continue;
         const int IL = TwoS1 ? xorp(bk.orb_irrep[ix], c.IM) : c.IM;
         const int d = bk.dim(ix, NL, TwoSL, IL);
         if (d > 0) { v.push_back({NL, TwoSL, IL, d, tot}); tot += d; }
      }
   }
   return v;
}
std::vector<Piece> right_pieces(const Bookkeeper& bk, int ix, const Center& c) {   // Sobject.cpp:335-348
   std::vector<Piece> v;
   int tot = 0;
   for (int NR = c.NM; NR <= c.NM + 2; NR++) {
      const int TwoS2 = (NR == c.NM + 1) ? 1 : 0;
      for (int TwoSR = c.TwoJM - TwoS2; TwoSR <= c.TwoJM + TwoS2; TwoSR += 2) {
         if (TwoSR < 0) continue;
         const int IR = TwoS2 ? xorp(bk.orb_irrep[ix + 1], c.IM) : c.IM;
         const int d = bk.dim(ix + 2, NR, TwoSR, IR);
         if (d > 0) { v.push_back({NR, TwoSR, IR, d, tot}); tot += d; }
      }
   }
   return v;
}
}   // namespace

double split_host(Bookkeeper& bk, int ix, const SLayout& S, const double* s_storage, int D, bool moving_right, bool change,
                  std::vector<double>& t_left, std::vector<double>& t_right, const SvdBatchFn& svd_batch) {
   auto now_s = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
   const double t_g0 = now_s();
   std::vector<Center> centers;
   bk.for_sectors(ix + 1, [&](int n, int ts, int ir) { if (bk.fcidim(ix + 1, n, ts, ir) > 0) centers.push_back({n, ts, ir}); });
   const int nc = (int)centers.size();
   std::vector<std::vector<double>> Lam(nc), Us(nc), VTs(nc);
   std::vector<int> cdim(nc, 0), dimLtot(nc, 0), dimRtot(nc, 0);
   std::vector<std::vector<Piece>> LP(nc), RP(nc);
   std::vector<std::vector<double>> mems(nc);
   std::vector<SvdJob> jobs;
   for (int ic = 0; ic < nc; ic++) {
      const Center& c = centers[ic];
      LP[ic] = left_pieces(bk, ix, c);
      RP[ic] = right_pieces(bk, ix, c);
      for (auto& p : LP[ic]) dimLtot[ic] += p.dim;
      for (auto& p : RP[ic]) dimRtot[ic] += p.dim;
      cdim[ic] = std::min(dimLtot[ic], dimRtot[ic]);
   }
   // centre sectors are independent (own matrix, own blocks of the new tensors): the host workers share them
   auto for_centers = [&](const std::function<void(int)>& body) {
      std::atomic<int> next{0};
      const int T = std::max(1, std::min(nc, plan_threads(8 * nc)));
      parallel_run(T, [&](int) { for (int ic; (ic = next.fetch_add(1)) < nc;) body(ic); });
   };
   for_centers([&](int ic) {
      const Center& c = centers[ic];
      if (cdim[ic] <= 0) return;
      const int M = dimLtot[ic], N = dimRtot[ic];
      std::vector<double>& mem = mems[ic];
      mem.assign((size_t)M * N, 0.0);
      for (auto& pl : LP[ic]) {
         const int TwoS1 = (pl.n + 1 == c.NM) ? 1 : 0;
         for (auto& pr : RP[ic]) {
            const int TwoS2 = (pr.n == c.NM + 1) ? 1 : 0;
            const int fase = phase(pl.ts + pr.ts + TwoS1 + TwoS2);
            const int jmin = std::max(std::abs(pr.ts - pl.ts), std::abs(TwoS2 - TwoS1)), jmax = std::min(TwoS1 + TwoS2, pl.ts + pr.ts);
            for (int TwoJ = jmin; TwoJ <= jmax; TwoJ += 2) {
               const int k = S.kappa(bk, pl.n, pl.ts, pl.ir, c.NM - pl.n, pr.n - c.NM, TwoJ, pr.n, pr.ts, pr.ir);
               if (k < 0) continue;
               const double pref = fase * std::sqrt(1.0 * (TwoJ + 1) * (pr.ts + 1)) * wigner6j(pl.ts, pr.ts, TwoJ, TwoS2, TwoS1, c.TwoJM);   // :378-381
               const double* blk = s_storage + S.blk[k].off;
               for (int r = 0; r < pr.dim; r++)
                  for (int l = 0; l < pl.dim; l++) mem[pl.start + l + (size_t)M * (pr.start + r)] += pref * blk[l + (size_t)pl.dim * r];
            }
         }
      }
      Lam[ic].resize(cdim[ic]); Us[ic].resize((size_t)cdim[ic] * M); VTs[ic].resize((size_t)cdim[ic] * N);
   });
   for (int ic = 0; ic < nc; ic++) {
      if (cdim[ic] <= 0) continue;
      SvdJob job;
      job.m = dimLtot[ic]; job.n = dimRtot[ic]; job.a = mems[ic].data(); job.s = Lam[ic].data(); job.u = Us[ic].data(); job.vt = VTs[ic].data();
      jobs.push_back(job);
   }
   // the decomposition itself (dgesdd_ per centre sector in the reference, Sobject.cpp:412-419): all sectors in one device batch
   const double t_g1 = now_s();
   if (!svd_batch || svd_batch(jobs) != 0) return -1.0;   // negative = failed (the caller reports the device error)
   mems.clear();
   const double t_g2 = now_s();

   double discarded = 0.0;
   if (change) {   // Sobject.cpp:437-499
      std::vector<int> newdim(cdim);
      long total = 0;
      for (int ic = 0; ic < nc; ic++) total += newdim[ic];
      if (total > D) {
         std::vector<double> values;
         for (int ic = 0; ic < nc; ic++) values.insert(values.end(), Lam[ic].begin(), Lam[ic].begin() + newdim[ic]);
         std::sort(values.begin(), values.end(), [](double a, double b) { return a > b; });
         const double lower = values[D];
         for (int ic = 0; ic < nc; ic++)
            for (int cnt = 0; cnt < newdim[ic]; cnt++)
               if (Lam[ic][cnt] <= lower) newdim[ic] = cnt;
         double tot = 0.0, dis = 0.0;
         for (int ic = 0; ic < nc; ic++)
            for (int il = 0; il < cdim[ic]; il++) {
               const double w = (centers[ic].TwoJM + 1) * Lam[ic][il] * Lam[ic][il];
               tot += w;
               if (Lam[ic][il] <= lower) dis += w;
            }
         discarded = dis / tot;
      }
      bool differs = false;
      for (int ic = 0; ic < nc; ic++)
         if (newdim[ic] != bk.dim(ix + 1, centers[ic].NM, centers[ic].TwoJM, centers[ic].IM)) differs = true;
      if (differs)
         for (int ic = 0; ic < nc; ic++) bk.set_dim(ix + 1, centers[ic].NM, centers[ic].TwoJM, centers[ic].IM, newdim[ic]);
   }

   // new site tensors in the (possibly changed) layouts  (Sobject.cpp:519-589)
   TLayout TL, TR;
   TL.build(bk, ix);
   TR.build(bk, ix + 1);
   t_left.assign((size_t)TL.size, 0.0);
   t_right.assign((size_t)TR.size, 0.0);
   for_centers([&](int ic) {
      const Center& c = centers[ic];
      const int dimM = bk.dim(ix + 1, c.NM, c.TwoJM, c.IM);
      if (dimM <= 0) return;
      const int lim = std::min(dimM, cdim[ic]);
      for (auto& pl : LP[ic]) {
         const int k = TL.kappa(bk, pl.n, pl.ts, pl.ir, c.NM, c.TwoJM, c.IM);
         if (k < 0) continue;
         double* blk = t_left.data() + TL.blk[k].off;
         for (int r = 0; r < lim; r++) {
            const double f = moving_right ? 1.0 : Lam[ic][r];
            for (int l = 0; l < pl.dim; l++) blk[l + (size_t)pl.dim * r] = f * Us[ic][pl.start + l + (size_t)dimLtot[ic] * r];
         }
      }
      for (auto& pr : RP[ic]) {
         const int k = TR.kappa(bk, c.NM, c.TwoJM, c.IM, pr.n, pr.ts, pr.ir);
         if (k < 0) continue;
         double* blk = t_right.data() + TR.blk[k].off;
         const double fb = std::sqrt((c.TwoJM + 1.0) / (pr.ts + 1));
         for (int l = 0; l < lim; l++) {
            const double f = fb * (moving_right ? Lam[ic][l] : 1.0);
            for (int r = 0; r < pr.dim; r++) blk[l + (size_t)dimM * r] = f * VTs[ic][l + (size_t)cdim[ic] * (pr.start + r)];
         }
      }
   });
   if (getenv("B2_TIMING")) fprintf(stderr, "split_host: %d centre sectors: gather %.4f s, decomposition %.4f s, truncation + scatter %.4f s\n", nc, t_g1 - t_g0, t_g2 - t_g1, now_s() - t_g2);
   return discarded;
}

// Left-normalise a site tensor in place (what TensorT::QR does with the R factor discarded, TensorT.cpp:188-289): for every right
// sector the stacked left blocks get orthonormal columns (modified Gram-Schmidt, twice).
void left_normalize_host(const Bookkeeper& bk, const TLayout& T, double* t) {
   const int ix = T.site;
   bk.for_sectors(ix + 1, [&](int NR, int TwoSR, int IR) {
      const int dR = bk.dim(ix + 1, NR, TwoSR, IR);
      if (dR <= 0) return;
      std::vector<int> ks;
      int rows = 0;
      for (int k = 0; k < T.nkappa(); k++)
         if (T.NR[k] == NR && T.twoSR[k] == TwoSR && T.IR[k] == IR) { ks.push_back(k); rows += T.blk[k].rows; }
      if (rows == 0) return;
      std::vector<double> A((size_t)rows * dR);
      int r0 = 0;
      for (int k : ks) {
         for (int c = 0; c < dR; c++)
            for (int r = 0; r < T.blk[k].rows; r++) A[r0 + r + (size_t)rows * c] = t[T.blk[k].off + r + (size_t)T.blk[k].rows * c];
         r0 += T.blk[k].rows;
      }
      // Householder QR with LAPACK's conventions (dgeqrf_ + dorgqr_ as TensorT::QR calls them, TensorT.cpp:227-252): reflector k
      // maps column k onto beta e_k with beta = -sign(alpha) * norm, so Q equals the reference's Q column by column (a Gram-Schmidt Q
      // differs from it by column signs, a different — equally valid — gauge).  Columns beyond min(rows, dR) are zero.
      const int kk = std::min(rows, dR);
      std::vector<double> tau(kk, 0.0);
      for (int c = 0; c < kk; c++) {
         double* v = &A[(size_t)rows * c];
         double xn = 0.0;
         for (int r = c + 1; r < rows; r++) xn += v[r] * v[r];
         const double alpha = v[c];
         if (xn == 0.0) { tau[c] = 0.0; continue; }          // H = I
         const double beta = -std::copysign(std::sqrt(alpha * alpha + xn), alpha);
         tau[c] = (beta - alpha) / beta;
         const double sc = 1.0 / (alpha - beta);
         for (int r = c + 1; r < rows; r++) v[r] *= sc;
         v[c] = beta;
         for (int j = c + 1; j < dR; j++) {                  // apply H = I - tau v v^T (v_c = 1) to the trailing columns
            double* w = &A[(size_t)rows * j];
            double d = w[c];
            for (int r = c + 1; r < rows; r++) d += v[r] * w[r];
            d *= tau[c];
            w[c] -= d;
            for (int r = c + 1; r < rows; r++) w[r] -= d * v[r];
         }
      }
      {  // Q = H_0 H_1 ... H_{kk-1} applied to the first kk columns of the identity, built backwards (dorg2r)
         std::vector<double> Q((size_t)rows * dR, 0.0);
         for (int c = kk - 1; c >= 0; c--) {
            const double* v = &A[(size_t)rows * c];
            double* qc = &Q[(size_t)rows * c];
            for (int j = c + 1; j < kk; j++) {               // columns already built: Q_j <- H_c Q_j (rows >= c)
               double* w = &Q[(size_t)rows * j];
               double d = w[c];
               for (int r = c + 1; r < rows; r++) d += v[r] * w[r];
               d *= tau[c];
               w[c] -= d;
               for (int r = c + 1; r < rows; r++) w[r] -= d * v[r];
            }
            qc[c] = 1.0 - tau[c];
            for (int r = c + 1; r < rows; r++) qc[r] = -tau[c] * v[r];
         }
         A.swap(Q);
      }
      r0 = 0;
      for (int k : ks) {
         for (int c = 0; c < dR; c++)
            for (int r = 0; r < T.blk[k].rows; r++) t[T.blk[k].off + r + (size_t)T.blk[k].rows * c] = A[r0 + r + (size_t)rows * c];
         r0 += T.blk[k].rows;
      }
   });
}

}   // namespace b2
