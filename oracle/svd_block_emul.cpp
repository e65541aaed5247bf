// TEST INFRASTRUCTURE (only tests/ may call this): CPU emulation of the block-Jacobi SVD kernel k_jacobi_block of
// chemps2_b200/csrc/b2_svd.cu, which stands for the dgesdd_ call of Sobject::Split (Sobject.cpp:412-419).
// The emulation follows the kernel phase by phase — every region between two __syncthreads() becomes a loop over the 128 threads
// of the CTA, with the same index arithmetic (round-robin block pairs, Gram entries per thread, rotation tasks, row-wise
// application of Q) — so that the schedule, the convergence rule and the accuracy of the device algorithm are pinned on the CPU
// against LAPACK (tests/test_svd_block_emul.py); the device kernel itself is compared with LAPACK in tests/test_svd_gpu.py.
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <vector>

namespace {
constexpr int NB = 8, NP = 16, BT = 128, RCH = 64;

void rr_pair(int n, int s, int k, int& p, int& q) {
   const int n1 = n - 1;
   if (k == 0) { p = n1; q = s; }
   else { p = (s + k) % n1; q = (s - k + n1) % n1; }
   if (p > q) std::swap(p, q);
}

// one CTA of one step; returns true when it rotated
bool cta(int R, int C, int nbe, int step, int k, double* W, double* V, double tiny) {
   static double panel[NP][RCH + 1], G[NP][NP + 1], Q[NP][NP + 1], rc[NB], rs[NB];
   static int rp[NB], rq[NB], cols[NP];
   int bi, bj;
   rr_pair(nbe, step % (nbe - 1), k, bi, bj);
   const int ni = std::max(0, std::min(NB, C - bi * NB)), nj = std::max(0, std::min(NB, C - bj * NB));
   const int np = ni + nj;
   if (np < 2) return false;
   for (int tid = 0; tid < NP; tid++) cols[tid] = tid < ni ? bi * NB + tid : (tid < np ? bj * NB + (tid - ni) : -1);
   double g0[BT] = {}, g1[BT] = {};
   for (int r0 = 0; r0 < R; r0 += RCH) {
      for (int tid = 0; tid < BT; tid++)
         for (int idx = tid; idx < NP * RCH; idx += BT) {
            const int c = idx / RCH, i = idx - c * RCH;
            panel[c][i] = (c < np && r0 + i < R) ? W[(size_t)R * cols[c] + r0 + i] : 0.0;
         }
      for (int tid = 0; tid < BT; tid++) {
         const int a = tid >> 4, b = tid & 15;
         for (int i = 0; i < RCH; i++) {
            const double y = panel[b][i];
            g0[tid] += panel[a][i] * y;
            g1[tid] += panel[a + NB][i] * y;
         }
      }
   }
   for (int tid = 0; tid < BT; tid++) {
      const int a = tid >> 4, b = tid & 15;
      G[a][b] = g0[tid]; G[a + NB][b] = g1[tid];
      Q[a][b] = (a == b) ? 1.0 : 0.0; Q[a + NB][b] = (a + NB == b) ? 1.0 : 0.0;
   }
   int need = 0;
   for (int tid = 0; tid < BT; tid++) {
      const int a = tid >> 4, b = tid & 15;
      if (a < b && b < np) need |= (std::fabs(g0[tid]) > 1e-15 * std::sqrt(G[a][a] * G[b][b]) && std::fabs(g0[tid]) > tiny);
      if (a + NB < b && b < np) need |= (std::fabs(g1[tid]) > 1e-15 * std::sqrt(G[a + NB][a + NB] * G[b][b]) && std::fabs(g1[tid]) > tiny);
   }
   if (!need) return false;
   const int ne = np + (np & 1);
   for (int st = 0; st < ne - 1; st++) {
      for (int tid = 0; tid < ne / 2; tid++) {
         int p, q;
         rr_pair(ne, st, tid, p, q);
         double c = 1.0, sn = 0.0;
         if (q < np) {
            const double gamma = G[p][q], alpha = G[p][p], beta = G[q][q];
            if (!(std::fabs(gamma) <= 1e-15 * std::sqrt(std::fabs(alpha * beta)) || std::fabs(gamma) <= tiny)) {
               const double zeta = (beta - alpha) / (2.0 * gamma);
               const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
               c = 1.0 / std::sqrt(1.0 + t * t); sn = c * t;
            }
         }
         rp[tid] = p; rq[tid] = q; rc[tid] = c; rs[tid] = sn;
      }
      for (int tid = 0; tid < BT; tid++) {
         const int kk = tid >> 4, r = tid & 15;
         const bool on = kk < ne / 2 && rs[kk] != 0.0;
         if (!on) continue;
         const int p = rp[kk], q = rq[kk];
         const double c = rc[kk], sn = rs[kk];
         const double x = G[r][p], y = G[r][q];
         G[r][p] = c * x - sn * y; G[r][q] = sn * x + c * y;
         const double u = Q[r][p], v = Q[r][q];
         Q[r][p] = c * u - sn * v; Q[r][q] = sn * u + c * v;
      }
      for (int tid = 0; tid < BT; tid++) {
         const int kk = tid >> 4, r = tid & 15;
         const bool on = kk < ne / 2 && rs[kk] != 0.0;
         if (!on) continue;
         const int p = rp[kk], q = rq[kk];
         const double c = rc[kk], sn = rs[kk];
         const double x = G[p][r], y = G[q][r];
         G[p][r] = c * x - sn * y; G[q][r] = sn * x + c * y;
      }
   }
   for (int pass = 0; pass < 2; pass++) {
      double* M = pass == 0 ? W : V;
      const int rows = pass == 0 ? R : C;
      for (int tid = 0; tid < BT; tid++)
         for (int i = tid; i < rows; i += BT) {
            double x[NP];
            for (int c = 0; c < NP; c++) x[c] = c < np ? M[(size_t)rows * cols[c] + i] : 0.0;
            for (int bb = 0; bb < np; bb++) {
               double y = 0.0;
               for (int c = 0; c < NP; c++) y += x[c] * Q[c][bb];
               M[(size_t)rows * cols[bb] + i] = y;
            }
         }
   }
   return true;
}
}   // namespace

// W: R x C column-major (R >= C) in, W with orthogonal columns out; V: C x C out (A = W V^T).  Returns the number of sweeps.
extern "C" int b2o_svd_block(int R, int C, double* W, double* V) {
   for (size_t e = 0; e < (size_t)C * C; e++) V[e] = 0.0;
   for (int c = 0; c < C; c++) V[c + (size_t)C * c] = 1.0;
   if (C < 2) return 0;
   double scale = 0.0;
   for (size_t e = 0; e < (size_t)R * C; e++) scale = std::max(scale, std::fabs(W[e]));
   const double tiny = scale * scale * 1e-300;
   const int nb = (C + NB - 1) / NB, nbe = std::max(2, nb + (nb & 1));
   const int nsteps = std::max(1, nbe - 1);
   for (int sweep = 0; sweep < 60; sweep++) {
      bool rotated = false;
      for (int st = 0; st < nsteps; st++)
         for (int k = 0; k < nbe / 2; k++) rotated = cta(R, C, nbe, st, k, W, V, tiny) || rotated;
      if (!rotated) return sweep + 1;
   }
   return 60;
}
