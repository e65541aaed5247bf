/* TEST INFRASTRUCTURE ONLY — never called by the product path (there is no CPU fallback).
 *
 * Executes the compiled device work lists (stage-1 tiles, stage-2 tiles, split-K reduce jobs, wave by wave) with plain
 * loops, following exactly the semantics documented in chemps2_b200/csrc/b2_device.h.  It lets the CPU test-suite check
 * the scheduler of b2_heff.cpp (waves, workspace reuse, split-K, deduplicated intermediates) against the reference's
 * golden sigma vectors without a GPU.  The arithmetic being scheduled is Heff::makeHeff's (Heff.cpp:43-248).
 */
#include <cstdint>
#include <cstring>
#include <vector>

#include "../chemps2_b200/csrc/b2_heff.h"
#include "../include/chemps2_b200.h"

using namespace b2;

static void run_tile(const Tile& t, const GemmItem* items, double* const* base) {
   std::vector<double> acc((size_t)t.mrem * t.nrem, 0.0);
   for (int it = t.item_begin; it < t.item_end; it++) {
      const GemmItem& I = items[it];
      const double* X = base[I.xs] + I.xoff;
      if (I.flags & IF_AXPY) {
         for (int c = 0; c < t.nrem; c++)
            for (int r = 0; r < t.mrem; r++)
               acc[r + (size_t)t.mrem * c] += I.alpha * ((I.flags & IF_TX) ? X[(size_t)(t.n0 + c) + (size_t)(t.m0 + r) * I.ldx] : X[(size_t)(t.m0 + r) + (size_t)(t.n0 + c) * I.ldx]);
         continue;
      }
      const double* Y = base[I.ys] + I.yoff;
      for (int c = 0; c < t.nrem; c++)
         for (int r = 0; r < t.mrem; r++) {
            double s = 0.0;
            for (int k = 0; k < I.k; k++) {
               const double x = (I.flags & IF_TX) ? X[(size_t)k + (size_t)(t.m0 + r) * I.ldx] : X[(size_t)(t.m0 + r) + (size_t)k * I.ldx];
               const double y = (I.flags & IF_TY) ? Y[(size_t)(t.n0 + c) + (size_t)k * I.ldy] : Y[(size_t)k + (size_t)(t.n0 + c) * I.ldy];
               s += x * y;
            }
            acc[r + (size_t)t.mrem * c] += I.alpha * s;
         }
   }
   double* C = base[t.cspace] + t.coff;
   for (int c = 0; c < t.nrem; c++)
      for (int r = 0; r < t.mrem; r++) {
         double* p = C + (size_t)(t.cm0 + r) + (size_t)(t.cn0 + c) * t.ldc;
         if (t.accumulate) *p += acc[r + (size_t)t.mrem * c]; else *p = acc[r + (size_t)t.mrem * c];
      }
}

static void run_lists(const b2_worklists* wl, double** base);

extern "C" void b2o_run_worklists(const b2_worklists* wl, const double* left, const double* right, const double* presum, const double* vin,
                                  double* vout, int64_t veclength) {
   double* base[SP_COUNT] = {nullptr, const_cast<double*>(left), const_cast<double*>(right), const_cast<double*>(presum), nullptr,
                             const_cast<double*>(vin), vout, nullptr};
   std::memset(vout, 0, sizeof(double) * (size_t)veclength);
   run_lists(wl, base);
}

/* operator update: spaces LEFT = old operator arena, RIGHT = MPS tensor, PRESUM, VOUT = new operator arena (zeroed by the caller
 * before pass 0, kept between the passes) */
extern "C" void b2o_run_update_pass(const b2_worklists* wl, const double* old_arena, const double* t, const double* presum, double* new_arena) {
   double* base[SP_COUNT] = {nullptr, const_cast<double*>(old_arena), const_cast<double*>(t), const_cast<double*>(presum), nullptr, nullptr, new_arena, nullptr};
   run_lists(wl, base);
}

static void run_lists(const b2_worklists* wl, double** base) {
   std::vector<double> work((size_t)wl->work_size + 1, 0.0), part((size_t)wl->part_size + 1, 0.0);
   base[SP_WORK] = work.data(); base[SP_PART] = part.data();
   const Wave* waves = (const Wave*)wl->waves;
   for (int64_t w = 0; w < wl->n_waves; w++) {
      const Wave& W = waves[w];
      std::fill(work.begin(), work.end(), 1e300);   // poison: a wave must not read intermediates of an earlier wave
      std::fill(part.begin(), part.end(), 1e300);
      for (int c = 0; c < kNumTileClasses; c++)
         for (int i = W.t1_begin[c]; i < W.t1_end[c]; i++) run_tile(((const Tile*)wl->tiles1[c])[i], (const GemmItem*)wl->items1, base);
      for (int c = 0; c < kNumTileClasses; c++)
         for (int i = W.t2_begin[c]; i < W.t2_end[c]; i++) run_tile(((const Tile*)wl->tiles2[c])[i], (const GemmItem*)wl->items2, base);
      for (int i = W.red_begin; i < W.red_end; i++) {
         const ReduceJob& j = ((const ReduceJob*)wl->reduces)[i];
         for (int e = 0; e < j.mrem * j.nrem; e++) {
            double v = 0.0;
            for (int p = 0; p < j.nparts; p++) v += part[(size_t)j.part_off + (size_t)p * j.part_stride + e];
            base[j.dst_space][(size_t)j.dst_off +  (size_t)(j.m0 + e % j.mrem) + (size_t)(j.n0 + e / j.mrem) * j.ldc] += v;
         }
      }
   }
}

/* diagonal of H_eff (Heff::fillHeffDiag, Heff.cpp:250-315 + HeffDiagonal.cpp:27-642): every list item is one reference term restricted
 * to the operator-block diagonals, diag[k](i, j) += f * A(i, i) * B(j, j) (a missing operator counts as 1); semantics of k_diag */
extern "C" void b2o_run_diag(const DiagItem* items, const DiagTile* tiles, int64_t ntiles, const double* left, const double* right, const double* presum,
                             double* out, int64_t veclength) {
   const double* base[SP_COUNT] = {nullptr, left, right, presum, nullptr, nullptr, nullptr, nullptr};
   std::memset(out, 0, sizeof(double) * (size_t)veclength);
   for (int64_t t = 0; t < ntiles; t++) {
      const DiagTile& T = tiles[t];
      for (int j = 0; j < T.nrem; j++)
         for (int i = 0; i < T.mrem; i++) {
            double v = 0.0;
            for (int it = T.item_begin; it < T.item_end; it++) {
               const DiagItem& I = items[it];
               const double a = I.as ? base[I.as][I.aoff + (size_t)(T.m0 + i) * (I.lda + 1)] : 1.0;
               const double b = I.bs ? base[I.bs][I.boff + (size_t)(T.n0 + j) * (I.ldb + 1)] : 1.0;
               v += I.f * a * b;
            }
            out[T.coff + (size_t)(T.m0 + i) + (size_t)(T.n0 + j) * T.ldc] = v;
         }
   }
}
