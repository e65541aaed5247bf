/* TEST INFRASTRUCTURE ONLY — CPU checker, never called by the product path (there is no CPU fallback).
 *
 * Restates the arithmetic of Heff::makeHeff (Heff.cpp:43-248) term by term: the reference zeroes every target block
 * (Heff.cpp:61) and then, for every diagram term, issues  dgemm_(op(A), S[src]) -> temp ; dgemm_(temp, op(B)) -> sigma
 * (e.g. HeffDiagrams2.cpp:70-74) or a single one-sided dgemm_ (HeffDiagrams1.cpp:33) or a daxpy_ (HeffDiagrams1.cpp:52).
 * Here the same products are done with plain triple loops on a flat term list (b2_flat_term of include/chemps2_b200.h),
 * independent of any BLAS.  The integral-weighted operator pre-sums (dcopy_+daxpy_ loops, HeffDiagrams3.cpp:64-75) are
 * restated by b2o_presum.
 *
 * Parity of this checker is pinned by tests/golden/*.npz, which were produced by the UNMODIFIED reference
 * (oracle/ref_driver.cpp -> Heff::makeHeff) — see tests/test_oracle_golden.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
   int32_t dst, src;
   int32_t a_rows, a_cols;
   int32_t b_rows, b_cols;
   int8_t a_space, a_trans;
   int8_t b_space, b_trans;
   int32_t owner;
   int64_t a_off, b_off;
   double factor;
} flat_term;

typedef struct {
   int64_t dst_off;
   int64_t src_off;
   int64_t size;
   int32_t space;
   double coef;
} flat_presum;

/* presum[dst_off + e] += coef * arena[space][src_off + e] */
void b2o_presum(const flat_presum* parts, int64_t nparts, const double* left, const double* right, double* presum, int64_t presum_size) {
   memset(presum, 0, sizeof(double) * (size_t)presum_size);
   for (int64_t p = 0; p < nparts; p++) {
      const double* src = (parts[p].space == 1 ? left : right) + parts[p].src_off;
      double* dst = presum + parts[p].dst_off;
      for (int64_t e = 0; e < parts[p].size; e++) dst[e] += parts[p].coef * src[e];
   }
}

static double elemA(const flat_term* t, const double* A, int i, int k) {   /* op(A)[i][k], i < dimL, k < dLs */
   return t->a_trans ? A[k + (size_t)t->a_rows * i] : A[i + (size_t)t->a_rows * k];
}
static double elemB(const flat_term* t, const double* B, int k, int j) {   /* op(B)[k][j], k < dRs, j < dimR */
   return t->b_trans ? B[j + (size_t)t->b_rows * k] : B[k + (size_t)t->b_rows * j];
}

/* vec_out = sum of all terms applied to vec_in.  blk_off/rows/cols describe the Sobject blocks.  Only terms with
 * owner == rank are applied when world > 1 (mirrors the MPI guards of Heff.cpp:66-239). */
void b2o_apply(const flat_term* terms, int64_t nterms, int nblk, const int64_t* blk_off, const int32_t* blk_rows, const int32_t* blk_cols,
               const double* left, const double* right, const double* presum, const double* vin, double* vout, int rank, int world) {
   int64_t total = nblk ? blk_off[nblk - 1] + (int64_t)blk_rows[nblk - 1] * blk_cols[nblk - 1] : 0;
   memset(vout, 0, sizeof(double) * (size_t)total);
   for (int64_t it = 0; it < nterms; it++) {
      const flat_term* t = terms + it;
      if (world > 1 && t->owner != rank) continue;
      const int dimL = blk_rows[t->dst], dimR = blk_cols[t->dst], dLs = blk_rows[t->src], dRs = blk_cols[t->src];
      const double* S = vin + blk_off[t->src];
      double* H = vout + blk_off[t->dst];
      const double* arenas[4] = {NULL, left, right, presum};
      const double* A = t->a_space ? arenas[t->a_space] + t->a_off : NULL;
      const double* B = t->b_space ? arenas[t->b_space] + t->b_off : NULL;
      if (A && B) {
         double* temp = (double*)malloc(sizeof(double) * (size_t)dimL * dRs);
         for (int j = 0; j < dRs; j++)
            for (int i = 0; i < dimL; i++) {
               double s = 0.0;
               for (int k = 0; k < dLs; k++) s += elemA(t, A, i, k) * S[k + (size_t)dLs * j];
               temp[i + (size_t)dimL * j] = s;
            }
         for (int j = 0; j < dimR; j++)
            for (int i = 0; i < dimL; i++) {
               double s = 0.0;
               for (int k = 0; k < dRs; k++) s += temp[i + (size_t)dimL * k] * elemB(t, B, k, j);
               H[i + (size_t)dimL * j] += t->factor * s;
            }
         free(temp);
      } else if (A) {
         for (int j = 0; j < dimR; j++)
            for (int i = 0; i < dimL; i++) {
               double s = 0.0;
               for (int k = 0; k < dLs; k++) s += elemA(t, A, i, k) * S[k + (size_t)dLs * j];
               H[i + (size_t)dimL * j] += t->factor * s;
            }
      } else if (B) {
         for (int j = 0; j < dimR; j++)
            for (int i = 0; i < dimL; i++) {
               double s = 0.0;
               for (int k = 0; k < dRs; k++) s += S[i + (size_t)dimL * k] * elemB(t, B, k, j);
               H[i + (size_t)dimL * j] += t->factor * s;
            }
      } else {
         for (int64_t e = 0; e < (int64_t)dimL * dimR; e++) H[e] += t->factor * S[e];
      }
   }
}
