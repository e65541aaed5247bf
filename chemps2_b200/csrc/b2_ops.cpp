#include "b2_ops.h"

namespace b2 {

std::shared_ptr<const OpLayout> OpSet::layout(const Bookkeeper& bk, int two_j, int n_elec, int irrep) {
   const int64_t k = ((int64_t)two_j << 40) | ((int64_t)n_elec << 20) | irrep;
   auto it = layouts.find(k);
   if (it != layouts.end()) return it->second;
   auto l = std::make_shared<OpLayout>();
   l->build(bk, boundary, two_j, n_elec, irrep);
   layouts[k] = l;
   return l;
}

int OpSet::add(const Bookkeeper& bk, int kind, int i, int j) {
   OpTensor t;
   t.kind = kind; t.i = i; t.j = j;
   if (kind == K_X || kind == K_G || kind == K_Y || kind == K_Z) t.irrep = 0;
   else if (i == j && (kind == K_L || kind == K_Q || kind == K_K || kind == K_M)) t.irrep = bk.orb_irrep[i];
   else t.irrep = xorp(bk.orb_irrep[i], bk.orb_irrep[j]);
   t.prime_last = (kind == K_F1 || kind == K_D) ? moving_right : true;
   t.lay = layout(bk, kind_two_j(kind), kind_nelec(kind), t.irrep);
   t.off = size;
   size += (t.lay->size + 15) / 16 * 16;   // 128-byte aligned operator starts
   index[key(kind, i, j)] = (int)ops.size();
   ops.push_back(t);
   return (int)ops.size() - 1;
}

void OpSet::build_all(const Bookkeeper& bk, int boundary_, bool moving_right_) {
   boundary = boundary_; moving_right = moving_right_;
   ops.clear(); index.clear(); layouts.clear(); size = 0;
   const int L = bk.L, b = boundary;
   // sites inside the renormalized block / outside it
   const int in_lo = moving_right ? 0 : b, in_hi = moving_right ? b - 1 : L - 1;
   const int out_lo = moving_right ? b : 0, out_hi = moving_right ? L - 1 : b - 1;
   for (int s = in_lo; s <= in_hi; s++) add(bk, K_L, s, s);
   for (int i = in_lo; i <= in_hi; i++)
      for (int j = i; j <= in_hi; j++) {
         add(bk, K_S0, i, j);
         if (j > i) add(bk, K_S1, i, j);
         add(bk, K_F0, i, j);
         add(bk, K_F1, i, j);
      }
   for (int i = out_lo; i <= out_hi; i++)
      for (int j = i; j <= out_hi; j++) {
         add(bk, K_A, i, j);
         if (j > i) add(bk, K_B, i, j);
         add(bk, K_C, i, j);
         add(bk, K_D, i, j);
      }
   for (int s = out_lo; s <= out_hi; s++) add(bk, K_Q, s, s);
   add(bk, K_X, -1, -1);
}

void OpSet::build_reduced(const Bookkeeper& bk, int boundary_, bool moving_right_, bool only_L) {
   boundary = boundary_; moving_right = moving_right_; reduced = true;
   ops.clear(); index.clear(); layouts.clear(); size = 0;
   const int L = bk.L, b = boundary;
   const int in_lo = moving_right ? 0 : b, in_hi = moving_right ? b - 1 : L - 1;
   for (int s = in_lo; s <= in_hi; s++) add(bk, K_L, s, s);
   if (only_L) return;
   for (int i = in_lo; i <= in_hi; i++)
      for (int j = i; j <= in_hi; j++) {
         add(bk, K_S0, i, j);
         if (j > i) add(bk, K_S1, i, j);
         add(bk, K_F0, i, j);
         add(bk, K_F1, i, j);
      }
}

void OpSet::build_correlation(const Bookkeeper& bk, int boundary_) {
   boundary = boundary_; moving_right = true; reduced = true;
   ops.clear(); index.clear(); layouts.clear(); size = 0;
   for (int s = 0; s < boundary; s++) {
      add(bk, K_G, s, s); add(bk, K_Y, s, s); add(bk, K_Z, s, s); add(bk, K_K, s, s); add(bk, K_M, s, s);
   }
}

}   // namespace b2
