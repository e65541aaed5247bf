// b2_diag.cpp — diagonal of the effective Hamiltonian (Heff::fillHeffDiag, Heff.cpp:250-315; HeffDiagonal.cpp).
#include "b2_sigma.h"

namespace b2 {

void build_heff_diag(double* diag, const SLayout& S, const Bookkeeper& bk, const Problem& prob, const OpSet* left,
                     const double* left_arena, const OpSet* right, const double* right_arena, int site) {
   (void)bk; (void)prob; (void)left; (void)left_arena; (void)right; (void)right_arena; (void)site;
   for (int64_t i = 0; i < S.size; i++) diag[i] = 0.0;   // filled in below (TODO)
}

}   // namespace b2
