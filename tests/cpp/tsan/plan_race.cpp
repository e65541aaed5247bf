// TSAN driver: what the sweep driver's prefetch does on the host — sigma plan of site s+1 on a helper thread while the update plan of
// site s is built on the main thread — on a planning-only context.
#include "chemps2_b200.h"
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include <random>
#define CK(x) do { int rc_ = (x); if (rc_) { std::printf("FAIL %s: %s\n", #x, b2_last_error()); std::exit(1); } } while (0)
int main(int argc, char** argv) {
   const int L = 12, D = argc > 1 ? atoi(argv[1]) : 120;
   std::vector<int> irr(L, 0);
   std::vector<double> t(L * L), v((size_t)L * L * L * L);
   std::mt19937 g(7);
   std::uniform_real_distribution<double> u(-1, 1);
   for (int i = 0; i < L; i++) for (int j = 0; j <= i; j++) t[i + L * j] = t[j + L * i] = u(g);
   auto V = [&](int a, int b, int c, int d) -> double& { return v[a + L * (b + L * (c + (size_t)L * d))]; };
   for (int a = 0; a < L; a++) for (int b = 0; b < L; b++) for (int c = 0; c < L; c++) for (int d = 0; d < L; d++) {
      // <ab|cd> = (ac|bd): 8-fold symmetry
      const double x = 0.1 * std::cos(1.0 * ((a + 1) * (c + 1) + (b + 1) * (d + 1)) + 0.3 * ((a + c) * (b + d)));
      V(a, b, c, d) = x;
   }
   for (int a = 0; a < L; a++) for (int b = 0; b < L; b++) for (int c = 0; c < L; c++) for (int d = 0; d < L; d++) {
      const double x = V(a, b, c, d);
      V(b, a, d, c) = x; V(c, d, a, b) = x; V(d, c, b, a) = x; V(c, b, a, d) = x; V(a, d, c, b) = x; V(b, c, d, a) = x; V(d, a, b, c) = x;
   }
   b2_ctx* ctx = nullptr;
   CK(b2_ctx_create(-1, &ctx));
   CK(b2_problem_set_integrals(ctx, L, 0, L, 0, 0, irr.data(), t.data(), v.data(), 0.0));
   CK(b2_bk_init(ctx, D));
   CK(b2_ctx_set_option(ctx, "parallel_min_terms", 64));
   for (int rep = 0; rep < 3; rep++)
      for (int s = 2; s < L - 4; s++) {
         b2_opset *old_l = nullptr, *fresh = nullptr, *right = nullptr;
         CK(b2_opset_create(ctx, s, 1, &old_l));
         CK(b2_opset_create(ctx, s + 1, 1, &fresh));
         CK(b2_opset_create(ctx, s + 3, 0, &right));
         b2_heff* h = nullptr;
         b2_update* up = nullptr;
         std::thread th([&] { CK(b2_heff_create(ctx, s + 1, fresh, right, 2, rep % 2, &h)); });
         CK(b2_update_create_sharded(ctx, s, 1, old_l, fresh, 2, rep % 2, &up));
         th.join();
         std::printf("site %d: sigma terms %lld\n", s, (long long)b2_heff_num_terms(h));
         b2_heff_destroy(h); b2_update_destroy(up);
         b2_opset_destroy(old_l); b2_opset_destroy(fresh); b2_opset_destroy(right);
      }
   b2_ctx_destroy(ctx);
   std::printf("ok\n");
}
