#include "b2_core.h"
#include <atomic>
#include <cstdio>
#include <thread>
#include <vector>
int main() {
   std::atomic<long long> total{0};
   auto caller = [&](int id) {
      for (int rep = 0; rep < 300; rep++) {
         const int n = 1 + (rep * 7 + id * 3) % 17;
         std::vector<int> hit(n, 0);
         b2::parallel_run(n, [&](int t) { hit[t]++; long long s = 0; for (int i = 0; i < 2000 * ((t % 3) + 1); i++) s += i % 7; total += s & 1; });
         for (int t = 0; t < n; t++) if (hit[t] != 1) { std::printf("BAD piece %d ran %d times\n", t, hit[t]); std::abort(); }
      }
   };
   std::vector<std::thread> th;
   for (int i = 0; i < 4; i++) th.emplace_back(caller, i);
   for (auto& t : th) t.join();
   // nested use: a piece that itself calls parallel_run
   b2::parallel_run(6, [&](int) { b2::parallel_run(5, [&](int) { total++; }); });
   std::printf("ok %lld\n", (long long)total);
}
