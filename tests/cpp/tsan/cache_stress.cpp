#include "b2_core.h"
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
int main() {
   setenv("B2_HOST_CACHE_GB", "0.05", 1);   // small cache: the overflow path runs too
   auto worker = [&](int id) {
      for (int rep = 0; rep < 2000; rep++) {
         const size_t n = ((size_t)1 << 20) + (size_t)((rep * 7919 + id * 104729) % 4000000);
         b2::BigVec<char> v(n);
         std::memset(v.data(), id, 64);
         b2::BigVec<char> w(n / 3 + 5);
         v.swap(w);
      }
   };
   std::vector<std::thread> th;
   for (int i = 0; i < 6; i++) th.emplace_back(worker, i);
   for (auto& t : th) t.join();
   std::printf("ok\n");
}
