// b2_core.h — host-side symmetry bookkeeping and packed tensor layouts for the B200 DMRG sweep path.
//
// Everything here reproduces the *enumeration orders and block layouts* of the reference so that packed host
// arrays are interchangeable with the reference's gStorage() arrays:
//   Bookkeeper  <-> CheMPS2::SyBookkeeper            (SyBookkeeper.cpp:29-309)
//   TLayout     <-> CheMPS2::TensorT sector table    (TensorT.cpp:38-104)
//   OpLayout    <-> CheMPS2::TensorOperator table    (TensorOperator.cpp:29-102)
//   SLayout     <-> CheMPS2::Sobject sector table    (Sobject.cpp:36-147)
//   wigner6j/9j <-> CheMPS2::Wigner                  (Wigner.cpp:294-368)
// The look-ups are O(1) hash look-ups instead of the reference's linear scans.
#pragma once
#include <memory>
#include <new>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <string>
#include <unordered_map>
#include <vector>

namespace b2 {

// Host worker pool of the plan builders: fn(0) runs on the caller, fn(1) .. fn(n-1) on parked worker threads; returns when all are
// done.  The workers are created once and sleep on a condition variable between calls.  Freshly created threads start on their
// parent's core and the load balancer spreads them only after tens of milliseconds, so a plan-building burst of 50-500 ms on new
// std::threads runs almost serially (measured: 8 x 41 ms of work in 340 ms on new threads, 55-70 ms on parked ones).  A call made
// while the pool is busy (another host thread, or a nested call) runs its n pieces sequentially on the caller.
void parallel_run(int n, const std::function<void(int)>& fn);

// Large host blocks of the plan builders (work lists, term arrays: hundreds of MB per site, built and dropped at every site).  malloc
// hands such sizes to mmap, so every plan pays the page faults of fresh memory and every release the munmap (measured: 20 ms per site at
// D = 2000, as much as the Split).  Freed blocks are kept (size classes with <= 12.5 % rounding, at most B2_HOST_CACHE_GB = 4 GB idle) and
// handed out again, dirty: callers must not expect zeros.
void* host_block_acquire(size_t bytes);
void host_block_release(void* p, size_t bytes);
// std::allocator whose blocks of 1 MB and more come from that cache
template <class T> struct CachedAlloc : std::allocator<T> {
   template <class U> struct rebind { using other = CachedAlloc<U>; };
   CachedAlloc() = default;
   template <class U> CachedAlloc(const CachedAlloc<U>&) {}
   static constexpr size_t kCached = (size_t)1 << 20;
   T* allocate(size_t n) {
      const size_t bytes = n * sizeof(T);
      return bytes >= kCached ? static_cast<T*>(host_block_acquire(bytes)) : static_cast<T*>(::operator new(bytes));
   }
   void deallocate(T* p, size_t n) {
      const size_t bytes = n * sizeof(T);
      if (bytes >= kCached) host_block_release(p, bytes); else ::operator delete(p);
   }
};
template <class T> using BigVec = std::vector<T, CachedAlloc<T>>;

// (-1)^{two_power/2}; same integer semantics as Special::phase (Special.h:36) / Heff::phase (Heff.h:75).
inline int phase(int two_power) { return (((two_power / 2) % 2) != 0) ? -1 : 1; }
inline int xorp(int a, int b) { return a ^ b; }   // abelian point-group product, Irreps.h:123

double wigner6j(int two_ja, int two_jb, int two_jc, int two_jd, int two_je, int two_jf);
double wigner9j(int two_ja, int two_jb, int two_jc, int two_jd, int two_je, int two_jf, int two_jg, int two_jh,
                int two_ji);

int num_irreps_of_group(int group);   // psi4 numbering 0..7 = c1 ci c2 cs d2 c2v c2h d2h

// The additive-feedback generator behind glibc's srand() / rand() (TYPE_3: x_i = x_{i-3} + x_{i-31} mod 2^32, output x_i >> 1,
// seeded through the Lehmer sequence 16807 x mod 2^31-1, first 310 outputs discarded; RAND_MAX = 2^31-1).  The reference draws its
// random MPS (TensorT::random, TensorT.cpp:167-173) and the noise of Sobject::addNoise (Sobject.cpp:652-659) from rand(); a private
// generator with the identical stream makes seeded runs comparable step by step with the reference without touching libc's global
// state (tests/test_rng.py pins it against libc).
struct GlibcRand {
   uint32_t r[34];
   int pos = 0;   // index of the next output modulo 34 (ring buffer of the last 34 values)
   static constexpr double RANDMAX = 2147483647.0;
   explicit GlibcRand(unsigned int seed = 1) { reseed(seed); }
   void reseed(unsigned int seed) {
      int32_t x[34];
      x[0] = (int32_t)(seed == 0 ? 1u : seed);
      for (int i = 1; i < 31; i++) {
         int64_t v = (16807LL * x[i - 1]) % 2147483647LL;
         if (v < 0) v += 2147483647LL;
         x[i] = (int32_t)v;
      }
      for (int i = 31; i < 34; i++) x[i] = x[i - 31];
      for (int i = 0; i < 34; i++) r[i] = (uint32_t)x[i];
      pos = 0;                      // r[k % 34] holds x_k; the next value to produce is x_34
      for (int i = 34; i < 344; i++) next_raw();
   }
   uint32_t next_raw() {            // x_k = x_{k-31} + x_{k-3}; stored over x_{k-34}
      const uint32_t v = r[(pos + 3) % 34] + r[(pos + 31) % 34];
      r[pos] = v;
      pos = (pos + 1) % 34;
      return v;
   }
   int next() { return (int)(next_raw() >> 1); }   // == rand()
};

// ---------------------------------------------------------------------------------------------------------
// Problem: target sector + dense two-body table with the one-body part folded in (Problem.cpp:351-384).
// Orbitals are already in DMRG order (the caller applies any reordering).
struct Problem {
   int L = 0, group = 0, N = 0, twoS = 0, irrep = 0;
   double econst = 0.0;
   std::vector<int> orb_irrep;
   std::vector<double> mx;   // mx[a + L*(b + L*(c + L*d))]
   double V(int a, int b, int c, int d) const { return mx[a + L * (b + L * (c + L * (size_t)d))]; }
   // tmat[L*L] (i + L*j), vmat[L^4] physicist <ab|cd> at a + L*(b + L*(c + L*d))
   void build(const double* tmat, const double* vmat);
};

// ---------------------------------------------------------------------------------------------------------
struct Bookkeeper {
   int L = 0, N = 0, twoS = 0, irrep = 0, nirr = 1;
   std::vector<int> orb_irrep;
   std::vector<int> Nmin, Nmax;                 // per boundary 0..L
   std::vector<std::vector<int>> tsmin, tsmax;  // [boundary][N - Nmin]
   std::vector<std::vector<int>> slot0;         // [boundary][N - Nmin] -> first slot of (N, tsmin, irrep 0)
   std::vector<int> nslots;                     // per boundary
   std::vector<std::vector<int>> fci, cur;      // [boundary][slot]

   void init(const Problem& p, int D);          // FCI dims + ceil-scaled current dims (SyBookkeeper.cpp:29-49)
   int slot(int b, int n, int two_s, int irr) const {   // -1 when outside the table (SyBookkeeper.cpp:271-280)
      if (b < 0 || b > L) return -1;
      if (n > Nmax[b] || n < Nmin[b]) return -1;
      const int lo = tsmin[b][n - Nmin[b]], hi = tsmax[b][n - Nmin[b]];
      if (((two_s - lo) & 1) || two_s < lo || two_s > hi) return -1;
      if (irr < 0 || irr >= nirr) return -1;
      return slot0[b][n - Nmin[b]] + ((two_s - lo) / 2) * nirr + irr;
   }
   int dim(int b, int n, int two_s, int irr) const { const int s = slot(b, n, two_s, irr); return s < 0 ? 0 : cur[b][s]; }
   int fcidim(int b, int n, int two_s, int irr) const { const int s = slot(b, n, two_s, irr); return s < 0 ? 0 : fci[b][s]; }
   void set_dim(int b, int n, int two_s, int irr, int value) {   // SyBookkeeper.cpp:163-169
      const int s = slot(b, n, two_s, irr);
      if (s >= 0 && fci[b][s] != 0) cur[b][s] = value;
   }
   int max_dim_at(int b) const;
   int tot_dim_at(int b) const;
   bool is_possible() const { return dim(L, N, twoS, irrep) == 1; }
   // visit populated-or-not sectors of one boundary in the reference order N up, 2S up, irrep up
   template <class F> void for_sectors(int b, F&& f) const {
      for (int n = Nmin[b]; n <= Nmax[b]; n++)
         for (int ts = tsmin[b][n - Nmin[b]]; ts <= tsmax[b][n - Nmin[b]]; ts += 2)
            for (int ir = 0; ir < nirr; ir++) f(n, ts, ir);
   }
private:
   void fill_fci();
};

// ---------------------------------------------------------------------------------------------------------
// Packed block layouts. A block is column-major rows x cols with ld = rows, no padding, offsets in doubles.
struct Block { int64_t off; int rows, cols; };

struct TLayout {   // MPS site tensor for site `site` (left boundary site, right boundary site+1)
   int site = 0;
   std::vector<int> NL, twoSL, IL, NR, twoSR, IR;
   std::vector<Block> blk;
   int64_t size = 0;
   std::unordered_map<uint64_t, int> index;
   void build(const Bookkeeper& bk, int site);
   int kappa(const Bookkeeper& bk, int nl, int tsl, int il, int nr, int tsr, int ir) const;
   int nkappa() const { return (int)blk.size(); }
};

struct OpLayout {  // renormalized operator at `boundary` with quantum numbers (two_j, n_elec, irrep)
   int boundary = 0, two_j = 0, n_elec = 0, irrep = 0;
   std::vector<int> Nup, twoSup, Iup, twoSdown;
   std::vector<Block> blk;
   int64_t size = 0;
   std::unordered_map<uint64_t, int> index;
   void build(const Bookkeeper& bk, int boundary, int two_j, int n_elec, int irrep);
   // block for up sector (n1,ts1,i1) -> down sector (n2,ts2,i2); -1 if absent (TensorOperator.cpp:119-137)
   int kappa(const Bookkeeper& bk, int n1, int ts1, int i1, int n2, int ts2, int i2) const;
   int nkappa() const { return (int)blk.size(); }
};

struct SLayout {   // two-site object at sites (site, site+1)
   int site = 0;
   std::vector<int> NL, twoSL, IL, N1, N2, twoJ, NR, twoSR, IR;
   std::vector<Block> blk;
   int64_t size = 0;
   std::unordered_map<uint64_t, int> index;
   void build(const Bookkeeper& bk, int site);
   int kappa(const Bookkeeper& bk, int nl, int tsl, int il, int n1, int n2, int tj, int nr, int tsr, int ir) const;
   int nkappa() const { return (int)blk.size(); }
};

}   // namespace b2
