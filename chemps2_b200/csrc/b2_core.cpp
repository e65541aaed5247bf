// b2_core.cpp — see b2_core.h. Host-only; no CUDA here.
#include "b2_core.h"

#include <algorithm>
#include <atomic>
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <condition_variable>
#include <map>
#include <mutex>
#include <new>
#include <thread>
#include <vector>
#include <unistd.h>

namespace b2 {

// ------------------------------------------------------------------------------------------------ host worker pool
namespace {
// Parked workers shared by every caller.  A parallel_run is a JOB of n independent pieces; several jobs may be open at the same time
// (the sweep driver builds the sigma plan of the next site while the operator-update plan of the current one is built on the calling
// thread): idle workers take the next unclaimed piece of the oldest open job, the caller of a job works on its own pieces too and then
// waits for the pieces other threads took.  Pieces never wait for one another, so any number of threads >= 1 finishes every job.
struct WorkerPool {
   struct Job {
      const std::function<void(int)>* fn = nullptr;
      int n = 0, next = 0, done = 0;      // pieces, first unclaimed piece, finished pieces (all under the pool mutex)
      std::condition_variable finished;
   };
   std::mutex m;
   std::condition_variable wake;
   std::vector<std::thread> workers;
   std::vector<Job*> open;               // jobs with unclaimed pieces, oldest first
   bool quit = false;

   // claims a piece of the oldest open job (caller holds the mutex)
   Job* claim(int& piece) {
      if (open.empty()) return nullptr;
      Job* j = open.front();
      piece = j->next++;
      if (j->next >= j->n) open.erase(open.begin());
      return j;
   }
   void worker_main() {
      std::unique_lock<std::mutex> lk(m);
      for (;;) {
         wake.wait(lk, [&] { return quit || !open.empty(); });
         if (quit) return;
         int piece = -1;
         Job* j = claim(piece);
         if (!j) continue;
         lk.unlock();
         (*j->fn)(piece);
         lk.lock();
         if (++j->done == j->n) j->finished.notify_all();   // the job object lives until its caller has seen done == n under this mutex
      }
   }
   void run(int n, const std::function<void(int)>& fn) {
      if (n <= 1) { for (int t = 0; t < n; t++) fn(t); return; }
      Job job;
      job.fn = &fn; job.n = n;
      std::unique_lock<std::mutex> lk(m);
      while ((int)workers.size() < n - 1) workers.emplace_back(&WorkerPool::worker_main, this);
      open.push_back(&job);
      wake.notify_all();
      while (job.next < job.n) {          // the caller takes pieces of ITS job only: it must not get stuck in somebody else's long piece
         const int piece = job.next++;
         if (job.next >= job.n) open.erase(std::find(open.begin(), open.end(), &job));
         lk.unlock();
         fn(piece);
         lk.lock();
         ++job.done;
      }
      job.finished.wait(lk, [&] { return job.done == job.n; });
   }
   void shutdown() {
      { std::lock_guard<std::mutex> lk(m); quit = true; }
      wake.notify_all();
      for (std::thread& th : workers) th.join();
      workers.clear();
   }
};

// The pool lives on the heap and belongs to one process: a forked child inherits the object but neither its threads nor a usable
// copy of its mutexes / condition variables (their waiter bookkeeping describes threads that do not exist there), so the child
// abandons the inherited object and starts a fresh one.
struct PoolHolder {
   WorkerPool* pool = nullptr;
   std::atomic<pid_t> owner{0};   // read before the guard is taken (fork detection), written under it
   std::mutex guard;
   WorkerPool* get() {
      const pid_t me = getpid();
      const pid_t seen = owner.load();
      if (seen != 0 && seen != me) new (&guard) std::mutex();   // first call in a forked child: the copy may be in the locked state
      std::lock_guard<std::mutex> lk(guard);
      if (owner != me) {
         pool = new WorkerPool;                                     // an inherited pool is leaked on purpose
         owner = me;
      }
      return pool;
   }
   ~PoolHolder() {
      if (pool && owner == getpid()) { pool->shutdown(); delete pool; }
   }
};
}   // namespace

void parallel_run(int n, const std::function<void(int)>& fn) {
   static PoolHolder holder;
   holder.get()->run(n, fn);
}

// ------------------------------------------------------------------------------------------------ host block cache
namespace {
struct HostBlockCache {
   std::mutex m;
   std::map<size_t, std::vector<void*>> idle;     // class size -> blocks
   size_t idle_bytes = 0;
   size_t keep = [] { const char* e = getenv("B2_HOST_CACHE_GB"); return (size_t)((e ? atof(e) : 4.0) * 1073741824.0); }();
};
HostBlockCache& host_cache() { static HostBlockCache* c = new HostBlockCache; return *c; }   // never destroyed: releases may come from static destructors
size_t host_class(size_t n) {   // next multiple of 2^(floor(log2 n) - 3)
   size_t step = 1;
   while ((step << 4) <= n) step <<= 1;
   return (n + step - 1) / step * step;
}
}   // namespace

void* host_block_acquire(size_t bytes) {
   const size_t cls = host_class(std::max<size_t>(bytes, 64));
   HostBlockCache& C = host_cache();
   {
      std::lock_guard<std::mutex> lk(C.m);
      auto it = C.idle.find(cls);
      if (it != C.idle.end() && !it->second.empty()) {
         void* p = it->second.back();
         it->second.pop_back();
         C.idle_bytes -= cls;
         return p;
      }
   }
   void* p = std::malloc(cls);
   if (!p) {   // give the cache back and try once more
      std::vector<void*> drop;
      {
         std::lock_guard<std::mutex> lk(C.m);
         for (auto& kv : C.idle) { drop.insert(drop.end(), kv.second.begin(), kv.second.end()); kv.second.clear(); }
         C.idle_bytes = 0;
      }
      for (void* q : drop) std::free(q);
      p = std::malloc(cls);
      if (!p) throw std::bad_alloc();
   }
   return p;
}
void host_block_release(void* p, size_t bytes) {
   if (!p) return;
   const size_t cls = host_class(std::max<size_t>(bytes, 64));
   HostBlockCache& C = host_cache();
   std::vector<void*> drop;
   {
      std::lock_guard<std::mutex> lk(C.m);
      if (cls > C.keep) drop.push_back(p);
      else {
         if (C.idle_bytes + cls > C.keep) {   // full of sizes nobody asks for any more (the bond dimension moved on): start over
            for (auto& kv : C.idle) { drop.insert(drop.end(), kv.second.begin(), kv.second.end()); kv.second.clear(); }
            C.idle_bytes = 0;
         }
         C.idle[cls].push_back(p);
         C.idle_bytes += cls;
      }
   }
   for (void* q : drop) std::free(q);
}

// ------------------------------------------------------------------------------------------------ Wigner
// Racah's closed form for the 6j symbol with long-double factorials (arguments are doubled spins).
// Same function as CheMPS2::Wigner::wigner6j (Wigner.cpp:294-342); evaluated independently.
namespace {
struct FactTable {
   long double f[400];
   FactTable() { f[0] = 1.0L; for (int i = 1; i < 400; i++) f[i] = f[i - 1] * (long double)i; }
};
const FactTable& facts() { static FactTable t; return t; }

inline bool triangle_fails(int a, int b, int c) {
   if ((a + b + c) % 2 != 0) return true;
   if (c > a + b || c < std::abs(a - b)) return true;
   return false;
}
inline long double delta2(int a, int b, int c) {   // squared triangle coefficient
   const long double* f = facts().f;
   return f[(a + b - c) / 2] * f[(a - b + c) / 2] * f[(-a + b + c) / 2] / f[(a + b + c) / 2 + 1];
}
}   // namespace

double wigner6j(int a, int b, int c, int d, int e, int f_) {
   if (a < 0 || b < 0 || c < 0 || d < 0 || e < 0 || f_ < 0) return 0.0;
   if (triangle_fails(a, b, c) || triangle_fails(d, e, c) || triangle_fails(a, e, f_) || triangle_fails(d, b, f_)) return 0.0;
   const int a1 = (a + b + c) / 2, a2 = (d + e + c) / 2, a3 = (a + e + f_) / 2, a4 = (d + b + f_) / 2;
   const int b1 = (a + b + d + e) / 2, b2 = (a + c + d + f_) / 2, b3 = (b + c + e + f_) / 2;
   const int kmin = std::max(std::max(a1, a2), std::max(a3, a4));
   const int kmax = std::min(b1, std::min(b2, b3));
   if (kmax < kmin) return 0.0;
   const long double* f = facts().f;
   long double sum = 0.0L;
   for (int k = kmin; k <= kmax; k++) {
      const long double term = f[k + 1] / (f[k - a1] * f[k - a2] * f[k - a3] * f[k - a4] * f[b1 - k] * f[b2 - k] * f[b3 - k]);
      sum += (k % 2 == 0) ? term : -term;
   }
   const long double pre = sqrtl(delta2(a, b, c) * delta2(d, e, c) * delta2(a, e, f_) * delta2(d, b, f_));
   return (double)(pre * sum);
}

// 9j as a sum over products of three 6j symbols (standard identity; Wigner.cpp:344-368 uses the same one).
double wigner9j(int a, int b, int c, int d, int e, int f_, int g, int h, int i) {
   if (triangle_fails(a, b, c) || triangle_fails(d, e, f_) || triangle_fails(g, h, i) || triangle_fails(a, d, g) ||
       triangle_fails(b, e, h) || triangle_fails(c, f_, i))
      return 0.0;
   const int lo = std::max(std::abs(a - i), std::max(std::abs(h - d), std::abs(b - f_)));
   const int hi = std::min(a + i, std::min(h + d, b + f_));
   double value = 0.0;
   for (int x = lo; x <= hi; x += 2)
      value += (x + 1) * wigner6j(a, b, c, f_, i, x) * wigner6j(d, e, f_, b, x, h) * wigner6j(g, h, i, x, a, d);
   return (lo % 2 == 0) ? value : -value;
}

int num_irreps_of_group(int group) {
   static const int n[8] = {1, 2, 2, 2, 4, 4, 4, 8};
   return (group >= 0 && group < 8) ? n[group] : -1;
}

// ------------------------------------------------------------------------------------------------ Problem
void Problem::build(const double* tmat, const double* vmat) {
   mx.assign((size_t)L * L * L * L, 0.0);
   const double pref = 1.0 / (N - 1);
   for (int d = 0; d < L; d++)
      for (int c = 0; c < L; c++)
         for (int b = 0; b < L; b++)
            for (int a = 0; a < L; a++) {
               const size_t idx = a + L * (b + L * (c + L * (size_t)d));
               double v = vmat[idx];
               if (a == c) v += pref * tmat[b + L * d];
               if (b == d) v += pref * tmat[a + L * c];
               mx[idx] = v;
            }
}

// ------------------------------------------------------------------------------------------------ Bookkeeper
static const int kDimCutoff = 262144;   // Options.h:76 SYBK_dimensionCutoff

void Bookkeeper::init(const Problem& p, int D) {
   L = p.L; N = p.N; twoS = p.twoS; irrep = p.irrep; nirr = num_irreps_of_group(p.group); orb_irrep = p.orb_irrep;
   Nmin.assign(L + 1, 0); Nmax.assign(L + 1, 0);
   tsmin.assign(L + 1, {}); tsmax.assign(L + 1, {}); slot0.assign(L + 1, {}); nslots.assign(L + 1, 0);
   fci.assign(L + 1, {}); cur.assign(L + 1, {});
   for (int b = 0; b <= L; b++) {
      Nmin[b] = std::max(std::max(0, N + 2 * (b - L)), b - L + (N + twoS) / 2);
      Nmax[b] = std::min(std::min(2 * b, N), b + (N - twoS) / 2);
      int count = 0;
      for (int n = Nmin[b]; n <= Nmax[b]; n++) {
         const int t = L - b - std::abs(N - n - L + b);
         const int lo = std::max(n % 2, twoS - t);
         const int hi = std::min(b - std::abs(b - n), twoS + t);
         tsmin[b].push_back(lo); tsmax[b].push_back(hi); slot0[b].push_back(count);
         if (hi >= lo) count += ((hi - lo) / 2 + 1) * nirr;
      }
      nslots[b] = count;
      fci[b].assign(count, 0); cur[b].assign(count, 0);
   }
   fill_fci();
   cur = fci;
   for (int b = 1; b <= L - 1; b++) {
      const int tot = tot_dim_at(b);
      if (tot > D) {
         const double factor = (1.0 * D) / tot;
         for_sectors(b, [&](int n, int ts, int ir) {
            const int value = (int)(std::ceil(factor * dim(b, n, ts, ir)) + 0.1);
            set_dim(b, n, ts, ir, value);
         });
      }
   }
}

void Bookkeeper::fill_fci() {
   fci[0][slot(0, Nmin[0], tsmin[0][0], 0)] = 1;
   for (int b = 1; b <= L; b++)
      for_sectors(b, [&](int n, int ts, int ir) {
         const int io = xorp(ir, orb_irrep[b - 1]);
         long v = (long)fcidim(b - 1, n, ts, ir) + fcidim(b - 1, n - 2, ts, ir) + fcidim(b - 1, n - 1, ts + 1, io) +
                  fcidim(b - 1, n - 1, ts - 1, io);
         fci[b][slot(b, n, ts, ir)] = (int)std::min<long>(kDimCutoff, v);
      });
   const int rhs = fcidim(L, N, twoS, irrep);
   std::fill(fci[L].begin(), fci[L].end(), 0);
   { const int s = slot(L, N, twoS, irrep); if (s >= 0) fci[L][s] = std::min(1, rhs); }
   for (int b = L - 1; b >= 0; b--)
      for_sectors(b, [&](int n, int ts, int ir) {
         const int io = xorp(ir, orb_irrep[b]);
         long v = (long)fcidim(b + 1, n, ts, ir) + fcidim(b + 1, n + 2, ts, ir) + fcidim(b + 1, n + 1, ts + 1, io) +
                  fcidim(b + 1, n + 1, ts - 1, io);
         const int s = slot(b, n, ts, ir);
         fci[b][s] = (int)std::min<long>(fci[b][s], std::min<long>(kDimCutoff, v));
      });
}

int Bookkeeper::max_dim_at(int b) const { int m = 0; for (int d : cur[b]) m = std::max(m, d); return m; }
int Bookkeeper::tot_dim_at(int b) const { int t = 0; for (int d : cur[b]) t += d; return t; }

// ------------------------------------------------------------------------------------------------ layouts
void TLayout::build(const Bookkeeper& bk, int site_) {
   site = site_;
   NL.clear(); twoSL.clear(); IL.clear(); NR.clear(); twoSR.clear(); IR.clear(); blk.clear(); index.clear();
   size = 0;
   bk.for_sectors(site, [&](int nl, int tsl, int il) {
      const int dl = bk.dim(site, nl, tsl, il);
      if (dl <= 0) return;
      for (int nr = nl; nr <= nl + 2; nr++) {
         const int tj = (nr == nl + 1) ? 1 : 0;
         for (int tsr = tsl - tj; tsr <= tsl + tj; tsr += 2) {
            if (tsr < 0) continue;
            const int ir = (nr == nl + 1) ? xorp(il, bk.orb_irrep[site]) : il;
            const int dr = bk.dim(site + 1, nr, tsr, ir);
            if (dr <= 0) continue;
            index[((uint64_t)bk.slot(site, nl, tsl, il) << 32) | (uint32_t)bk.slot(site + 1, nr, tsr, ir)] = (int)blk.size();
            NL.push_back(nl); twoSL.push_back(tsl); IL.push_back(il); NR.push_back(nr); twoSR.push_back(tsr); IR.push_back(ir);
            blk.push_back({size, dl, dr});
            size += (int64_t)dl * dr;
         }
      }
   });
}

int TLayout::kappa(const Bookkeeper& bk, int nl, int tsl, int il, int nr, int tsr, int ir) const {
   const int sl = bk.slot(site, nl, tsl, il), sr = bk.slot(site + 1, nr, tsr, ir);
   if (sl < 0 || sr < 0) return -1;
   auto it = index.find(((uint64_t)sl << 32) | (uint32_t)sr);
   return it == index.end() ? -1 : it->second;
}

void OpLayout::build(const Bookkeeper& bk, int boundary_, int two_j_, int n_elec_, int irrep_) {
   boundary = boundary_; two_j = two_j_; n_elec = n_elec_; irrep = irrep_;
   Nup.clear(); twoSup.clear(); Iup.clear(); twoSdown.clear(); blk.clear(); index.clear();
   size = 0;
   bk.for_sectors(boundary, [&](int nu, int tsu, int iu) {
      const int du = bk.dim(boundary, nu, tsu, iu);
      if (du <= 0) return;
      const int id = xorp(irrep, iu), nd = nu + n_elec;
      for (int tsd = tsu - two_j; tsd <= tsu + two_j; tsd += 2) {
         if (tsd < 0) continue;
         const int dd = bk.dim(boundary, nd, tsd, id);
         if (dd <= 0) continue;
         index[((uint64_t)bk.slot(boundary, nu, tsu, iu) << 16) | (uint32_t)tsd] = (int)blk.size();
         Nup.push_back(nu); twoSup.push_back(tsu); Iup.push_back(iu); twoSdown.push_back(tsd);
         blk.push_back({size, du, dd});
         size += (int64_t)du * dd;
      }
   });
}

int OpLayout::kappa(const Bookkeeper& bk, int n1, int ts1, int i1, int n2, int ts2, int i2) const {
   if (xorp(i1, irrep) != i2 || n2 != n1 + n_elec || std::abs(ts1 - ts2) > two_j || ts2 < 0) return -1;
   const int su = bk.slot(boundary, n1, ts1, i1);
   if (su < 0) return -1;
   auto it = index.find(((uint64_t)su << 16) | (uint32_t)ts2);
   return it == index.end() ? -1 : it->second;
}

static inline uint64_t skey(int sl, int n1, int n2, int tj, int sr) {
   return ((uint64_t)sl << 36) | ((uint64_t)sr << 8) | (uint64_t)((n1 << 4) | (n2 << 2) | tj);
}

void SLayout::build(const Bookkeeper& bk, int site_) {
   site = site_;
   NL.clear(); twoSL.clear(); IL.clear(); N1.clear(); N2.clear(); twoJ.clear(); NR.clear(); twoSR.clear(); IR.clear();
   blk.clear(); index.clear(); size = 0;
   const int i1 = bk.orb_irrep[site], i2 = bk.orb_irrep[site + 1];
   bk.for_sectors(site, [&](int nl, int tsl, int il) {
      const int dl = bk.dim(site, nl, tsl, il);
      if (dl <= 0) return;
      for (int n1 = 0; n1 <= 2; n1++)
         for (int n2 = 0; n2 <= 2; n2++) {
            const int nr = nl + n1 + n2;
            const int im = (n1 == 1) ? xorp(il, i1) : il;
            const int ir = (n2 == 1) ? xorp(im, i2) : im;
            const int tjmin = (n1 + n2) % 2;
            const int tjmax = (n1 == 1 && n2 == 1) ? 2 : tjmin;
            for (int tj = tjmin; tj <= tjmax; tj += 2)
               for (int tsr = tsl - tj; tsr <= tsl + tj; tsr += 2) {
                  if (tsr < 0) continue;
                  const int dr = bk.dim(site + 2, nr, tsr, ir);
                  if (dr <= 0) continue;
                  index[skey(bk.slot(site, nl, tsl, il), n1, n2, tj, bk.slot(site + 2, nr, tsr, ir))] = (int)blk.size();
                  NL.push_back(nl); twoSL.push_back(tsl); IL.push_back(il); N1.push_back(n1); N2.push_back(n2);
                  twoJ.push_back(tj); NR.push_back(nr); twoSR.push_back(tsr); IR.push_back(ir);
                  blk.push_back({size, dl, dr});
                  size += (int64_t)dl * dr;
               }
         }
   });
}

int SLayout::kappa(const Bookkeeper& bk, int nl, int tsl, int il, int n1, int n2, int tj, int nr, int tsr, int ir) const {
   if (n1 < 0 || n1 > 2 || n2 < 0 || n2 > 2 || tj < 0 || tj > 2) return -1;
   const int sl = bk.slot(site, nl, tsl, il), sr = bk.slot(site + 2, nr, tsr, ir);
   if (sl < 0 || sr < 0) return -1;
   auto it = index.find(skey(sl, n1, n2, tj, sr));
   return it == index.end() ? -1 : it->second;
}

}   // namespace b2
