// b2_pool.cpp — see b2_pool.h
#define B2_POOL_IMPL
#include "b2_pool.h"

#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace b2 {

namespace {

struct Block { void* p; size_t bytes; int device; unsigned long long stamp; std::vector<cudaEvent_t> events; };

struct Pool {
   std::mutex mtx;
   std::unordered_map<void*, std::pair<size_t, int>> live;        // pointer -> (class size, device)
   std::multimap<std::pair<int, size_t>, Block> cache;            // (device, class size) -> block
   std::vector<std::pair<int, cudaStream_t>> streams;             // registered streams with their device
   std::vector<cudaEvent_t> spare_events;
   size_t cached = 0;
   unsigned long long clock = 0;
   bool disabled = std::getenv("B2_NO_POOL") != nullptr;
};
Pool& pool() { static Pool* p = new Pool; return *p; }   // never destroyed: frees may arrive from static destructors

size_t size_class(size_t n) {   // next multiple of 2^(floor(log2 n) - 3): at most 12.5 % above the request, 512-byte granularity below 4 KiB
   if (n <= 4096) return (n + 511) / 512 * 512;
   size_t step = 1;
   while ((step << 4) <= n) step <<= 1;   // step = 2^(floor(log2 n) - 3)
   return (n + step - 1) / step * step;
}

size_t cache_budget(int device) {
   static size_t budget[64] = {};
   if (device < 0 || device >= 64) return 0;
   if (!budget[device]) {
      size_t free_b = 0, total_b = 0;
      if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) total_b = (size_t)16 << 30;
      const char* e = std::getenv("B2_POOL_CACHE_GB");
      budget[device] = e ? (size_t)(std::atof(e) * 1073741824.0) : total_b / 4;
   }
   return budget[device];
}

void release_block(Pool& P, Block& b) {   // mtx held
   for (cudaEvent_t ev : b.events) { cudaEventSynchronize(ev); P.spare_events.push_back(ev); }
   b.events.clear();
   int cur = 0;
   cudaGetDevice(&cur);
   if (cur != b.device) cudaSetDevice(b.device);
   ::cudaFree(b.p);
   if (cur != b.device) cudaSetDevice(cur);
   P.cached -= b.bytes;
}

void trim_locked(Pool& P, int device, size_t keep_bytes) {   // evict least-recently-freed blocks of `device` (all devices: -1) down to keep_bytes
   while (P.cached > keep_bytes && !P.cache.empty()) {
      auto victim = P.cache.end();
      for (auto it = P.cache.begin(); it != P.cache.end(); ++it)
         if ((device < 0 || it->second.device == device) && (victim == P.cache.end() || it->second.stamp < victim->second.stamp)) victim = it;
      if (victim == P.cache.end()) break;
      release_block(P, victim->second);
      P.cache.erase(victim);
   }
}

}   // namespace

void pool_register_stream(cudaStream_t s) {
   Pool& P = pool();
   std::lock_guard<std::mutex> g(P.mtx);
   int dev = 0;
   cudaGetDevice(&dev);
   for (auto& x : P.streams) if (x.second == s && x.first == dev) return;
   P.streams.push_back({dev, s});
}
void pool_unregister_stream(cudaStream_t s) {
   Pool& P = pool();
   std::lock_guard<std::mutex> g(P.mtx);
   P.streams.erase(std::remove_if(P.streams.begin(), P.streams.end(), [&](const std::pair<int, cudaStream_t>& x) { return x.second == s; }), P.streams.end());
}

cudaError_t pool_malloc(void** p, size_t bytes) {
   Pool& P = pool();
   if (P.disabled) return ::cudaMalloc(p, bytes);
   *p = nullptr;
   if (bytes == 0) bytes = 1;
   const size_t cls = size_class(bytes);
   int dev = 0;
   cudaError_t e = cudaGetDevice(&dev);
   if (e != cudaSuccess) return e;
   std::lock_guard<std::mutex> g(P.mtx);
   // smallest cached block of this device that holds the request and is at most twice its size: workspaces and work lists change size
   // from site to site, and with exact-class matching a D = 2000 sweep filled the cache with near misses — every free then evicted a
   // GB-sized block through the driver (device-wide synchronising cudaFree) and every allocation went to cudaMalloc
   // (profiles/r2y_sweep_timing.log: 0.8 s "release" + 0.6-1.2 s "update epilogue" per half sweep, both ~0 at D = 1000)
   auto it = P.cache.lower_bound({dev, cls});
   if (it != P.cache.end() && it->first.first == dev && it->first.second <= 2 * cls) {
      Block b = std::move(it->second);
      P.cache.erase(it);
      P.cached -= b.bytes;
      for (cudaEvent_t ev : b.events) {   // work that was in flight when the block was freed must finish before any registered stream touches it
         for (auto& st : P.streams) if (st.first == dev) cudaStreamWaitEvent(st.second, ev, 0);
         P.spare_events.push_back(ev);
      }
      P.live[b.p] = {b.bytes, dev};   // the block keeps its own size class
      *p = b.p;
      return cudaSuccess;
   }
   e = ::cudaMalloc(p, cls);
   if (e != cudaSuccess) {   // out of memory with blocks in the cache: give them back and try once more
      cudaGetLastError();
      trim_locked(P, dev, 0);
      e = ::cudaMalloc(p, cls);
      if (e != cudaSuccess) return e;
   }
   P.live[*p] = {cls, dev};
   return cudaSuccess;
}

cudaError_t pool_free(void* p) {
   if (!p) return cudaSuccess;
   Pool& P = pool();
   std::unique_lock<std::mutex> g(P.mtx);
   auto it = P.live.find(p);
   if (it == P.live.end()) { g.unlock(); return ::cudaFree(p); }   // not ours (allocated with the pool disabled)
   Block b;
   b.p = p; b.bytes = it->second.first; b.device = it->second.second; b.stamp = ++P.clock;
   P.live.erase(it);
   int cur = 0;
   cudaGetDevice(&cur);
   if (cur != b.device) cudaSetDevice(b.device);
   for (auto& st : P.streams) {
      if (st.first != b.device) continue;
      cudaEvent_t ev = nullptr;
      if (!P.spare_events.empty()) { ev = P.spare_events.back(); P.spare_events.pop_back(); }
      else if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { ev = nullptr; cudaGetLastError(); }
      if (ev && cudaEventRecord(ev, st.second) == cudaSuccess) b.events.push_back(ev);
      else {   // cannot order the reuse behind this stream: fall back to a synchronising free
         if (ev) P.spare_events.push_back(ev);
         cudaGetLastError();
         for (cudaEvent_t x : b.events) P.spare_events.push_back(x);
         if (cur != b.device) cudaSetDevice(cur);
         g.unlock();
         return ::cudaFree(p);
      }
   }
   if (cur != b.device) cudaSetDevice(cur);
   const size_t budget = cache_budget(b.device);
   if (b.bytes > budget) { P.cached += b.bytes; release_block(P, b); return cudaSuccess; }
   P.cached += b.bytes;
   const int dev = b.device;
   P.cache.insert({{b.device, b.bytes}, std::move(b)});
   if (P.cached > budget) trim_locked(P, dev, budget);
   return cudaSuccess;
}

void pool_trim() {
   Pool& P = pool();
   std::lock_guard<std::mutex> g(P.mtx);
   trim_locked(P, -1, 0);
}
namespace {
struct PinnedCache {
   std::mutex mtx;
   std::vector<std::pair<void*, size_t>> idle;                 // (buffer, capacity)
   std::unordered_map<void*, size_t> out;                      // handed out -> capacity
   size_t idle_bytes = 0;
};
PinnedCache& pinned() { static PinnedCache* c = new PinnedCache; return *c; }
constexpr size_t kPinnedKeep = (size_t)2 << 30;               // idle pinned memory kept at most
}   // namespace
void* pinned_acquire(size_t bytes) {
   PinnedCache& C = pinned();
   bytes = std::max<size_t>(bytes, 4096);
   {
      std::lock_guard<std::mutex> g(C.mtx);
      int best = -1;
      for (int i = 0; i < (int)C.idle.size(); i++)              // smallest idle buffer that fits
         if (C.idle[i].second >= bytes && (best < 0 || C.idle[i].second < C.idle[best].second)) best = i;
      if (best >= 0) {
         std::pair<void*, size_t> b = C.idle[best];
         C.idle.erase(C.idle.begin() + best);
         C.idle_bytes -= b.second;
         C.out[b.first] = b.second;
         return b.first;
      }
   }
   const size_t cap = size_class(bytes + bytes / 4);           // head room: the next Split is rarely exactly this size
   void* p = nullptr;
   if (cudaMallocHost(&p, cap) != cudaSuccess) { cudaGetLastError(); return nullptr; }
   std::lock_guard<std::mutex> g(C.mtx);
   C.out[p] = cap;
   return p;
}
void pinned_release(void* p) {
   if (!p) return;
   PinnedCache& C = pinned();
   std::vector<void*> drop;
   {
      std::lock_guard<std::mutex> g(C.mtx);
      auto it = C.out.find(p);
      if (it == C.out.end()) return;
      C.idle.push_back({p, it->second});
      C.idle_bytes += it->second;
      C.out.erase(it);
      while (C.idle_bytes > kPinnedKeep && !C.idle.empty()) {   // oldest first
         drop.push_back(C.idle.front().first);
         C.idle_bytes -= C.idle.front().second;
         C.idle.erase(C.idle.begin());
      }
   }
   for (void* q : drop) cudaFreeHost(q);
}
size_t pool_cached_bytes() {
   Pool& P = pool();
   std::lock_guard<std::mutex> g(P.mtx);
   return P.cached;
}

}   // namespace b2
