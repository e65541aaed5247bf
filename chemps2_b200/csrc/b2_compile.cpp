// b2_compile.cpp — see b2_compile.h.  Host only; everything it emits is deterministic (no atomics on the device).
//
// Scheduling:
//   * a three-factor term is split into a stage-1 product W (order chosen per term to minimise FLOPs) kept in a workspace,
//     and a stage-2 product accumulated into the destination tile; identical stage-1 products inside a wave are shared;
//   * the term list is cut into WAVES so that the stage-1 workspace stays below CompileOptions::work_budget;
//   * inside a wave the terms of one destination tile are cut into split-K CHUNKS of ~chunk_k accumulated inner dimension,
//     one CTA each; a tile with one chunk adds straight into the destination, otherwise the chunks write partial slots
//     and a reduce job sums them in a fixed order.
#include "b2_compile.h"
#include "b2_core.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <functional>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>

namespace b2 {

namespace {

int tile_class_for(int m, int n) {
   const int d = std::min(m, n);
   if (d > 32) return 0;
   if (d > 16) return 1;
   if (d > 8) return 2;
   return 3;
}

// balanced tiling: ceil(M / edge) tiles of equal size (rounded up to the 8-row MMA granule) instead of full tiles + a sliver
inline int balanced_step(int M, int edge) {
   const int nt = (M + edge - 1) / edge;
   const int s = ((M + nt - 1) / nt + 7) / 8 * 8;
   return std::min(s, edge);
}

// Stage-1 products shared inside a wave: open-addressing table keyed by the two operand blocks (space, transposition and
// arena offset identify a block).  Generation-stamped slots make "clear" free; millions of look-ups per plan, so no node
// allocations and one cache line per probe.
struct WInfo { int64_t off; int rows, cols; };
class WTable {
   struct Slot { uint64_t k1, k2; uint32_t gen, idx; };
   std::vector<Slot> slots_;
   std::vector<WInfo> vals_;
   uint32_t gen_ = 1;
   size_t mask_ = 0, used_ = 0;
   static uint64_t key_of(const MatRef& m) { return (uint64_t)m.off | ((uint64_t)m.space << 56) | ((uint64_t)m.trans << 63); }
   static size_t hash(uint64_t a, uint64_t b) {
      uint64_t h = a * 0x9E3779B97F4A7C15ULL ^ (b + 0x7F4A7C15ULL) * 0xC2B2AE3D27D4EB4FULL;
      return (size_t)(h ^ (h >> 31));
   }
   void grow() {
      std::vector<Slot> old;
      old.swap(slots_);
      slots_.assign(old.empty() ? (size_t)1 << 16 : old.size() * 2, Slot{0, 0, 0, 0});
      mask_ = slots_.size() - 1;
      for (const Slot& s : old)
         if (s.gen == gen_) {
            size_t i = hash(s.k1, s.k2) & mask_;
            while (slots_[i].gen == gen_) i = (i + 1) & mask_;
            slots_[i] = s;
         }
   }
public:
   void clear() {
      if (++gen_ == 0) { for (Slot& s : slots_) s.gen = 0; gen_ = 1; }
      used_ = 0; vals_.clear();
   }
   // returns the stored product or nullptr; `slot` receives the insertion position for put()
   const WInfo* find(const MatRef& a, const MatRef& b, size_t& slot) {
      if ((used_ + 1) * 10 >= slots_.size() * 7) grow();
      const uint64_t k1 = key_of(a), k2 = key_of(b);
      size_t i = hash(k1, k2) & mask_;
      while (slots_[i].gen == gen_) {
         if (slots_[i].k1 == k1 && slots_[i].k2 == k2) return &vals_[slots_[i].idx];
         i = (i + 1) & mask_;
      }
      slot = i;
      return nullptr;
   }
   void put(size_t slot, const MatRef& a, const MatRef& b, const WInfo& w) {
      slots_[slot] = Slot{key_of(a), key_of(b), gen_, (uint32_t)vals_.size()};
      vals_.push_back(w);
      used_++;
   }
};

inline void set_x(GemmItem& g, const MatRef& m) { g.xs = m.space; g.xoff = m.off; g.ldx = m.rows; if (m.trans) g.flags |= IF_TX; }
inline void set_y(GemmItem& g, const MatRef& m) { g.ys = m.space; g.yoff = m.off; g.ldy = m.rows; if (m.trans) g.flags |= IF_TY; }

}   // namespace

double CompiledWork::launches() const {
   double n = 0.0;
   for (const Wave& w : waves) {
      for (int c = 0; c < kNumTileClasses; c++) n += (w.t1_end[c] > w.t1_begin[c]) + (w.t2_end[c] > w.t2_begin[c]);
      n += (w.red_end > w.red_begin);
   }
   return n;
}
double CompiledWork::bytes() const {
   double b = sizeof(GemmItem) * (double)(items1.size() + items2.size()) + sizeof(ReduceJob) * (double)reduces.size();
   for (int c = 0; c < kNumTileClasses; c++) b += sizeof(Tile) * (double)(tiles1[c].size() + tiles2[c].size());
   return b;
}

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
namespace {

// the launch order of the tiles [b, e) of one wave: see weight_sort in compile_range
void group_sort(ListVec<Tile>& v, int b, int e, const ListVec<GemmItem>& items);

// compiles terms[0, n) (grouped by dst) into `out`; sort_waves = false leaves the launch order to the caller
void compile_range(CompiledWork& out, Term3* terms, size_t nterms, const std::vector<DstBlock>& dst, uint8_t dst_space, const CompileOptions& opt,
                   bool sort_waves);

}   // namespace

void compile_terms(CompiledWork& out, std::vector<Term3>& terms, const std::vector<DstBlock>& dst, uint8_t dst_space, const CompileOptions& opt) {
   compile_terms(out, terms.data(), terms.size(), dst, dst_space, opt);
}

void compile_terms(CompiledWork& out, Term3* term_data, size_t nterms, const std::vector<DstBlock>& dst, uint8_t dst_space, const CompileOptions& opt) {
   out = CompiledWork();
   const double T_total = now_s();
   struct Span {   // what the code below needs of a vector
      Term3* p; size_t n;
      size_t size() const { return n; }
      Term3* data() const { return p; }
      Term3& operator[](size_t i) const { return p[i]; }
      Term3* begin() const { return p; }
      Term3* end() const { return p + n; }
   } terms{term_data, nterms};
   {  // every destination block must form ONE contiguous group: two groups would become two CTAs that read-modify-write the same
      // tile in one launch.  Generators that visit a block twice (e.g. TensorQ/TensorX: update + AddTerms) are regrouped here.
      std::vector<char> seen(dst.size(), 0);
      bool grouped = true;
      for (size_t i = 0; i < terms.size() && grouped; i++) {
         if (i > 0 && terms[i].dst == terms[i - 1].dst) continue;
         if (seen[terms[i].dst]) grouped = false;
         seen[terms[i].dst] = 1;
      }
      if (!grouped) {   // stable counting sort by destination block (ids are dense): one pass to count, one to scatter
         std::vector<size_t> pos(dst.size() + 1, 0);
         for (const Term3& t : terms) pos[t.dst + 1]++;
         for (size_t k = 0; k < dst.size(); k++) pos[k + 1] += pos[k];
         static_assert(std::is_trivially_copyable<Term3>::value, "Term3 is moved with memcpy");
         const size_t bytes = sizeof(Term3) * terms.size();
         Term3* sorted = static_cast<Term3*>(host_block_acquire(bytes));   // recycled block: no page faults, no constructor pass
         for (const Term3& t : terms) std::memcpy(static_cast<void*>(sorted + pos[t.dst]++), &t, sizeof(Term3));
         std::memcpy(static_cast<void*>(terms.data()), sorted, bytes);
         host_block_release(sorted, bytes);
      }
   }
   const int T = std::max(1, std::min<int>(opt.threads, (int)(terms.size() / std::max<int64_t>(opt.parallel_min_terms, 1))));
   if (T <= 1) {
      compile_range(out, terms.data(), terms.size(), dst, dst_space, opt, true);
   } else {
      // Small plans are dominated by the time to BUILD them (the tiny-block regime of a first sweep): the term list is cut at
      // destination-block boundaries into T segments that are compiled concurrently, each with 1/T of the workspace budget, and
      // merged wave by wave (workspace / partial-slot / item offsets shifted per segment).  Stage-1 products are then shared only
      // inside a segment, which is why large plans (where the FLOPs matter, not the planning) keep the sequential path.
      std::vector<size_t> cut(T + 1, terms.size());
      cut[0] = 0;
      for (int t = 1; t < T; t++) {
         size_t i = std::max(cut[t - 1], terms.size() * t / T);
         while (i < terms.size() && i > 0 && terms[i].dst == terms[i - 1].dst) i++;
         cut[t] = i;
      }
      std::vector<CompiledWork> seg(T);
      CompileOptions o = opt;
      o.work_budget = std::max<int64_t>(opt.work_budget / T, 1 << 20);
      auto run_all = [&](const std::function<void(int)>& fn) { parallel_run(T, fn); };
      // every segment orders its own waves (heaviest group first); the merged wave is the concatenation of the segments' waves
      const double t_seg0 = now_s();
      run_all([&](int t) { compile_range(seg[t], terms.data() + cut[t], cut[t + 1] - cut[t], dst, dst_space, o, true); });
      const double t_seg1 = now_s();
      std::vector<int64_t> wbase(T + 1, 0), pbase(T + 1, 0);
      std::vector<int> i1base(T + 1, 0), i2base(T + 1, 0);
      size_t nwaves = 0;
      for (int t = 0; t < T; t++) {
         wbase[t + 1] = wbase[t] + (seg[t].work_size + 15) / 16 * 16;
         pbase[t + 1] = pbase[t] + (seg[t].part_size + 15) / 16 * 16;
         i1base[t + 1] = i1base[t] + (int)seg[t].items1.size();
         i2base[t + 1] = i2base[t] + (int)seg[t].items2.size();
         nwaves = std::max(nwaves, seg[t].waves.size());
         out.flops_exec += seg[t].flops_exec; out.n_stage1 += seg[t].n_stage1;
      }
      out.work_size = wbase[T]; out.part_size = pbase[T];
      // positions of every (wave, segment) slice in the merged tile / reduce lists
      struct Pos { int t1[kNumTileClasses], t2[kNumTileClasses], red; };
      std::vector<std::vector<Pos>> pos(nwaves, std::vector<Pos>(T));
      int n1[kNumTileClasses] = {}, n2[kNumTileClasses] = {}, nred = 0;
      out.waves.resize(nwaves);
      for (size_t w = 0; w < nwaves; w++) {
         Wave& gw = out.waves[w];
         for (int c = 0; c < kNumTileClasses; c++) { gw.t1_begin[c] = n1[c]; gw.t2_begin[c] = n2[c]; }
         gw.red_begin = nred;
         for (int t = 0; t < T; t++) {
            Pos& ps = pos[w][t];
            for (int c = 0; c < kNumTileClasses; c++) { ps.t1[c] = n1[c]; ps.t2[c] = n2[c]; }
            ps.red = nred;
            if (w >= seg[t].waves.size()) continue;
            const Wave& sw = seg[t].waves[w];
            for (int c = 0; c < kNumTileClasses; c++) { n1[c] += sw.t1_end[c] - sw.t1_begin[c]; n2[c] += sw.t2_end[c] - sw.t2_begin[c]; }
            nred += sw.red_end - sw.red_begin;
         }
         for (int c = 0; c < kNumTileClasses; c++) { gw.t1_end[c] = n1[c]; gw.t2_end[c] = n2[c]; }
         gw.red_end = nred;
      }
      {  // the merged lists are tens of MB of fresh memory: size them on several threads so the first-touch page faults overlap
         std::vector<std::function<void()>> sizing;
         sizing.push_back([&] { out.items1.resize(i1base[T]); });
         sizing.push_back([&] { out.items2.resize(i2base[T]); });
         sizing.push_back([&] { out.reduces.resize(nred); });
         for (int c = 0; c < kNumTileClasses; c++) {
            sizing.push_back([&, c] { out.tiles1[c].resize(n1[c]); });
            sizing.push_back([&, c] { out.tiles2[c].resize(n2[c]); });
         }
         std::atomic<int> next_sizing{0};
         run_all([&](int) { for (int i; (i = next_sizing.fetch_add(1)) < (int)sizing.size();) sizing[i](); });
      }
      const double t_seg2 = now_s();
      run_all([&](int t) {   // every segment copies its own slices, shifting item indices and workspace / partial-slot offsets
         std::copy(seg[t].items1.begin(), seg[t].items1.end(), out.items1.begin() + i1base[t]);
         for (size_t i = 0; i < seg[t].items2.size(); i++) {
            GemmItem g = seg[t].items2[i];
            if (g.xs == SP_WORK) g.xoff += wbase[t];
            if (g.ys == SP_WORK) g.yoff += wbase[t];
            out.items2[i2base[t] + i] = g;
         }
         for (size_t w = 0; w < seg[t].waves.size(); w++) {
            const Wave& sw = seg[t].waves[w];
            const Pos& ps = pos[w][t];
            for (int c = 0; c < kNumTileClasses; c++) {
               for (int i = sw.t1_begin[c]; i < sw.t1_end[c]; i++) {
                  Tile x = seg[t].tiles1[c][i];
                  x.item_begin += i1base[t]; x.item_end += i1base[t];
                  if (x.cspace == SP_WORK) x.coff += wbase[t];
                  out.tiles1[c][ps.t1[c] + (i - sw.t1_begin[c])] = x;
               }
               for (int i = sw.t2_begin[c]; i < sw.t2_end[c]; i++) {
                  Tile x = seg[t].tiles2[c][i];
                  x.item_begin += i2base[t]; x.item_end += i2base[t];
                  if (x.cspace == SP_PART) x.coff += pbase[t];
                  out.tiles2[c][ps.t2[c] + (i - sw.t2_begin[c])] = x;
               }
            }
            for (int i = sw.red_begin; i < sw.red_end; i++) {
               ReduceJob r = seg[t].reduces[i];
               r.part_off += pbase[t];
               out.reduces[ps.red + (i - sw.red_begin)] = r;
            }
         }
         seg[t] = CompiledWork();
      });
      if (getenv("B2_TIMING")) fprintf(stderr, "compile_terms: regroup %.3f s, segments %.3f s, merge %.3f s (allocation %.3f s)\n", t_seg0 - T_total, t_seg1 - t_seg0, now_s() - t_seg1, t_seg2 - t_seg1);
   }
   for (int c = 0; c < kNumTileClasses; c++) out.n_tiles += (long long)out.tiles1[c].size() + (long long)out.tiles2[c].size();
   if (getenv("B2_TIMING")) fprintf(stderr, "compile_terms: total %.3f s, %d thread(s)\n", now_s() - T_total, T);
}

namespace {

void group_sort(ListVec<Tile>& v, int b, int e, const ListVec<GemmItem>& items) {
   // Launch order = heaviest GROUPS first (static load balance across the SMs), where a group is the set of CTAs that
   // stream the same items (all tiles of one stage-1 product / of one split-K chunk of a destination block): they read the
   // same operand panels, so they must be resident together for the panels to be served by L2 instead of DRAM.
   // Total order: (group weight descending, first item of the group ascending, emission index ascending).  Millions of tiles
   // per plan, so no comparison sort: the groups (dense in their first item) are ordered by a stable LSD radix sort on the
   // weight, the tiles of a group keep their emission order.
   if (e <= b) return;
   const int n = e - b;
   int ib_lo = v[b].item_begin, ib_hi = v[b].item_begin;
   for (int i = b; i < e; i++) { ib_lo = std::min(ib_lo, v[i].item_begin); ib_hi = std::max(ib_hi, v[i].item_begin); }
   const size_t G = (size_t)(ib_hi - ib_lo) + 1;
   std::vector<long long> wmax(G, 0);      // group weight = its heaviest tile, indexed by the group's first item
   std::vector<int> count(G, 0);           // tiles per group
   {
      std::vector<long long> ksum(G, -1);  // accumulated inner dimension of the group's items (the same for every tile of a group)
      std::vector<int> kend(G, 0);
      for (int i = b; i < e; i++) {
         const Tile& t = v[i];
         const size_t g = (size_t)(t.item_begin - ib_lo);
         long long ks;
         if (ksum[g] >= 0 && kend[g] == t.item_end) ks = ksum[g];
         else {
            ks = 0;
            for (int it = t.item_begin; it < t.item_end; it++) ks += items[it].k + 4;
            ksum[g] = ks; kend[g] = t.item_end;
         }
         const long long w = ks * ((long long)((t.mrem + 7) / 8) * ((t.nrem + 7) / 8));
         wmax[g] = std::max(wmax[g], w);
         count[g]++;
      }
   }
   std::vector<uint32_t> ord, tmp;         // the non-empty groups, ascending in their first item
   ord.reserve(G);
   long long wtop = 0;
   for (size_t g = 0; g < G; g++) if (count[g]) { ord.push_back((uint32_t)g); wtop = std::max(wtop, wmax[g]); }
   tmp.resize(ord.size());
   constexpr int kBits = 11;
   for (int shift = 0; shift < 63 && (wtop >> shift) != 0; shift += kBits) {   // stable, descending in the current digit
      size_t hist[(1 << kBits) + 1] = {};
      for (uint32_t g : ord) hist[(1 << kBits) - 1 - (size_t)((wmax[g] >> shift) & ((1 << kBits) - 1)) + 1]++;
      for (int d = 0; d < (1 << kBits); d++) hist[d + 1] += hist[d];
      for (uint32_t g : ord) tmp[hist[(1 << kBits) - 1 - (size_t)((wmax[g] >> shift) & ((1 << kBits) - 1))]++] = g;
      ord.swap(tmp);
   }
   std::vector<int> start(G, 0);
   int at = 0;
   for (uint32_t g : ord) { start[g] = at; at += count[g]; }
   ListVec<Tile> sorted(n);
   for (int i = b; i < e; i++) sorted[start[(size_t)(v[i].item_begin - ib_lo)]++] = v[i];
   std::copy(sorted.begin(), sorted.end(), v.begin() + b);
}

void compile_range(CompiledWork& out, Term3* terms, size_t nterms, const std::vector<DstBlock>& dst, uint8_t dst_space, const CompileOptions& opt,
                   bool sort_waves) {
   out = CompiledWork();
   WTable wmap;
   out.items2.reserve(nterms);
   out.items1.reserve(nterms / 2);
   int64_t wave_work = 0, wave_part = 0;
   Wave wave{};
   auto open_wave = [&]() {
      for (int c = 0; c < kNumTileClasses; c++) {
         wave.t1_begin[c] = (int)out.tiles1[c].size();
         wave.t2_begin[c] = (int)out.tiles2[c].size();
      }
      wave.red_begin = (int)out.reduces.size();
      wave_work = 0; wave_part = 0;
      wmap.clear();
   };
   auto weight_sort = [&](ListVec<Tile>& v, int b, int e, const ListVec<GemmItem>& items) { if (sort_waves) group_sort(v, b, e, items); };
   auto close_wave = [&]() {
      bool any = (int)out.reduces.size() > wave.red_begin;
      for (int c = 0; c < kNumTileClasses; c++) {
         wave.t1_end[c] = (int)out.tiles1[c].size();
         wave.t2_end[c] = (int)out.tiles2[c].size();
         any = any || wave.t1_end[c] > wave.t1_begin[c] || wave.t2_end[c] > wave.t2_begin[c];
         weight_sort(out.tiles1[c], wave.t1_begin[c], wave.t1_end[c], out.items1);
         weight_sort(out.tiles2[c], wave.t2_begin[c], wave.t2_end[c], out.items2);
      }
      wave.red_end = (int)out.reduces.size();
      out.work_size = std::max(out.work_size, wave_work);
      out.part_size = std::max(out.part_size, wave_part);
      if (any) out.waves.push_back(wave);
   };

   // W = op(a) * op(b), shared inside the wave
   auto get_w = [&](const MatRef& a, const MatRef& b) -> WInfo {
      size_t slot = 0;
      if (const WInfo* hit = wmap.find(a, b, slot)) return *hit;
      WInfo w{};
      w.rows = a.op_rows(); w.cols = b.op_cols();
      GemmItem g{};
      g.alpha = 1.0; g.k = a.op_cols();
      set_x(g, a); set_y(g, b);
      w.off = wave_work;
      wave_work += ((int64_t)w.rows * w.cols + 15) / 16 * 16;
      const int ib = (int)out.items1.size();
      out.items1.push_back(g);
      const int cls = tile_class_for(w.rows, w.cols), e = kTileEdge[cls];
      const int sm = balanced_step(w.rows, e), sn = balanced_step(w.cols, e);
      for (int n0 = 0; n0 < w.cols; n0 += sn)
         for (int m0 = 0; m0 < w.rows; m0 += sm) {
            Tile t{};
            t.coff = w.off; t.ldc = w.rows; t.m0 = t.cm0 = m0; t.n0 = t.cn0 = n0;
            t.mrem = std::min(sm, w.rows - m0); t.nrem = std::min(sn, w.cols - n0);
            t.item_begin = ib; t.item_end = ib + 1; t.cspace = SP_WORK; t.accumulate = 0;
            out.tiles1[cls].push_back(t);
         }
      out.flops_exec += 2.0 * w.rows * w.cols * g.k;
      out.n_stage1++;
      wmap.put(slot, a, b, w);
      return w;
   };

   // emit the stage-2 CTAs of items2[ib, ie) for destination block d (all inside the current wave)
   auto emit_block = [&](const DstBlock& db, int ib, int ie) {
      if (ie <= ib) return;
      const int M = db.rows, N = db.cols;
      std::vector<int> cuts{ib};
      int64_t acc = 0;
      for (int i = ib; i < ie; i++) {
         acc += out.items2[i].k + 4;
         if (acc >= opt.chunk_k && i + 1 < ie) { cuts.push_back(i + 1); acc = 0; }
      }
      cuts.push_back(ie);
      const int nchunks = (int)cuts.size() - 1;
      const int cls = tile_class_for(M, N), e = kTileEdge[cls];
      const int sm = balanced_step(M, e), sn = balanced_step(N, e);
      for (int n0 = 0; n0 < N; n0 += sn)
         for (int m0 = 0; m0 < M; m0 += sm) {
            const int mrem = std::min(sm, M - m0), nrem = std::min(sn, N - n0);
            if (nchunks == 1) {
               Tile t{};
               t.coff = db.off; t.ldc = M; t.m0 = t.cm0 = m0; t.n0 = t.cn0 = n0; t.mrem = mrem; t.nrem = nrem;
               t.item_begin = ib; t.item_end = ie; t.cspace = dst_space; t.accumulate = 1;
               out.tiles2[cls].push_back(t);
            } else {
               const int64_t stride = ((int64_t)mrem * nrem + 15) / 16 * 16;
               ReduceJob r{};
               r.dst_off = db.off; r.ldc = M; r.m0 = m0; r.n0 = n0; r.mrem = mrem; r.nrem = nrem;
               r.part_off = wave_part; r.nparts = nchunks; r.part_stride = stride; r.dst_space = dst_space;
               out.reduces.push_back(r);
               for (int c = 0; c < nchunks; c++) {
                  Tile t{};
                  t.coff = wave_part + c * stride; t.ldc = mrem; t.m0 = m0; t.n0 = n0; t.cm0 = 0; t.cn0 = 0; t.mrem = mrem; t.nrem = nrem;
                  t.item_begin = cuts[c]; t.item_end = cuts[c + 1]; t.cspace = SP_PART; t.accumulate = 0;
                  out.tiles2[cls].push_back(t);
               }
               wave_part += stride * nchunks;
            }
         }
   };

   const bool prof = getenv("B2_TIMING2") != nullptr;
   double t_terms = 0.0, t_emit = 0.0, t_close = 0.0;
   open_wave();
   size_t i0 = 0;
   while (i0 < nterms) {
      const double tq0 = prof ? now_s() : 0.0;
      size_t i1 = i0;
      while (i1 < nterms && terms[i1].dst == terms[i0].dst) i1++;
      // block-axpy terms first inside every destination block: the kernel consumes them before it starts its GEMM pipeline
      std::stable_partition(terms + i0, terms + i1, [](const Term3& t) { return !t.p.present() && !t.r.present(); });
      const DstBlock& db = dst[terms[i0].dst];
      const int M = db.rows, N = db.cols;
      int ib = (int)out.items2.size();
      for (size_t i = i0; i < i1; i++) {
         const Term3& t = terms[i];
         GemmItem g{};
         g.alpha = t.f;
         const bool hp = t.p.present(), hq = t.q.present(), hr = t.r.present();
         if (hp && hq && hr) {
            const double kq1 = t.q.op_rows(), kq2 = t.q.op_cols();
            const double f_left = (double)M * kq1 * kq2 + (double)M * kq2 * N;     // (P*Q)*R
            const double f_right = kq1 * kq2 * N + (double)M * kq1 * N;            // P*(Q*R)
            if (f_left <= f_right) {
               const WInfo w = get_w(t.p, t.q);
               g.xs = SP_WORK; g.xoff = w.off; g.ldx = w.rows; set_y(g, t.r); g.k = w.cols;
            } else {
               const WInfo w = get_w(t.q, t.r);
               set_x(g, t.p); g.ys = SP_WORK; g.yoff = w.off; g.ldy = w.rows; g.k = w.rows;
            }
         } else if (hp && hq) { set_x(g, t.p); set_y(g, t.q); g.k = t.p.op_cols(); }
         else if (hq && hr) { set_x(g, t.q); set_y(g, t.r); g.k = t.q.op_cols(); }
         else if (hp && hr) { set_x(g, t.p); set_y(g, t.r); g.k = t.p.op_cols(); }
         else if (hq) { set_x(g, t.q); g.flags |= IF_AXPY; g.k = 0; }     // f * op(Q)
         else continue;                                                   // nothing to add
         out.flops_exec += (g.flags & IF_AXPY) ? 2.0 * M * N : 2.0 * M * N * g.k;
         out.items2.push_back(g);
         if (wave_work >= opt.work_budget) {   // workspace full: flush what this block has so far and start a new wave
            emit_block(db, ib, (int)out.items2.size());
            close_wave();
            open_wave();
            ib = (int)out.items2.size();
         }
      }
      const double tq1 = prof ? now_s() : 0.0;
      emit_block(db, ib, (int)out.items2.size());
      if (prof) { t_terms += tq1 - tq0; t_emit += now_s() - tq1; }
      i0 = i1;
   }
   const double tq2 = prof ? now_s() : 0.0;
   close_wave();
   if (prof) { t_close = now_s() - tq2; fprintf(stderr, "compile_range: terms %.3f s, emit %.3f s, close (sort) %.3f s\n", t_terms, t_emit, t_close); }
}

}   // namespace

}   // namespace b2
