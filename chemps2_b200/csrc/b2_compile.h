// b2_compile.h — generic scheduler: a list of three-factor block contractions -> device work lists.
//
// Both hot loops of the sweep are lists of the same primitive ("one canonical arithmetic form"):
//     C[dst] += f * op(P) * op(Q) * op(R)           any factor may be absent (= identity)
//   sigma build      : P = left renormalized operator block, Q = S[src], R = right operator block   (Heff::makeHeff)
//   operator update  : P = MPS block (transposed when moving right), Q = old operator block, R = MPS block   (TensorOperator::update)
// compile_terms() turns such a list into stage-1 tiles (shared intermediates in a bounded workspace), stage-2 tiles
// (K-concatenated over all terms of a target tile, split-K chunked) and deterministic reduce jobs, wave by wave.
#pragma once
#include <memory>
#include <utility>
#include <vector>

#include "b2_core.h"
#include "b2_device.h"

namespace b2 {

// Work lists are plain structs by the million: a vector that default-initialises (= leaves untouched) what resize() adds, so that a list
// sized for a parallel fill is first touched by the threads that write it instead of being zero-filled by one.
// Blocks of 1 MB and more come from the host block cache (CachedAlloc, b2_core.h): no page faults on re-use, no munmap on release.
template <class T> struct NoInitAlloc : CachedAlloc<T> {
   template <class U> struct rebind { using other = NoInitAlloc<U>; };
   NoInitAlloc() = default;
   template <class U> NoInitAlloc(const NoInitAlloc<U>&) {}
   template <class U> void construct(U* p) { ::new ((void*)p) U; }
   template <class U, class... A> void construct(U* p, A&&... a) { ::new ((void*)p) U(std::forward<A>(a)...); }
};
template <class T> using ListVec = std::vector<T, NoInitAlloc<T>>;

struct MatRef {            // a stored column-major matrix (ld = rows); trans: it enters the product transposed
   uint8_t space = SP_NONE, trans = 0;
   int32_t rows = 0, cols = 0;
   int64_t off = 0;
   bool present() const { return space != SP_NONE; }
   int op_rows() const { return trans ? cols : rows; }
   int op_cols() const { return trans ? rows : cols; }
};

struct Term3 {
   int32_t dst = -1;       // index into the destination block table
   MatRef p, q, r;
   double f = 0.0;
};

struct DstBlock { int64_t off; int32_t rows, cols; };

// One wave = a contiguous slice of the term list whose stage-1 intermediates fit the workspace budget.
// Launch order per wave: stage-1 tiles (all classes) -> stage-2 tiles (all classes) -> reduce jobs.
struct Wave {
   int t1_begin[kNumTileClasses], t1_end[kNumTileClasses];
   int t2_begin[kNumTileClasses], t2_end[kNumTileClasses];
   int red_begin, red_end;
};

struct CompileOptions {
   int64_t work_budget = (int64_t)6 << 30;   // doubles of stage-1 workspace per wave (48 GiB of the 180 GB, capped by half of the free HBM): bigger waves share more stage-1 products (profiles/r1_tuning.md)
   int64_t chunk_k = 2048;                   // split-K: accumulated inner dimension per CTA
   int threads = 1;                          // host threads compile_terms may use (small plans only, see b2_compile.cpp)
   int64_t parallel_min_terms = 20000;       // ... and only when every thread gets at least this many terms
};

struct CompiledWork {
   ListVec<GemmItem> items1, items2;
   ListVec<Tile> tiles1[kNumTileClasses];   // stage 1: W = op(P)*op(Q)  or  op(Q)*op(R)
   ListVec<Tile> tiles2[kNumTileClasses];   // stage 2: destination tiles / split-K partial slots
   ListVec<ReduceJob> reduces;
   std::vector<Wave> waves;
   int64_t work_size = 0;                       // doubles (max over waves)
   int64_t part_size = 0;                       // doubles (max over waves)
   double flops_exec = 0.0;
   long long n_stage1 = 0, n_tiles = 0;
   double launches() const;
   double bytes() const;
};

// `terms` must be grouped by dst (all terms of one destination block contiguous, blocks in any order).
// dst_space: address space of the destination blocks (SP_VOUT for sigma, SP_NEW for operator updates).
void compile_terms(CompiledWork& out, std::vector<Term3>& terms, const std::vector<DstBlock>& dst, uint8_t dst_space,
                   const CompileOptions& opt);
// the same on a raw array (re-ordered in place): for callers that fill the term list on several threads
void compile_terms(CompiledWork& out, Term3* terms, size_t nterms, const std::vector<DstBlock>& dst, uint8_t dst_space,
                   const CompileOptions& opt);

}   // namespace b2
