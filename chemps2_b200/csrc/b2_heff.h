// b2_heff.h — compiles a SigmaPlan into device work lists (stage-1 intermediates, stage-2 sigma tiles, split-K reduces).
#pragma once
#include <vector>

#include "b2_device.h"
#include "b2_sigma.h"

namespace b2 {

// One wave = a contiguous slice of the term list whose stage-1 intermediates fit the workspace budget.
// Launch order per wave: stage-1 tiles (all classes) -> stage-2 tiles (all classes) -> reduce jobs.
struct Wave {
   int t1_begin[kNumTileClasses], t1_end[kNumTileClasses];
   int t2_begin[kNumTileClasses], t2_end[kNumTileClasses];
   int red_begin, red_end;
};

struct CompileOptions {
   int64_t work_budget = (int64_t)1 << 25;   // doubles of stage-1 workspace per wave (256 MiB)
   int64_t chunk_k = 4096;                   // split-K: accumulated inner dimension per CTA
};

struct CompiledSigma {
   std::vector<GemmItem> items1, items2;
   std::vector<Tile> tiles1[kNumTileClasses];   // stage 1: W = op(A)*S[src]  or  S[src]*op(B)
   std::vector<Tile> tiles2[kNumTileClasses];   // stage 2: sigma tiles / partial slots
   std::vector<ReduceJob> reduces;
   std::vector<Wave> waves;
   std::vector<DiagItem> diag_items;            // terms with dst == src: they are the diagonal of H_eff
   std::vector<DiagTile> diag_tiles;
   std::vector<PresumJob> presum_jobs;
   std::vector<PresumPart> presum_parts;
   int64_t work_size = 0;                       // doubles (max over waves)
   int64_t part_size = 0;                       // doubles (max over waves)
   double flops_exec = 0.0;
   long long n_stage1 = 0, n_tiles = 0, n_terms_used = 0;
};

// rank/world: keep only the terms whose owner == rank
void compile_sigma(CompiledSigma& out, const SigmaPlan& plan, const OpSet* left, const OpSet* right, int rank, int world,
                   const CompileOptions& opt = CompileOptions());

}   // namespace b2
