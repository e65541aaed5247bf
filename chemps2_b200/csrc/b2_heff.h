// b2_heff.h — compiles a SigmaPlan into device work lists (stage-1 intermediates, stage-2 sigma tiles).
#pragma once
#include <vector>

#include "b2_device.h"
#include "b2_sigma.h"

namespace b2 {

struct CompiledSigma {
   std::vector<GemmItem> items;
   std::vector<Tile> tiles1[kNumTileClasses];   // stage 1: W = op(A)*S[src]  or  S[src]*op(B)
   std::vector<Tile> tiles2[kNumTileClasses];   // stage 2: sigma tiles
   std::vector<PresumJob> presum_jobs;
   std::vector<PresumPart> presum_parts;
   int64_t work_size = 0;                       // doubles
   double flops_exec = 0.0;
   long long n_stage1 = 0, n_tiles = 0, n_terms_used = 0;
};

// rank/world: keep only the terms whose owner == rank (all sigma tiles are kept so that sigma is fully written)
void compile_sigma(CompiledSigma& out, const SigmaPlan& plan, const OpSet* left, const OpSet* right, int rank, int world);

}   // namespace b2
