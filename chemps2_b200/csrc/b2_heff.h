// b2_heff.h — compiles a SigmaPlan into device work lists (through the generic scheduler of b2_compile.h).
#pragma once
#include <vector>

#include "b2_compile.h"
#include "b2_sigma.h"

namespace b2 {

struct CompiledSigma : CompiledWork {
   std::vector<DiagItem> diag_items;            // terms with dst == src: they are the diagonal of H_eff
   std::vector<DiagTile> diag_tiles;
   std::vector<PresumJob> presum_jobs;
   std::vector<PresumPart> presum_parts;
   long long n_terms_used = 0;
};

// rank/world: keep only the terms whose owner == rank
void compile_sigma(CompiledSigma& out, const SigmaPlan& plan, const OpSet* left, const OpSet* right, int rank, int world,
                   const CompileOptions& opt = CompileOptions());

}   // namespace b2
