// b2_capi_internal.h — what the translation units of the C ABI (b2_capi*.cpp) share: the opaque handle types behind
// include/chemps2_b200.h, error reporting and the few helpers used across files.  Not installed, not part of the ABI.
#pragma once
#include <cuda_runtime.h>

#include "b2_pool.h"

#include <cstdarg>
#include <cstdio>
#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/chemps2_b200.h"
#include "b2_core.h"
#include "b2_davidson.h"
#include "b2_device.h"
#include "b2_heff.h"
#include "b2_ops.h"
#include "b2_sigma.h"
#include "b2_sobject.h"
#include "b2_twodm.h"
#include "b2_update.h"

using namespace b2;             // internal header, included only by b2_capi*.cpp

namespace b2capi {
inline double wall_seconds() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int fail(int code, const char* fmt, ...);      // records the message b2_last_error() returns (per thread) and passes the code through
}   // namespace b2capi
using namespace b2capi;

#define CUDA_TRY(call)                                                                         \
   do {                                                                                        \
      cudaError_t e_ = (call);                                                                 \
      if (e_ != cudaSuccess) return fail(B2_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
   } while (0)

struct b2_ctx {
   int device = -1;
   cudaStream_t stream = nullptr;
   bool own_stream = true;
   Problem prob;
   Bookkeeper bk;
   bool have_problem = false, have_bk = false;
   CompileOptions copt;
   int simulate_oom = 0;                // test hook: the next N device allocations of operator sets / plans report B2_ERR_CUDA
   // plans below this many reference FLOPs per apply are scheduled on all host cores: segment-wise scheduling shares stage-1 products
   // only inside a segment (+1-3 % executed FLOPs, measured on the N2/cc-pVDZ D=2000 and tetracene D=3000 shapes) but builds the plan
   // 2-4x faster, which wins as long as a Davidson solve (~15 sigma builds) is shorter than the planning it saves
   double parallel_plan_flops = 1e13;
   int davidson_max_matvec = 5000;          // safety net of the device Davidson (the reference loops until convergence)
};

struct b2_opset {
   b2_ctx* ctx = nullptr;
   OpSet set;
   std::vector<double> host;   // host mirror, allocated on first use (upload / download / planning-only contexts)
   double* dev = nullptr;
   double* spill = nullptr;    // pinned host copy while the set is offloaded (b2_opset_offload): the device arena is released
   bool offloaded = false;
   std::string spill_file;     // second tier (b2_opset_offload_file): the arena lives in this file, neither HBM nor host memory is held
   void ensure_host() { if (host.size() != (size_t)set.size) host.assign((size_t)set.size, 0.0); }
   ~b2_opset() { if (spill) cudaFreeHost(spill); if (dev) cudaFree(dev); if (!spill_file.empty()) std::remove(spill_file.c_str()); }
};

struct b2_heff {
   b2_ctx* ctx = nullptr;
   b2_opset *left = nullptr, *right = nullptr;
   SigmaPlan plan;
   CompiledSigma comp;
   // device copies
   GemmItem *d_items1 = nullptr, *d_items2 = nullptr;
   ReduceJob* d_reduces = nullptr;
   DiagItem* d_diag_items = nullptr;
   DiagTile* d_diag_tiles = nullptr;
   int64_t* d_blk_off = nullptr;                 // Sobject block offsets (nkappa + 1)
   double *d_p2s = nullptr, *d_s2p = nullptr;    // sqrt(2SR+1) and its inverse per block (Sobject.cpp:624-650)
   double* d_part = nullptr;
   Tile* d_tiles1[kNumTileClasses] = {nullptr, nullptr, nullptr, nullptr};
   Tile* d_tiles2[kNumTileClasses] = {nullptr, nullptr, nullptr, nullptr};
   PresumJob* d_jobs = nullptr;
   PresumPart* d_parts = nullptr;
   double *d_presum = nullptr, *d_work = nullptr, *d_vin = nullptr, *d_vout = nullptr;
   double *h_vin = nullptr, *h_vout = nullptr;   // pinned staging
   cudaEvent_t ev0 = nullptr, ev1 = nullptr;
   b2_allreduce_fn allreduce = nullptr;          // sums partial sigma / diag vectors over the GPUs (NCCL in the caller)
   void* allreduce_user = nullptr;
   double last_kernel_s = 0.0;
   long long launches = 0;
   int world = 1, rank = 0;
   double list_bytes = 0.0;                      // size of the device work lists (decides whether the sweep driver keeps the plan)
   // excited states (Heff::addDiagramExcitations): n_exc level-shifted lower states, one vector of veclength doubles each
   int n_exc = 0;
   double *d_exc = nullptr, *d_exc_coef = nullptr, *d_exc_scratch = nullptr;
   ~b2_heff() {   // also runs when b2_heff_create bails out half-way (e.g. out of HBM): nothing leaks
      cudaFree(d_diag_items); cudaFree(d_diag_tiles); cudaFree(d_blk_off); cudaFree(d_p2s); cudaFree(d_s2p);
      cudaFree(d_items1); cudaFree(d_items2); cudaFree(d_reduces); cudaFree(d_part);
      for (int c = 0; c < kNumTileClasses; c++) { cudaFree(d_tiles1[c]); cudaFree(d_tiles2[c]); }
      cudaFree(d_jobs); cudaFree(d_parts); cudaFree(d_presum); cudaFree(d_work); cudaFree(d_vin); cudaFree(d_vout);
      cudaFree(d_exc); cudaFree(d_exc_coef); cudaFree(d_exc_scratch);
      if (h_vin) cudaFreeHost(h_vin);
      if (h_vout) cudaFreeHost(h_vout);
      if (ev0) cudaEventDestroy(ev0);
      if (ev1) cudaEventDestroy(ev1);
   }
};

struct b2_update {
   b2_ctx* ctx = nullptr;
   b2_opset *old_set = nullptr, *new_set = nullptr;
   UpdatePlan plan;
   CompiledWork pass[2];
   std::vector<PresumJob> presum_jobs;
   std::vector<PresumPart> presum_parts;
   GemmItem *d_items1[2] = {nullptr, nullptr}, *d_items2[2] = {nullptr, nullptr};
   ReduceJob* d_reduces[2] = {nullptr, nullptr};
   Tile* d_tiles1[2][kNumTileClasses] = {};
   Tile* d_tiles2[2][kNumTileClasses] = {};
   PresumJob* d_jobs = nullptr;
   PresumPart* d_parts = nullptr;
   double *d_presum = nullptr, *d_work = nullptr, *d_part = nullptr, *d_t = nullptr, *h_t = nullptr;
   int world = 1, rank = 0;
   double list_bytes[2] = {0.0, 0.0};      // device work-list bytes per pass
   std::vector<int> op_owner;              // GPU that computes new operator i in pass 0
   bool mix_all_axpy = false;              // pass 1 holds block axpys only: it runs on k_axpy_tiles
   // whole-operator mixing (UpdatePlan::mix_flat) grouped by layout: Dst[elem, d] += sum_s Src[elem, s] * coef[s][d]
   struct MixGroup { int64_t size; int nd, ns; int64_t dst_begin, src_begin, coef_begin; };
   std::vector<MixGroup> mix_groups;
   std::vector<int64_t> mix_dst_off, mix_src_off;   // offsets in the new arena (dst) / in the arena of mix_src_space (src)
   std::vector<uint8_t> mix_src_space;
   std::vector<double> mix_coef;
   int64_t *d_mix_dst_off = nullptr, *d_mix_src_off = nullptr;
   uint8_t* d_mix_src_space = nullptr;
   double* d_mix_coef = nullptr;
   int64_t temp_begin = 0, temp_size = 0;           // region of the pre-sum arena that holds the transposed copies
   b2_allreduce_fn allreduce = nullptr;
   void* allreduce_user = nullptr;
   ~b2_update() {
      for (int p = 0; p < 2; p++) {
         cudaFree(d_items1[p]); cudaFree(d_items2[p]); cudaFree(d_reduces[p]);
         for (int c = 0; c < kNumTileClasses; c++) { cudaFree(d_tiles1[p][c]); cudaFree(d_tiles2[p][c]); }
      }
      cudaFree(d_jobs); cudaFree(d_parts); cudaFree(d_presum); cudaFree(d_work); cudaFree(d_part); cudaFree(d_t);
      cudaFree(d_mix_dst_off); cudaFree(d_mix_src_off); cudaFree(d_mix_src_space); cudaFree(d_mix_coef);
      if (h_t) cudaFreeHost(h_t);
   }
};

namespace b2capi {
// stage-1 workspace budget of a plan: the configured value, capped by half of the HBM that is free right now
CompileOptions budgeted(const b2_ctx* ctx);
template <class T, class A> int upload_vec(T** dptr, const std::vector<T, A>& v, cudaStream_t s) {
   *dptr = nullptr;
   if (v.empty()) return B2_OK;
   CUDA_TRY(cudaMalloc(dptr, sizeof(T) * v.size()));
   CUDA_TRY(cudaMemcpyAsync(*dptr, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice, s));
   return B2_OK;
}
// operator sets of the 2-RDM chain: L only (only_L) or L, S0, S1, F0, F1 (DMRG::updateMovingLeftSafe2DM)
int opset_create_reduced(b2_ctx* ctx, int boundary, bool mr, bool only_L, b2_opset** out);
// the two halves of b2_heff_create (h->ctx, left, right, world, rank set by the caller): host = enumeration + scheduling without any
// CUDA call, device = uploads, workspaces, pre-sums
void heff_build_host(b2_heff* h, int site, const CompileOptions& budget);
int heff_setup_device(b2_heff* h);
// a plan parked in the sweep driver's cache keeps only its device work lists; unpark re-binds it to the operator sets of the new visit
void heff_park(b2_heff* h);
int heff_unpark(b2_heff* h, b2_opset* left, b2_opset* right);
void update_park(b2_update* u);
int update_unpark(b2_update* u, b2_opset* old_set, b2_opset* new_set);
void fill_worklists(const CompiledWork& c, b2_worklists* o);
// upload a compiled work list, run it once on the context stream, free it (small one-shot contractions: Join, gauge moves, 2-RDM)
int run_compiled_once(b2_ctx* ctx, const CompiledWork& w, DevBases b);
}   // namespace b2capi
