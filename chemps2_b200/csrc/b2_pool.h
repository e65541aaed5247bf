// b2_pool.h — caching device allocator behind every cudaMalloc / cudaFree of the library.
//
// A sweep allocates and frees ~60 device buffers per site (work lists, workspaces, Davidson basis, Join, SVD); cudaMalloc of gigabytes and
// cudaFree (which synchronises the whole device) were a measurable part of every half sweep (profiles/r1_tuning.md, r2 timing logs).  Freed
// blocks are kept in size classes (<= 12.5 % rounding) and handed out again; nothing is returned to the driver until the cache exceeds its
// budget or an allocation fails.  Safety across streams: a freed block carries one event per registered stream (the streams of the live
// contexts), and every registered stream waits for those events before the block is reused — so a block freed while kernels still read it
// (b2_heff_destroy right after an asynchronous apply) cannot be overwritten early, without any host synchronisation.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace b2 {
cudaError_t pool_malloc(void** p, size_t bytes);
cudaError_t pool_free(void* p);
void pool_register_stream(cudaStream_t s);
void pool_unregister_stream(cudaStream_t s);
void pool_trim();                 // give every cached block back to the driver
size_t pool_cached_bytes();       // bytes held in the cache (free for the library's purposes)
// Pinned host staging buffers (cudaMallocHost costs milliseconds and every Split / plan upload needs tens of MB): a handful of buffers
// is kept and handed out again; a caller must be done with its transfers (stream synchronised) before it releases one.
void* pinned_acquire(size_t bytes);   // nullptr when the allocation fails (callers fall back to pageable memory)
void pinned_release(void* p);
}   // namespace b2

#ifndef B2_POOL_IMPL
#define cudaMalloc(p, n) b2::pool_malloc((void**)(p), (size_t)(n))
#define cudaFree(p) b2::pool_free((void*)(p))
#endif
