// Shared helpers for libclc_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <utility>

#include "../../include/clc_b200.h"

namespace clc {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// Thread-local text of the last CUDA failure, surfaced through clc_last_cuda_error().
extern thread_local char g_last_cuda_error[256];

inline int cuda_fail(cudaError_t e, const char* where) {
  snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s: %s", where, cudaGetErrorString(e));
  return CLC_ERR_CUDA;
}

// Process-wide count of kernels enqueued by this library (statistics only; bench.py reports it).
extern std::atomic<unsigned long long> g_kernel_launches;

// Bring-up hooks exist only in the debug build of the library (libclc_b200_dbg.so, -DCLC_DEBUG_ABI): a bit
// mask of the kernels a multi-kernel entry point enqueues (clc_debug_set_stage_mask), kernel-specific
// experiment bits and a PDL switch.  The production library has no mutable global state on its call paths.
#ifdef CLC_DEBUG_ABI
extern std::atomic<int> g_stage_mask;
inline bool stage_on(int bit) { return (g_stage_mask.load(std::memory_order_relaxed) >> bit) & 1; }
inline int dbg_bits() { return (g_stage_mask.load(std::memory_order_relaxed) >> 8) & 0xff; }
extern std::atomic<int> g_pdl;   // 1 = on (default); CLC_NO_PDL=1 in the environment turns it off
inline bool pdl_on() { return g_pdl.load(std::memory_order_relaxed) != 0; }
#else
inline constexpr bool stage_on(int) { return true; }
inline constexpr int dbg_bits() { return 0; }
inline constexpr bool pdl_on() { return true; }
#endif

// Per-kernel tracing (clc_trace_*): when on, one cudaEvent is recorded after every kernel launch.
extern std::atomic<bool> g_trace_on;
void trace_record(const char* name);

#define CLC_CHECK_LAUNCH(where)                                  \
  do {                                                           \
    cudaError_t e__ = cudaGetLastError();                        \
    if (e__ != cudaSuccess) return ::clc::cuda_fail(e__, where); \
    ::clc::g_kernel_launches.fetch_add(1, std::memory_order_relaxed); \
    if (::clc::g_trace_on.load(std::memory_order_relaxed)) ::clc::trace_record(where); \
  } while (0)

#define CLC_CUDA(call)                                           \
  do {                                                           \
    cudaError_t e__ = (call);                                    \
    if (e__ != cudaSuccess) return ::clc::cuda_fail(e__, #call); \
  } while (0)

// ---- programmatic dependent launch (PDL): every kernel of the serial chains is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so its launch + prologue overlap the tail of the
// kernel before it (also inside captured CUDA graphs).  A kernel calls pdl_wait() before it touches
// global memory (returns once the preceding grid has completed and its writes are visible); the
// implicit trigger at CTA exit is used (pdl_trigger() compiles to nothing unless CLC_PDL_EARLY=1).
// Both are no-ops without the attribute.
#ifndef CLC_PDL_EARLY
#define CLC_PDL_EARLY 0   /* measured on B200: the explicit early trigger costs ~6% on the cfg2 chain */
#endif
__device__ __forceinline__ void pdl_trigger() {
#if CLC_PDL_EARLY
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_on() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(std::forward<Args>(args))...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; result valid in thread 0.  `red` is >= 32 floats of shared memory.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect `red` against a previous use
  if (lane == 0) red[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (wid == 0) v = warp_sum(v);
  return v;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// Streaming (read-once / write-once) variants: do not pollute L1.
__device__ __forceinline__ float4 ld4_stream(const float* p) {
  return __ldcs(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void st4_stream(float* p, float4 v) {
  __stcs(reinterpret_cast<float4*>(p), v);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Grid size for a grid-stride elementwise kernel: enough CTAs for `work_items` at `per_block`
// items per pass, capped at `waves` full waves of the 148 SMs x `ctas_per_sm` resident CTAs.
inline int grid_for(int64_t work_items, int per_block, int ctas_per_sm = 8, int waves = 1) {
  int64_t need = (work_items + per_block - 1) / per_block;
  int64_t cap = (int64_t)kNumSMs * ctas_per_sm * waves;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// 2-D grid for [B rows] x [items per row] elementwise kernels: x covers one row in `per_block` chunks, y the rows,
// the whole grid capped at ~ctas_per_sm CTAs per SM (kernels stride over both dimensions).
inline dim3 grid_rows(int64_t rows, int64_t items_per_row, int per_block, int ctas_per_sm = 8) {
  int64_t gy = rows < 1 ? 1 : (rows > 65535 ? 65535 : rows);
  int64_t need = (items_per_row + per_block - 1) / per_block;
  if (need < 1) need = 1;
  int64_t cap = ((int64_t)kNumSMs * ctas_per_sm + gy - 1) / gy;
  if (cap < 1) cap = 1;
  return dim3((unsigned)(need < cap ? need : cap), (unsigned)gy, 1);
}

}  // namespace clc
