// Data-parallel exchange of the latent path over NVLink 5 / NVSwitch peer memory.
//
// The path's only collective is one small all-reduce per training step: the 2-double bpp
// statistic and the EntropyBottleneck parameter gradients (11 136 floats), 89 KB as doubles.  At that size an
// NCCL ring all-reduce is pure latency (~25-30 us at 8 ranks); here ONE kernel per rank does it in one shot:
//   1. pack the local contribution (as doubles) into this rank's peer-visible buffer, slot = step parity;
//   2. the last CTA to finish packing publishes the step number to every rank's flag array (st.release.sys
//      over NVLink);
//   3. every CTA waits until all ranks have published this step (ld.acquire.sys on its own flags), then
//      sums the world's buffers in FIXED rank order (deterministic, identical bits on every rank) with
//      cache-bypassing peer loads and writes the result: sums for the statistic, mean for the gradients
//      (the reference's DataParallel gradient is the mean over the global batch, train_CLC.py:472-473).
// Two slots make the buffers safe without a second handshake: a rank can only reach step s+2's pack after
// every rank has published s+1, i.e. after every rank has finished reading step s.
// The step counter lives in device memory and is advanced by the kernel itself, so the launch is captured
// into the step's CUDA graph like any other kernel.
//
// Peer-visible memory is a plain cudaMalloc region exported through CUDA IPC; the 64-byte handles travel over
// torch.distributed (plumbing), the library opens them (clc_peer_*).
#include "common.cuh"

namespace clc {

constexpr int kMaxPeers = 8;

struct PeerArgs {
  double* bufs[kMaxPeers];                 // rank r's data region: 2 slots x n doubles
  unsigned long long* flags[kMaxPeers];    // rank r's flag array: one step number per source rank
  int rank, world;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_peer_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// state[0] = completed steps, state[1] = CTAs that finished packing, state[2] = CTAs that finished reducing
__global__ void __launch_bounds__(256)
peer_allreduce_kernel(const PeerArgs a, double* __restrict__ stat, int n_stat, float* __restrict__ grads,
                      int n_grads, float grad_scale, unsigned long long* __restrict__ state) {
  const unsigned long long seq = state[0] + 1;
  const int n = n_stat + n_grads;
  const size_t slot = (size_t)(seq & 1ull) * (size_t)n;
  double* mine = a.bufs[a.rank] + slot;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    mine[i] = i < n_stat ? stat[i] : (double)grads[i - n_stat];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long arrived = atomicAdd(&state[1], 1ull);
    if (arrived == gridDim.x - 1) {            // the whole contribution is in place: publish the step everywhere
      state[1] = 0;
      __threadfence_system();
      for (int r = 0; r < a.world; ++r) st_release_sys(a.flags[r] + a.rank, seq);
    }
  }
  if ((int)threadIdx.x < a.world) {
    const unsigned long long* f = a.flags[a.rank] + threadIdx.x;
    while (ld_acquire_sys(f) < seq) { }
  }
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < a.world; ++r) s += ld_peer_f64(a.bufs[r] + slot + i);
    if (i < n_stat) stat[i] = s;
    else grads[i - n_stat] = (float)(s * (double)grad_scale);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long done = atomicAdd(&state[2], 1ull);
    if (done == gridDim.x - 1) {
      state[2] = 0;
      state[0] = seq;
    }
  }
}

}  // namespace clc

using namespace clc;

// ---- peer-visible memory (CUDA IPC) ----
extern "C" int clc_peer_alloc(size_t bytes, void** ptr) {
  if (!ptr || bytes == 0) return CLC_ERR_INVALID_ARGUMENT;
  CLC_CUDA(cudaMalloc(ptr, bytes));
  CLC_CUDA(cudaMemset(*ptr, 0, bytes));
  CLC_CUDA(cudaDeviceSynchronize());
  return CLC_OK;
}

extern "C" int clc_peer_free(void* ptr) {
  if (ptr) CLC_CUDA(cudaFree(ptr));
  return CLC_OK;
}

extern "C" int clc_peer_export(void* ptr, uint8_t handle[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (!ptr || !handle) return CLC_ERR_INVALID_ARGUMENT;
  cudaIpcMemHandle_t h;
  CLC_CUDA(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle, &h, 64);
  return CLC_OK;
}

extern "C" int clc_peer_open(const uint8_t handle[64], void** ptr) {
  if (!ptr || !handle) return CLC_ERR_INVALID_ARGUMENT;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  CLC_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return CLC_OK;
}

extern "C" int clc_peer_close(void* ptr) {
  if (ptr) CLC_CUDA(cudaIpcCloseMemHandle(ptr));
  return CLC_OK;
}

extern "C" size_t clc_peer_allreduce_bytes(int32_t n_stat, int32_t n_grads, int32_t world) {
  if (n_stat < 0 || n_grads < 0 || world < 1 || world > kMaxPeers) return 0;
  const size_t data = 2 * (size_t)(n_stat + n_grads) * sizeof(double);
  return (data + 255) / 256 * 256 + 256;     // data slots | flags (one 256-byte line)
}

extern "C" int clc_peer_allreduce(void* const* regions, int32_t rank, int32_t world, double* stat, int32_t n_stat,
                                  float* grads, int32_t n_grads, float grad_scale, uint64_t* state, void* stream) {
  if (!regions || !state || rank < 0 || rank >= world || n_stat < 0 || n_grads < 0) return CLC_ERR_INVALID_ARGUMENT;
  if (world > kMaxPeers) return CLC_ERR_UNSUPPORTED;
  if ((n_stat && !stat) || (n_grads && !grads)) return CLC_ERR_INVALID_ARGUMENT;
  const int n = n_stat + n_grads;
  if (n == 0) return CLC_OK;
  const size_t data = (2 * (size_t)n * sizeof(double) + 255) / 256 * 256;
  PeerArgs a;
  a.rank = rank;
  a.world = world;
  for (int r = 0; r < world; ++r) {
    if (!regions[r]) return CLC_ERR_INVALID_ARGUMENT;
    a.bufs[r] = reinterpret_cast<double*>(regions[r]);
    a.flags[r] = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(regions[r]) + data);
  }
  int grid = (n + 1023) / 1024;               // ~4 elements per thread; a handful of CTAs, all co-resident
  if (grid > 16) grid = 16;
  if (grid < 1) grid = 1;
  peer_allreduce_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, stat, n_stat, grads, n_grads, grad_scale,
                                                                reinterpret_cast<unsigned long long*>(state));
  CLC_CHECK_LAUNCH("clc_peer_allreduce");
  return CLC_OK;
}
