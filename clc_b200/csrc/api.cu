// Version / error-string entry points of the C ABI.
#include "common.cuh"

#include <stdlib.h>

#include <mutex>
#include <vector>

namespace clc {
thread_local char g_last_cuda_error[256] = {0};
std::atomic<unsigned long long> g_kernel_launches{0};
#ifdef CLC_DEBUG_ABI
std::atomic<int> g_stage_mask{0xff};
std::atomic<int> g_pdl{getenv("CLC_NO_PDL") ? 0 : 1};
#endif

// ---- per-kernel tracing: CUDA events on the traced stream, one after every launch ----
std::atomic<bool> g_trace_on{false};
namespace {
struct TraceRec { const char* name; cudaEvent_t ev; };
std::mutex g_trace_mu;
std::vector<TraceRec> g_trace;
cudaStream_t g_trace_stream = nullptr;
void trace_clear() {
  for (auto& r : g_trace) cudaEventDestroy(r.ev);
  g_trace.clear();
}
}  // namespace
void trace_record(const char* name) {
  std::lock_guard<std::mutex> lk(g_trace_mu);
  cudaEvent_t ev;
  if (cudaEventCreate(&ev) != cudaSuccess) return;
  cudaEventRecord(ev, g_trace_stream);
  g_trace.push_back({name, ev});
}
}

extern "C" int clc_version(void) { return 2; }

extern "C" const char* clc_strerror(int status) {
  switch (status) {
    case CLC_OK: return "ok";
    case CLC_ERR_INVALID_ARGUMENT: return "invalid argument";
    case CLC_ERR_UNSUPPORTED: return "unsupported shape or mode";
    case CLC_ERR_WORKSPACE: return "workspace too small";
    case CLC_ERR_CUDA: return "CUDA error (see clc_last_cuda_error)";
    case CLC_ERR_ARCH: return "device is not sm_100 (B200) class";
    default: return "unknown status";
  }
}

#ifdef CLC_DEBUG_ABI
extern "C" CLC_API void clc_debug_set_stage_mask(int mask) { clc::g_stage_mask.store(mask); }
#endif

extern "C" const char* clc_last_cuda_error(void) { return clc::g_last_cuda_error; }

extern "C" uint64_t clc_kernel_launch_count(void) {
  return clc::g_kernel_launches.load(std::memory_order_relaxed);
}

extern "C" int clc_trace_start(void* stream) {
  std::lock_guard<std::mutex> lk(clc::g_trace_mu);
  clc::trace_clear();
  clc::g_trace_stream = (cudaStream_t)stream;
  cudaEvent_t ev;
  CLC_CUDA(cudaEventCreate(&ev));
  CLC_CUDA(cudaEventRecord(ev, clc::g_trace_stream));
  clc::g_trace.push_back({"(start)", ev});
  clc::g_trace_on.store(true);
  return CLC_OK;
}

extern "C" int clc_trace_mark(void) {
  if (!clc::g_trace_on.load()) return CLC_ERR_INVALID_ARGUMENT;
  clc::trace_record("(mark)");
  return CLC_OK;
}

extern "C" int clc_trace_stop(void) {
  clc::g_trace_on.store(false);
  std::lock_guard<std::mutex> lk(clc::g_trace_mu);
  if (!clc::g_trace.empty()) CLC_CUDA(cudaEventSynchronize(clc::g_trace.back().ev));
  return CLC_OK;
}

extern "C" int clc_trace_count(void) {
  std::lock_guard<std::mutex> lk(clc::g_trace_mu);
  return clc::g_trace.empty() ? 0 : (int)clc::g_trace.size() - 1;
}

extern "C" int clc_trace_get(int i, const char** name, float* ms) {
  std::lock_guard<std::mutex> lk(clc::g_trace_mu);
  if (i < 0 || i + 1 >= (int)clc::g_trace.size() || !name || !ms) return CLC_ERR_INVALID_ARGUMENT;
  *name = clc::g_trace[i + 1].name;
  CLC_CUDA(cudaEventElapsedTime(ms, clc::g_trace[i].ev, clc::g_trace[i + 1].ev));
  return CLC_OK;
}
