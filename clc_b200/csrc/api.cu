// Version / error-string entry points of the C ABI.
#include "common.cuh"

namespace clc {
thread_local char g_last_cuda_error[256] = {0};
std::atomic<unsigned long long> g_kernel_launches{0};
}

extern "C" int clc_version(void) { return 1; }

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

extern "C" const char* clc_last_cuda_error(void) { return clc::g_last_cuda_error; }

extern "C" uint64_t clc_kernel_launch_count(void) {
  return clc::g_kernel_launches.load(std::memory_order_relaxed);
}
