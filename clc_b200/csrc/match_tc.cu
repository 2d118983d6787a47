// tcgen05 screening path of the match stage (placeholder until the kernel lands).
#include "match.cuh"

extern "C" size_t clc_match_topk_tc_workspace_bytes(int64_t, int32_t, int32_t, int32_t, int32_t, int32_t,
                                                    int32_t, int32_t) {
  return 0;
}

extern "C" int clc_match_topk_tc(const float*, const float*, int64_t, int32_t, int32_t, int32_t, int32_t,
                                 int32_t, int32_t, int32_t, int32_t, float*, int32_t*, int32_t*, void*, size_t,
                                 void*) {
  return CLC_ERR_UNSUPPORTED;
}
