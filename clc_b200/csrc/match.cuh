// Shared pieces of the matching kernels (match.cu: exact fp32 path; match_tc.cu: tcgen05 path).
#pragma once
#include "common.cuh"

namespace clc {

// Device-side mirror of clc_patch_view (include/clc_b200.h).
struct PatchAddr {
  const float* q;
  int64_t sn, spy, spx, sc, sy;
  int npx, repeat;
  __host__ __device__ __forceinline__ int64_t patch_off(int patch) const {
    const int py = patch / npx, px = patch - py * npx;
    return (int64_t)py * spy + (int64_t)px * spx;
  }
};

struct PosStat {
  float ym;  // y_mean = conv2d(y, ones/K)                          (Patch_Matching.py:872-874)
  float dY;  // denominator_y = sum_y_square - y_mean*y_mean*K      (:889-893)
};

__device__ __forceinline__ PosStat pos_stat(float box_s1, float box_s2, float inv_k, float Kf) {
  PosStat s;
  s.ym = box_s1 * inv_k;
  s.dY = box_s2 - s.ym * s.ym * Kf;
  return s;
}

// out = numerator / sqrt(denominator), operation order of Patch_Matching.py:880-905.
__device__ __forceinline__ float pearson(float xy, PosStat ps, float xs, float sxx, float Kf) {
  const float num = xy - ps.ym * xs;   // numerator = xy - y_mean * x_sum
  const float xm = xs / Kf;            // x_mean
  const float dX = sxx - xm * xs;      // denominator_x
  const float den = ps.dY * dX;        // denominator
  return num / sqrtf(den);
}

int launch_channel_sums(const float* r, float* s1, float* s2, int64_t NP, int C, int64_t HW,
                        cudaStream_t st);
int launch_patch_stats(const PatchAddr& qa, float* xs, float* sxx, int64_t NQ, int P, int C, int ph,
                       int pw, cudaStream_t st);

}  // namespace clc
