// CLM conditional fusion, elementwise part of SimpleCLM.forward (models/CLM.py:170-182):
//   w_r   = softmax_r(att_r)            (per pixel, over the R references)
//   out_c = sum_r w_r * sigmoid(att_r) * ref_t[r, c] + y_c
// The 1x1 / 3x3 convolutions around it stay nn.Conv2d.  HBM-bound: float4 coalesced accesses,
// the per-pixel coefficients are formed once per thread and reused over a channel group.
#include "common.cuh"

namespace clc {

constexpr int kMaxRefs = 8;
constexpr int kChanGroup = 16;  // channels handled by one thread of the forward kernel

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// coef[r] = softmax_r(a)[r] * sigmoid(a[r]);  also returns w[r], s[r] when asked (backward).
__device__ __forceinline__ void clm_coef(const float* a, int R, float* coef, float* w, float* s) {
  float mx = a[0];
  for (int r = 1; r < R; ++r) mx = fmaxf(mx, a[r]);
  float den = 0.f;
  float e[kMaxRefs];
  for (int r = 0; r < R; ++r) { e[r] = expf(a[r] - mx); den += e[r]; }
  for (int r = 0; r < R; ++r) {
    const float wr = e[r] / den, sr = sigmoidf_(a[r]);
    coef[r] = wr * sr;
    if (w) { w[r] = wr; s[r] = sr; }
  }
}

// grid = (ceil(S/VW/128), ceil(C/kChanGroup), B); thread = VW consecutive pixels.
template <int VW>
__global__ void __launch_bounds__(128)
clm_fuse_fwd_kernel(const float* __restrict__ ref_t, int64_t ref_sr, int64_t ref_sb,
                    const float* __restrict__ att, int64_t att_sr, int64_t att_sb,
                    const float* __restrict__ y, float* __restrict__ out, int R, int64_t B, int C,
                    int64_t S) {
  const int64_t s0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VW;
  if (s0 >= S) return;
  const int64_t b = blockIdx.z;
  float coef[VW][kMaxRefs];
  {
    float a[VW][kMaxRefs];
    for (int r = 0; r < R; ++r) {
      const float* ap = att + (int64_t)r * att_sr + b * att_sb + s0;
      if constexpr (VW == 4) {
        const float4 v = ld4(ap);
        a[0][r] = v.x; a[1][r] = v.y; a[2][r] = v.z; a[3][r] = v.w;
      } else {
        a[0][r] = ap[0];
      }
    }
#pragma unroll
    for (int v = 0; v < VW; ++v) clm_coef(a[v], R, coef[v], nullptr, nullptr);
  }
  const int c0 = blockIdx.y * kChanGroup;
  const int c1 = min(C, c0 + kChanGroup);
  for (int c = c0; c < c1; ++c) {
    const int64_t o = (b * C + c) * S + s0;
    if constexpr (VW == 4) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < R; ++r) {
        const float4 t = ld4_stream(ref_t + (int64_t)r * ref_sr + b * ref_sb + (int64_t)c * S + s0);
        // (aligned_stack * attention_weights).sum(dim=1): products summed left to right over r
        acc.x += t.x * coef[0][r]; acc.y += t.y * coef[1][r];
        acc.z += t.z * coef[2][r]; acc.w += t.w * coef[3][r];
      }
      const float4 yv = ld4_stream(y + o);
      st4_stream(out + o, make_float4(acc.x + yv.x, acc.y + yv.y, acc.z + yv.z, acc.w + yv.w));
    } else {
      float acc = 0.f;
      for (int r = 0; r < R; ++r) acc += ref_t[(int64_t)r * ref_sr + b * ref_sb + (int64_t)c * S + s0] * coef[0][r];
      out[o] = acc + y[o];
    }
  }
}

// Backward.  block = (32 pixel-threads, 8 channel lanes); channel lanes split C and reduce
// G_r = sum_c g_c * ref_t[r,c] through shared memory.
//   g_ref_t[r,c] = g_c * coef_r
//   g_att[m]     = w_m s_m G_m - w_m * sum_r G_r s_r w_r + G_m w_m s_m (1 - s_m)
template <int VW>
__global__ void __launch_bounds__(256)
clm_fuse_bwd_kernel(const float* __restrict__ ref_t, int64_t ref_sr, int64_t ref_sb,
                    const float* __restrict__ att, int64_t att_sr, int64_t att_sb,
                    const float* __restrict__ g_out, float* __restrict__ g_ref_t,
                    float* __restrict__ g_att, int R, int64_t B, int C, int64_t S) {
  __shared__ float Gs[8][kMaxRefs][32 * VW + 1];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int64_t s0 = ((int64_t)blockIdx.x * 32 + tx) * VW;
  const int64_t b = blockIdx.z;
  const bool live = s0 < S;
  float coef[VW][kMaxRefs], w[VW][kMaxRefs], sg[VW][kMaxRefs];
  float G[VW][kMaxRefs];
#pragma unroll
  for (int v = 0; v < VW; ++v)
    for (int r = 0; r < kMaxRefs; ++r) G[v][r] = 0.f;
  if (live) {
    float a[VW][kMaxRefs];
    for (int r = 0; r < R; ++r) {
      const float* ap = att + (int64_t)r * att_sr + b * att_sb + s0;
      if constexpr (VW == 4) {
        const float4 v = ld4(ap);
        a[0][r] = v.x; a[1][r] = v.y; a[2][r] = v.z; a[3][r] = v.w;
      } else {
        a[0][r] = ap[0];
      }
    }
#pragma unroll
    for (int v = 0; v < VW; ++v) clm_coef(a[v], R, coef[v], w[v], sg[v]);
    for (int c = ty; c < C; c += 8) {
      const int64_t o = (b * C + c) * S + s0;
      if constexpr (VW == 4) {
        const float4 g = ld4_stream(g_out + o);
        for (int r = 0; r < R; ++r) {
          const int64_t ro = (int64_t)r * ref_sr + b * ref_sb + (int64_t)c * S + s0;
          const float4 t = ld4_stream(ref_t + ro);
          G[0][r] = fmaf(g.x, t.x, G[0][r]); G[1][r] = fmaf(g.y, t.y, G[1][r]);
          G[2][r] = fmaf(g.z, t.z, G[2][r]); G[3][r] = fmaf(g.w, t.w, G[3][r]);
          st4_stream(g_ref_t + ro, make_float4(g.x * coef[0][r], g.y * coef[1][r], g.z * coef[2][r], g.w * coef[3][r]));
        }
      } else {
        const float g = g_out[o];
        for (int r = 0; r < R; ++r) {
          const int64_t ro = (int64_t)r * ref_sr + b * ref_sb + (int64_t)c * S + s0;
          G[0][r] = fmaf(g, ref_t[ro], G[0][r]);
          g_ref_t[ro] = g * coef[0][r];
        }
      }
    }
  }
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int v = 0; v < VW; ++v) Gs[ty][r][tx * VW + v] = G[v][r];
  __syncthreads();
  if (ty == 0 && live) {
#pragma unroll
    for (int v = 0; v < VW; ++v) {
      float Gt[kMaxRefs];
      float mix = 0.f;
      for (int r = 0; r < R; ++r) {
        float t = 0.f;
        for (int l = 0; l < 8; ++l) t += Gs[l][r][tx * VW + v];
        Gt[r] = t;
        mix = fmaf(t, coef[v][r], mix);  // sum_r G_r s_r w_r
      }
      for (int m = 0; m < R; ++m) {
        const float ga = coef[v][m] * Gt[m] - w[v][m] * mix + Gt[m] * coef[v][m] * (1.f - sg[v][m]);
        g_att[(int64_t)m * att_sr + b * att_sb + s0 + v] = ga;
      }
    }
  }
}

}  // namespace clc

using namespace clc;

static bool clm_args_ok(int R, int64_t B, int C, int64_t S) {
  return R >= 1 && R <= kMaxRefs && B >= 0 && C >= 1 && S >= 1 && B <= 65535;
}

extern "C" int clc_clm_fuse_fwd(const float* ref_t, int64_t ref_sr, int64_t ref_sb, const float* att,
                                int64_t att_sr, int64_t att_sb, const float* y, float* out,
                                int32_t R, int64_t B, int32_t C, int64_t S, void* stream) {
  if (!ref_t || !att || !y || !out) return CLC_ERR_INVALID_ARGUMENT;
  if (R > kMaxRefs || B > 65535) return CLC_ERR_UNSUPPORTED;
  if (!clm_args_ok(R, B, C, S)) return CLC_ERR_INVALID_ARGUMENT;
  if (B == 0) return CLC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (S % 4 == 0) && aligned16(ref_t) && aligned16(att) && aligned16(y) && aligned16(out) &&
                   !((ref_sr | ref_sb | att_sr | att_sb) & 3);
  const unsigned gy = (C + kChanGroup - 1) / kChanGroup;
  if (vec) {
    dim3 grid((unsigned)((S / 4 + 127) / 128), gy, (unsigned)B);
    clm_fuse_fwd_kernel<4><<<grid, 128, 0, st>>>(ref_t, ref_sr, ref_sb, att, att_sr, att_sb, y, out, R, B, C, S);
  } else {
    dim3 grid((unsigned)((S + 127) / 128), gy, (unsigned)B);
    clm_fuse_fwd_kernel<1><<<grid, 128, 0, st>>>(ref_t, ref_sr, ref_sb, att, att_sr, att_sb, y, out, R, B, C, S);
  }
  CLC_CHECK_LAUNCH("clc_clm_fuse_fwd");
  return CLC_OK;
}

extern "C" int clc_clm_fuse_bwd(const float* ref_t, int64_t ref_sr, int64_t ref_sb, const float* att,
                                int64_t att_sr, int64_t att_sb, const float* g_out, float* g_ref_t,
                                float* g_att, int32_t R, int64_t B, int32_t C, int64_t S, void* stream) {
  if (!ref_t || !att || !g_out || !g_ref_t || !g_att) return CLC_ERR_INVALID_ARGUMENT;
  if (R > kMaxRefs || B > 65535) return CLC_ERR_UNSUPPORTED;
  if (!clm_args_ok(R, B, C, S)) return CLC_ERR_INVALID_ARGUMENT;
  if (B == 0) return CLC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (S % 4 == 0) && aligned16(ref_t) && aligned16(att) && aligned16(g_out) &&
                   aligned16(g_ref_t) && aligned16(g_att) && !((ref_sr | ref_sb | att_sr | att_sb) & 3);
  dim3 block(32, 8);
  if (vec) {
    dim3 grid((unsigned)((S / 4 + 31) / 32), 1, (unsigned)B);
    clm_fuse_bwd_kernel<4><<<grid, block, 0, st>>>(ref_t, ref_sr, ref_sb, att, att_sr, att_sb, g_out, g_ref_t, g_att, R, B, C, S);
  } else {
    dim3 grid((unsigned)((S + 31) / 32), 1, (unsigned)B);
    clm_fuse_bwd_kernel<1><<<grid, block, 0, st>>>(ref_t, ref_sr, ref_sb, att, att_sr, att_sb, g_out, g_ref_t, g_att, R, B, C, S);
  }
  CLC_CHECK_LAUNCH("clc_clm_fuse_bwd");
  return CLC_OK;
}
