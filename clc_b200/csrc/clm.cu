// CLM conditional fusion, elementwise part of SimpleCLM.forward (models/CLM.py:170-182):
//   w_r   = softmax_r(att_r)            (per pixel, over the R references)
//   out_c = sum_r w_r * sigmoid(att_r) * ref_t[r, c] + y_c
// The 1x1 / 3x3 convolutions around it stay nn.Conv2d.  HBM-bound: float4 coalesced accesses,
// the per-pixel coefficients are formed once per thread and reused over a channel group.
#include "common.cuh"

#include <cooperative_groups.h>

namespace clc {

constexpr int kMaxRefs = 8;

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

// Forward.  block = 256 threads = 8 warps; a lane owns VW consecutive pixels, a warp owns the
// channels c0 + warp + 8*i (i < CPW) of its CTA.  grid = (ceil(S/(32*VW)), ceil(C/(8*CPW)), B).
// The per-pixel coefficients (R exps + R sigmoids) are formed once per thread and reused over
// its CPW channels; CPW is chosen by the host so the grid fills the 148 SMs.
template <int VW, int RT>
__global__ void __launch_bounds__(256)
clm_fuse_fwd_kernel(const float* __restrict__ ref_t, int64_t ref_sr, int64_t ref_sb,
                    const float* __restrict__ att, int64_t att_sr, int64_t att_sb,
                    const float* __restrict__ y, float* __restrict__ out, int R, int C, int64_t S,
                    int CPW) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t s0 = ((int64_t)blockIdx.x * 32 + lane) * VW;
  pdl_trigger();
  pdl_wait();
  if (s0 >= S) return;
  const int64_t b = blockIdx.z;
  float coef[VW][kMaxRefs];
  {
    float a[VW][kMaxRefs];
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      if (r >= R) break;
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
  const int cbase = blockIdx.y * 8 * CPW + warp;
  if constexpr (VW == 4) {
    // two channels per iteration, all 2 * (R + 1) loads issued before the first use
    for (int i = 0; i < CPW; i += 2) {
      float4 t[2][RT], yv[2];
      bool ok[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int c = cbase + 8 * (i + u);
        ok[u] = (i + u < CPW) && c < C;
        if (ok[u]) {
#pragma unroll
          for (int r = 0; r < RT; ++r)
            if (r < R) t[u][r] = ld4_stream(ref_t + (int64_t)r * ref_sr + b * ref_sb + (int64_t)c * S + s0);
          yv[u] = ld4_stream(y + (b * C + c) * S + s0);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (!ok[u]) continue;
        const int c = cbase + 8 * (i + u);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        // (aligned_stack * attention_weights).sum(dim=1): products summed left to right over r
#pragma unroll
        for (int r = 0; r < RT; ++r)
          if (r < R) {
            acc.x += t[u][r].x * coef[0][r]; acc.y += t[u][r].y * coef[1][r];
            acc.z += t[u][r].z * coef[2][r]; acc.w += t[u][r].w * coef[3][r];
          }
        st4_stream(out + (b * C + c) * S + s0,
                   make_float4(acc.x + yv[u].x, acc.y + yv[u].y, acc.z + yv[u].z, acc.w + yv[u].w));
      }
    }
  } else {
    for (int i = 0; i < CPW; ++i) {
      const int c = cbase + 8 * i;
      if (c >= C) break;
      const int64_t o = (b * C + c) * S + s0;
      float acc = 0.f;
      for (int r = 0; r < R; ++r) acc += ref_t[(int64_t)r * ref_sr + b * ref_sb + (int64_t)c * S + s0] * coef[0][r];
      out[o] = acc + y[o];
    }
  }
}

// Backward.  block = (32 pixel lanes, 8 channel lanes); a thread-block CLUSTER of CS CTAs along
// grid.y splits the C channels, every CTA reduces its partial G_r = sum_c g_c * ref_t[r,c] over
// its channel lanes in shared memory, and cluster rank 0 combines the CS partials through
// distributed shared memory in fixed rank order (deterministic, no atomics, no workspace).
//   g_ref_t[r,c] = g_c * coef_r
//   g_att[m]     = w_m s_m G_m - w_m * sum_r G_r s_r w_r + G_m w_m s_m (1 - s_m)
template <int VW, int RT>
__global__ void __launch_bounds__(256)
clm_fuse_bwd_kernel(const float* __restrict__ ref_t, int64_t ref_sr, int64_t ref_sb,
                    const float* __restrict__ att, int64_t att_sr, int64_t att_sb,
                    const float* __restrict__ g_out, float* __restrict__ g_ref_t,
                    float* __restrict__ g_att, int R, int C, int64_t S) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float Gs[8][kMaxRefs][32 * VW + 1];
  __shared__ float Gp[kMaxRefs][32 * VW];  // this CTA's partial, read by rank 0 through DSMEM
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int64_t s0 = ((int64_t)blockIdx.x * 32 + tx) * VW;
  const int64_t b = blockIdx.z;
  const bool live = s0 < S;
  const unsigned CS = cluster.dim_blocks().y, rank = cluster.block_rank();
  pdl_trigger();
  pdl_wait();
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
    // channels of this CTA: c = blockIdx.y + gridDim.y * m  (interleaved), channel lane ty takes every 8th
#pragma unroll 2
    for (int c = blockIdx.y + gridDim.y * ty; c < C; c += gridDim.y * 8) {
      const int64_t o = (b * C + c) * S + s0;
      if constexpr (VW == 4) {
        // all R + 1 loads of the channel are issued before the first use
        const float4 g = ld4_stream(g_out + o);
        float4 t[RT];
#pragma unroll
        for (int r = 0; r < RT; ++r)
          if (r < R) t[r] = ld4_stream(ref_t + (int64_t)r * ref_sr + b * ref_sb + (int64_t)c * S + s0);
#pragma unroll
        for (int r = 0; r < RT; ++r)
          if (r < R) {
            const int64_t ro = (int64_t)r * ref_sr + b * ref_sb + (int64_t)c * S + s0;
            G[0][r] = fmaf(g.x, t[r].x, G[0][r]); G[1][r] = fmaf(g.y, t[r].y, G[1][r]);
            G[2][r] = fmaf(g.z, t[r].z, G[2][r]); G[3][r] = fmaf(g.w, t[r].w, G[3][r]);
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
  if (ty == 0) {
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int v = 0; v < VW; ++v) {
        float t = 0.f;
        for (int l = 0; l < 8; ++l) t += Gs[l][r][tx * VW + v];
        Gp[r][tx * VW + v] = t;
      }
  }
  cluster.sync();  // partials of every CTA of the cluster are in place
  if (rank == 0 && ty == 0 && live) {
#pragma unroll
    for (int v = 0; v < VW; ++v) {
      float Gt[kMaxRefs];
      float mix = 0.f;
      for (int r = 0; r < R; ++r) {
        float t = 0.f;
        for (unsigned k = 0; k < CS; ++k) {
          const float* remote = cluster.map_shared_rank(&Gp[0][0], k);
          t += remote[r * (32 * VW) + tx * VW + v];
        }
        Gt[r] = t;
        mix = fmaf(t, coef[v][r], mix);  // sum_r G_r s_r w_r
      }
      for (int m = 0; m < R; ++m) {
        const float ga = coef[v][m] * Gt[m] - w[v][m] * mix + Gt[m] * coef[v][m] * (1.f - sg[v][m]);
        g_att[(int64_t)m * att_sr + b * att_sb + s0 + v] = ga;
      }
    }
  }
  cluster.sync();  // keep every CTA's shared memory alive until rank 0 has read it
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
  const int VW = vec ? 4 : 1;
  const unsigned gx = (unsigned)((S + 32 * VW - 1) / (32 * VW));
  // channels per warp: as many as keeps >= 2 CTAs per SM in flight (coefficients amortised over them)
  int CPW = 16;
  while (CPW > 1 && (int64_t)gx * B * ((C + 8 * CPW - 1) / (8 * CPW)) < 2 * kNumSMs) CPW >>= 1;
  const unsigned gy = (unsigned)((C + 8 * CPW - 1) / (8 * CPW));
  if (gy > 65535) return CLC_ERR_UNSUPPORTED;
  dim3 grid(gx, gy, (unsigned)B);
  if (vec && R <= 4)
    CLC_CUDA(launch_pdl(clm_fuse_fwd_kernel<4, 4>, grid, dim3(256), 0, st, ref_t, ref_sr, ref_sb, att, att_sr, att_sb, y, out,
                        R, C, S, CPW));
  else if (vec)
    CLC_CUDA(launch_pdl(clm_fuse_fwd_kernel<4, 8>, grid, dim3(256), 0, st, ref_t, ref_sr, ref_sb, att, att_sr, att_sb, y, out,
                        R, C, S, CPW));
  else
    CLC_CUDA(launch_pdl(clm_fuse_fwd_kernel<1, 8>, grid, dim3(256), 0, st, ref_t, ref_sr, ref_sb, att, att_sr, att_sb, y, out,
                        R, C, S, CPW));
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
  // pixels per lane: float4 only when that still leaves enough CTAs; cluster size: up to 8 CTAs
  // (portable limit) split the channels of one pixel tile.
  const bool vec4 = vec && (int64_t)((S / 4 + 31) / 32) * B * 8 >= 2 * kNumSMs;
  const unsigned gx = (unsigned)(vec4 ? (S / 4 + 31) / 32 : (S + 31) / 32);
  unsigned CS = 8;
  while (CS > 1 && (unsigned)C < 8 * CS) CS >>= 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(gx, CS, (unsigned)B);
  cfg.blockDim = dim3(32, 8, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = CS;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_on() ? 2 : 1;
  if (vec4 && R <= 4)
    CLC_CUDA(cudaLaunchKernelEx(&cfg, clm_fuse_bwd_kernel<4, 4>, ref_t, ref_sr, ref_sb, att, att_sr, att_sb, g_out,
                                g_ref_t, g_att, (int)R, (int)C, S));
  else if (vec4)
    CLC_CUDA(cudaLaunchKernelEx(&cfg, clm_fuse_bwd_kernel<4, 8>, ref_t, ref_sr, ref_sb, att, att_sr, att_sb, g_out,
                                g_ref_t, g_att, (int)R, (int)C, S));
  else
    CLC_CUDA(cudaLaunchKernelEx(&cfg, clm_fuse_bwd_kernel<1, 8>, ref_t, ref_sr, ref_sb, att, att_sr, att_sb, g_out,
                                g_ref_t, g_att, (int)R, (int)C, S));
  CLC_CHECK_LAUNCH("clc_clm_fuse_bwd");
  return CLC_OK;
}
