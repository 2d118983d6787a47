// CLM variant (a): similarity-softmax alignment of a reference latent (models/CLM.py:5-128).
//
// The reference forms, per (image, reference), the HW x HW similarity  sim = y_t^T ref_t / T, its row softmax
// (CLM.py:104-107), and then -- in DeformableAlignment.forward, :16-20 -- accumulates `sim[:, i, j] * x` over EVERY
// query position (i, j): the only thing that survives of the HW x HW map is its COLUMN SUMS
//     colsum[p] = sum_q softmax_p(sim[q, :])[p],          weighted_x[c, p] = x[c, p] * colsum[p].
// clc_clm_sim_colsum computes them flash-style in two passes over 64 x 64 tiles of the similarity that live in
// registers only (row max / row sum first, then the normalised column sums); the map (419 MB per reference at
// 1280 x 2048) is never written.  The contraction runs in fp32 on the CUDA cores: with T = 0.5 the
// logits reach +-100 and the softmax is sharply peaked, so bf16 / single-pass tf32 tensor-core products (1e-2 ..
// 1e-3 absolute logit error) do not meet a 1e-5 parity bar; the sizes this variant is usable at (the reference
// walks H*W*B*9 Python iterations per call) make the kernel a few tens of microseconds.
//
// clc_clm_deform_fwd is the reference's hand-rolled "deformable" sampling (:35-60): 9 bilinear taps around
// (h, w) + offset, modulated, summed -- with its quirks kept (no kernel-tap base grid, taps outside the image
// dropped, int() truncation, clamped +1 neighbours).  clc_clm_attention_sum_fwd is :117-126.
// Forward only: the reference's variant (a) is not trainable in practice (25 s per 2 x 64 x 32 x 32 forward).
#include "common.cuh"

namespace clc {

constexpr int kSimTile = 64;      // similarity tile edge (queries x reference positions)
constexpr int kSimCK = 16;        // channels per shared-memory stage

// acc[i][j] += sum_c yt[c, q0 + 4*ty + i] * rt[c, p0 + 4*tx + j] over all channels; 256 threads, (ty, tx) = 16 x 16.
__device__ __forceinline__ void sim_tile(const float* __restrict__ yt, const float* __restrict__ rt, int C, int64_t HW,
                                         int64_t q0, int64_t p0, float (&acc)[4][4], float (*sA)[kSimTile],
                                         float (*sB)[kSimTile]) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int lc = tid >> 4, l4 = (tid & 15) * 4;          // this thread's load slot: channel lc, positions l4..l4+3
  const bool vec = (HW & 3) == 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int c0 = 0; c0 < C; c0 += kSimCK) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    const int c = c0 + lc;
    if (c < C) {
      const float* ya = yt + (int64_t)c * HW + q0 + l4;
      const float* rb = rt + (int64_t)c * HW + p0 + l4;
      if (vec && q0 + l4 + 3 < HW) a = ld4(ya);
      else {
        if (q0 + l4 + 0 < HW) a.x = ya[0];
        if (q0 + l4 + 1 < HW) a.y = ya[1];
        if (q0 + l4 + 2 < HW) a.z = ya[2];
        if (q0 + l4 + 3 < HW) a.w = ya[3];
      }
      if (vec && p0 + l4 + 3 < HW) b = ld4(rb);
      else {
        if (p0 + l4 + 0 < HW) b.x = rb[0];
        if (p0 + l4 + 1 < HW) b.y = rb[1];
        if (p0 + l4 + 2 < HW) b.z = rb[2];
        if (p0 + l4 + 3 < HW) b.w = rb[3];
      }
    }
    __syncthreads();                                       // previous stage fully consumed
    *reinterpret_cast<float4*>(&sA[lc][l4]) = a;
    *reinterpret_cast<float4*>(&sB[lc][l4]) = b;
    __syncthreads();
#pragma unroll
    for (int cc = 0; cc < kSimCK; ++cc) {
      const float4 av = *reinterpret_cast<const float4*>(&sA[cc][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&sB[cc][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }
}

// max / sum over the 16 lanes (tx) that share the same query rows: xor 1, 2, 4, 8 stays inside the half-warp
__device__ __forceinline__ float half_warp_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Pass 1: stats[nb][q] = {row max, row sum of exp(s - max)} of s[q, :] = sim[q, :] / T.   grid = (query tiles, NB)
__global__ void __launch_bounds__(256)
sim_row_stats_kernel(const float* __restrict__ y_t, const float* __restrict__ ref_t, int64_t B_y, int C, int64_t HW,
                     float inv_T, float2* __restrict__ stats) {
  __shared__ __align__(16) float sA[kSimCK][kSimTile];
  __shared__ __align__(16) float sB[kSimCK][kSimTile];
  const int64_t nb = blockIdx.y, q0 = (int64_t)blockIdx.x * kSimTile;
  const float* yt = y_t + (nb % B_y) * (int64_t)C * HW;
  const float* rt = ref_t + nb * (int64_t)C * HW;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float m[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { m[i] = -INFINITY; l[i] = 0.f; }
  for (int64_t p0 = 0; p0 < HW; p0 += kSimTile) {
    float acc[4][4];
    sim_tile(yt, rt, C, HW, q0, p0, acc, sA, sB);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float s[4], tm = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[j] = (p0 + tx * 4 + j < HW) ? acc[i][j] * inv_T : -INFINITY;
        tm = fmaxf(tm, s[j]);
      }
      tm = half_warp_max(tm);
      const float mn = fmaxf(m[i], tm);
      float ts = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) ts += (s[j] == -INFINITY) ? 0.f : expf(s[j] - mn);
      ts = half_warp_sum(ts);
      l[i] = (m[i] == -INFINITY ? 0.f : l[i] * expf(m[i] - mn)) + ts;
      m[i] = mn;
    }
  }
  if (tx == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t q = q0 + ty * 4 + i;
      if (q < HW) stats[nb * HW + q] = make_float2(m[i], l[i]);
    }
  }
}

// Pass 2: colsum[nb][p] = sum_q exp(s[q, p] - max_q) / sum_q.   grid = (reference-position tiles, NB)
__global__ void __launch_bounds__(256)
sim_colsum_kernel(const float* __restrict__ y_t, const float* __restrict__ ref_t, int64_t B_y, int C, int64_t HW,
                  float inv_T, const float2* __restrict__ stats, float* __restrict__ colsum) {
  __shared__ __align__(16) float sA[kSimCK][kSimTile];
  __shared__ __align__(16) float sB[kSimCK][kSimTile];
  __shared__ float red[16][kSimTile];
  const int64_t nb = blockIdx.y, p0 = (int64_t)blockIdx.x * kSimTile;
  const float* yt = y_t + (nb % B_y) * (int64_t)C * HW;
  const float* rt = ref_t + nb * (int64_t)C * HW;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float col[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t q0 = 0; q0 < HW; q0 += kSimTile) {
    float2 st[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t q = q0 + ty * 4 + i;
      st[i] = q < HW ? __ldg(stats + nb * HW + q) : make_float2(0.f, 1.f);
    }
    float acc[4][4];
    sim_tile(yt, rt, C, HW, q0, p0, acc, sA, sB);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (q0 + ty * 4 + i >= HW) continue;
      const float inv_l = 1.0f / st[i].y;
#pragma unroll
      for (int j = 0; j < 4; ++j) col[j] += expf(acc[i][j] * inv_T - st[i].x) * inv_l;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) red[ty][tx * 4 + j] = col[j];
  __syncthreads();
  if (threadIdx.x < kSimTile && p0 + threadIdx.x < HW) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 16; ++r) s += red[r][threadIdx.x];     // fixed order: deterministic
    colsum[nb * HW + p0 + threadIdx.x] = s;
  }
}

// out[nb, 0:C] = x,  out[nb, C:2C] = x * colsum  (the `concat_feat` of CLM.py:22).   grid = (chunks, C, NB)
__global__ void __launch_bounds__(256)
weighted_concat_kernel(const float* __restrict__ x, const float* __restrict__ colsum, float* __restrict__ out, int C,
                       int64_t HW) {
  const int64_t nb = blockIdx.z;
  const int c = blockIdx.y;
  const float* xp = x + (nb * C + c) * HW;
  const float* cs = colsum + nb * HW;
  float* o0 = out + (nb * 2 * C + c) * HW;
  float* o1 = o0 + (int64_t)C * HW;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (int64_t)gridDim.x * blockDim.x) {
    const float v = xp[p];
    o0[p] = v;
    o1[p] = v * cs[p];
  }
}

// deform_conv (CLM.py:35-60).  One thread per pixel and group of kDefCh channels: taps outer, channels inner.
constexpr int kDefCh = 16;
__global__ void __launch_bounds__(128)
deform_fwd_kernel(const float* __restrict__ x, const float* __restrict__ offset, const float* __restrict__ modulation,
                  int mod_is_logit, float* __restrict__ out, int C, int H, int W) {
  const int64_t nb = blockIdx.z;
  const int HW = H * W;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const int c0 = blockIdx.y * kDefCh;
  const int h = p / W, w = p - h * W;
  const float* xb = x + (nb * C + c0) * HW;
  const float* ob = offset + nb * 18 * HW + p;
  const float* mb = modulation + nb * 9 * HW + p;
  float acc[kDefCh];
#pragma unroll
  for (int c = 0; c < kDefCh; ++c) acc[c] = 0.f;
  const float hmax = (float)(H - 1), wmax = (float)(W - 1);
#pragma unroll 1
  for (int k = 0; k < 9; ++k) {
    const float off_h = (float)h + ob[(2 * k) * HW];
    const float off_w = (float)w + ob[(2 * k + 1) * HW];
    if (!(off_h >= 0.f && off_h <= hmax && off_w >= 0.f && off_w <= wmax)) continue;   // (NaN offsets drop out too)
    const int h0 = (int)off_h, w0 = (int)off_w;
    const int h1 = min(h0 + 1, H - 1), w1 = min(w0 + 1, W - 1);
    const float lh = off_h - (float)h0, lw = off_w - (float)w0;
    // operation order of :52-55 -- weights first, then the four products summed left to right, no contraction
    const float w00 = __fmul_rn(1.f - lh, 1.f - lw), w10 = __fmul_rn(lh, 1.f - lw);
    const float w01 = __fmul_rn(1.f - lh, lw), w11 = __fmul_rn(lh, lw);
    float md = mb[k * HW];
    if (mod_is_logit) md = 1.0f / (1.0f + expf(-md));
    const int i00 = h0 * W + w0, i10 = h1 * W + w0, i01 = h0 * W + w1, i11 = h1 * W + w1;
#pragma unroll
    for (int c = 0; c < kDefCh; ++c) {
      if (c0 + c < C) {
        const float* xc = xb + (int64_t)c * HW;
        float v = __fadd_rn(__fmul_rn(w00, xc[i00]), __fmul_rn(w10, xc[i10]));
        v = __fadd_rn(v, __fmul_rn(w01, xc[i01]));
        v = __fadd_rn(v, __fmul_rn(w11, xc[i11]));
        acc[c] = __fadd_rn(acc[c], __fmul_rn(v, md));
      }
    }
  }
#pragma unroll
  for (int c = 0; c < kDefCh; ++c)
    if (c0 + c < C) out[(nb * C + c0 + c) * HW + p] = acc[c];
}

// CLM.py:117-126: out = sum_r softmax_r(att)[r] * aligned[r] + y.   aligned [R, B, C, S], att [R, B, S].
constexpr int kAttMaxRefs = 8;
__global__ void __launch_bounds__(128)
attention_sum_kernel(const float* __restrict__ aligned, const float* __restrict__ att, const float* __restrict__ y,
                     float* __restrict__ out, int R, int64_t B, int C, int64_t S) {
  const int64_t b = blockIdx.z;
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= S) return;
  const int c0 = blockIdx.y * kDefCh;
  float wgt[kAttMaxRefs], mx = -INFINITY, den = 0.f;
  for (int r = 0; r < R; ++r) { wgt[r] = att[((int64_t)r * B + b) * S + p]; mx = fmaxf(mx, wgt[r]); }
  for (int r = 0; r < R; ++r) { wgt[r] = expf(wgt[r] - mx); den += wgt[r]; }
  for (int r = 0; r < R; ++r) wgt[r] = wgt[r] / den;
  for (int c = c0; c < c0 + kDefCh && c < C; ++c) {
    float acc = 0.f;
    for (int r = 0; r < R; ++r) acc = __fadd_rn(acc, __fmul_rn(aligned[(((int64_t)r * B + b) * C + c) * S + p], wgt[r]));
    out[(b * C + c) * S + p] = acc + y[(b * C + c) * S + p];
  }
}

}  // namespace clc

using namespace clc;

extern "C" size_t clc_clm_sim_colsum_workspace_bytes(int64_t NB, int64_t HW) {
  if (NB < 0 || HW < 0) return 0;
  return (size_t)NB * (size_t)HW * sizeof(float2);
}

extern "C" int clc_clm_sim_colsum(const float* y_t, const float* ref_t, int64_t NB, int64_t B_y, int32_t C, int64_t HW,
                                  float temperature, float* colsum, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  if (!y_t || !ref_t || !colsum || NB < 0 || B_y < 1 || C < 1 || HW < 1 || !(temperature > 0.f))
    return CLC_ERR_INVALID_ARGUMENT;
  if (NB == 0) return CLC_OK;
  if (NB > 65535) return CLC_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < clc_clm_sim_colsum_workspace_bytes(NB, HW)) return CLC_ERR_WORKSPACE;
  if ((HW & 3) == 0 && (!aligned16(y_t) || !aligned16(ref_t))) return CLC_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t tiles = (HW + kSimTile - 1) / kSimTile;
  if (tiles > 0x7fffffff) return CLC_ERR_UNSUPPORTED;
  dim3 grid((unsigned)tiles, (unsigned)NB);
  float2* stats = reinterpret_cast<float2*>(workspace);
  const float inv_T = 1.0f / temperature;
  sim_row_stats_kernel<<<grid, 256, 0, st>>>(y_t, ref_t, B_y, C, HW, inv_T, stats);
  CLC_CHECK_LAUNCH("clc_clm_sim_colsum(row stats)");
  sim_colsum_kernel<<<grid, 256, 0, st>>>(y_t, ref_t, B_y, C, HW, inv_T, stats, colsum);
  CLC_CHECK_LAUNCH("clc_clm_sim_colsum(column sums)");
  return CLC_OK;
}

extern "C" int clc_clm_weighted_concat(const float* x, const float* colsum, float* out, int64_t NB, int32_t C,
                                       int64_t HW, void* stream) {
  if (!x || !colsum || !out || NB < 0 || C < 1 || HW < 1) return CLC_ERR_INVALID_ARGUMENT;
  if (NB == 0) return CLC_OK;
  if (NB > 65535 || C > 65535) return CLC_ERR_UNSUPPORTED;
  int64_t gx = (HW + 255) / 256;
  if (gx > 64) gx = 64;
  weighted_concat_kernel<<<dim3((unsigned)gx, (unsigned)C, (unsigned)NB), 256, 0, (cudaStream_t)stream>>>(x, colsum, out,
                                                                                                           C, HW);
  CLC_CHECK_LAUNCH("clc_clm_weighted_concat");
  return CLC_OK;
}

extern "C" int clc_clm_deform_fwd(const float* x, const float* offset, const float* modulation,
                                  int32_t modulation_is_logit, float* out, int64_t NB, int32_t C, int32_t H, int32_t W,
                                  void* stream) {
  if (!x || !offset || !modulation || !out || NB < 0 || C < 1 || H < 1 || W < 1) return CLC_ERR_INVALID_ARGUMENT;
  if (NB == 0) return CLC_OK;
  if (NB > 65535 || (int64_t)H * W > 0x7fffffff) return CLC_ERR_UNSUPPORTED;
  const unsigned gx = (unsigned)(((int64_t)H * W + 127) / 128), gy = (unsigned)((C + kDefCh - 1) / kDefCh);
  if (gy > 65535) return CLC_ERR_UNSUPPORTED;
  deform_fwd_kernel<<<dim3(gx, gy, (unsigned)NB), 128, 0, (cudaStream_t)stream>>>(x, offset, modulation,
                                                                                  modulation_is_logit, out, C, H, W);
  CLC_CHECK_LAUNCH("clc_clm_deform_fwd");
  return CLC_OK;
}

extern "C" int clc_clm_attention_sum_fwd(const float* aligned, const float* att, const float* y, float* out, int32_t R,
                                         int64_t B, int32_t C, int64_t S, void* stream) {
  if (!aligned || !att || !y || !out || R < 1 || B < 0 || C < 1 || S < 1) return CLC_ERR_INVALID_ARGUMENT;
  if (R > kAttMaxRefs || B > 65535) return CLC_ERR_UNSUPPORTED;
  if (B == 0) return CLC_OK;
  const int64_t gx = (S + 127) / 128;
  const unsigned gy = (unsigned)((C + kDefCh - 1) / kDefCh);
  if (gx > 0x7fffffff || gy > 65535) return CLC_ERR_UNSUPPORTED;
  attention_sum_kernel<<<dim3((unsigned)gx, gy, (unsigned)B), 128, 0, (cudaStream_t)stream>>>(aligned, att, y, out, R, B,
                                                                                              C, S);
  CLC_CHECK_LAUNCH("clc_clm_attention_sum_fwd");
  return CLC_OK;
}
