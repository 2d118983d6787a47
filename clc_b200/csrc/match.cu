// Reference matching, exact-fp32 path ("fp32 mode"): Pearson correlation map on CUDA cores
// with sequential fp32 FMA accumulation, warp-shuffle top-k, Gaussian mask, gather/blend and
// the backward kernels.  The tcgen05 screening path lives in match_tc.cu and shares the
// statistics / re-scoring helpers declared in match.cuh.
//
// Reference arithmetic followed (file:line into the reference tree):
//   L2_or_pearson_corr ........ models/Patch_Matching.py:854-910
//   create_gaussian_masks ..... models/Patch_Matching.py:779-807
//   SI_Wraper ................. models/Patch_Matching.py:218-240
#include "match.cuh"

#include <cooperative_groups.h>

namespace clc {

// ------------------------------------------------------------------------------------------
// Per-pixel channel sums of the reference-side features: S1 = sum_c r, S2 = sum_c r^2.
// One thread per pixel, looping over channels (coalesced across pixels).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
channel_sums_kernel(const float* __restrict__ r, float* __restrict__ s1, float* __restrict__ s2,
                    int64_t NP, int C, int64_t HW) {
  const int64_t total = NP * HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / HW, px = i - n * HW;
    const float* p = r + n * C * HW + px;
    float a = 0.f, b = 0.f;
    for (int c = 0; c < C; ++c) {
      const float v = p[(int64_t)c * HW];
      a += v;
      b = fmaf(v, v, b);
    }
    s1[i] = a;
    s2[i] = b;
  }
}

// Per-patch query statistics: xs = sum q, sxx = sum q^2 over the C*ph*pw patch elements.
// One CTA per (query image, patch): thread-strided partial sums over the C*ph patch rows,
// fixed-order block reduction (deterministic).
__global__ void __launch_bounds__(256)
patch_stats_kernel(PatchAddr qa, float* __restrict__ xs, float* __restrict__ sxx, int P, int C, int ph,
                   int pw) {
  __shared__ float red[32];
  const int w = blockIdx.x;
  const int nq = w / P, patch = w - nq * P;
  const float* base = qa.q + (int64_t)nq * qa.sn + qa.patch_off(patch);
  const int rows = C * ph;
  float a = 0.f, b = 0.f;
  for (int e = threadIdx.x; e < rows; e += blockDim.x) {
    const int c = e / ph, dy = e - c * ph;
    const float* p = base + (int64_t)c * qa.sc + (int64_t)dy * qa.sy;
    for (int dx = 0; dx < pw; ++dx) {
      const float v = p[dx];
      a += v;
      b = fmaf(v, v, b);
    }
  }
  a = block_sum(a, red);
  b = block_sum(b, red);
  if (threadIdx.x == 0) { xs[w] = a; sxx[w] = b; }
}

// ------------------------------------------------------------------------------------------
// Correlation GEMM (fp32 FMA, 64x64 tile, 4x4 micro-tile) with the Pearson epilogue.
//   xy[patch, pos] = sum_{c,dy,dx} q[patch,c,dy,dx] * r[c, oy+dy, ox+dx]
// accumulated in ascending k = (c, dy, dx) order with one fp32 FMA chain per output.
// ------------------------------------------------------------------------------------------
constexpr int TM = 64, TN = 64, TK = 16;

__global__ void __launch_bounds__(256)
pearson_corr_kernel(PatchAddr qa, const float* __restrict__ r, const float* __restrict__ s1,
                    const float* __restrict__ s2, const float* __restrict__ xs,
                    const float* __restrict__ sxx, const float* __restrict__ mask,
                    float* __restrict__ corr, int P, int C, int ph, int pw, int fh, int fw) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int ch = fh - ph + 1, cw = fw - pw + 1, L = ch * cw;
  const int pp = ph * pw, K = C * pp;
  const int64_t HW = (int64_t)fh * fw;
  const int n = blockIdx.z;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int nq = n / qa.repeat;
  const float* qbase = qa.q + (int64_t)nq * qa.sn;
  const float* rbase = r + (int64_t)n * C * HW;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;

  // Per-thread load coordinates (fixed across the K loop).
  //   A: 4 elements, e = tid + i*256 -> k_local = e % 16, patch_local = e / 16
  //   B: 4 elements, e = tid + i*256 -> pos_local = e % 64, k_local = e / 64
  int64_t a_off[4];
  bool a_ok[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int e = tid + i * 256;
    const int pl = e >> 4;
    a_ok[i] = (m0 + pl) < P;
    a_off[i] = a_ok[i] ? qa.patch_off(m0 + pl) : 0;
  }
  int64_t b_off[4];
  bool b_ok[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int e = tid + i * 256;
    const int pos = n0 + (e & 63);
    b_ok[i] = pos < L;
    const int oy = b_ok[i] ? pos / cw : 0, ox = b_ok[i] ? pos - oy * cw : 0;
    b_off[i] = (int64_t)oy * fw + ox;
  }

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += TK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;
      {  // A
        const int kl = e & 15, pl = e >> 4;
        const int k = k0 + kl;
        float v = 0.f;
        if (a_ok[i] && k < K) {
          const int c = k / pp, rem = k - c * pp;
          const int dy = rem / pw, dx = rem - dy * pw;
          v = qbase[a_off[i] + (int64_t)c * qa.sc + (int64_t)dy * qa.sy + dx];
        }
        As[kl][pl] = v;
      }
      {  // B
        const int kl = e >> 6, pl = e & 63;
        const int k = k0 + kl;
        float v = 0.f;
        if (b_ok[i] && k < K) {
          const int c = k / pp, rem = k - c * pp;
          const int dy = rem / pw, dx = rem - dy * pw;
          v = rbase[(int64_t)c * HW + b_off[i] + (int64_t)dy * fw + dx];
        }
        Bs[kl][pl] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // Pearson epilogue, in the reference's operation order (Patch_Matching.py:880-905).
  const float inv_k = 1.0f / (float)K;  // kernel_mean = ones / patch_size
  const float* s1n = s1 + (int64_t)n * HW;
  const float* s2n = s2 + (int64_t)n * HW;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int pos = n0 + tx * 4 + j;
    if (pos >= L) continue;
    const int oy = pos / cw, ox = pos - oy * cw;
    float b1 = 0.f, b2 = 0.f;
    for (int dy = 0; dy < ph; ++dy)
      for (int dx = 0; dx < pw; ++dx) {
        b1 += s1n[(oy + dy) * fw + ox + dx];
        b2 += s2n[(oy + dy) * fw + ox + dx];
      }
    const PosStat ps = pos_stat(b1, b2, inv_k, (float)K);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int patch = m0 + ty * 4 + i;
      if (patch >= P) continue;
      const int64_t qi = (int64_t)nq * P + patch;
      float v = pearson(acc[i][j], ps, xs[qi], sxx[qi], (float)K);
      if (mask) v *= mask[(int64_t)patch * L + pos];
      corr[((int64_t)n * P + patch) * L + pos] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Row-wise top-k by k passes of a warp-shuffle arg-max, each pass excluding everything at or
// above the previous winner in (value desc, index asc) order.  Deterministic; ties resolve to
// the lowest index; NaN ranks above every number (as torch.topk).  One warp per row.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool key_before(float av, int ai, float bv, int bi) {
  // true if (av, ai) ranks strictly before (bv, bi)
  const bool an = av != av, bn = bv != bv;
  if (an != bn) return an;
  if (!an && av != bv) return av > bv;
  return ai < bi;
}

__global__ void __launch_bounds__(256)
topk_rows_kernel(const float* __restrict__ x, int64_t R, int64_t L, int k, float* __restrict__ val,
                 int32_t* __restrict__ idx) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= R) return;
  const float* xr = x + row * L;
  float pv = 0.f;
  int pi = -1;  // previous winner; pi < 0 means none yet
  for (int t = 0; t < k; ++t) {
    float bv = 0.f;
    int bi = -1;
    for (int64_t i = lane; i < L; i += 32) {
      const float v = xr[i];
      if (pi >= 0 && !key_before(pv, pi, v, (int)i)) continue;  // not after the previous winner
      if (bi < 0 || key_before(v, (int)i, bv, bi)) { bv = v; bi = (int)i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi >= 0 && (bi < 0 || key_before(ov, oi, bv, bi))) { bv = ov; bi = oi; }
    }
    if (lane == 0) {
      val[row * k + t] = bv;
      idx[row * k + t] = bi;
    }
    pv = bv;
    pi = bi;
    if (bi < 0) break;  // fewer than k elements
  }
}

// ------------------------------------------------------------------------------------------
// create_gaussian_masks, fp64 on the device, rounded to fp32 (the reference builds it in
// numpy float64 and casts).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gaussian_mask_kernel(float* __restrict__ mask, int img_h, int img_w, int ph, int pw) {
  const int ch = img_h - ph + 1, cw = img_w - pw + 1;
  const int P = (img_h * img_w) / (ph * pw);
  const int64_t total = (int64_t)P * ch * cw;
  const double patch_img_w = (double)img_w / (double)pw;
  const double sig_h = 0.5 * img_h, sig_w = 0.5 * img_w;
  const int r0 = (ph + 1) / 2 - 1, c0 = (pw + 1) / 2 - 1;
  const double kNeg4Ln2 = -4.0 * 0.693147180559945309417232121458;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(i / ((int64_t)ch * cw));
    const int rem = (int)(i - (int64_t)p * ch * cw);
    const int ii = rem / cw, jj = rem - ii * cw;
    const double center_h = (floor((double)p / patch_img_w) + 0.5) * ph;
    const double center_w = (fmod((double)p, patch_img_w) + 0.5) * pw;
    const double hv = (double)(ii + r0 + 1) - (double)(ph % 2) / 2.0;
    const double wv = (double)(jj + c0 + 1) - (double)(pw % 2) / 2.0;
    const double rg = ((hv - center_h) * (hv - center_h)) / (sig_h * sig_h);
    const double cg = ((wv - center_w) * (wv - center_w)) / (sig_w * sig_w);
    mask[i] = (float)exp(kNeg4Ln2 * (rg + cg));
  }
}

// ------------------------------------------------------------------------------------------
// Gather + blend (SI_Wraper :226-238).  One CTA per (problem, patch row, channel chunk): the
// softmax(val * T) weights and source offsets of the row's patches are formed once in shared
// memory, then every thread produces output pixels (coalesced along x) from the k gathered windows.
// ------------------------------------------------------------------------------------------
constexpr int kMaxK = 32;

template <bool STACK>
__global__ void __launch_bounds__(256)
gather_blend_fwd_kernel(const float* __restrict__ feat, const int32_t* __restrict__ idx,
                        const float* __restrict__ val, float temperature, float* __restrict__ out,
                        float* __restrict__ weights, int C, int fh, int fw, int gh, int gw, int corr_w,
                        int k, int CH, int cch) {
  extern __shared__ float gsm[];  // w_s[npx*k], off_s[npx*k]
  const int npx = fw / gw, P = (fh / gh) * npx;
  float* w_s = gsm;
  int* off_s = reinterpret_cast<int*>(gsm + npx * k);
  const int n = blockIdx.y;
  const int py = blockIdx.x / cch, c0 = (blockIdx.x - py * cch) * CH;
  const int c1 = min(C, c0 + CH);
  const int HW = fh * fw;
  for (int px = threadIdx.x; px < npx; px += blockDim.x) {
    const int patch = py * npx + px;
    const int64_t o = ((int64_t)n * P + patch) * k;
    float mx = -INFINITY, den = 0.f;
    if (!STACK) {
      for (int j = 0; j < k; ++j) mx = fmaxf(mx, __fmul_rn(val[o + j], temperature));
      for (int j = 0; j < k; ++j) den += expf(__fsub_rn(__fmul_rn(val[o + j], temperature), mx));   // torch: exp(v*T - max), no FMA
    }
    for (int j = 0; j < k; ++j) {
      const int id = idx[o + j];
      const int sy = id / corr_w, sx = id - sy * corr_w;
      off_s[px * k + j] = sy * fw + sx;
      if (!STACK) {
        const float wj = expf(__fsub_rn(__fmul_rn(val[o + j], temperature), mx)) / den;
        w_s[px * k + j] = wj;
        if (weights && c0 == 0) weights[o + j] = wj;
      }
    }
  }
  __syncthreads();
  const int items = (c1 - c0) * gh * fw;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int x = it % fw, rowi = it / fw;
    const int cl = rowi / gh, dy = rowi - cl * gh;
    const int c = c0 + cl;
    const int px = x / gw, dx = x - px * gw;
    const float* fp = feat + ((int64_t)n * C + c) * HW + dy * fw + dx;
    const int oo = (py * gh + dy) * fw + x;
    if (STACK) {
      for (int j = 0; j < k; ++j) out[(((int64_t)n * k + j) * C + c) * HW + oo] = fp[off_s[px * k + j]];
    } else {
      float acc = 0.f;
      // torch.sum over the k axis: sequential left-to-right accumulation of y_patch * weight
      for (int j = 0; j < k; ++j) acc += fp[off_s[px * k + j]] * w_s[px * k + j];
      out[((int64_t)n * C + c) * HW + oo] = acc;
    }
  }
}

// Block-wide sums of three values at once; results valid in every thread.
__device__ __forceinline__ void block_sum3(float& a, float& b, float& c, float (*red)[3]) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
  __syncthreads();
  if (lane == 0) { red[wid][0] = a; red[wid][1] = b; red[wid][2] = c; }
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  a = b = c = 0.f;
  for (int w = 0; w < nw; ++w) { a += red[w][0]; b += red[w][1]; c += red[w][2]; }
}

// Backward: one CTA per (n, patch).  g_feat scatter-add (windows overlap -> atomics);
// g_w[j] = sum g_out * patch_j reduced over the CTA; g_val through the softmax.
__global__ void __launch_bounds__(256)
gather_blend_bwd_kernel(const float* __restrict__ feat, const int32_t* __restrict__ idx,
                        const float* __restrict__ weights, float temperature,
                        const float* __restrict__ g_out, float* __restrict__ g_feat,
                        float* __restrict__ g_val, int C, int fh, int fw, int gh, int gw, int corr_w,
                        int k, int is_stack) {
  __shared__ float red[32];
  __shared__ float gw_s[kMaxK];
  const int npx = fw / gw, P = (fh / gh) * npx;
  const int HW = fh * fw;
  const int n = blockIdx.x / P;
  const int patch = blockIdx.x - n * P;
  const int py = patch / npx, px = patch - py * npx;
  const int rows = C * gh, pp = gh * gw;
  const int32_t* ip = idx + ((int64_t)n * P + patch) * k;
  const int dst0 = py * gh * fw + px * gw;
  for (int j = 0; j < k; ++j) {
    const int id = ip[j];
    const int sy = id / corr_w, sx = id - sy * corr_w;
    const int src0 = sy * fw + sx;
    const float wj = is_stack ? 1.f : weights[((int64_t)n * P + patch) * k + j];
    const float* gbase = is_stack ? g_out + ((int64_t)n * k + j) * C * HW : g_out + (int64_t)n * C * HW;
    float part = 0.f;
    for (int e = threadIdx.x; e < rows * gw; e += blockDim.x) {  // flat (c, dy, dx), dx fastest
      const int c = e / pp, rem = e - c * pp;
      const int dy = rem / gw, dx = rem - dy * gw;
      const float g = gbase[(int64_t)c * HW + dst0 + dy * fw + dx];
      const int64_t f = ((int64_t)n * C + c) * HW + src0 + dy * fw + dx;
      if (!is_stack) part = fmaf(g, feat[f], part);
      atomicAdd(&g_feat[f], wj * g);
    }
    if (!is_stack) {
      const float tot = block_sum(part, red);
      if (threadIdx.x == 0) gw_s[j] = tot;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float* gv = g_val + ((int64_t)n * P + patch) * k;
    if (is_stack) {
      for (int j = 0; j < k; ++j) gv[j] = 0.f;
    } else {
      // w = softmax(v*T): dL/dv_j = T * w_j * (g_w_j - sum_i w_i g_w_i)
      const float* w = weights + ((int64_t)n * P + patch) * k;
      float dot = 0.f;
      for (int j = 0; j < k; ++j) dot = fmaf(w[j], gw_s[j], dot);
      for (int j = 0; j < k; ++j) gv[j] = temperature * w[j] * (gw_s[j] - dot);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Backward of the masked Pearson value at the k selected positions.  One CTA per (n, patch).
// out = num / sqrt(D), num = xy - ym*xs, D = dY*dX, dY = syy - ym^2 K, dX = sxx - xm*xs.
// xy's query operand is detached in the reference (conv2d weights = x.data), every other use
// of x (xs, sxx, xm) is attached.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pearson_topk_bwd_kernel(PatchAddr qa, const float* __restrict__ r, const float* __restrict__ mask,
                        const int32_t* __restrict__ idx, const float* __restrict__ g_val,
                        float* __restrict__ g_r, float* __restrict__ g_q, int P, int C, int ph, int pw,
                        int fh, int fw, int k) {
  __shared__ float red[8][3];
  const int cw = fw - pw + 1, L = (fh - ph + 1) * cw;
  const int HW = fh * fw;
  const int n = blockIdx.x / P;
  const int patch = blockIdx.x - n * P;
  const int nq = n / qa.repeat;
  const float* qb = qa.q + (int64_t)nq * qa.sn + qa.patch_off(patch);
  const float* rb = r + (int64_t)n * C * HW;
  const int rows = C * ph, K = rows * pw, pp = ph * pw;
  const float Kf = (float)K, inv_k = 1.0f / Kf;

  // patch statistics
  float a = 0.f, b = 0.f, z = 0.f;
  for (int e = threadIdx.x; e < rows; e += blockDim.x) {
    const int c = e / ph, dy = e - c * ph;
    const float* qp = qb + (int64_t)c * qa.sc + (int64_t)dy * qa.sy;
    for (int dx = 0; dx < pw; ++dx) {
      const float v = qp[dx];
      a += v;
      b = fmaf(v, v, b);
    }
  }
  block_sum3(a, b, z, red);
  const float xs = a, sxx = b;
  const float xm = xs / Kf;
  const float dX = sxx - xm * xs;

  float t_gxs = 0.f, t_gsxx = 0.f, t_gxm = 0.f;  // accumulated over the k positions
  for (int j = 0; j < k; ++j) {
    const int id = idx[((int64_t)n * P + patch) * k + j];
    const int oy = id / cw, ox = id - oy * cw;
    const int src0 = oy * fw + ox;
    float s1 = 0.f, s2 = 0.f, xy = 0.f;
    for (int e = threadIdx.x; e < rows; e += blockDim.x) {
      const int c = e / ph, dy = e - c * ph;
      const float* qp = qb + (int64_t)c * qa.sc + (int64_t)dy * qa.sy;
      const float* rp = rb + (int64_t)c * HW + src0 + dy * fw;
      for (int dx = 0; dx < pw; ++dx) {
        const float qv = qp[dx], rv = rp[dx];
        s1 += rv;
        s2 = fmaf(rv, rv, s2);
        xy = fmaf(qv, rv, xy);
      }
    }
    block_sum3(s1, s2, xy, red);
    const float ym = s1 * inv_k;
    const float dY = s2 - ym * ym * Kf;
    const float D = dY * dX;
    const float num = xy - ym * xs;
    const float rs = rsqrtf(D);
    float g = g_val[((int64_t)n * P + patch) * k + j];
    if (mask) g *= mask[(int64_t)patch * L + id];
    const float g_num = g * rs;
    const float g_D = -0.5f * g * num * rs / D;
    const float g_dY = g_D * dX, g_dX = g_D * dY;
    const float g_ym = -g_num * xs - 2.f * g_dY * ym * Kf;
    const float g_xy = g_num;
    t_gxs += -g_num * ym - g_dX * xm;
    t_gsxx += g_dX;
    t_gxm += -g_dX * xs;
    const float c_mean = g_ym * inv_k;
    for (int e = threadIdx.x; e < K; e += blockDim.x) {  // flat (c, dy, dx), dx fastest
      const int c = e / pp, rem = e - c * pp;
      const int dy = rem / pw, dx = rem - dy * pw;
      const float qv = qb[(int64_t)c * qa.sc + (int64_t)dy * qa.sy + dx];
      const int64_t ro = (int64_t)c * HW + src0 + dy * fw + dx;
      atomicAdd(&g_r[(int64_t)n * C * HW + ro], fmaf(g_xy, qv, fmaf(2.f * g_dY, rb[ro], c_mean)));
    }
  }
  if (g_q) {
    float* gqb = g_q + (int64_t)nq * qa.sn + qa.patch_off(patch);
    const float cst = t_gxs + t_gxm * inv_k;
    for (int e = threadIdx.x; e < rows; e += blockDim.x) {
      const int c = e / ph, dy = e - c * ph;
      const int64_t o = (int64_t)c * qa.sc + (int64_t)dy * qa.sy;
      for (int dx = 0; dx < pw; ++dx) atomicAdd(&gqb[o + dx], fmaf(2.f * t_gsxx, qb[o + dx], cst));
    }
  }
}

// ------------------------------------------------------------------------------------------
// Fused backward of match + gather/blend for the single-scale case in which the gathered
// feature map IS the matched reference (feat == r, gather patch == match patch): the two
// scatters hit the same k windows, so one pass forms every reduction (g_w, s1, s2, xy per
// window), one thread turns them into the softmax / Pearson coefficients, and a single pass
// issues ONE atomic per window element:
//   g_r[win_j] += w_j * g_out[patch]  +  g_xy_j * q[patch] + 2 g_dY_j * r[win_j] + c_mean_j
// Elements are indexed flat (c, dy, dx) with dx fastest so the 32 lanes of a warp-wide atomic
// fall into 32-byte sectors pw at a time.  One CTA per (problem, patch).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_sum4(float& a, float& b, float& c, float& d, float (*red)[4]) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c); d = warp_sum(d);
  __syncthreads();
  if (lane == 0) { red[wid][0] = a; red[wid][1] = b; red[wid][2] = c; red[wid][3] = d; }
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  a = b = c = d = 0.f;
  for (int w = 0; w < nw; ++w) { a += red[w][0]; b += red[w][1]; c += red[w][2]; d += red[w][3]; }
}

__global__ void __launch_bounds__(256)
match_bwd_kernel(PatchAddr qa, const float* __restrict__ r, const float* __restrict__ mask,
                 const int32_t* __restrict__ idx, const float* __restrict__ weights, float temperature,
                 const float* __restrict__ g_out, float* __restrict__ g_r, float* __restrict__ g_q,
                 float* __restrict__ g_val_out, int P, int C, int ph, int pw, int fh, int fw, int k) {
  __shared__ float red[8][4];
  __shared__ float sums[kMaxK][4];   // per window: g_w, s1, s2, xy
  __shared__ float coef[kMaxK][4];   // per window: w_j, g_xy, 2*g_dY, c_mean
  __shared__ int src_s[kMaxK];
  __shared__ float qcoef[2];         // g_q[e] += qcoef[0] * q[e] + qcoef[1]
  const int cw = fw - pw + 1, L = (fh - ph + 1) * cw;
  const int HW = fh * fw;
  const int n = blockIdx.x / P;
  const int patch = blockIdx.x - n * P;
  const int nq = n / qa.repeat;
  const int npx = fw / pw;
  const int py = patch / npx, px = patch - py * npx;
  const float* qb = qa.q + (int64_t)nq * qa.sn + qa.patch_off(patch);
  const float* rb = r + (int64_t)n * C * HW;
  const float* gb = g_out + (int64_t)n * C * HW + (py * ph) * fw + px * pw;
  const int pp = ph * pw, K = C * pp;
  const float Kf = (float)K, inv_k = 1.0f / Kf;
  const int64_t po = ((int64_t)n * P + patch) * k;
  if (threadIdx.x < k) {
    const int id = idx[po + threadIdx.x];
    const int oy = id / cw, ox = id - oy * cw;
    src_s[threadIdx.x] = oy * fw + ox;
  }
  // patch statistics
  float a = 0.f, b = 0.f, z0 = 0.f, z1 = 0.f;
  for (int e = threadIdx.x; e < K; e += blockDim.x) {
    const int c = e / pp, rem = e - c * pp;
    const int dy = rem / pw, dx = rem - dy * pw;
    const float v = qb[(int64_t)c * qa.sc + (int64_t)dy * qa.sy + dx];
    a += v;
    b = fmaf(v, v, b);
  }
  block_sum4(a, b, z0, z1, red);  // (also orders the src_s writes before their use)
  const float xs = a, sxx = b;
  // per-window reductions
  for (int j = 0; j < k; ++j) {
    const int src0 = src_s[j];
    float gw = 0.f, s1 = 0.f, s2 = 0.f, xy = 0.f;
    for (int e = threadIdx.x; e < K; e += blockDim.x) {
      const int c = e / pp, rem = e - c * pp;
      const int dy = rem / pw, dx = rem - dy * pw;
      const float qv = qb[(int64_t)c * qa.sc + (int64_t)dy * qa.sy + dx];
      const float g = gb[(int64_t)c * HW + dy * fw + dx];
      const float rv = rb[(int64_t)c * HW + src0 + dy * fw + dx];
      gw = fmaf(g, rv, gw);
      s1 += rv;
      s2 = fmaf(rv, rv, s2);
      xy = fmaf(qv, rv, xy);
    }
    block_sum4(gw, s1, s2, xy, red);
    if (threadIdx.x == 0) { sums[j][0] = gw; sums[j][1] = s1; sums[j][2] = s2; sums[j][3] = xy; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float xm = xs / Kf;
    const float dX = sxx - xm * xs;
    // w = softmax(v*T): dL/dv_j = T * w_j * (g_w_j - sum_i w_i g_w_i)
    float dot = 0.f;
    for (int j = 0; j < k; ++j) dot = fmaf(weights[po + j], sums[j][0], dot);
    float t_gxs = 0.f, t_gsxx = 0.f, t_gxm = 0.f;
    for (int j = 0; j < k; ++j) {
      const float wj = weights[po + j];
      const float gv = temperature * wj * (sums[j][0] - dot);
      if (g_val_out) g_val_out[po + j] = gv;
      const float s1 = sums[j][1], s2 = sums[j][2], xy = sums[j][3];
      const float ym = s1 * inv_k;
      const float dY = s2 - ym * ym * Kf;
      const float D = dY * dX;
      const float num = xy - ym * xs;
      const float rs = rsqrtf(D);
      float g = gv;
      if (mask) g *= mask[(int64_t)patch * L + idx[po + j]];
      const float g_num = g * rs;
      const float g_D = -0.5f * g * num * rs / D;
      const float g_dY = g_D * dX, g_dX = g_D * dY;
      const float g_ym = -g_num * xs - 2.f * g_dY * ym * Kf;
      t_gxs += -g_num * ym - g_dX * xm;
      t_gsxx += g_dX;
      t_gxm += -g_dX * xs;
      coef[j][0] = wj; coef[j][1] = g_num; coef[j][2] = 2.f * g_dY; coef[j][3] = g_ym * inv_k;
    }
    qcoef[0] = 2.f * t_gsxx;
    qcoef[1] = t_gxs + t_gxm * inv_k;
  }
  __syncthreads();
  float* grb = g_r + (int64_t)n * C * HW;
  for (int j = 0; j < k; ++j) {
    const int src0 = src_s[j];
    const float cw_ = coef[j][0], cxy = coef[j][1], cdy = coef[j][2], cm = coef[j][3];
    for (int e = threadIdx.x; e < K; e += blockDim.x) {
      const int c = e / pp, rem = e - c * pp;
      const int dy = rem / pw, dx = rem - dy * pw;
      const float qv = qb[(int64_t)c * qa.sc + (int64_t)dy * qa.sy + dx];
      const float g = gb[(int64_t)c * HW + dy * fw + dx];
      const int64_t ro = (int64_t)c * HW + src0 + dy * fw + dx;
      const float rv = rb[ro];
      atomicAdd(&grb[ro], fmaf(cw_, g, fmaf(cxy, qv, fmaf(cdy, rv, cm))));
    }
  }
  if (g_q) {
    float* gqb = g_q + (int64_t)nq * qa.sn + qa.patch_off(patch);
    const float c0 = qcoef[0], c1 = qcoef[1];
    for (int e = threadIdx.x; e < K; e += blockDim.x) {
      const int c = e / pp, rem = e - c * pp;
      const int dy = rem / pw, dx = rem - dy * pw;
      const int64_t o = (int64_t)c * qa.sc + (int64_t)dy * qa.sy + dx;
      atomicAdd(&gqb[o], fmaf(c0, qb[o], c1));
    }
  }
}

// ------------------------------------------------------------------------------------------
// Channels-last variant of the fused backward (the fast path).  With the reference latents
// transposed to [pixel][C] every window row is one contiguous run of C floats, so all loads are
// coalesced float4 and the scatter is a float4 vector atomic (red.global.add.v4.f32) into a
// channels-last gradient scratch that a tiled transpose then adds into the NCHW output.
// The query and upstream-gradient patches are staged once per CTA in shared memory as [shift][C].
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
nchw_to_cl_kernel(const float* __restrict__ x, float* __restrict__ xT, int C, int HW) {
  __shared__ float t[32][33];
  pdl_trigger();
  pdl_wait();
  const int n = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* xn = x + (int64_t)n * C * HW;
  float* xTn = xT + (int64_t)n * C * HW;
  float v[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int c = c0 + threadIdx.y + 8 * u, p = p0 + threadIdx.x;
    v[u] = (c < C && p < HW) ? xn[(int64_t)c * HW + p] : 0.f;
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) t[threadIdx.y + 8 * u][threadIdx.x] = v[u];
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (c < C && p < HW) xTn[(int64_t)p * C + c] = t[threadIdx.x][i];
  }
}

// SimpleCLM attention-logit gradient from the per-reference G_r = sum_c g_fused_c * aligned_r,c the CLM-fused match
// backward published ([NQ*P][R][16], 4 x 4 patches):
//   g_att_m = w_m s_m G_m - w_m sum_r G_r s_r w_r + G_m w_m s_m (1 - s_m),   w = softmax_r(att), s = sigmoid(att)
struct ClmAttGrad {
  const float* g_scratch;   // NULL = nothing to do
  const float* att;
  float* g_att;
  int64_t att_sr, att_sb;
  int R, P, fw;
};

// x[n][c][p] (+)= xT[n][p][c].  With `ga`, the blocks (*, 0, first reference of an image) also form g_att for their
// 32 pixels: the R CTAs of a patch in the backward kernel only publish their G_r -- no fence, no arrival counter,
// no cluster barrier on that kernel's critical path (each of those cost 2-3.4 us per CTA there).
template <bool ADD>
__global__ void __launch_bounds__(256)
cl_to_nchw_kernel(const float* __restrict__ xT, float* __restrict__ x, int C, int HW, const ClmAttGrad ga) {
  __shared__ float t[32][33];
  pdl_trigger();
  pdl_wait();
  const int n = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  if (ga.g_scratch != nullptr && blockIdx.y == 0 && n % ga.R == 0 && threadIdx.y == 0 && p0 + (int)threadIdx.x < HW) {
    const int nq = n / ga.R, pix = p0 + threadIdx.x;
    const int y = pix / ga.fw, xx = pix - y * ga.fw;
    const int patch = (y >> 2) * (ga.fw >> 2) + (xx >> 2), o = (y & 3) * 4 + (xx & 3);
    const float* gs = ga.g_scratch + ((int64_t)nq * ga.P + patch) * ga.R * 16 + o;
    float av[8], w[8], sg[8], Gt[8];
    float mx = -INFINITY, den = 0.f, mix = 0.f;
    for (int r = 0; r < ga.R; ++r) {
      av[r] = ga.att[(int64_t)r * ga.att_sr + (int64_t)nq * ga.att_sb + pix];
      Gt[r] = gs[r * 16];
      mx = fmaxf(mx, av[r]);
    }
    for (int r = 0; r < ga.R; ++r) { w[r] = expf(av[r] - mx); den += w[r]; }
    for (int r = 0; r < ga.R; ++r) {
      w[r] = w[r] / den;
      sg[r] = 1.0f / (1.0f + expf(-av[r]));
      mix = fmaf(Gt[r], w[r] * sg[r], mix);          // sum_r G_r s_r w_r
    }
    for (int m = 0; m < ga.R; ++m) {
      const float cm = w[m] * sg[m];
      ga.g_att[(int64_t)m * ga.att_sr + (int64_t)nq * ga.att_sb + pix] = cm * Gt[m] - w[m] * mix + Gt[m] * cm * (1.f - sg[m]);
    }
  }
  const float* xTn = xT + (int64_t)n * C * HW;
  float* xn = x + (int64_t)n * C * HW;
  float v[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int p = p0 + threadIdx.y + 8 * u, c = c0 + threadIdx.x;
    v[u] = (c < C && p < HW) ? xTn[(int64_t)p * C + c] : 0.f;
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) t[threadIdx.y + 8 * u][threadIdx.x] = v[u];
  __syncthreads();
  if (ADD) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = c0 + threadIdx.y + 8 * u, p = p0 + threadIdx.x;
      v[u] = (c < C && p < HW) ? xn[(int64_t)c * HW + p] : 0.f;
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = threadIdx.y + 8 * u;
    const int c = c0 + i, p = p0 + threadIdx.x;
    if (c < C && p < HW) xn[(int64_t)c * HW + p] = ADD ? v[u] + t[threadIdx.x][i] : t[threadIdx.x][i];
  }
}

__device__ __forceinline__ void red_add4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// Requires C % 4 == 0 and pw == 4 with 16-byte aligned patch rows in q / g_out / g_q.
__global__ void __launch_bounds__(256)
match_bwd_cl_kernel(PatchAddr qa, const float* __restrict__ rT, const float* __restrict__ mask,
                    const int32_t* __restrict__ idx, const float* __restrict__ weights, float temperature,
                    const float* __restrict__ g_out, float* __restrict__ g_rT, float* __restrict__ g_q,
                    float* __restrict__ g_val_out, int P, int C, int ph, int pw, int fh, int fw, int k) {
  extern __shared__ float4 sm4[];           // Q[S][C/4], G[S][C/4]
  __shared__ float red[8][4];
  __shared__ float sums[kMaxK][4];          // per window: g_w, s1, s2, xy
  __shared__ float coef[kMaxK][4];          // per window: w_j, g_xy, 2*g_dY, c_mean
  __shared__ int src_s[kMaxK];
  __shared__ float qcoef[2];
  const int cw = fw - pw + 1, L = (fh - ph + 1) * cw;
  const int HW = fh * fw;
  const int n = blockIdx.x / P;
  const int patch = blockIdx.x - n * P;
  const int nq = n / qa.repeat;
  const int npx = fw / pw;
  const int py = patch / npx, px = patch - py * npx;
  const int S = ph * pw, c4n = C >> 2, K = C * S;
  const int items = S * c4n;                // float4 items of one patch / window
  float4* Q = sm4;
  float4* G = sm4 + items;
  float* Qf = reinterpret_cast<float*>(Q);
  float* Gf = reinterpret_cast<float*>(G);
  const float Kf = (float)K, inv_k = 1.0f / Kf;
  const int64_t po = ((int64_t)n * P + patch) * k;
  if (threadIdx.x < k) {
    const int id = idx[po + threadIdx.x];
    const int oy = id / cw, ox = id - oy * cw;
    src_s[threadIdx.x] = oy * fw + ox;
  }
  // stage q and g patches: global rows (c, dy) of pw = 4 floats -> shared [s = dy*4+dx][c]
  const float* qb = qa.q + (int64_t)nq * qa.sn + qa.patch_off(patch);
  const float* gb = g_out + (int64_t)n * C * HW + (py * ph) * fw + px * pw;
  for (int e = threadIdx.x; e < C * ph; e += blockDim.x) {
    const int dy = e / C, c = e - dy * C;   // c fastest: conflict-free shared stores
    const float4 qv = ld4(qb + (int64_t)c * qa.sc + (int64_t)dy * qa.sy);
    const float4 gv = ld4(gb + (int64_t)c * HW + dy * fw);
    const int s0 = dy * 4;
    Qf[(s0 + 0) * C + c] = qv.x; Qf[(s0 + 1) * C + c] = qv.y; Qf[(s0 + 2) * C + c] = qv.z; Qf[(s0 + 3) * C + c] = qv.w;
    Gf[(s0 + 0) * C + c] = gv.x; Gf[(s0 + 1) * C + c] = gv.y; Gf[(s0 + 2) * C + c] = gv.z; Gf[(s0 + 3) * C + c] = gv.w;
  }
  __syncthreads();
  // patch statistics
  float a = 0.f, b = 0.f, z0 = 0.f, z1 = 0.f;
  for (int f = threadIdx.x; f < items; f += blockDim.x) {
    const float4 v = Q[f];
    a += (v.x + v.y) + (v.z + v.w);
    b = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, b))));
  }
  block_sum4(a, b, z0, z1, red);
  const float xs = a, sxx = b;
  const float* rTn = rT + (int64_t)n * HW * C;
  // per-window reductions
  for (int j = 0; j < k; ++j) {
    const int src0 = src_s[j];
    float gw = 0.f, s1 = 0.f, s2 = 0.f, xy = 0.f;
    for (int f = threadIdx.x; f < items; f += blockDim.x) {
      const int s = f / c4n, c4 = f - s * c4n;
      const int dy = s >> 2, dx = s & 3;
      const float4 rv = ld4(rTn + (int64_t)(src0 + dy * fw + dx) * C + 4 * c4);
      const float4 qv = Q[f], gv = G[f];
      gw = fmaf(gv.x, rv.x, fmaf(gv.y, rv.y, fmaf(gv.z, rv.z, fmaf(gv.w, rv.w, gw))));
      s1 += (rv.x + rv.y) + (rv.z + rv.w);
      s2 = fmaf(rv.x, rv.x, fmaf(rv.y, rv.y, fmaf(rv.z, rv.z, fmaf(rv.w, rv.w, s2))));
      xy = fmaf(qv.x, rv.x, fmaf(qv.y, rv.y, fmaf(qv.z, rv.z, fmaf(qv.w, rv.w, xy))));
    }
    block_sum4(gw, s1, s2, xy, red);
    if (threadIdx.x == 0) { sums[j][0] = gw; sums[j][1] = s1; sums[j][2] = s2; sums[j][3] = xy; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float xm = xs / Kf;
    const float dX = sxx - xm * xs;
    float dot = 0.f;
    for (int j = 0; j < k; ++j) dot = fmaf(weights[po + j], sums[j][0], dot);
    float t_gxs = 0.f, t_gsxx = 0.f, t_gxm = 0.f;
    for (int j = 0; j < k; ++j) {
      const float wj = weights[po + j];
      const float gv = temperature * wj * (sums[j][0] - dot);
      if (g_val_out) g_val_out[po + j] = gv;
      const float s1 = sums[j][1], s2 = sums[j][2], xy = sums[j][3];
      const float ym = s1 * inv_k;
      const float dY = s2 - ym * ym * Kf;
      const float D = dY * dX;
      const float num = xy - ym * xs;
      const float rs = rsqrtf(D);
      float g = gv;
      if (mask) g *= mask[(int64_t)patch * L + idx[po + j]];
      const float g_num = g * rs;
      const float g_D = -0.5f * g * num * rs / D;
      const float g_dY = g_D * dX, g_dX = g_D * dY;
      const float g_ym = -g_num * xs - 2.f * g_dY * ym * Kf;
      t_gxs += -g_num * ym - g_dX * xm;
      t_gsxx += g_dX;
      t_gxm += -g_dX * xs;
      coef[j][0] = wj; coef[j][1] = g_num; coef[j][2] = 2.f * g_dY; coef[j][3] = g_ym * inv_k;
    }
    qcoef[0] = 2.f * t_gsxx;
    qcoef[1] = t_gxs + t_gxm * inv_k;
  }
  __syncthreads();
  float* g_rTn = g_rT + (int64_t)n * HW * C;
  for (int j = 0; j < k; ++j) {
    const int src0 = src_s[j];
    const float cw_ = coef[j][0], cxy = coef[j][1], cdy = coef[j][2], cm = coef[j][3];
    for (int f = threadIdx.x; f < items; f += blockDim.x) {
      const int s = f / c4n, c4 = f - s * c4n;
      const int dy = s >> 2, dx = s & 3;
      const int64_t o = (int64_t)(src0 + dy * fw + dx) * C + 4 * c4;
      const float4 rv = ld4(rTn + o);
      const float4 qv = Q[f], gv = G[f];
      float4 v;
      v.x = fmaf(cw_, gv.x, fmaf(cxy, qv.x, fmaf(cdy, rv.x, cm)));
      v.y = fmaf(cw_, gv.y, fmaf(cxy, qv.y, fmaf(cdy, rv.y, cm)));
      v.z = fmaf(cw_, gv.z, fmaf(cxy, qv.z, fmaf(cdy, rv.z, cm)));
      v.w = fmaf(cw_, gv.w, fmaf(cxy, qv.w, fmaf(cdy, rv.w, cm)));
      red_add4(g_rTn + o, v);
    }
  }
  if (g_q) {
    float* gqb = g_q + (int64_t)nq * qa.sn + qa.patch_off(patch);
    const float c0 = qcoef[0], c1 = qcoef[1];
    for (int e = threadIdx.x; e < C * ph; e += blockDim.x) {
      const int dy = e / C, c = e - dy * C;
      const int s0 = dy * 4;
      float4 v;
      v.x = fmaf(c0, Qf[(s0 + 0) * C + c], c1); v.y = fmaf(c0, Qf[(s0 + 1) * C + c], c1);
      v.z = fmaf(c0, Qf[(s0 + 2) * C + c], c1); v.w = fmaf(c0, Qf[(s0 + 3) * C + c], c1);
      red_add4(gqb + (int64_t)c * qa.sc + (int64_t)dy * qa.sy, v);
    }
  }
}

// Register-resident variant for k <= 4 (the wiring's k): the k windows are loaded ONCE, all k*ITEMS
// float4 loads of a thread are in flight together, every reduction of the CTA (patch statistics +
// 4 sums per window) goes through ONE block reduction, and the scatter is issued from registers.
// Three dependent global-memory phases per CTA instead of 2k+2.  ITEMS = ceil(S*C/4 / 256).
template <int ITEMS, bool REREAD>
__global__ void __launch_bounds__(256, REREAD ? 3 : 2)
match_bwd_cl_reg_kernel(PatchAddr qa, const float* __restrict__ rT, const float* __restrict__ mask,
                        const int32_t* __restrict__ idx, const float* __restrict__ weights, float temperature,
                        const float* __restrict__ g_out, float* __restrict__ g_rT, float* __restrict__ g_q,
                        float* __restrict__ g_val_out, int P, int C, int ph, int pw, int fh, int fw, int k, int dbg) {
  constexpr int KK = 4, NV = 2 + 4 * KK;
  extern __shared__ float4 sm4[];           // Q[S][C/4], G[S][C/4]
  __shared__ float red[8][NV];
  __shared__ float tot[NV];                 // xs, sxx, then per window: g_w, s1, s2, xy
  __shared__ float coef[KK][4];             // per window: w_j, g_xy, 2*g_dY, c_mean
  __shared__ float part[KK][3];             // per window: its terms of t_gxs, t_gsxx, t_gxm
  __shared__ int src_s[KK];
  __shared__ float w_s[KK], m_s[KK];        // softmax weight and mask value of each window
  const int cw = fw - pw + 1, L = (fh - ph + 1) * cw;
  const int HW = fh * fw;
  const int n = blockIdx.x / P;
  const int patch = blockIdx.x - n * P;
  const int nq = n / qa.repeat;
  const int npx = fw / pw;
  const int py = patch / npx, px = patch - py * npx;
  const int S = ph * pw, c4n = C >> 2, K = C * S;
  const int items = S * c4n;
  float4* Q = sm4;
  float4* G = sm4 + items;
  float* Qf = reinterpret_cast<float*>(Q);
  float* Gf = reinterpret_cast<float*>(G);
  const float Kf = (float)K, inv_k = 1.0f / Kf;
  const int64_t po = ((int64_t)n * P + patch) * k;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid < KK) {
    int src = 0;
    float wj = 0.f, mj = 1.f;
    if (tid < k) {
      const int id = idx[po + tid];
      wj = weights[po + tid];
      const int oy = id / cw, ox = id - oy * cw;
      src = oy * fw + ox;
      if (mask) mj = mask[(int64_t)patch * L + id];
    }
    src_s[tid] = src; w_s[tid] = wj; m_s[tid] = mj;
  }
  // stage q and g patches: global rows (c, dy) of pw = 4 floats -> shared [s = dy*4+dx][c]; all of a
  // thread's loads are issued before the first shared store (items == C*ph when pw == 4)
  const float* qb = qa.q + (int64_t)nq * qa.sn + qa.patch_off(patch);
  const float* gb = g_out + (int64_t)n * C * HW + (py * ph) * fw + px * pw;
  {
    float4 qv[ITEMS], gv[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int e = tid + i * 256;
      const int dy = e / C, c = e - dy * C;   // c fastest: conflict-free shared stores
      if (e < C * ph) {
        qv[i] = ld4(qb + (int64_t)c * qa.sc + (int64_t)dy * qa.sy);
        gv[i] = ld4(gb + (int64_t)c * HW + dy * fw);
      }
    }
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int e = tid + i * 256;
      const int dy = e / C, c = e - dy * C;
      if (e < C * ph) {
        const int s0 = dy * 4;
        Qf[(s0 + 0) * C + c] = qv[i].x; Qf[(s0 + 1) * C + c] = qv[i].y; Qf[(s0 + 2) * C + c] = qv[i].z; Qf[(s0 + 3) * C + c] = qv[i].w;
        Gf[(s0 + 0) * C + c] = gv[i].x; Gf[(s0 + 1) * C + c] = gv[i].y; Gf[(s0 + 2) * C + c] = gv[i].z; Gf[(s0 + 3) * C + c] = gv[i].w;
      }
    }
  }
  __syncthreads();
  // ---- all window loads of this thread, issued together ----
  const float* rTn = rT + (int64_t)n * HW * C;
  int woff[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int f = tid + i * 256;
    const int s = f / c4n, c4 = f - s * c4n;
    woff[i] = (f < items) ? ((s >> 2) * fw + (s & 3)) * C + 4 * c4 : -1;
  }
  // REREAD = false: all KK windows stay in registers between the reduction and the scatter (128
  // registers, 2 CTAs / SM).  REREAD = true: windows are processed two at a time and read again for
  // the scatter (L1 / L2 hits), 85 registers, 3 CTAs / SM -- chosen when it saves a wave of CTAs.
  constexpr int WB = REREAD ? 2 : KK;        // windows in flight
  float4 rv[WB][ITEMS];
  float v[NV];
#pragma unroll
  for (int t = 0; t < NV; ++t) v[t] = 0.f;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int f = tid + i * 256;
    if (f >= items) continue;
    const float4 q4 = Q[f];
    v[0] += (q4.x + q4.y) + (q4.z + q4.w);
    v[1] = fmaf(q4.x, q4.x, fmaf(q4.y, q4.y, fmaf(q4.z, q4.z, fmaf(q4.w, q4.w, v[1]))));
  }
#pragma unroll
  for (int j0 = 0; j0 < KK; j0 += WB) {
#pragma unroll
    for (int jj = 0; jj < WB; ++jj)
#pragma unroll
      for (int i = 0; i < ITEMS; ++i)
        rv[jj][i] = (j0 + jj < k && woff[i] >= 0) ? ld4(rTn + (int64_t)src_s[j0 + jj] * C + woff[i])
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int f = tid + i * 256;
      if (f >= items) continue;
      const float4 q4 = Q[f], g4 = G[f];
#pragma unroll
      for (int jj = 0; jj < WB; ++jj) {
        const int j = j0 + jj;
        const float4 r4 = rv[jj][i];
        v[2 + 4 * j] = fmaf(g4.x, r4.x, fmaf(g4.y, r4.y, fmaf(g4.z, r4.z, fmaf(g4.w, r4.w, v[2 + 4 * j]))));
        v[3 + 4 * j] += (r4.x + r4.y) + (r4.z + r4.w);
        v[4 + 4 * j] = fmaf(r4.x, r4.x, fmaf(r4.y, r4.y, fmaf(r4.z, r4.z, fmaf(r4.w, r4.w, v[4 + 4 * j]))));
        v[5 + 4 * j] = fmaf(q4.x, r4.x, fmaf(q4.y, r4.y, fmaf(q4.z, r4.z, fmaf(q4.w, r4.w, v[5 + 4 * j]))));
      }
    }
  }
#pragma unroll
  for (int t = 0; t < NV; ++t) {
    v[t] = warp_sum(v[t]);
    if (lane == 0) red[wid][t] = v[t];
  }
  __syncthreads();
  if (tid < NV) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += red[w][tid];
    tot[tid] = a;
  }
  __syncthreads();
  // ---- per-window coefficients: thread j <-> window j ----
  if (tid < k) {
    const int j = tid;
    const float xs = tot[0], sxx = tot[1];
    const float xm = xs / Kf;
    const float dX = sxx - xm * xs;
    float dot = 0.f;   // w = softmax(v*T): dL/dv_j = T * w_j * (g_w_j - sum_i w_i g_w_i)
    for (int i = 0; i < k; ++i) dot = fmaf(w_s[i], tot[2 + 4 * i], dot);
    const float wj = w_s[j];
    const float gv = temperature * wj * (tot[2 + 4 * j] - dot);
    if (g_val_out) g_val_out[po + j] = gv;
    const float s1 = tot[3 + 4 * j], s2 = tot[4 + 4 * j], xy = tot[5 + 4 * j];
    const float ym = s1 * inv_k;
    const float dY = s2 - ym * ym * Kf;
    const float D = dY * dX;
    const float num = xy - ym * xs;
    const float rs = rsqrtf(D);
    float g = gv;
    if (mask) g *= m_s[j];
    const float g_num = g * rs;
    const float g_D = -0.5f * g * num * rs / D;
    const float g_dY = g_D * dX, g_dX = g_D * dY;
    const float g_ym = -g_num * xs - 2.f * g_dY * ym * Kf;
    part[j][0] = -g_num * ym - g_dX * xm;
    part[j][1] = g_dX;
    part[j][2] = -g_dX * xs;
    coef[j][0] = wj; coef[j][1] = g_num; coef[j][2] = 2.f * g_dY; coef[j][3] = g_ym * inv_k;
  }
  __syncthreads();
  // ---- scatter: ONE vector atomic per window element ----
  float* g_rTn = g_rT + (int64_t)n * HW * C;
#pragma unroll
  for (int j0 = 0; j0 < KK; j0 += WB) {
    if (j0 >= k) break;
    if (REREAD) {
#pragma unroll
      for (int jj = 0; jj < WB; ++jj)
#pragma unroll
        for (int i = 0; i < ITEMS; ++i)
          rv[jj][i] = (j0 + jj < k && woff[i] >= 0) ? ld4(rTn + (int64_t)src_s[j0 + jj] * C + woff[i])
                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int jj = 0; jj < WB; ++jj) {
      const int j = j0 + jj;
      if (j >= k) break;
      const float cw_ = coef[j][0], cxy = coef[j][1], cdy = coef[j][2], cm = coef[j][3];
      float* dst = g_rTn + (int64_t)src_s[j] * C;
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        const int f = tid + i * 256;
        if (f >= items) continue;
        const float4 q4 = Q[f], g4 = G[f], r4 = rv[jj][i];
        float4 o;
        o.x = fmaf(cw_, g4.x, fmaf(cxy, q4.x, fmaf(cdy, r4.x, cm)));
        o.y = fmaf(cw_, g4.y, fmaf(cxy, q4.y, fmaf(cdy, r4.y, cm)));
        o.z = fmaf(cw_, g4.z, fmaf(cxy, q4.z, fmaf(cdy, r4.z, cm)));
        o.w = fmaf(cw_, g4.w, fmaf(cxy, q4.w, fmaf(cdy, r4.w, cm)));
        if (!(dbg & 1)) red_add4(dst + woff[i], o);
        else if (o.x == 1.2345e33f) dst[woff[i]] = o.y + o.z + o.w;
      }
    }
  }
  if (g_q && !(dbg & 2)) {
    float t_gxs = 0.f, t_gsxx = 0.f, t_gxm = 0.f;
    for (int j = 0; j < k; ++j) { t_gxs += part[j][0]; t_gsxx += part[j][1]; t_gxm += part[j][2]; }
    const float c0 = 2.f * t_gsxx, c1 = t_gxs + t_gxm * inv_k;
    float* gqb = g_q + (int64_t)nq * qa.sn + qa.patch_off(patch);
    for (int e = tid; e < C * ph; e += 256) {
      const int dy = e / C, c = e - dy * C;
      const int s0 = dy * 4;
      float4 o;
      o.x = fmaf(c0, Qf[(s0 + 0) * C + c], c1); o.y = fmaf(c0, Qf[(s0 + 1) * C + c], c1);
      o.z = fmaf(c0, Qf[(s0 + 2) * C + c], c1); o.w = fmaf(c0, Qf[(s0 + 3) * C + c], c1);
      red_add4(gqb + (int64_t)c * qa.sc + (int64_t)dy * qa.sy, o);
    }
  }
}

// "Thread owns its items" variant (pw == 4, k <= 4, blockDim = ph * C/4 a multiple of 32): thread
// (dy, c4) owns the 4 x 4 block {dx = 0..3} x {channels 4*c4 .. 4*c4+3} of the patch.  Its q / g values
// come straight from the NCHW tensors as four float4 rows (one per channel) -- the [shift][channel]
// transposition happens in register naming, there is no shared-memory staging -- and every window
// address is base + dx * C, so the kernel has no per-item index arithmetic.  Windows are processed
// two at a time and read again for the scatter (L1 / L2 hits).
// Fusion of the SimpleCLM elementwise backward (models/CLM.py:170-182) into the match backward: the kernel
// then reads g_fused (the gradient of the FUSED feature) instead of g_aligned, forms
//   g_aligned_r = g_fused * coef_r,  coef_r = softmax_r(att) * sigmoid(att_r)        (never written to HBM)
// on the fly, and also produces g_att.  g_att_m = coef_m G_m - w_m sum_r G_r coef_r + G_m coef_m (1 - s_m) needs
// G_r = sum_c g_fused_c * aligned_r,c of ALL references at a pixel.  Each of the R CTAs of one (image, patch) reduces
// its own G_r over the channels and publishes the 16 values to a global scratch; g_att itself is formed by a few
// blocks of the cl_to_nchw launch that follows (ClmAttGrad), in fixed order r = 0..R-1.  Nothing on this kernel's
// critical path waits for a peer: the first version ran the R CTAs as a thread-block cluster and exchanged G_r
// through DSMEM (its two cluster barriers were 36 % of the stall samples at cfg2), the second let the last CTA to
// arrive do it (fence + arrival atomic: 2-3.4 us of warp 0's time ahead of a block-wide barrier).
#ifdef CLC_DEBUG_ABI
// bring-up: wall-clock (ns) phase stamps of the first 64 CTAs of the thread-owns-items backward kernel
__device__ long long g_bwd_stamps[64][16];
#define BW_STAMP(i) do { if (threadIdx.x == 0 && blockIdx.x < 64) { long long t_; \
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_bwd_stamps[blockIdx.x][(i)] = t_; } } while (0)
#else
#define BW_STAMP(i) do { } while (0)
#endif

constexpr int kClmMaxRefs = 8;
struct ClmBwdArgs {
  const float* g_fused;     // [NQ, C, fh*fw]
  const float* att;         // plane (r, b) of [fh*fw] logits at att + r*att_sr + b*att_sb
  int64_t att_sr, att_sb;
  const float* aligned;     // [NP, C, fh*fw] blended references of the forward pass
  float* g_att;             // same addressing as att
  int R;
  float* g_scratch;         // [NQ*P][R][16] published G_r
};

template <int NT, bool CLM>
__global__ void __launch_bounds__(NT, 2)
match_bwd_own_kernel(PatchAddr qa, const float* __restrict__ rT, const float* __restrict__ mask,
                     const int32_t* __restrict__ idx, const float* __restrict__ weights, float temperature,
                     const float* __restrict__ g_out, float* __restrict__ g_rT, float* __restrict__ g_q,
                     float* __restrict__ g_val_out, int P, int C, int ph, int fh, int fw, int k, int dbg,
                     const ClmBwdArgs ca) {
  constexpr int KK = 4, NV = 2 + 4 * KK, NW = NT / 32, pw = 4;
  __shared__ float red[NW][NV];
  __shared__ float gpart[CLM ? NT : 1][4];     // per-thread partial G (its 4 channels) for the 4 pixels of its row
  __shared__ float tot[NV];                 // xs, sxx, then per window: g_w, s1, s2, xy
  __shared__ float coef[KK][4];             // per window: w_j, g_xy, 2*g_dY, c_mean
  __shared__ float part[KK][3];             // per window: its terms of t_gxs, t_gsxx, t_gxm
  __shared__ int src_s[KK];
  __shared__ float w_s[KK], m_s[KK];
  const int cw = fw - pw + 1, L = (fh - ph + 1) * cw;
  const int HW = fh * fw;
  // CLM: the R problems (references) of one (image, patch) are consecutive blocks = one cluster
  const int n = CLM ? (int)((blockIdx.x / ca.R / P) * ca.R + blockIdx.x % ca.R) : (int)(blockIdx.x / P);
  const int patch = CLM ? (int)((blockIdx.x / ca.R) % P) : (int)(blockIdx.x - n * P);
  const int nq = n / qa.repeat;
  const int npx = fw / pw;
  const int py = patch / npx, px = patch - py * npx;
  const int c4n = C >> 2, K = C * ph * pw;
  const float Kf = (float)K, inv_k = 1.0f / Kf;
  const int64_t po = ((int64_t)n * P + patch) * k;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int dy = tid / c4n, c4 = tid - dy * c4n;
  BW_STAMP(0);
  pdl_trigger();
  pdl_wait();
  BW_STAMP(1);
  if (tid < KK) {
    int src = 0;
    float wj = 0.f, mj = 1.f;
    if (tid < k) {
      const int id = idx[po + tid];
      wj = weights[po + tid];
      const int oy = id / cw, ox = id - oy * cw;
      src = oy * fw + ox;
      if (mask) mj = mask[(int64_t)patch * L + id];
    }
    src_s[tid] = src; w_s[tid] = wj; m_s[tid] = mj;
  }
  // q / g: 4 channel rows of 4 floats each; q[dx][i] = channel 4*c4+i at (dy, dx)
  const float* qp = qa.q + (int64_t)nq * qa.sn + qa.patch_off(patch) + (int64_t)(4 * c4) * qa.sc + (int64_t)dy * qa.sy;
  const float* gp = g_out + ((int64_t)n * C + 4 * c4) * HW + (py * ph + dy) * fw + px * pw;
  float4 qc[4], gc[4];
  if constexpr (!CLM) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      qc[i] = ld4(qp + (int64_t)i * qa.sc);
      gc[i] = ld4(gp + (int64_t)i * HW);
    }
  } else {
    // g_aligned = g_fused * coef_r on the 4 pixels of this thread's patch row; partial G_r over its 4 channels
    const int r_own = n - nq * ca.R;
    const int64_t s0 = (int64_t)(py * ph + dy) * fw + px * pw;
    const float* gf = ca.g_fused + ((int64_t)nq * C + 4 * c4) * HW + s0;
    const float* al = ca.aligned + ((int64_t)n * C + 4 * c4) * HW + s0;
    float4 av[4], a4[kClmMaxRefs];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      qc[i] = ld4(qp + (int64_t)i * qa.sc);
      gc[i] = ld4(gf + (int64_t)i * HW);
      av[i] = ld4(al + (int64_t)i * HW);
    }
#pragma unroll
    for (int r = 0; r < kClmMaxRefs; ++r)
      if (r < ca.R) a4[r] = ld4(ca.att + (int64_t)r * ca.att_sr + (int64_t)nq * ca.att_sb + s0);
    float gp4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      gp4[0] = fmaf(gc[i].x, av[i].x, gp4[0]); gp4[1] = fmaf(gc[i].y, av[i].y, gp4[1]);
      gp4[2] = fmaf(gc[i].z, av[i].z, gp4[2]); gp4[3] = fmaf(gc[i].w, av[i].w, gp4[3]);
    }
#pragma unroll
    for (int dx = 0; dx < 4; ++dx) gpart[tid][dx] = gp4[dx];
    float cf[4];
#pragma unroll
    for (int dx = 0; dx < 4; ++dx) {
      // coef_r = softmax_r(att)[r] * sigmoid(att_r), operation order of clm.cu::clm_coef
      float mx = -INFINITY, den = 0.f, e_own = 0.f, a_own = 0.f;
#pragma unroll
      for (int r = 0; r < kClmMaxRefs; ++r)
        if (r < ca.R) mx = fmaxf(mx, dx == 0 ? a4[r].x : dx == 1 ? a4[r].y : dx == 2 ? a4[r].z : a4[r].w);
#pragma unroll
      for (int r = 0; r < kClmMaxRefs; ++r)
        if (r < ca.R) {
          const float a = dx == 0 ? a4[r].x : dx == 1 ? a4[r].y : dx == 2 ? a4[r].z : a4[r].w;
          const float e = expf(a - mx);
          den += e;
          if (r == r_own) { e_own = e; a_own = a; }
        }
      cf[dx] = (e_own / den) * (1.0f / (1.0f + expf(-a_own)));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { gc[i].x *= cf[0]; gc[i].y *= cf[1]; gc[i].z *= cf[2]; gc[i].w *= cf[3]; }
  }
  float4 q[4], g[4];   // [dx] -> float4 over the 4 channels
  q[0] = make_float4(qc[0].x, qc[1].x, qc[2].x, qc[3].x); q[1] = make_float4(qc[0].y, qc[1].y, qc[2].y, qc[3].y);
  q[2] = make_float4(qc[0].z, qc[1].z, qc[2].z, qc[3].z); q[3] = make_float4(qc[0].w, qc[1].w, qc[2].w, qc[3].w);
  g[0] = make_float4(gc[0].x, gc[1].x, gc[2].x, gc[3].x); g[1] = make_float4(gc[0].y, gc[1].y, gc[2].y, gc[3].y);
  g[2] = make_float4(gc[0].z, gc[1].z, gc[2].z, gc[3].z); g[3] = make_float4(gc[0].w, gc[1].w, gc[2].w, gc[3].w);
  __syncthreads();    // src_s (and gpart) visible
  BW_STAMP(2);
  if constexpr (CLM) {
    // G_r at the 16 pixels (dy, dx): fixed-order sum of the c4n per-thread partials of row dy
    if (tid < 64) {
      const int o = tid >> 2, qq = tid & 3, ody = o >> 2, odx = o & 3;
      const int per = (c4n + 3) / 4;
      float a = 0.f;
      for (int t = qq * per; t < (qq + 1) * per && t < c4n; ++t) a += gpart[ody * c4n + t][odx];
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      const int r_own = n - nq * ca.R;
      const int64_t grp = (int64_t)nq * P + patch;
      float* gs = ca.g_scratch + grp * ca.R * 16;
      if (qq == 0) gs[r_own * 16 + o] = a;                 // published; g_att is formed by the cl_to_nchw launch
    }
  }
  BW_STAMP(3);
  const float* rbase = rT + ((int64_t)n * HW + dy * fw) * C + 4 * c4;    // + (src + dx) * C
  float* gbase = g_rT + ((int64_t)n * HW + dy * fw) * C + 4 * c4;
  float v[NV];
#pragma unroll
  for (int t = 0; t < NV; ++t) v[t] = 0.f;
#pragma unroll
  for (int dx = 0; dx < 4; ++dx) {
    v[0] += (q[dx].x + q[dx].y) + (q[dx].z + q[dx].w);
    v[1] = fmaf(q[dx].x, q[dx].x, fmaf(q[dx].y, q[dx].y, fmaf(q[dx].z, q[dx].z, fmaf(q[dx].w, q[dx].w, v[1]))));
  }
  float4 rv[2][4];
#pragma unroll
  for (int j0 = 0; j0 < KK; j0 += 2) {
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const float* wp = rbase + (int64_t)src_s[j0 + jj] * C;
#pragma unroll
      for (int dx = 0; dx < 4; ++dx)
        rv[jj][dx] = (j0 + jj < k) ? ld4(wp + dx * C) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int j = j0 + jj;
#pragma unroll
      for (int dx = 0; dx < 4; ++dx) {
        const float4 r4 = rv[jj][dx], g4 = g[dx], q4 = q[dx];
        v[2 + 4 * j] = fmaf(g4.x, r4.x, fmaf(g4.y, r4.y, fmaf(g4.z, r4.z, fmaf(g4.w, r4.w, v[2 + 4 * j]))));
        v[3 + 4 * j] += (r4.x + r4.y) + (r4.z + r4.w);
        v[4 + 4 * j] = fmaf(r4.x, r4.x, fmaf(r4.y, r4.y, fmaf(r4.z, r4.z, fmaf(r4.w, r4.w, v[4 + 4 * j]))));
        v[5 + 4 * j] = fmaf(q4.x, r4.x, fmaf(q4.y, r4.y, fmaf(q4.z, r4.z, fmaf(q4.w, r4.w, v[5 + 4 * j]))));
      }
    }
  }
  BW_STAMP(4);
#pragma unroll
  for (int t = 0; t < NV; ++t) {
    v[t] = warp_sum(v[t]);
    if (lane == 0) red[wid][t] = v[t];
  }
  __syncthreads();
  if (tid < NV) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) a += red[w][tid];
    tot[tid] = a;
  }
  __syncthreads();
  BW_STAMP(5);
  if (tid < k) {
    const int j = tid;
    const float xs = tot[0], sxx = tot[1];
    const float xm = xs / Kf;
    const float dX = sxx - xm * xs;
    float dot = 0.f;   // w = softmax(v*T): dL/dv_j = T * w_j * (g_w_j - sum_i w_i g_w_i)
    for (int i = 0; i < k; ++i) dot = fmaf(w_s[i], tot[2 + 4 * i], dot);
    const float wj = w_s[j];
    const float gv = temperature * wj * (tot[2 + 4 * j] - dot);
    if (g_val_out) g_val_out[po + j] = gv;
    const float s1 = tot[3 + 4 * j], s2 = tot[4 + 4 * j], xy = tot[5 + 4 * j];
    const float ym = s1 * inv_k;
    const float dY = s2 - ym * ym * Kf;
    const float D = dY * dX;
    const float num = xy - ym * xs;
    const float rs = rsqrtf(D);
    float gg = gv;
    if (mask) gg *= m_s[j];
    const float g_num = gg * rs;
    const float g_D = -0.5f * gg * num * rs / D;
    const float g_dY = g_D * dX, g_dX = g_D * dY;
    const float g_ym = -g_num * xs - 2.f * g_dY * ym * Kf;
    part[j][0] = -g_num * ym - g_dX * xm;
    part[j][1] = g_dX;
    part[j][2] = -g_dX * xs;
    coef[j][0] = wj; coef[j][1] = g_num; coef[j][2] = 2.f * g_dY; coef[j][3] = g_ym * inv_k;
  }
  __syncthreads();
  BW_STAMP(6);
#pragma unroll
  for (int j0 = 0; j0 < KK; j0 += 2) {
    if (j0 >= k) break;
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const float* wp = rbase + (int64_t)src_s[j0 + jj] * C;
#pragma unroll
      for (int dx = 0; dx < 4; ++dx)
        rv[jj][dx] = (j0 + jj < k) ? ld4(wp + dx * C) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int j = j0 + jj;
      if (j >= k) break;
      const float cw_ = coef[j][0], cxy = coef[j][1], cdy = coef[j][2], cm = coef[j][3];
      float* dst = gbase + (int64_t)src_s[j] * C;
#pragma unroll
      for (int dx = 0; dx < 4; ++dx) {
        const float4 q4 = q[dx], g4 = g[dx], r4 = rv[jj][dx];
        float4 o;
        o.x = fmaf(cw_, g4.x, fmaf(cxy, q4.x, fmaf(cdy, r4.x, cm)));
        o.y = fmaf(cw_, g4.y, fmaf(cxy, q4.y, fmaf(cdy, r4.y, cm)));
        o.z = fmaf(cw_, g4.z, fmaf(cxy, q4.z, fmaf(cdy, r4.z, cm)));
        o.w = fmaf(cw_, g4.w, fmaf(cxy, q4.w, fmaf(cdy, r4.w, cm)));
        if (!(dbg & 1)) red_add4(dst + dx * C, o);
        else if (o.x == 1.2345e33f) dst[dx * C] = o.y + o.z + o.w;
      }
    }
  }
  BW_STAMP(7);
  if (g_q && !(dbg & 2)) {
    float t_gxs = 0.f, t_gsxx = 0.f, t_gxm = 0.f;
    for (int j = 0; j < k; ++j) { t_gxs += part[j][0]; t_gsxx += part[j][1]; t_gxm += part[j][2]; }
    const float c0 = 2.f * t_gsxx, c1 = t_gxs + t_gxm * inv_k;
    float* gqp = g_q + (int64_t)nq * qa.sn + qa.patch_off(patch) + (int64_t)(4 * c4) * qa.sc + (int64_t)dy * qa.sy;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 o;
      o.x = fmaf(c0, qc[i].x, c1); o.y = fmaf(c0, qc[i].y, c1); o.z = fmaf(c0, qc[i].z, c1); o.w = fmaf(c0, qc[i].w, c1);
      red_add4(gqp + (int64_t)i * qa.sc, o);
    }
  }
  BW_STAMP(8);
}

template <int NT>
static int launch_bwd_own(unsigned blocks, cudaStream_t st, PatchAddr qa, const float* rT, const float* mask,
                          const int32_t* idx, const float* weights, float temperature, const float* g_out,
                          float* g_rT, float* g_q, float* g_val, int P, int C, int ph, int fh, int fw, int k) {
  CLC_CUDA(launch_pdl(match_bwd_own_kernel<NT, false>, dim3(blocks), dim3(NT), 0, st, qa, rT, mask, idx, weights, temperature,
                      g_out, g_rT, g_q, g_val, P, C, ph, fh, fw, k, dbg_bits(), ClmBwdArgs{}));
  CLC_CHECK_LAUNCH("clc_match_bwd(main)");
  return CLC_OK;
}
// CLM-fused variant: the R CTAs of one (image, patch) are consecutive blocks; PDL.
template <int NT>
static int launch_bwd_own_clm(unsigned blocks, cudaStream_t st, PatchAddr qa, const float* rT, const float* mask,
                              const int32_t* idx, const float* weights, float temperature, float* g_rT, float* g_q,
                              float* g_val, int P, int C, int ph, int fh, int fw, int k, const ClmBwdArgs& ca) {
  const float* no_g_out = nullptr;
  CLC_CUDA(launch_pdl(match_bwd_own_kernel<NT, true>, dim3(blocks), dim3(NT), 0, st, qa, rT, mask, idx, weights, temperature,
                      no_g_out, g_rT, g_q, g_val, P, C, ph, fh, fw, k, dbg_bits(), ca));
  CLC_CHECK_LAUNCH("clc_match_clm_bwd(main)");
  return CLC_OK;
}

template <int ITEMS, bool REREAD>
static int launch_bwd_reg2(unsigned blocks, size_t smem, cudaStream_t st, PatchAddr qa, const float* rT,
                           const float* mask, const int32_t* idx, const float* weights, float temperature,
                           const float* g_out, float* g_rT, float* g_q, float* g_val, int P, int C, int ph,
                           int pw, int fh, int fw, int k) {
  auto kern = match_bwd_cl_reg_kernel<ITEMS, REREAD>;
  if (smem > 48 * 1024) CLC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  kern<<<blocks, 256, smem, st>>>(qa, rT, mask, idx, weights, temperature, g_out, g_rT, g_q, g_val, P, C, ph, pw,
                                  fh, fw, k, dbg_bits());
  CLC_CHECK_LAUNCH("clc_match_bwd(main)");
  return CLC_OK;
}

template <int ITEMS>
static int launch_bwd_reg(unsigned blocks, size_t smem, cudaStream_t st, PatchAddr qa, const float* rT,
                          const float* mask, const int32_t* idx, const float* weights, float temperature,
                          const float* g_out, float* g_rT, float* g_q, float* g_val, int P, int C, int ph,
                          int pw, int fh, int fw, int k) {
  // CTAs resident per SM: 2 with the windows kept in registers, 3 when they are re-read; re-read
  // when that saves a wave (and shared memory allows 3 CTAs)
  const unsigned w2 = (blocks + 2 * kNumSMs - 1) / (2 * kNumSMs), w3 = (blocks + 3 * kNumSMs - 1) / (3 * kNumSMs);
  const bool reread = (w3 < w2 && 3 * (smem + 1024) <= 220 * 1024) || (dbg_bits() & 4);
  if (reread)
    return launch_bwd_reg2<ITEMS, true>(blocks, smem, st, qa, rT, mask, idx, weights, temperature, g_out, g_rT, g_q,
                                        g_val, P, C, ph, pw, fh, fw, k);
  return launch_bwd_reg2<ITEMS, false>(blocks, smem, st, qa, rT, mask, idx, weights, temperature, g_out, g_rT, g_q,
                                       g_val, P, C, ph, pw, fh, fw, k);
}

// Host helpers shared with match_tc.cu --------------------------------------------------------
int launch_channel_sums(const float* r, float* s1, float* s2, int64_t NP, int C, int64_t HW,
                        cudaStream_t st) {
  channel_sums_kernel<<<grid_for(NP * HW, 256), 256, 0, st>>>(r, s1, s2, NP, C, HW);
  CLC_CHECK_LAUNCH("channel_sums");
  return CLC_OK;
}

int launch_patch_stats(const PatchAddr& qa, float* xs, float* sxx, int64_t NQ, int P, int C, int ph,
                       int pw, cudaStream_t st) {
  if (NQ * P > 0x7fffffffLL) return CLC_ERR_UNSUPPORTED;
  patch_stats_kernel<<<(unsigned)(NQ * P), 256, 0, st>>>(qa, xs, sxx, P, C, ph, pw);
  CLC_CHECK_LAUNCH("patch_stats");
  return CLC_OK;
}

}  // namespace clc

using namespace clc;

static bool view_ok(const clc_patch_view* v) { return v && v->q && v->npx >= 1 && v->q_repeat >= 1; }

static PatchAddr make_addr(const clc_patch_view* v) {
  PatchAddr a;
  a.q = v->q; a.sn = v->q_sn; a.spy = v->q_spy; a.spx = v->q_spx; a.sc = v->q_sc; a.sy = v->q_sy;
  a.npx = v->npx; a.repeat = v->q_repeat;
  return a;
}

extern "C" size_t clc_pearson_corr_workspace_bytes(int64_t NP, int32_t P, int32_t C, int32_t ph,
                                                   int32_t pw, int32_t fh, int32_t fw) {
  (void)C; (void)ph; (void)pw;
  // s1, s2: NP*fh*fw each;  xs, sxx: NP*P each (upper bound: one query per problem)
  return sizeof(float) * (size_t)(2 * NP * (int64_t)fh * fw + 2 * NP * (int64_t)P) + 256;
}

extern "C" int clc_pearson_corr(const clc_patch_view* qv, const float* r, const float* mask, float* corr,
                                int64_t NP, int32_t P, int32_t C, int32_t ph, int32_t pw,
                                int32_t fh, int32_t fw, void* workspace, size_t workspace_bytes,
                                void* stream) {
  if (!view_ok(qv) || !r || !corr || NP < 0 || P < 1 || C < 1 || ph < 1 || pw < 1) return CLC_ERR_INVALID_ARGUMENT;
  if (fh < ph || fw < pw) return CLC_ERR_INVALID_ARGUMENT;
  if (NP == 0) return CLC_OK;
  if (NP > 65535) return CLC_ERR_UNSUPPORTED;
  if (NP % qv->q_repeat) return CLC_ERR_INVALID_ARGUMENT;
  if (!workspace || workspace_bytes < clc_pearson_corr_workspace_bytes(NP, P, C, ph, pw, fh, fw))
    return CLC_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t HW = (int64_t)fh * fw;
  const int64_t NQ = NP / qv->q_repeat;
  float* s1 = (float*)workspace;
  float* s2 = s1 + NP * HW;
  float* xs = s2 + NP * HW;
  float* sxx = xs + NQ * P;
  const PatchAddr qa = make_addr(qv);
  int rc;
  if ((rc = launch_channel_sums(r, s1, s2, NP, C, HW, st))) return rc;
  if ((rc = launch_patch_stats(qa, xs, sxx, NQ, P, C, ph, pw, st))) return rc;
  const int L = (fh - ph + 1) * (fw - pw + 1);
  dim3 grid((L + TN - 1) / TN, (P + TM - 1) / TM, (unsigned)NP);
  pearson_corr_kernel<<<grid, 256, 0, st>>>(qa, r, s1, s2, xs, sxx, mask, corr, P, C, ph, pw, fh, fw);
  CLC_CHECK_LAUNCH("clc_pearson_corr");
  return CLC_OK;
}

extern "C" int clc_topk_rows(const float* x, int64_t R, int64_t L, int32_t k, float* val, int32_t* idx,
                             void* stream) {
  if (!x || !val || !idx || R < 0 || L < 1 || k < 1) return CLC_ERR_INVALID_ARGUMENT;
  if (k > L || L > 2147483647LL) return CLC_ERR_INVALID_ARGUMENT;
  if (k > kMaxK) return CLC_ERR_UNSUPPORTED;
  if (R == 0) return CLC_OK;
  topk_rows_kernel<<<(unsigned)((R * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, R, L, k, val, idx);
  CLC_CHECK_LAUNCH("clc_topk_rows");
  return CLC_OK;
}

extern "C" int clc_gaussian_mask(float* mask, int32_t img_h, int32_t img_w, int32_t ph, int32_t pw,
                                 void* stream) {
  if (!mask || ph < 1 || pw < 1 || img_h < ph || img_w < pw) return CLC_ERR_INVALID_ARGUMENT;
  if (img_w % pw) return CLC_ERR_UNSUPPORTED;  // patch_img_w must be integral (as in the reference use)
  const int64_t total = (int64_t)((img_h * img_w) / (ph * pw)) * (img_h - ph + 1) * (img_w - pw + 1);
  gaussian_mask_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(mask, img_h, img_w, ph, pw);
  CLC_CHECK_LAUNCH("clc_gaussian_mask");
  return CLC_OK;
}

static bool gather_args_ok(int64_t NP, int C, int fh, int fw, int gh, int gw, int corr_w, int k) {
  return NP >= 0 && C >= 1 && gh >= 1 && gw >= 1 && fh >= gh && fw >= gw && fh % gh == 0 &&
         fw % gw == 0 && corr_w >= 1 && k >= 1 && k <= kMaxK;
}

extern "C" int clc_gather_blend_fwd(const float* feat, const int32_t* idx, const float* val, float temperature,
                                    float* out, float* weights, int64_t NP, int32_t C, int32_t fh, int32_t fw,
                                    int32_t gh, int32_t gw, int32_t corr_w, int32_t k, int32_t is_stack,
                                    void* stream) {
  if (!feat || !idx || !out || (!is_stack && !val)) return CLC_ERR_INVALID_ARGUMENT;
  if (!gather_args_ok(NP, C, fh, fw, gh, gw, corr_w, k)) return CLC_ERR_INVALID_ARGUMENT;
  if (NP == 0) return CLC_OK;
  if (NP > 65535 || (int64_t)C * fh * fw > 0x7fffffffLL) return CLC_ERR_UNSUPPORTED;
  const int npy = fh / gh, npx = fw / gw;
  // channel chunk per CTA: aim for >= 4 CTAs per SM, at least 8 channels per CTA
  int CH = 64;
  while (CH > 8 && (int64_t)NP * npy * ((C + CH - 1) / CH) < 4 * kNumSMs) CH >>= 1;
  const int cch = (C + CH - 1) / CH;
  dim3 grid((unsigned)(npy * cch), (unsigned)NP);
  const size_t sm = (size_t)npx * k * 8;
  if (sm > 48 * 1024) return CLC_ERR_UNSUPPORTED;
  if (is_stack)
    gather_blend_fwd_kernel<true><<<grid, 256, sm, (cudaStream_t)stream>>>(
        feat, idx, val, temperature, out, weights, C, fh, fw, gh, gw, corr_w, k, CH, cch);
  else
    gather_blend_fwd_kernel<false><<<grid, 256, sm, (cudaStream_t)stream>>>(
        feat, idx, val, temperature, out, weights, C, fh, fw, gh, gw, corr_w, k, CH, cch);
  CLC_CHECK_LAUNCH("clc_gather_blend_fwd");
  return CLC_OK;
}

extern "C" int clc_gather_blend_bwd(const float* feat, const int32_t* idx, const float* weights,
                                    float temperature, const float* g_out, float* g_feat, float* g_val,
                                    int64_t NP, int32_t C, int32_t fh, int32_t fw, int32_t gh, int32_t gw,
                                    int32_t corr_w, int32_t k, int32_t is_stack, void* stream) {
  if (!feat || !idx || !g_out || !g_feat || !g_val || (!is_stack && !weights)) return CLC_ERR_INVALID_ARGUMENT;
  if (!gather_args_ok(NP, C, fh, fw, gh, gw, corr_w, k)) return CLC_ERR_INVALID_ARGUMENT;
  if (NP == 0) return CLC_OK;
  const int64_t blocks = NP * (fh / gh) * (fw / gw);
  if (blocks > 2147483647LL) return CLC_ERR_UNSUPPORTED;
  if ((int64_t)C * fh * fw > 0x7fffffffLL) return CLC_ERR_UNSUPPORTED;
  gather_blend_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      feat, idx, weights, temperature, g_out, g_feat, g_val, C, fh, fw, gh, gw, corr_w, k, is_stack);
  CLC_CHECK_LAUNCH("clc_gather_blend_bwd");
  return CLC_OK;
}

extern "C" int clc_pearson_topk_bwd(const clc_patch_view* qv, const float* r, const float* mask,
                                    const int32_t* idx, const float* g_val, float* g_r, float* g_q,
                                    int64_t NP, int32_t P, int32_t C, int32_t ph, int32_t pw,
                                    int32_t fh, int32_t fw, int32_t k, void* stream) {
  if (!view_ok(qv) || !r || !idx || !g_val || !g_r) return CLC_ERR_INVALID_ARGUMENT;
  if (NP < 0 || P < 1 || C < 1 || ph < 1 || pw < 1 || fh < ph || fw < pw || k < 1) return CLC_ERR_INVALID_ARGUMENT;
  if (NP == 0) return CLC_OK;
  if (NP * P > 2147483647LL) return CLC_ERR_UNSUPPORTED;
  if ((int64_t)C * fh * fw > 0x7fffffffLL) return CLC_ERR_UNSUPPORTED;
  pearson_topk_bwd_kernel<<<(unsigned)(NP * P), 256, 0, (cudaStream_t)stream>>>(
      make_addr(qv), r, mask, idx, g_val, g_r, g_q, P, C, ph, pw, fh, fw, k);
  CLC_CHECK_LAUNCH("clc_pearson_topk_bwd");
  return CLC_OK;
}

// (a 256-byte-aligned pad of NP*fh*fw ints sits behind the gradient scratch: reserved, cleared with it)
static inline size_t bwd_ws_counter_bytes(int64_t NP, int fh, int fw) {
  return ((size_t)NP * fh * fw * sizeof(int) + 255) / 256 * 256;
}
static inline size_t bwd_ws_zero_bytes(int64_t NP, int C, int fh, int fw) {
  return sizeof(float) * (size_t)NP * C * fh * fw + bwd_ws_counter_bytes(NP, fh, fw);
}

#ifdef CLC_DEBUG_ABI
extern "C" CLC_API int clc_debug_bwd_stamps(long long* host_out /* [64][16] */) {
  if (!host_out) return CLC_ERR_INVALID_ARGUMENT;
  CLC_CUDA(cudaMemcpyFromSymbol(host_out, clc::g_bwd_stamps, sizeof(long long) * 64 * 16));
  return CLC_OK;
}
#endif

extern "C" size_t clc_match_bwd_workspace_bytes(int64_t NP, int32_t C, int32_t fh, int32_t fw) {
  if (NP < 0 || C < 1 || fh < 1 || fw < 1) return 0;
  // gradient scratch (channels-last g_r) | reserved pad | channels-last copy of r (or, when the caller supplies
  // that copy, the CLM-fused kernel's published G_r)
  return bwd_ws_counter_bytes(NP, fh, fw) + 2 * sizeof(float) * (size_t)NP * C * fh * fw + 512;
}

extern "C" int clc_match_bwd_zero_workspace(void* workspace, size_t workspace_bytes, int64_t NP, int32_t C,
                                           int32_t fh, int32_t fw, void* stream) {
  if (!workspace || NP < 0 || C < 1 || fh < 1 || fw < 1) return CLC_ERR_INVALID_ARGUMENT;
  if (workspace_bytes < clc_match_bwd_workspace_bytes(NP, C, fh, fw)) return CLC_ERR_WORKSPACE;
  float* g_rT = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  CLC_CUDA(cudaMemsetAsync(g_rT, 0, bwd_ws_zero_bytes(NP, C, fh, fw), (cudaStream_t)stream));
  return CLC_OK;
}

extern "C" int clc_match_bwd(const clc_patch_view* qv, const float* r, const float* r_cl, const float* mask,
                             const int32_t* idx, const float* weights, float temperature, const float* g_out,
                             float* g_r, float* g_q, float* g_val, int64_t NP, int32_t P, int32_t C, int32_t ph,
                             int32_t pw, int32_t fh, int32_t fw, int32_t k, int32_t flags, void* workspace,
                             size_t workspace_bytes, void* stream) {
  if (!view_ok(qv) || !r || !idx || !weights || !g_out || !g_r) return CLC_ERR_INVALID_ARGUMENT;
  if (NP < 0 || P < 1 || C < 1 || ph < 1 || pw < 1 || fh < ph || fw < pw || k < 1) return CLC_ERR_INVALID_ARGUMENT;
  if (fh % ph || fw % pw || P != (fh / ph) * (fw / pw)) return CLC_ERR_INVALID_ARGUMENT;
  if (k > kMaxK) return CLC_ERR_UNSUPPORTED;
  if (NP == 0) return CLC_OK;
  if (NP * P > 2147483647LL || (int64_t)C * fh * fw > 0x7fffffffLL || NP > 65535) return CLC_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const PatchAddr qa = make_addr(qv);
  const bool overwrite = (flags & CLC_MATCH_BWD_OVERWRITE_G_R) != 0;
  const size_t smem = (size_t)2 * ph * pw * C * sizeof(float);
  const int HW = fh * fw;
  const size_t plane = sizeof(float) * (size_t)NP * C * HW;
  // fast path: channels-last operands, float4 everywhere (needs a workspace and 16-byte aligned patch rows)
  const bool cl_ok = workspace && pw == 4 && C % 4 == 0 && fw % 4 == 0 && smem <= 200 * 1024 && aligned16(qv->q) &&
                     aligned16(g_out) && (!g_q || aligned16(g_q)) && qa.sy % 4 == 0 && qa.sc % 4 == 0 &&
                     qa.spx % 4 == 0 && qa.spy % 4 == 0 && qa.sn % 4 == 0 && (!r_cl || aligned16(r_cl));
  if (!cl_ok) {
    if (overwrite) CLC_CUDA(cudaMemsetAsync(g_r, 0, plane, st));
    match_bwd_kernel<<<(unsigned)(NP * P), 256, 0, st>>>(qa, r, mask, idx, weights, temperature, g_out, g_r, g_q,
                                                        g_val, P, C, ph, pw, fh, fw, k);
    CLC_CHECK_LAUNCH("clc_match_bwd");
    return CLC_OK;
  }
  if (workspace_bytes < clc_match_bwd_workspace_bytes(NP, C, fh, fw)) return CLC_ERR_WORKSPACE;
  float* g_rT = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  float* rT_own = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(g_rT) + bwd_ws_zero_bytes(NP, C, fh, fw));
  dim3 tgrid((HW + 31) / 32, (C + 31) / 32, (unsigned)NP), tblock(32, 8);
  if (!(flags & CLC_MATCH_BWD_WS_ZEROED)) CLC_CUDA(cudaMemsetAsync(g_rT, 0, plane, st));
  const float* rT = r_cl;
  if (!rT) {  // no channels-last copy supplied (e.g. from clc_match_topk_tc_ref_cl): make one
    CLC_CUDA(launch_pdl(nchw_to_cl_kernel, tgrid, tblock, 0, st, r, rT_own, C, HW));
    CLC_CHECK_LAUNCH("clc_match_bwd(nchw_to_cl)");
    rT = rT_own;
  }
  const int items = ph * pw * (C / 4), per = (items + 255) / 256;
  const unsigned blocks = (unsigned)(NP * P);
  int rc = CLC_OK;
  const int nt_own = ph * (C / 4);
  const bool own_ok = k <= 4 && !(dbg_bits() & 8) &&
                      (nt_own == 128 || nt_own == 192 || nt_own == 256 || nt_own == 320 || nt_own == 384);
  if (!stage_on(0)) {
  } else if (own_ok) {
#define CLC_OWN_CASE(N) case N: rc = launch_bwd_own<N>(blocks, st, qa, rT, mask, idx, weights, temperature, g_out, \
                                                        g_rT, g_q, g_val, P, C, ph, fh, fw, k); break;
    switch (nt_own) { CLC_OWN_CASE(128) CLC_OWN_CASE(192) CLC_OWN_CASE(256) CLC_OWN_CASE(320) CLC_OWN_CASE(384) }
#undef CLC_OWN_CASE
    if (rc) return rc;
  } else if (k <= 4 && per <= 8) {
#define CLC_BWD_CASE(I) case I: rc = launch_bwd_reg<I>(blocks, smem, st, qa, rT, mask, idx, weights, temperature, \
                                                        g_out, g_rT, g_q, g_val, P, C, ph, pw, fh, fw, k); break;
    switch (per) {
      CLC_BWD_CASE(1) CLC_BWD_CASE(2) CLC_BWD_CASE(3) CLC_BWD_CASE(4)
      CLC_BWD_CASE(5) CLC_BWD_CASE(6) CLC_BWD_CASE(7) CLC_BWD_CASE(8)
    }
#undef CLC_BWD_CASE
    if (rc) return rc;
  } else {
    if (smem > 48 * 1024)
      CLC_CUDA(cudaFuncSetAttribute(match_bwd_cl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    match_bwd_cl_kernel<<<blocks, 256, smem, st>>>(qa, rT, mask, idx, weights, temperature, g_out, g_rT, g_q, g_val,
                                                   P, C, ph, pw, fh, fw, k);
    CLC_CHECK_LAUNCH("clc_match_bwd(main)");
  }
  if (!stage_on(1)) return CLC_OK;
  if (overwrite) CLC_CUDA(launch_pdl(cl_to_nchw_kernel<false>, tgrid, tblock, 0, st, g_rT, g_r, C, HW, ClmAttGrad{}));
  else CLC_CUDA(launch_pdl(cl_to_nchw_kernel<true>, tgrid, tblock, 0, st, g_rT, g_r, C, HW, ClmAttGrad{}));
  CLC_CHECK_LAUNCH("clc_match_bwd(cl_to_nchw)");
  return CLC_OK;
}

extern "C" int clc_match_clm_bwd(const clc_patch_view* qv, const float* r_cl, const float* mask, const int32_t* idx,
                                 const float* weights, float temperature, const float* g_fused, const float* att,
                                 int64_t att_sr, int64_t att_sb, const float* aligned, float* g_r, float* g_q,
                                 float* g_val, float* g_att, int64_t NP, int32_t R, int32_t P, int32_t C, int32_t ph,
                                 int32_t pw, int32_t fh, int32_t fw, int32_t k, int32_t flags, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  if (!view_ok(qv) || !r_cl || !idx || !weights || !g_fused || !att || !aligned || !g_r || !g_att || !workspace)
    return CLC_ERR_INVALID_ARGUMENT;
  if (NP < 0 || R < 1 || P < 1 || C < 1 || ph < 1 || pw < 1 || fh < ph || fw < pw || k < 1) return CLC_ERR_INVALID_ARGUMENT;
  if (fh % ph || fw % pw || P != (fh / ph) * (fw / pw) || NP % R || qv->q_repeat != R) return CLC_ERR_INVALID_ARGUMENT;
  if (NP == 0) return CLC_OK;
  if (NP * P > 2147483647LL || (int64_t)C * fh * fw > 0x7fffffffLL || NP > 65535) return CLC_ERR_UNSUPPORTED;
  // the fused kernel is the thread-owns-items variant: 4x4 patches, k <= 4, one thread per (patch row, 4 channels)
  const int nt_own = ph * (C / 4);
  const PatchAddr qa = make_addr(qv);
  const bool ok = R <= kClmMaxRefs && k <= 4 && pw == 4 && C % 4 == 0 && fw % 4 == 0 && ph == 4 &&
                  (nt_own == 128 || nt_own == 192 || nt_own == 256 || nt_own == 320 || nt_own == 384) &&
                  aligned16(qv->q) && aligned16(g_fused) && aligned16(att) && aligned16(aligned) && aligned16(r_cl) &&
                  (!g_q || aligned16(g_q)) && qa.sy % 4 == 0 && qa.sc % 4 == 0 && qa.spx % 4 == 0 && qa.spy % 4 == 0 &&
                  qa.sn % 4 == 0 && !((att_sr | att_sb) & 3);
  if (!ok) return CLC_ERR_UNSUPPORTED;
  if (workspace_bytes < clc_match_bwd_workspace_bytes(NP, C, fh, fw)) return CLC_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = fh * fw;
  float* g_rT = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  if (!(flags & CLC_MATCH_BWD_WS_ZEROED)) CLC_CUDA(cudaMemsetAsync(g_rT, 0, bwd_ws_zero_bytes(NP, C, fh, fw), st));
  ClmBwdArgs ca;
  ca.g_fused = g_fused; ca.att = att; ca.att_sr = att_sr; ca.att_sb = att_sb; ca.aligned = aligned; ca.g_att = g_att;
  ca.R = R;
  ca.g_scratch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(g_rT) + bwd_ws_zero_bytes(NP, C, fh, fw));
  const unsigned blocks = (unsigned)(NP * P);
  int rc = CLC_OK;
#define CLC_OWN_CASE(N) case N: rc = launch_bwd_own_clm<N>(blocks, st, qa, r_cl, mask, idx, weights, temperature, g_rT, \
                                                            g_q, g_val, P, C, ph, fh, fw, k, ca); break;
  if (stage_on(0)) switch (nt_own) { CLC_OWN_CASE(128) CLC_OWN_CASE(192) CLC_OWN_CASE(256) CLC_OWN_CASE(320) CLC_OWN_CASE(384) }
#undef CLC_OWN_CASE
  if (rc) return rc;
  if (!stage_on(1)) return CLC_OK;
  dim3 tgrid((HW + 31) / 32, (C + 31) / 32, (unsigned)NP), tblock(32, 8);
  ClmAttGrad ga;
  ga.g_scratch = ca.g_scratch; ga.att = att; ga.g_att = g_att; ga.att_sr = att_sr; ga.att_sb = att_sb;
  ga.R = R; ga.P = P; ga.fw = fw;
  if (flags & CLC_MATCH_BWD_OVERWRITE_G_R) CLC_CUDA(launch_pdl(cl_to_nchw_kernel<false>, tgrid, tblock, 0, st, g_rT, g_r, C, HW, ga));
  else CLC_CUDA(launch_pdl(cl_to_nchw_kernel<true>, tgrid, tblock, 0, st, g_rT, g_r, C, HW, ga));
  CLC_CHECK_LAUNCH("clc_match_clm_bwd(cl_to_nchw)");
  return CLC_OK;
}
