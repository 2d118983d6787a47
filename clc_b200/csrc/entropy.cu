// Entropy stage of the CLC latent path: GaussianConditional (+STE round, +bpp partial),
// LRP add, symbols/indexes, EntropyBottleneck (factorised prior) and the log2-sum of
// RateDistortionLoss -- forward and backward.  All kernels are HBM-bandwidth bound:
// float4 coalesced streaming accesses, grid-stride over a grid sized in multiples of the
// 148 SMs, warp-shuffle + one double atomic per CTA for the bpp partial sums.
//
// Reference arithmetic followed (file:line into the reference tree):
//   likelihood ............ models/CLC_run.py:718-736 (in-tree restatement of GaussianConditional)
//   ste_round ............. models/CLC_run.py:35-36, :571, :528-530
//   LRP add ............... models/CLC_run.py:582-583
//   bpp ................... train_CLC.py:48-51
//   EntropyBottleneck ..... compressai (un-vendored) as called at models/CLC_run.py:526
#include "common.cuh"

namespace clc {

constexpr float kInvSqrt2 = 0.70710678118654752440f;     // -const, const = -(2 ** -0.5) of CLC_run.py:729
constexpr float kInvSqrt2Pi = 0.39894228040143267794f;

// ------------------------------------------------------------------------------------------
// In-kernel quantisation noise (training-mode `inputs + U(-1/2, 1/2)`, compressai EntropyModel.quantize
// "noise" as reached from CLC_run.py:526 / :569): Philox4x32-10, counter = (element index / 4, stream offset),
// key = seed.  The state {seed, base offset} lives in DEVICE memory, so a captured CUDA graph draws fresh noise
// on every replay once its owner advances the base offset (clc_rng_advance); `offset` separates the calls of
// one step.  Forward and backward regenerate the same sample from (state, offset, element index).
// ------------------------------------------------------------------------------------------
struct RngArg {
  const unsigned long long* state;   // device: {seed, base offset}; NULL = no in-kernel noise
  unsigned long long offset;
};

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}
// 24-bit mantissa sample in the OPEN interval (-1/2, 1/2), symmetric around 0
__device__ __forceinline__ float u_centered(uint32_t r) { return ((float)(r >> 8) + 0.5f) * 5.9604644775390625e-8f - 0.5f; }

struct RngCtx {
  uint2 key;
  uint32_t off_lo, off_hi;
  __device__ __forceinline__ explicit RngCtx(const RngArg& a) {
    const unsigned long long seed = a.state ? a.state[0] : 0ull, off = (a.state ? a.state[1] : 0ull) + a.offset;
    key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    off_lo = (uint32_t)off;
    off_hi = (uint32_t)(off >> 32);
  }
  // the four samples of elements 4*i4 .. 4*i4+3
  __device__ __forceinline__ float4 noise4(unsigned long long i4) const {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)i4, (uint32_t)(i4 >> 32), off_lo, off_hi), key);
    return make_float4(u_centered(r.x), u_centered(r.y), u_centered(r.z), u_centered(r.w));
  }
  __device__ __forceinline__ float noise1(unsigned long long e) const {
    const float4 n = noise4(e >> 2);
    const int c = (int)(e & 3);
    return c == 0 ? n.x : c == 1 ? n.y : c == 2 ? n.z : n.w;
  }
};

__global__ void rng_advance_kernel(unsigned long long* state, unsigned long long n) { state[1] += n; }

// bpp = -(sum_i log2_sums[i]) / num_pixels  (train_CLC.py:48-51 with the log2 sums the kernels accumulated);
// optionally advances the noise stream for the next step.  One thread.
__global__ void bpp_finalize_kernel(const double* log2_sums, int n, double num_pixels, double* bpp,
                                    unsigned long long* rng_state, unsigned long long rng_advance) {
  double a = 0.0;
  for (int i = 0; i < n; ++i) a += log2_sums[i];
  *bpp = -a / num_pixels;
  if (rng_state) rng_state[1] += rng_advance;
}

// ------------------------------------------------------------------------------------------
// GaussianConditional
// ------------------------------------------------------------------------------------------
// Phi(u) - Phi(l) with u - l = 1/s loses ~1.25*s ulps to cancellation when formed as a
// difference of two erfc values (the reference's own fp32 result is off by up to 3.3e-5
// relative at s = 256, SURVEY.md 7.3-4 / tests/golden/gaussian.npz).  For s >= 4 the mass of the
// unit-width bin is evaluated instead by its midpoint Taylor series, which has no cancellation:
//   int_{m-d/2}^{m+d/2} phi = d*phi(m)*[1 + d^2 He2(m)/24 + d^4 He4(m)/1920 + d^6 He6(m)/322560],
//   d = 1/s, m = v/s  (even Hermite polynomials), accurate to 3e-6 relative against fp64 for
// every s >= 4 wherever the result is above the 1e-9 floor (and below the floor where it is not).
constexpr float kSeriesMinScale = 4.0f;

__device__ __forceinline__ float gauss_mass_series(float v, float d) {   // d = 1 / scale
  const float t = v * d;
  const float a = t * t, d2 = d * d;
  const float phi = kInvSqrt2Pi * __expf(-0.5f * a);
  const float h2 = a - 1.f;
  const float h4 = (a - 6.f) * a + 3.f;
  const float h6 = ((a - 15.f) * a + 45.f) * a - 15.f;
  const float poly = 1.f + d2 * (h2 * (1.f / 24.f) + d2 * (h4 * (1.f / 1920.f) + d2 * (h6 * (1.f / 322560.f))));
  return d * phi * poly;
}

// erfc(z) for z >= 0 with fractional error < 1.1e-7 in exact arithmetic everywhere (so the far tail keeps its
// relative accuracy, which a fit of erfc itself would not): erfc(z) = t * exp(-z^2 + P9(t)), t = 1 / (1 + z/2),
// the classic Chebyshev fit of log(erfc(z) e^{z^2} / t) (Numerical Recipes "erfcc").  One MUFU.RCP, one MUFU.EX2,
// eleven FMAs -- a third of the instructions of libdevice's erfcf, which made the forward kernel issue-bound
// (profiles/r2_full_cfg4: 81 % issue slots busy).  In fp32 with ex2.approx the relative error is < 4e-6 down to
// erfc = 1e-19, the fp64 leg of tests/test_entropy_gpu.py::test_gc_golden_reference_vectors.
__device__ __forceinline__ float erfc_pos(float z) {
  const float t = __fdividef(1.0f, fmaf(0.5f, z, 1.0f));
  float p = 0.17087277f;
  p = fmaf(p, t, -0.82215223f);
  p = fmaf(p, t, 1.48851587f);
  p = fmaf(p, t, -1.13520398f);
  p = fmaf(p, t, 0.27886807f);
  p = fmaf(p, t, -0.18628806f);
  p = fmaf(p, t, 0.09678418f);
  p = fmaf(p, t, 0.37409196f);
  p = fmaf(p, t, 1.00002368f);
  p = fmaf(p, t, -1.26551223f);
  return t * __expf(fmaf(-z, z, p));
}

struct GcOut {
  float lik_raw, lik, y_hat, outputs;
};

// One element of GaussianConditional.forward + ste_round.  Argument formation order follows
// the reference exactly: values = outputs - means; (half - |values|) / scales as a true
// division; const * inputs; half * erfc(.); upper - lower  (CLC_run.py:720-736).
__device__ __forceinline__ GcOut gc_elem(float y, float s, float m, float n, bool train,
                                         float scale_bound, float lik_bound) {
  GcOut o;
  const float q = rintf(y - m);  // torch.round: half-to-even
  o.y_hat = q + m;
  o.outputs = train ? (y + n) : o.y_hat;
  // values = outputs - means, formed literally in fp32 in both modes (in eval mode
  // (round(y-m)+m)-m is not always exactly round(y-m)).
  const float values = o.outputs - m;
  const float sc = fmaxf(s, scale_bound);
  const float v = fabsf(values);
  // one reciprocal (MUFU.RCP, 1 ulp) shared by both bin edges and by the series: (half - v) * (1/s) differs
  // from the reference's true division (half - v) / s by ~2 ulp, i.e. < 5e-6 relative in the likelihood even at
  // the 1e-9 floor -- more than an order below the 1e-4 bar -- on a kernel that is instruction-issue-bound
  const float inv = __fdividef(1.0f, sc);
  if (sc >= kSeriesMinScale) {
    o.lik_raw = gauss_mass_series(v, inv);
  } else {
    // Phi(up) - Phi(lo) with both erfc arguments taken POSITIVE (lo < 0 always; up < 0 unless |values| < 1/2),
    // which is also what the reference's half * erfc(-x / sqrt 2) evaluates for the tail side
    const float up = (0.5f - v) * inv;
    const float lo = (-0.5f - v) * inv;
    const float eu = 0.5f * erfc_pos(fabsf(up) * kInvSqrt2);
    const float L = 0.5f * erfc_pos(-lo * kInvSqrt2);
    const float U = up < 0.f ? eu : 1.0f - eu;
    o.lik_raw = U - L;
  }
  o.lik = fmaxf(o.lik_raw, lik_bound);
  return o;
}

struct GcFwdParams {
  const float *y, *scale, *mean, *noise;
  float *lik, *y_hat, *outputs;
  double* log2_sum;
  int64_t y_bs, scale_bs, mean_bs, noise_bs, lik_bs, y_hat_bs, outputs_bs;
  int64_t B, CS;
  float scale_bound, lik_bound;
  RngArg rng;
};

template <bool VEC>
__global__ void __launch_bounds__(256) gc_fwd_kernel(const GcFwdParams p) {
  __shared__ float red[32];
  const bool use_rng = p.rng.state != nullptr;
  const bool train = p.noise != nullptr || use_rng;
  const RngCtx rc(p.rng);
  constexpr int W = VEC ? 4 : 1;
  const int64_t per_b = p.CS / W;
  float acc = 0.f;
  // grid = (blocks per batch row, batch rows): no 64-bit division per item
  for (int64_t b = blockIdx.y; b < p.B; b += gridDim.y)
  for (int64_t jj = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; jj < per_b; jj += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = jj * W;
    if constexpr (VEC) {
      const float4 y4 = ld4_stream(p.y + b * p.y_bs + j);
      const float4 s4 = ld4_stream(p.scale + b * p.scale_bs + j);
      const float4 m4 = p.mean ? ld4_stream(p.mean + b * p.mean_bs + j) : make_float4(0, 0, 0, 0);
      const float4 n4 = use_rng ? rc.noise4((unsigned long long)(b * p.CS + j) >> 2)
                                : (train ? ld4_stream(p.noise + b * p.noise_bs + j) : make_float4(0, 0, 0, 0));
      const GcOut o0 = gc_elem(y4.x, s4.x, m4.x, n4.x, train, p.scale_bound, p.lik_bound);
      const GcOut o1 = gc_elem(y4.y, s4.y, m4.y, n4.y, train, p.scale_bound, p.lik_bound);
      const GcOut o2 = gc_elem(y4.z, s4.z, m4.z, n4.z, train, p.scale_bound, p.lik_bound);
      const GcOut o3 = gc_elem(y4.w, s4.w, m4.w, n4.w, train, p.scale_bound, p.lik_bound);
      st4_stream(p.lik + b * p.lik_bs + j, make_float4(o0.lik, o1.lik, o2.lik, o3.lik));
      if (p.y_hat)
        st4(p.y_hat + b * p.y_hat_bs + j, make_float4(o0.y_hat, o1.y_hat, o2.y_hat, o3.y_hat));
      if (p.outputs)
        st4_stream(p.outputs + b * p.outputs_bs + j,
                   make_float4(o0.outputs, o1.outputs, o2.outputs, o3.outputs));
      // (bpp partial: MUFU log2, 2 ulp -- the sum is good to ~1e-7 relative, the bar on bpp is 1e-3 absolute)
      if (p.log2_sum) acc += (__log2f(o0.lik) + __log2f(o1.lik)) + (__log2f(o2.lik) + __log2f(o3.lik));
    } else {
      const float y = p.y[b * p.y_bs + j];
      const float s = p.scale[b * p.scale_bs + j];
      const float m = p.mean ? p.mean[b * p.mean_bs + j] : 0.f;
      const float n = use_rng ? rc.noise1((unsigned long long)(b * p.CS + j)) : (train ? p.noise[b * p.noise_bs + j] : 0.f);
      const GcOut o = gc_elem(y, s, m, n, train, p.scale_bound, p.lik_bound);
      p.lik[b * p.lik_bs + j] = o.lik;
      if (p.y_hat) p.y_hat[b * p.y_hat_bs + j] = o.y_hat;
      if (p.outputs) p.outputs[b * p.outputs_bs + j] = o.outputs;
      if (p.log2_sum) acc += log2f(o.lik);
    }
  }
  if (p.log2_sum) {
    const float tot = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(p.log2_sum, (double)tot);
  }
}

struct GcBwdParams {
  const float *y, *scale, *mean, *noise, *lik, *g_lik, *g_y_hat;
  float *g_y, *g_scale, *g_mean;
  int64_t y_bs, scale_bs, mean_bs, noise_bs, lik_bs, g_lik_bs, g_y_hat_bs, g_y_bs, g_scale_bs, g_mean_bs;
  int64_t B, CS;
  float bpp_coef, scale_bound, lik_bound;
  RngArg rng;
};

struct GcGrad {
  float g_y, g_scale, g_mean;
};

// Analytic backward of gc_elem.  lik = Phi(up) - Phi(lo); dPhi(t)/dt = phi(t).
//   dlik/dv = (phi(lo) - phi(up)) / sc;  dlik/dsc = (lo*phi(lo) - up*phi(up)) / sc
// LowerBound gates (compressai.ops.LowerBound): pass iff x >= bound or grad < 0, for both the
// likelihood floor and the scale floor.  Eval mode: round() has zero gradient, so only the
// scale receives gradient through the likelihood.
__device__ __forceinline__ GcGrad gc_elem_bwd(float y, float s, float m, float n, bool train,
                                              float lik, float g_lik, bool has_g_lik, float bpp_coef,
                                              float g_y_hat, float scale_bound, float lik_bound) {
  GcGrad g;
  float values;
  if (train) {
    values = (y + n) - m;
  } else {
    const float q = rintf(y - m);
    values = (q + m) - m;
  }
  const float sc = fmaxf(s, scale_bound);
  const float v = fabsf(values);
  const float inv = __fdividef(1.0f, sc);          // same shared reciprocal as the forward (gc_elem)
  const float up = (0.5f - v) * inv;
  const float lo = (-0.5f - v) * inv;
  const float pu = kInvSqrt2Pi * __expf(-0.5f * up * up);
  const float pl = kInvSqrt2Pi * __expf(-0.5f * lo * lo);
  float gl = has_g_lik ? g_lik : (bpp_coef / lik);
  // likelihood LowerBound gate: lik > bound means the raw value was above the floor.
  const bool pass_l = (lik > lik_bound) || (gl < 0.f);
  gl = pass_l ? gl : 0.f;
  const float dv = (pl - pu) * inv;
  const float ds = (lo * pl - up * pu) * inv;
  const float sgn = (values > 0.f) ? 1.f : ((values < 0.f) ? -1.f : 0.f);
  const float gv = gl * dv * sgn;
  g.g_y = g_y_hat + (train ? gv : 0.f);
  g.g_mean = train ? -gv : 0.f;
  const float gs = gl * ds;
  g.g_scale = ((s >= scale_bound) || (gs < 0.f)) ? gs : 0.f;
  return g;
}

template <bool VEC>
__global__ void __launch_bounds__(256) gc_bwd_kernel(const GcBwdParams p) {
  const bool use_rng = p.rng.state != nullptr;
  const bool train = p.noise != nullptr || use_rng;
  const RngCtx rc(p.rng);
  const bool has_gl = p.g_lik != nullptr;
  constexpr int W = VEC ? 4 : 1;
  const int64_t per_b = p.CS / W;
  // grid = (blocks per batch row, batch rows): no 64-bit division per item
  for (int64_t b = blockIdx.y; b < p.B; b += gridDim.y)
  for (int64_t jj = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; jj < per_b; jj += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = jj * W;
    if constexpr (VEC) {
      const float4 z4 = make_float4(0, 0, 0, 0);
      const float4 y4 = ld4_stream(p.y + b * p.y_bs + j);
      const float4 s4 = ld4_stream(p.scale + b * p.scale_bs + j);
      const float4 m4 = p.mean ? ld4_stream(p.mean + b * p.mean_bs + j) : z4;
      const float4 n4 = use_rng ? rc.noise4((unsigned long long)(b * p.CS + j) >> 2)
                                : (train ? ld4_stream(p.noise + b * p.noise_bs + j) : z4);
      const float4 l4 = ld4_stream(p.lik + b * p.lik_bs + j);
      const float4 gl4 = has_gl ? ld4_stream(p.g_lik + b * p.g_lik_bs + j) : z4;
      const float4 gy4 = p.g_y_hat ? ld4_stream(p.g_y_hat + b * p.g_y_hat_bs + j) : z4;
      const GcGrad g0 = gc_elem_bwd(y4.x, s4.x, m4.x, n4.x, train, l4.x, gl4.x, has_gl, p.bpp_coef, gy4.x, p.scale_bound, p.lik_bound);
      const GcGrad g1 = gc_elem_bwd(y4.y, s4.y, m4.y, n4.y, train, l4.y, gl4.y, has_gl, p.bpp_coef, gy4.y, p.scale_bound, p.lik_bound);
      const GcGrad g2 = gc_elem_bwd(y4.z, s4.z, m4.z, n4.z, train, l4.z, gl4.z, has_gl, p.bpp_coef, gy4.z, p.scale_bound, p.lik_bound);
      const GcGrad g3 = gc_elem_bwd(y4.w, s4.w, m4.w, n4.w, train, l4.w, gl4.w, has_gl, p.bpp_coef, gy4.w, p.scale_bound, p.lik_bound);
      st4_stream(p.g_y + b * p.g_y_bs + j, make_float4(g0.g_y, g1.g_y, g2.g_y, g3.g_y));
      st4_stream(p.g_scale + b * p.g_scale_bs + j, make_float4(g0.g_scale, g1.g_scale, g2.g_scale, g3.g_scale));
      if (p.g_mean)
        st4_stream(p.g_mean + b * p.g_mean_bs + j, make_float4(g0.g_mean, g1.g_mean, g2.g_mean, g3.g_mean));
    } else {
      const float y = p.y[b * p.y_bs + j];
      const float s = p.scale[b * p.scale_bs + j];
      const float m = p.mean ? p.mean[b * p.mean_bs + j] : 0.f;
      const float n = use_rng ? rc.noise1((unsigned long long)(b * p.CS + j)) : (train ? p.noise[b * p.noise_bs + j] : 0.f);
      const float l = p.lik[b * p.lik_bs + j];
      const float gl = has_gl ? p.g_lik[b * p.g_lik_bs + j] : 0.f;
      const float gy = p.g_y_hat ? p.g_y_hat[b * p.g_y_hat_bs + j] : 0.f;
      const GcGrad g = gc_elem_bwd(y, s, m, n, train, l, gl, has_gl, p.bpp_coef, gy, p.scale_bound, p.lik_bound);
      p.g_y[b * p.g_y_bs + j] = g.g_y;
      p.g_scale[b * p.g_scale_bs + j] = g.g_scale;
      if (p.g_mean) p.g_mean[b * p.g_mean_bs + j] = g.g_mean;
    }
  }
}

// ------------------------------------------------------------------------------------------
// LRP add
// ------------------------------------------------------------------------------------------
template <bool VEC, bool BWD>
__global__ void __launch_bounds__(256)
lrp_kernel(float* __restrict__ io, int64_t io_bs, const float* __restrict__ lrp, int64_t lrp_bs,
           const float* __restrict__ g, int64_t g_bs, int64_t B, int64_t CS) {
  constexpr int W = VEC ? 4 : 1;
  const int64_t per_b = CS / W;
  // grid = (blocks per batch row, batch rows): no 64-bit division per item
  for (int64_t b = blockIdx.y; b < B; b += gridDim.y)
  for (int64_t jj = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; jj < per_b; jj += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = jj * W;
    if constexpr (VEC) {
      const float4 l4 = ld4_stream(lrp + b * lrp_bs + j);
      const float t0 = tanhf(l4.x), t1 = tanhf(l4.y), t2 = tanhf(l4.z), t3 = tanhf(l4.w);
      if constexpr (!BWD) {
        float4 v = ld4(io + b * io_bs + j);
        v.x += 0.5f * t0; v.y += 0.5f * t1; v.z += 0.5f * t2; v.w += 0.5f * t3;
        st4(io + b * io_bs + j, v);
      } else {
        const float4 g4 = ld4_stream(g + b * g_bs + j);
        st4_stream(io + b * io_bs + j,
                   make_float4(g4.x * (0.5f * (1.f - t0 * t0)), g4.y * (0.5f * (1.f - t1 * t1)),
                               g4.z * (0.5f * (1.f - t2 * t2)), g4.w * (0.5f * (1.f - t3 * t3))));
      }
    } else {
      const float t = tanhf(lrp[b * lrp_bs + j]);
      if constexpr (!BWD) io[b * io_bs + j] += 0.5f * t;
      else io[b * io_bs + j] = g[b * g_bs + j] * (0.5f * (1.f - t * t));
    }
  }
}

// ------------------------------------------------------------------------------------------
// symbols + scale-table indexes
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gc_symbols_indexes_kernel(const float* __restrict__ y, int64_t y_bs, const float* __restrict__ scale,
                          int64_t scale_bs, const float* __restrict__ mean, int64_t mean_bs,
                          const float* __restrict__ table, int T, int32_t* __restrict__ symbols,
                          int64_t symbols_bs, int32_t* __restrict__ indexes, int64_t indexes_bs,
                          int64_t B, int64_t CS, float scale_bound) {
  __shared__ float tab[256];
  for (int t = threadIdx.x; t < T; t += blockDim.x) tab[t] = table[t];
  __syncthreads();
  const int64_t total = B * CS;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / CS, j = i - b * CS;
    if (symbols) {
      const float m = mean ? mean[b * mean_bs + j] : 0.f;
      symbols[b * symbols_bs + j] = (int32_t)rintf(y[b * y_bs + j] - m);
    }
    if (indexes) {
      const float sc = fmaxf(scale[b * scale_bs + j], scale_bound);
      // index = (T-1) - #{t < T-1 : sc <= tab[t]}.  The table ascends, so the count is
      // (T-1) - first t in [0, T-1] with sc <= tab[t] (T-1 if none): a binary search
      // instead of the reference's T-1 full-tensor passes.
      int lo = 0, hi = T - 1;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (sc <= tab[mid]) hi = mid; else lo = mid + 1;
      }
      indexes[b * indexes_bs + j] = lo;
    }
  }
}

// ------------------------------------------------------------------------------------------
// EntropyBottleneck, filters (3,3,3,3).  Per-channel parameter pack in shared memory:
//   [0..2]   softplus(M0) (3x1)   [3..11]  softplus(M1) (3x3) [12..20] M2 [21..29] M3 [30..32] M4 (1x3)
//   [33..35] b0 [36..38] b1 [39..41] b2 [42..44] b3 [45] b4
//   [46..48] tanh(f0) [49..51] tanh(f1) [52..54] tanh(f2) [55..57] tanh(f3)
// ------------------------------------------------------------------------------------------
constexpr int kEbPack = 58;

struct EbPtrs {
  const float* matrix[5];
  const float* bias[5];
  const float* factor[4];
};
struct EbGradPtrs {
  float* matrix[5];
  float* bias[5];
  float* factor[4];
};

__device__ __forceinline__ float softplusf(float x) {
  // torch.nn.functional.softplus, beta=1, threshold=20
  return x > 20.f ? x : log1pf(expf(x));
}

// raw -> transformed parameters of channel c (58 values), done by the first 58 threads.  The source address is
// SELECTED per thread and loaded once after the selection: a 14-way if-chain with one load per arm serialises
// 14 global-load latencies inside the two warps (it was ~6 us of the 10 us this kernel took at cfg2).
// Returns the thread's raw parameter (the backward's softplus / tanh chain rule needs it again).
__device__ __forceinline__ float eb_load_pack(const EbPtrs& P, int c, float* pk) {
  const int t = threadIdx.x;
  float raw = 0.f;
  if (t < kEbPack) {
    const float* src = P.matrix[0];
    int i = c * 3 + t;
    if (t >= 3) { src = P.matrix[1]; i = c * 9 + (t - 3); }
    if (t >= 12) { src = P.matrix[2]; i = c * 9 + (t - 12); }
    if (t >= 21) { src = P.matrix[3]; i = c * 9 + (t - 21); }
    if (t >= 30) { src = P.matrix[4]; i = c * 3 + (t - 30); }
    if (t >= 33) { src = P.bias[0]; i = c * 3 + (t - 33); }
    if (t >= 36) { src = P.bias[1]; i = c * 3 + (t - 36); }
    if (t >= 39) { src = P.bias[2]; i = c * 3 + (t - 39); }
    if (t >= 42) { src = P.bias[3]; i = c * 3 + (t - 42); }
    if (t >= 45) { src = P.bias[4]; i = c; }
    if (t >= 46) { src = P.factor[0]; i = c * 3 + (t - 46); }
    if (t >= 49) { src = P.factor[1]; i = c * 3 + (t - 49); }
    if (t >= 52) { src = P.factor[2]; i = c * 3 + (t - 52); }
    if (t >= 55) { src = P.factor[3]; i = c * 3 + (t - 55); }
    raw = __ldg(src + i);
    pk[t] = t < 33 ? softplusf(raw) : (t < 46 ? raw : tanhf(raw));
  }
  return raw;
}

// logits = _logits_cumulative(x) for one scalar input.  Operation order follows upstream:
// matmul(softplus(M), x) accumulated left to right, + bias, + tanh(factor) * tanh(logits).
// When KEEP, th[l][j] = tanh(pre-gate activation) of layers 0..3 and the layer inputs are kept for backward (the
// backward needs the activation only through its tanh: keeping that saves 24 tanhf per element).
template <bool KEEP>
__device__ __forceinline__ float eb_logits(const float* pk, float x, float (*a)[3], float (*in)[3]) {
  float h[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float t = pk[j] * x + pk[33 + j];
    const float th = tanhf(t);
    if (KEEP) a[0][j] = th;
    h[j] = t + pk[46 + j] * th;
  }
#pragma unroll
  for (int l = 1; l < 4; ++l) {
    float o[3];
    if (KEEP) { in[l][0] = h[0]; in[l][1] = h[1]; in[l][2] = h[2]; }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float* mrow = pk + 3 + (l - 1) * 9 + j * 3;
      float t = mrow[0] * h[0];
      t = fmaf(mrow[1], h[1], t);
      t = fmaf(mrow[2], h[2], t);
      t += pk[33 + 3 * l + j];
      const float th = tanhf(t);
      if (KEEP) a[l][j] = th;
      o[j] = t + pk[46 + 3 * l + j] * th;
    }
    h[0] = o[0]; h[1] = o[1]; h[2] = o[2];
  }
  if (KEEP) { in[4][0] = h[0]; in[4][1] = h[1]; in[4][2] = h[2]; }
  float t = pk[30] * h[0];
  t = fmaf(pk[31], h[1], t);
  t = fmaf(pk[32], h[2], t);
  return t + pk[45];
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

struct EbFwdParams {
  const float *z, *noise, *quantiles;
  float *lik, *z_hat, *outputs;
  double* log2_sum;
  int64_t B, C, S;
  float lik_bound;
  EbPtrs P;
  RngArg rng;
};

// grid = (C, chunks): CTA (c, k) handles elements k, k+chunks, ... of channel c's B*S values.
__global__ void __launch_bounds__(128) eb_fwd_kernel(const EbFwdParams p) {
  __shared__ float pk[64];
  __shared__ float red[32];
  const int c = blockIdx.x;
  // everything that does not depend on the parameter pack is loaded first, so its latency overlaps the pack's
  const float med = __ldg(p.quantiles + c * 3 + 1);
  const bool use_rng = p.rng.state != nullptr;
  const bool train = p.noise != nullptr || use_rng;
  const RngCtx rc(p.rng);
  const int64_t n = p.B * p.S;
  const int64_t e0 = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
  float z_pre = 0.f, n_pre = 0.f;
  if (e0 < n) {
    const int64_t off0 = ((e0 / p.S) * p.C + c) * p.S + e0 % p.S;
    z_pre = p.z[off0];
    if (p.noise) n_pre = p.noise[off0];
  }
  eb_load_pack(p.P, c, pk);
  __syncthreads();
  float acc = 0.f;
  for (int64_t e = e0; e < n; e += (int64_t)gridDim.y * blockDim.x) {
    const int64_t b = e / p.S, s = e - b * p.S;
    const int64_t off = (b * p.C + c) * p.S + s;
    const float z = e == e0 ? z_pre : p.z[off];
    const float zh = rintf(z - med) + med;
    const float x = train ? (z + (use_rng ? rc.noise1((unsigned long long)off) : (e == e0 ? n_pre : p.noise[off]))) : zh;
    const float lo = eb_logits<false>(pk, x - 0.5f, nullptr, nullptr);
    const float up = eb_logits<false>(pk, x + 0.5f, nullptr, nullptr);
    const float t = lo + up;
    const float sg = (t > 0.f) ? -1.f : ((t < 0.f) ? 1.f : 0.f);  // -sign(lower + upper)
    const float lik = fmaxf(fabsf(sigmoidf_(sg * up) - sigmoidf_(sg * lo)), p.lik_bound);
    p.lik[off] = lik;
    if (p.z_hat) p.z_hat[off] = zh;
    if (p.outputs) p.outputs[off] = x;
    acc += log2f(lik);
  }
  if (p.log2_sum) {
    const float tot = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(p.log2_sum, (double)tot);
  }
}

struct EbBwdParams {
  const float *z, *noise, *quantiles, *lik, *g_lik, *g_z_hat;
  float* g_z;
  int64_t B, C, S;
  float bpp_coef, lik_bound;
  bool param_grads;
  EbPtrs P;
  EbGradPtrs G;
  RngArg rng;
};

// Reverse-mode through one logits chain.  g_out = dL/dlogit.  Accumulates transformed-parameter
// gradients into gp[58] (w.r.t. softplus(M), b, tanh(f)) and returns dL/dx.
__device__ __forceinline__ float eb_logits_bwd(const float* pk, float x, const float (*a)[3],
                                               const float (*in)[3], float g_out, float* gp) {
  float gh[3];
  // layer 4: out = M4 . in4 + b4
  gp[45] += g_out;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    gp[30 + i] += g_out * in[4][i];
    gh[i] = g_out * pk[30 + i];
  }
#pragma unroll
  for (int l = 3; l >= 1; --l) {
    float gin[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float th = a[l][j];                         // tanh of the activation, kept by eb_logits<true>
      const float tf = pk[46 + 3 * l + j];
      gp[46 + 3 * l + j] += gh[j] * th;                 // d/d tanh(f)
      const float ga = gh[j] * (1.f + tf * (1.f - th * th));  // through h = a + tf*tanh(a)
      gp[33 + 3 * l + j] += ga;                         // bias
      const float* mrow = pk + 3 + (l - 1) * 9 + j * 3;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        gp[3 + (l - 1) * 9 + j * 3 + i] += ga * in[l][i];
        gin[i] = fmaf(ga, mrow[i], gin[i]);
      }
    }
    gh[0] = gin[0]; gh[1] = gin[1]; gh[2] = gin[2];
  }
  float gx = 0.f;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float th = a[0][j];
    const float tf = pk[46 + j];
    gp[46 + j] += gh[j] * th;
    const float ga = gh[j] * (1.f + tf * (1.f - th * th));
    gp[33 + j] += ga;
    gp[j] += ga * x;
    gx = fmaf(ga, pk[j], gx);
  }
  return gx;
}

__global__ void __launch_bounds__(128) eb_bwd_kernel(const EbBwdParams p) {
  __shared__ float pk[64];
  __shared__ float gsum[2][64];
  __shared__ float gred[kEbPack][129];         // [parameter][thread], odd pitch: conflict-free both ways
  const int c = blockIdx.x;
  const float med = __ldg(p.quantiles + c * 3 + 1);
  const bool use_rng = p.rng.state != nullptr;
  const bool train = p.noise != nullptr || use_rng;
  const RngCtx rc(p.rng);
  const int64_t n = p.B * p.S;
  const int64_t e0 = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
  float z_pre = 0.f, n_pre = 0.f, lik_pre = 1.f, gl_pre = 0.f, gzh_pre = 0.f;   // first element: loads ahead of the pack
  if (e0 < n) {
    const int64_t off0 = ((e0 / p.S) * p.C + c) * p.S + e0 % p.S;
    z_pre = p.z[off0];
    if (p.noise) n_pre = p.noise[off0];
    lik_pre = p.lik[off0];
    if (p.g_lik) gl_pre = p.g_lik[off0];
    if (p.g_z_hat) gzh_pre = p.g_z_hat[off0];
  }
  const float raw = eb_load_pack(p.P, c, pk);
  __syncthreads();
  float gp[kEbPack];
#pragma unroll
  for (int i = 0; i < kEbPack; ++i) gp[i] = 0.f;
  for (int64_t e = e0; e < n; e += (int64_t)gridDim.y * blockDim.x) {
    const int64_t b = e / p.S, s = e - b * p.S;
    const int64_t off = (b * p.C + c) * p.S + s;
    const bool first = e == e0;
    const float z = first ? z_pre : p.z[off];
    const float x = train ? (z + (use_rng ? rc.noise1((unsigned long long)off) : (first ? n_pre : p.noise[off]))) : (rintf(z - med) + med);
    float a_lo[4][3], in_lo[5][3], a_up[4][3], in_up[5][3];
    const float lo = eb_logits<true>(pk, x - 0.5f, a_lo, in_lo);
    const float up = eb_logits<true>(pk, x + 0.5f, a_up, in_up);
    const float t = lo + up;
    const float sg = (t > 0.f) ? -1.f : ((t < 0.f) ? 1.f : 0.f);
    const float su = sigmoidf_(sg * up), sl = sigmoidf_(sg * lo);
    const float D = su - sl;
    const float lik = first ? lik_pre : p.lik[off];
    float gl = p.g_lik ? (first ? gl_pre : p.g_lik[off]) : (p.bpp_coef / lik);
    gl = ((lik > p.lik_bound) || (gl < 0.f)) ? gl : 0.f;
    const float sd = (D > 0.f) ? 1.f : ((D < 0.f) ? -1.f : 0.f);  // d|D|/dD
    const float g_up = gl * sd * su * (1.f - su) * sg;
    const float g_lo = -gl * sd * sl * (1.f - sl) * sg;
    float gx = eb_logits_bwd(pk, x + 0.5f, a_up, in_up, g_up, gp);
    gx += eb_logits_bwd(pk, x - 0.5f, a_lo, in_lo, g_lo, gp);
    const float gzh = p.g_z_hat ? (first ? gzh_pre : p.g_z_hat[off]) : 0.f;  // STE: d z_hat / d z = 1
    p.g_z[off] = gzh + (train ? gx : 0.f);
  }
  if (!p.param_grads) return;
  // CTA reduction of the 58 transformed-parameter gradients through shared memory: every thread parks its 58
  // partials, then thread (half, parameter) adds 64 of the 128 columns in fixed order (58 STS + 64 LDS/FADD per
  // thread instead of 58 five-step shuffle trees; deterministic within the CTA).
#pragma unroll
  for (int i = 0; i < kEbPack; ++i) gred[i][threadIdx.x] = gp[i];
  __syncthreads();
  {
    const int prm = threadIdx.x & 63, half = threadIdx.x >> 6;
    if (prm < kEbPack) {
      float a = 0.f;
#pragma unroll 8
      for (int i = 0; i < 64; ++i) a += gred[prm][half * 64 + i];
      gsum[half][prm] = a;
    }
  }
  __syncthreads();
  // Chain through softplus / tanh of the raw parameters and add to global (one atomic each).
  const int t = threadIdx.x;
  if (t < kEbPack) {
    const float g = gsum[0][t] + gsum[1][t];
    if (t < 33) {
      int l, k;
      if (t < 3) { l = 0; k = t; } else if (t < 12) { l = 1; k = t - 3; } else if (t < 21) { l = 2; k = t - 12; }
      else if (t < 30) { l = 3; k = t - 21; } else { l = 4; k = t - 30; }
      const int per = (l == 0 || l == 4) ? 3 : 9;
      atomicAdd(&p.G.matrix[l][c * per + k], g * sigmoidf_(raw));  // d softplus = sigmoid
    } else if (t < 46) {
      const int l = (t - 33) / 3, k = (t - 33) % 3;
      if (t == 45) atomicAdd(&p.G.bias[4][c], g);
      else atomicAdd(&p.G.bias[l][c * 3 + k], g);
    } else {
      const int l = (t - 46) / 3, k = (t - 46) % 3;
      const float tf = pk[t];
      atomicAdd(&p.G.factor[l][c * 3 + k], g * (1.f - tf * tf));
    }
  }
}

// ------------------------------------------------------------------------------------------
// log2-sum (rate term) forward / backward
// ------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(256)
log2_sum_kernel(const float* __restrict__ lik, int64_t n, double* __restrict__ out) {
  __shared__ float red[32];
  float acc = 0.f;
  if constexpr (VEC) {
    const int64_t n4 = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (int64_t)gridDim.x * blockDim.x) {
      const float4 v = ld4_stream(lik + 4 * i);
      acc += (log2f(v.x) + log2f(v.y)) + (log2f(v.z) + log2f(v.w));
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) acc += log2f(lik[(n4 << 2) + threadIdx.x]);
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
      acc += log2f(lik[i]);
  }
  const float tot = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(out, (double)tot);
}

__global__ void __launch_bounds__(256)
log2_sum_bwd_kernel(const float* __restrict__ lik, float coef, const double* __restrict__ coef_dev,
                    float* __restrict__ g, int64_t n) {
  if (coef_dev) coef *= (float)(*coef_dev);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    g[i] = coef / lik[i];
}

}  // namespace clc

// ==========================================================================================
// C ABI
// ==========================================================================================
using namespace clc;

static bool vec_ok(int64_t CS, std::initializer_list<const void*> ptrs, std::initializer_list<int64_t> strides) {
  if (CS % 4) return false;
  for (const void* q : ptrs) if (q && !aligned16(q)) return false;
  for (int64_t s : strides) if (s % 4) return false;
  return true;
}

static int gc_fwd_impl(const float* y, int64_t y_bs, const float* scale, int64_t scale_bs,
                       const float* mean, int64_t mean_bs, const float* noise, int64_t noise_bs, RngArg rng,
                       float* lik, int64_t lik_bs, float* y_hat, int64_t y_hat_bs,
                       float* outputs, int64_t outputs_bs, double* log2_sum,
                       int64_t B, int64_t CS, float scale_bound, float lik_bound, void* stream) {
  if (!y || !scale || !lik || B < 0 || CS < 0) return CLC_ERR_INVALID_ARGUMENT;
  if (B == 0 || CS == 0) return CLC_OK;
  GcFwdParams p{y, scale, mean, noise, lik, y_hat, outputs, log2_sum,
                y_bs, scale_bs, mean_bs, noise_bs, lik_bs, y_hat_bs, outputs_bs, B, CS, scale_bound, lik_bound, rng};
  const bool vec = vec_ok(CS, {y, scale, mean, noise, lik, y_hat, outputs},
                          {y_bs, scale_bs, mean_bs, noise_bs, lik_bs, y_hat_bs, outputs_bs});
  cudaStream_t st = (cudaStream_t)stream;
  if (vec) gc_fwd_kernel<true><<<grid_rows(B, CS / 4, 256), 256, 0, st>>>(p);
  else gc_fwd_kernel<false><<<grid_rows(B, CS, 256), 256, 0, st>>>(p);
  CLC_CHECK_LAUNCH("clc_gc_fwd");
  return CLC_OK;
}

extern "C" int clc_gc_fwd(const float* y, int64_t y_bs, const float* scale, int64_t scale_bs,
                          const float* mean, int64_t mean_bs, const float* noise, int64_t noise_bs,
                          float* lik, int64_t lik_bs, float* y_hat, int64_t y_hat_bs,
                          float* outputs, int64_t outputs_bs, double* log2_sum,
                          int64_t B, int64_t CS, float scale_bound, float lik_bound, void* stream) {
  return gc_fwd_impl(y, y_bs, scale, scale_bs, mean, mean_bs, noise, noise_bs, RngArg{nullptr, 0}, lik, lik_bs, y_hat,
                     y_hat_bs, outputs, outputs_bs, log2_sum, B, CS, scale_bound, lik_bound, stream);
}

extern "C" int clc_gc_fwd_rng(const float* y, int64_t y_bs, const float* scale, int64_t scale_bs,
                              const float* mean, int64_t mean_bs, const uint64_t* rng_state, uint64_t rng_offset,
                              float* lik, int64_t lik_bs, float* y_hat, int64_t y_hat_bs,
                              float* outputs, int64_t outputs_bs, double* log2_sum,
                              int64_t B, int64_t CS, float scale_bound, float lik_bound, void* stream) {
  if (!rng_state) return CLC_ERR_INVALID_ARGUMENT;
  return gc_fwd_impl(y, y_bs, scale, scale_bs, mean, mean_bs, nullptr, 0,
                     RngArg{reinterpret_cast<const unsigned long long*>(rng_state), rng_offset}, lik, lik_bs, y_hat,
                     y_hat_bs, outputs, outputs_bs, log2_sum, B, CS, scale_bound, lik_bound, stream);
}

extern "C" int clc_bpp_finalize(const double* log2_sums, int32_t n, double num_pixels, double* bpp,
                                uint64_t* rng_state, uint64_t rng_advance, void* stream) {
  if (!log2_sums || !bpp || n < 0 || !(num_pixels > 0)) return CLC_ERR_INVALID_ARGUMENT;
  bpp_finalize_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(log2_sums, n, num_pixels, bpp,
                                                         reinterpret_cast<unsigned long long*>(rng_state), rng_advance);
  CLC_CHECK_LAUNCH("clc_bpp_finalize");
  return CLC_OK;
}

extern "C" int clc_zero(void* p, size_t bytes, void* stream) {
  if (!p && bytes) return CLC_ERR_INVALID_ARGUMENT;
  if (bytes) CLC_CUDA(cudaMemsetAsync(p, 0, bytes, (cudaStream_t)stream));
  return CLC_OK;
}

extern "C" int clc_rng_advance(uint64_t* rng_state, uint64_t n, void* stream) {
  if (!rng_state) return CLC_ERR_INVALID_ARGUMENT;
  rng_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned long long*>(rng_state), n);
  CLC_CHECK_LAUNCH("clc_rng_advance");
  return CLC_OK;
}

static int gc_bwd_impl(const float* y, int64_t y_bs, const float* scale, int64_t scale_bs,
                       const float* mean, int64_t mean_bs, const float* noise, int64_t noise_bs, RngArg rng,
                       const float* lik, int64_t lik_bs, const float* g_lik, int64_t g_lik_bs,
                       float bpp_coef, const float* g_y_hat, int64_t g_y_hat_bs,
                       float* g_y, int64_t g_y_bs, float* g_scale, int64_t g_scale_bs,
                       float* g_mean, int64_t g_mean_bs,
                       int64_t B, int64_t CS, float scale_bound, float lik_bound, void* stream) {
  if (!y || !scale || !lik || !g_y || !g_scale || B < 0 || CS < 0) return CLC_ERR_INVALID_ARGUMENT;
  if (mean && !g_mean) return CLC_ERR_INVALID_ARGUMENT;
  if (B == 0 || CS == 0) return CLC_OK;
  GcBwdParams p{y, scale, mean, noise, lik, g_lik, g_y_hat, g_y, g_scale, g_mean,
                y_bs, scale_bs, mean_bs, noise_bs, lik_bs, g_lik_bs, g_y_hat_bs, g_y_bs, g_scale_bs, g_mean_bs,
                B, CS, bpp_coef, scale_bound, lik_bound, rng};
  const bool vec = vec_ok(CS, {y, scale, mean, noise, lik, g_lik, g_y_hat, g_y, g_scale, g_mean},
                          {y_bs, scale_bs, mean_bs, noise_bs, lik_bs, g_lik_bs, g_y_hat_bs, g_y_bs, g_scale_bs, g_mean_bs});
  cudaStream_t st = (cudaStream_t)stream;
  if (vec) gc_bwd_kernel<true><<<grid_rows(B, CS / 4, 256), 256, 0, st>>>(p);
  else gc_bwd_kernel<false><<<grid_rows(B, CS, 256), 256, 0, st>>>(p);
  CLC_CHECK_LAUNCH("clc_gc_bwd");
  return CLC_OK;
}

extern "C" int clc_gc_bwd(const float* y, int64_t y_bs, const float* scale, int64_t scale_bs,
                          const float* mean, int64_t mean_bs, const float* noise, int64_t noise_bs,
                          const float* lik, int64_t lik_bs, const float* g_lik, int64_t g_lik_bs,
                          float bpp_coef, const float* g_y_hat, int64_t g_y_hat_bs,
                          float* g_y, int64_t g_y_bs, float* g_scale, int64_t g_scale_bs,
                          float* g_mean, int64_t g_mean_bs,
                          int64_t B, int64_t CS, float scale_bound, float lik_bound, void* stream) {
  return gc_bwd_impl(y, y_bs, scale, scale_bs, mean, mean_bs, noise, noise_bs, RngArg{nullptr, 0}, lik, lik_bs, g_lik,
                     g_lik_bs, bpp_coef, g_y_hat, g_y_hat_bs, g_y, g_y_bs, g_scale, g_scale_bs, g_mean, g_mean_bs, B, CS,
                     scale_bound, lik_bound, stream);
}

extern "C" int clc_gc_bwd_rng(const float* y, int64_t y_bs, const float* scale, int64_t scale_bs,
                              const float* mean, int64_t mean_bs, const uint64_t* rng_state, uint64_t rng_offset,
                              const float* lik, int64_t lik_bs, const float* g_lik, int64_t g_lik_bs,
                              float bpp_coef, const float* g_y_hat, int64_t g_y_hat_bs,
                              float* g_y, int64_t g_y_bs, float* g_scale, int64_t g_scale_bs,
                              float* g_mean, int64_t g_mean_bs,
                              int64_t B, int64_t CS, float scale_bound, float lik_bound, void* stream) {
  if (!rng_state) return CLC_ERR_INVALID_ARGUMENT;
  return gc_bwd_impl(y, y_bs, scale, scale_bs, mean, mean_bs, nullptr, 0,
                     RngArg{reinterpret_cast<const unsigned long long*>(rng_state), rng_offset}, lik, lik_bs, g_lik,
                     g_lik_bs, bpp_coef, g_y_hat, g_y_hat_bs, g_y, g_y_bs, g_scale, g_scale_bs, g_mean, g_mean_bs, B, CS,
                     scale_bound, lik_bound, stream);
}

extern "C" int clc_lrp_add_fwd(float* y_hat, int64_t y_hat_bs, const float* lrp, int64_t lrp_bs,
                               int64_t B, int64_t CS, void* stream) {
  if (!y_hat || !lrp || B < 0 || CS < 0) return CLC_ERR_INVALID_ARGUMENT;
  if (B == 0 || CS == 0) return CLC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (vec_ok(CS, {y_hat, lrp}, {y_hat_bs, lrp_bs}))
    lrp_kernel<true, false><<<grid_rows(B, CS / 4, 256), 256, 0, st>>>(y_hat, y_hat_bs, lrp, lrp_bs, nullptr, 0, B, CS);
  else
    lrp_kernel<false, false><<<grid_rows(B, CS, 256), 256, 0, st>>>(y_hat, y_hat_bs, lrp, lrp_bs, nullptr, 0, B, CS);
  CLC_CHECK_LAUNCH("clc_lrp_add_fwd");
  return CLC_OK;
}

extern "C" int clc_lrp_add_bwd(const float* g, int64_t g_bs, const float* lrp, int64_t lrp_bs,
                               float* g_lrp, int64_t g_lrp_bs, int64_t B, int64_t CS, void* stream) {
  if (!g || !lrp || !g_lrp || B < 0 || CS < 0) return CLC_ERR_INVALID_ARGUMENT;
  if (B == 0 || CS == 0) return CLC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (vec_ok(CS, {g, lrp, g_lrp}, {g_bs, lrp_bs, g_lrp_bs}))
    lrp_kernel<true, true><<<grid_rows(B, CS / 4, 256), 256, 0, st>>>(g_lrp, g_lrp_bs, lrp, lrp_bs, g, g_bs, B, CS);
  else
    lrp_kernel<false, true><<<grid_rows(B, CS, 256), 256, 0, st>>>(g_lrp, g_lrp_bs, lrp, lrp_bs, g, g_bs, B, CS);
  CLC_CHECK_LAUNCH("clc_lrp_add_bwd");
  return CLC_OK;
}

extern "C" int clc_gc_symbols_indexes(const float* y, int64_t y_bs, const float* scale, int64_t scale_bs,
                                      const float* mean, int64_t mean_bs, const float* scale_table, int T,
                                      int32_t* symbols, int64_t symbols_bs, int32_t* indexes, int64_t indexes_bs,
                                      int64_t B, int64_t CS, float scale_bound, void* stream) {
  if (B < 0 || CS < 0) return CLC_ERR_INVALID_ARGUMENT;
  if (symbols && !y) return CLC_ERR_INVALID_ARGUMENT;
  if (indexes && (!scale || !scale_table || T < 1)) return CLC_ERR_INVALID_ARGUMENT;
  if (T > 256) return CLC_ERR_UNSUPPORTED;
  if (B == 0 || CS == 0 || (!symbols && !indexes)) return CLC_OK;
  gc_symbols_indexes_kernel<<<grid_for(B * CS, 256), 256, 0, (cudaStream_t)stream>>>(
      y, y_bs, scale, scale_bs, mean, mean_bs, scale_table, T, symbols, symbols_bs, indexes, indexes_bs,
      B, CS, scale_bound);
  CLC_CHECK_LAUNCH("clc_gc_symbols_indexes");
  return CLC_OK;
}

static int eb_grid_y(int64_t C, int64_t n) {
  // enough CTAs per channel to cover n elements at 128 threads, but keep the whole grid
  // near 148 x 8 CTAs so per-channel parameter packs are not re-derived needlessly.
  int64_t chunks = (n + 127) / 128;
  int64_t cap = ((int64_t)kNumSMs * 16 + C - 1) / C;
  if (cap < 1) cap = 1;
  if (chunks > cap) chunks = cap;
  if (chunks > 65535) chunks = 65535;
  return (int)(chunks < 1 ? 1 : chunks);
}

static int eb_fwd_impl(const float* z, const float* noise, RngArg rng, const float* const matrix[5],
                       const float* const bias[5], const float* const factor[4], const float* quantiles,
                       float* lik, float* z_hat, float* outputs, double* log2_sum,
                       int64_t B, int64_t C, int64_t S, float lik_bound, void* stream) {
  if (!z || !matrix || !bias || !factor || !quantiles || !lik || B < 0 || C < 0 || S < 0)
    return CLC_ERR_INVALID_ARGUMENT;
  if (B == 0 || C == 0 || S == 0) return CLC_OK;
  if (C > 2147483647LL) return CLC_ERR_UNSUPPORTED;
  EbFwdParams p;
  p.z = z; p.noise = noise; p.quantiles = quantiles; p.lik = lik; p.z_hat = z_hat; p.outputs = outputs;
  p.log2_sum = log2_sum; p.B = B; p.C = C; p.S = S; p.lik_bound = lik_bound; p.rng = rng;
  for (int i = 0; i < 5; ++i) { p.P.matrix[i] = matrix[i]; p.P.bias[i] = bias[i]; if (!matrix[i] || !bias[i]) return CLC_ERR_INVALID_ARGUMENT; }
  for (int i = 0; i < 4; ++i) { p.P.factor[i] = factor[i]; if (!factor[i]) return CLC_ERR_INVALID_ARGUMENT; }
  dim3 grid((unsigned)C, (unsigned)eb_grid_y(C, B * S));
  eb_fwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(p);
  CLC_CHECK_LAUNCH("clc_eb_fwd");
  return CLC_OK;
}

extern "C" int clc_eb_fwd(const float* z, const float* noise, const float* const matrix[5],
                          const float* const bias[5], const float* const factor[4], const float* quantiles,
                          float* lik, float* z_hat, float* outputs, double* log2_sum,
                          int64_t B, int64_t C, int64_t S, float lik_bound, void* stream) {
  return eb_fwd_impl(z, noise, RngArg{nullptr, 0}, matrix, bias, factor, quantiles, lik, z_hat, outputs, log2_sum, B, C, S,
                     lik_bound, stream);
}

extern "C" int clc_eb_fwd_rng(const float* z, const uint64_t* rng_state, uint64_t rng_offset,
                              const float* const matrix[5], const float* const bias[5], const float* const factor[4],
                              const float* quantiles, float* lik, float* z_hat, float* outputs, double* log2_sum,
                              int64_t B, int64_t C, int64_t S, float lik_bound, void* stream) {
  if (!rng_state) return CLC_ERR_INVALID_ARGUMENT;
  return eb_fwd_impl(z, nullptr, RngArg{reinterpret_cast<const unsigned long long*>(rng_state), rng_offset}, matrix, bias,
                     factor, quantiles, lik, z_hat, outputs, log2_sum, B, C, S, lik_bound, stream);
}

static int eb_bwd_impl(const float* z, const float* noise, RngArg rng, const float* const matrix[5],
                       const float* const bias[5], const float* const factor[4], const float* quantiles,
                       const float* lik, const float* g_lik, float bpp_coef, const float* g_z_hat,
                       float* g_z, float* const g_matrix[5], float* const g_bias[5], float* const g_factor[4],
                       int64_t B, int64_t C, int64_t S, float lik_bound, void* stream) {
  if (!z || !matrix || !bias || !factor || !quantiles || !lik || !g_z || B < 0 || C < 0 || S < 0)
    return CLC_ERR_INVALID_ARGUMENT;
  if (B == 0 || C == 0 || S == 0) return CLC_OK;
  EbBwdParams p;
  p.z = z; p.noise = noise; p.quantiles = quantiles; p.lik = lik; p.g_lik = g_lik; p.g_z_hat = g_z_hat;
  p.g_z = g_z; p.B = B; p.C = C; p.S = S; p.bpp_coef = bpp_coef; p.lik_bound = lik_bound; p.rng = rng;
  p.param_grads = g_matrix && g_bias && g_factor;
  for (int i = 0; i < 5; ++i) {
    p.P.matrix[i] = matrix[i]; p.P.bias[i] = bias[i];
    if (!matrix[i] || !bias[i]) return CLC_ERR_INVALID_ARGUMENT;
    p.G.matrix[i] = p.param_grads ? g_matrix[i] : nullptr;
    p.G.bias[i] = p.param_grads ? g_bias[i] : nullptr;
    if (p.param_grads && (!g_matrix[i] || !g_bias[i])) return CLC_ERR_INVALID_ARGUMENT;
  }
  for (int i = 0; i < 4; ++i) {
    p.P.factor[i] = factor[i];
    if (!factor[i]) return CLC_ERR_INVALID_ARGUMENT;
    p.G.factor[i] = p.param_grads ? g_factor[i] : nullptr;
    if (p.param_grads && !g_factor[i]) return CLC_ERR_INVALID_ARGUMENT;
  }
  dim3 grid((unsigned)C, (unsigned)eb_grid_y(C, B * S));
  eb_bwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(p);
  CLC_CHECK_LAUNCH("clc_eb_bwd");
  return CLC_OK;
}

extern "C" int clc_eb_bwd(const float* z, const float* noise, const float* const matrix[5],
                          const float* const bias[5], const float* const factor[4], const float* quantiles,
                          const float* lik, const float* g_lik, float bpp_coef, const float* g_z_hat,
                          float* g_z, float* const g_matrix[5], float* const g_bias[5], float* const g_factor[4],
                          int64_t B, int64_t C, int64_t S, float lik_bound, void* stream) {
  return eb_bwd_impl(z, noise, RngArg{nullptr, 0}, matrix, bias, factor, quantiles, lik, g_lik, bpp_coef, g_z_hat, g_z,
                     g_matrix, g_bias, g_factor, B, C, S, lik_bound, stream);
}

extern "C" int clc_eb_bwd_rng(const float* z, const uint64_t* rng_state, uint64_t rng_offset,
                              const float* const matrix[5], const float* const bias[5], const float* const factor[4],
                              const float* quantiles, const float* lik, const float* g_lik, float bpp_coef,
                              const float* g_z_hat, float* g_z, float* const g_matrix[5], float* const g_bias[5],
                              float* const g_factor[4], int64_t B, int64_t C, int64_t S, float lik_bound, void* stream) {
  if (!rng_state) return CLC_ERR_INVALID_ARGUMENT;
  return eb_bwd_impl(z, nullptr, RngArg{reinterpret_cast<const unsigned long long*>(rng_state), rng_offset}, matrix, bias,
                     factor, quantiles, lik, g_lik, bpp_coef, g_z_hat, g_z, g_matrix, g_bias, g_factor, B, C, S, lik_bound,
                     stream);
}

extern "C" int clc_log2_sum_fwd(const float* lik, int64_t n, double* log2_sum, void* stream) {
  if (!lik || !log2_sum || n < 0) return CLC_ERR_INVALID_ARGUMENT;
  if (n == 0) return CLC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (aligned16(lik)) log2_sum_kernel<true><<<grid_for(n / 4 + 1, 256, 4), 256, 0, st>>>(lik, n, log2_sum);
  else log2_sum_kernel<false><<<grid_for(n, 256, 4), 256, 0, st>>>(lik, n, log2_sum);
  CLC_CHECK_LAUNCH("clc_log2_sum_fwd");
  return CLC_OK;
}

extern "C" int clc_log2_sum_bwd(const float* lik, float coef, const double* coef_dev, float* g_lik,
                                int64_t n, void* stream) {
  if (!lik || !g_lik || n < 0) return CLC_ERR_INVALID_ARGUMENT;
  if (n == 0) return CLC_OK;
  log2_sum_bwd_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(lik, coef, coef_dev, g_lik, n);
  CLC_CHECK_LAUNCH("clc_log2_sum_bwd");
  return CLC_OK;
}
