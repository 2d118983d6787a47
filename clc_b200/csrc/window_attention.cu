// Fused (shifted-)window multi-head self-attention core of the Swin blocks (the WMSA inside the
// per-slice SWAtten parameter networks, CLC_run.py:399-409, and inside g_a / g_s / h_a / h_s; class WMSA,
// CLC_run.py:107-193 == tcm.py).  The reference runs, per block: torch.roll, window partition (view + permute +
// reshape copy), the qkv Linear, three permuted slices, matmul, scale, + relative-position bias (an advanced-
// indexing gather of the table per call), masked_fill with a 6-D mask it allocates per call, softmax, matmul,
// permute + reshape copy, the output Linear, window reverse and torch.roll back -- ~14 kernels and ~8 full passes
// over the activations.  Here the two Linear layers stay cuBLAS and everything between them is ONE kernel that reads
// qkv [B, H, W, 3C] in place (token <-> pixel index arithmetic replaces roll / partition / reverse) and writes
// out [B, H, W, C]:
//   CTA = (window, head, image), one thread per query token of the 8 x 8 window; K and V of the window in shared
//   memory (broadcast reads), scores in shared memory transposed ([key][query]: conflict-free), the relative-position
//   bias looked up in the head's 15 x 15 table (staged in shared memory), the shift mask evaluated from coordinates.
// fp32 throughout; same operation order as the reference per element (dot, * scale, + bias, mask, softmax with
// max subtraction, weighted sum).  Forward only (inference); training keeps the torch path.
#include "common.cuh"

namespace clc {

constexpr int kWin = 8, kTok = kWin * kWin;

template <int HD>
__global__ void __launch_bounds__(kTok)
window_attention_kernel(const float* __restrict__ qkv, const float* __restrict__ rel_table, float* __restrict__ out,
                        int H, int W, int C, int heads, int shifted, float scale) {
  __shared__ __align__(16) float Ks[kTok][HD];
  __shared__ __align__(16) float Vs[kTok][HD];
  __shared__ float Ss[kTok][kTok + 1];                 // [key][query]
  __shared__ float tab[(2 * kWin - 1) * (2 * kWin - 1)];
  const int nw = W / kWin, nh = H / kWin;
  const int win = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int wy = win / nw, wx = win - wy * nw;
  const int i = threadIdx.x, yi = i >> 3, xi = i & 7;
  const int sh = shifted ? kWin / 2 : 0;
  int py = wy * kWin + yi + sh, px = wx * kWin + xi + sh;     // pixel of this token in the un-rolled image
  if (py >= H) py -= H;
  if (px >= W) px -= W;
  const float* tok = qkv + (((int64_t)b * H + py) * W + px) * (3 * C);
  float q[HD];
#pragma unroll
  for (int d = 0; d < HD; d += 4) {
    const float4 a = ld4(tok + head * HD + d);
    q[d] = a.x; q[d + 1] = a.y; q[d + 2] = a.z; q[d + 3] = a.w;
    *reinterpret_cast<float4*>(&Ks[i][d]) = ld4(tok + (heads + head) * HD + d);
    *reinterpret_cast<float4*>(&Vs[i][d]) = ld4(tok + (2 * heads + head) * HD + d);
  }
  for (int t = i; t < (2 * kWin - 1) * (2 * kWin - 1); t += kTok) tab[t] = rel_table[head * (2 * kWin - 1) * (2 * kWin - 1) + t];
  __syncthreads();
  // shift mask (tcm.py generate_mask): in the last window row / column, tokens from the two sides of the wrap
  // do not attend to each other
  const int s_ = kWin - kWin / 2;
  const bool last_r = shifted && wy == nh - 1, last_c = shifted && wx == nw - 1;
  float mx = -INFINITY;
  for (int j = 0; j < kTok; ++j) {
    const int yj = j >> 3, xj = j & 7;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) s = fmaf(q[d], Ks[j][d], s);
    s = s * scale + tab[(yi - yj + kWin - 1) * (2 * kWin - 1) + (xi - xj + kWin - 1)];
    const bool masked = (last_r && ((yi < s_) != (yj < s_))) || (last_c && ((xi < s_) != (xj < s_)));
    s = masked ? -INFINITY : s;
    Ss[j][i] = s;
    mx = fmaxf(mx, s);
  }
  float den = 0.f;
  for (int j = 0; j < kTok; ++j) {
    const float e = expf(Ss[j][i] - mx);               // (a token always attends to itself: mx is finite)
    Ss[j][i] = e;
    den += e;
  }
  const float inv = 1.0f / den;
  float acc[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) acc[d] = 0.f;
  for (int j = 0; j < kTok; ++j) {
    const float p = Ss[j][i] * inv;
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] = fmaf(p, Vs[j][d], acc[d]);
  }
  float* o = out + (((int64_t)b * H + py) * W + px) * C + head * HD;
#pragma unroll
  for (int d = 0; d < HD; d += 4) st4(o + d, make_float4(acc[d], acc[d + 1], acc[d + 2], acc[d + 3]));
}

}  // namespace clc

using namespace clc;

extern "C" int clc_window_attention_fwd(const float* qkv, const float* rel_table, float* out, int64_t B, int32_t H,
                                        int32_t W, int32_t C, int32_t head_dim, int32_t window, int32_t shifted,
                                        float scale, void* stream) {
  if (!qkv || !rel_table || !out || B < 0 || H < 1 || W < 1 || C < 1 || head_dim < 1) return CLC_ERR_INVALID_ARGUMENT;
  if (B == 0) return CLC_OK;
  if (window != kWin || H % kWin || W % kWin || C % head_dim || B > 65535 || C / head_dim > 65535 ||
      !aligned16(qkv) || !aligned16(out))
    return CLC_ERR_UNSUPPORTED;
  const int heads = C / head_dim;
  dim3 grid((unsigned)((H / kWin) * (W / kWin)), (unsigned)heads, (unsigned)B);
  cudaStream_t st = (cudaStream_t)stream;
  switch (head_dim) {
    case 8: window_attention_kernel<8><<<grid, kTok, 0, st>>>(qkv, rel_table, out, H, W, C, heads, shifted, scale); break;
    case 16: window_attention_kernel<16><<<grid, kTok, 0, st>>>(qkv, rel_table, out, H, W, C, heads, shifted, scale); break;
    case 32: window_attention_kernel<32><<<grid, kTok, 0, st>>>(qkv, rel_table, out, H, W, C, heads, shifted, scale); break;
    default: return CLC_ERR_UNSUPPORTED;
  }
  CLC_CHECK_LAUNCH("clc_window_attention_fwd");
  return CLC_OK;
}
