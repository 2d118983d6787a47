// Reference retrieval, the search half: the reference finds the n_refs nearest dictionary
// entries of every training sample with sklearn's ball-tree NearestNeighbors on the host, one sample at a time
// inside Dataset.__getitem__ (dataloader_ref_cluster.py:64 fit, :162 kneighbors) -- which is what forces
// num_workers = 0 in its run scripts.  The dictionary is small (n_clusters = 3 000 ResNet-50 features of 2 048
// floats = 24.6 MB, train_CLC.py:366), so the B200 version is a brute-force scan: one pass over the dictionary
// from HBM per batch of queries, exact squared Euclidean distances (formed as sum (x - y)^2, no |x|^2 + |y|^2 -
// 2xy cancellation), then the row-wise top-k kernel of the match stage on the negated distances.
//   queries live in shared memory (QT = 8 per pass, 64 KB at D = 2 048); a WARP owns a dictionary row, lanes run
//   over float4 chunks of the feature dimension; HBM-bound: N*D*4 bytes per pass.
#include "common.cuh"

namespace clc {

constexpr int kKnnQT = 8;     // queries per pass

__global__ void __launch_bounds__(256)
knn_neg_sqdist_kernel(const float* __restrict__ X, const float* __restrict__ Y, int Q, int N, int D,
                      float* __restrict__ out) {
  extern __shared__ float4 xs4[];                  // [QT][D/4]
  const int q0 = blockIdx.y * kKnnQT;
  const int nq = min(kKnnQT, Q - q0);
  const int d4 = D >> 2;
  for (int i = threadIdx.x; i < kKnnQT * d4; i += 256) {
    const int q = i / d4, j = i - q * d4;
    xs4[i] = q < nq ? ld4(X + (int64_t)(q0 + q) * D + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int row = blockIdx.x * 8 + warp; row < N; row += gridDim.x * 8) {
    const float4* yr = reinterpret_cast<const float4*>(Y + (int64_t)row * D);
    float acc[kKnnQT];
#pragma unroll
    for (int q = 0; q < kKnnQT; ++q) acc[q] = 0.f;
    for (int j0 = lane; j0 < d4; j0 += 128) {      // four 16-byte row loads in flight per lane
      float4 y[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + 32 * u;
        y[u] = j < d4 ? __ldcs(yr + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + 32 * u;
        if (j >= d4) continue;
#pragma unroll
        for (int q = 0; q < kKnnQT; ++q) {
          const float4 x = xs4[q * d4 + j];
          const float a = x.x - y[u].x, b = x.y - y[u].y, c = x.z - y[u].z, d = x.w - y[u].w;
          acc[q] = fmaf(a, a, fmaf(b, b, fmaf(c, c, fmaf(d, d, acc[q]))));
        }
      }
    }
#pragma unroll
    for (int q = 0; q < kKnnQT; ++q) {
      const float s = warp_sum(acc[q]);
      if (lane == 0 && q < nq) out[(int64_t)(q0 + q) * N + row] = -s;
    }
  }
}

}  // namespace clc

using namespace clc;

extern "C" int clc_knn_neg_sqdist(const float* queries, const float* dict, int32_t Q, int32_t N, int32_t D,
                                  float* neg_d2, void* stream) {
  if (!queries || !dict || !neg_d2 || Q < 0 || N < 1 || D < 1) return CLC_ERR_INVALID_ARGUMENT;
  if (Q == 0) return CLC_OK;
  if (D % 4 || !aligned16(queries) || !aligned16(dict)) return CLC_ERR_UNSUPPORTED;
  const size_t smem = (size_t)kKnnQT * D * sizeof(float);
  if (smem > 200 * 1024) return CLC_ERR_UNSUPPORTED;          // D <= 6 400
  const int passes = (Q + kKnnQT - 1) / kKnnQT;
  if (passes > 65535) return CLC_ERR_UNSUPPORTED;
  if (smem > 48 * 1024)
    CLC_CUDA(cudaFuncSetAttribute(knn_neg_sqdist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  int gx = (N + 7) / 8;
  const int cap = kNumSMs * 2;                                 // persistent: every CTA stages the queries once
  if (gx > cap) gx = cap;
  knn_neg_sqdist_kernel<<<dim3((unsigned)gx, (unsigned)passes), 256, smem, (cudaStream_t)stream>>>(queries, dict, Q, N, D,
                                                                                                   neg_d2);
  CLC_CHECK_LAUNCH("clc_knn_neg_sqdist");
  return CLC_OK;
}
