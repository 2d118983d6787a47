// Bring-up microbenchmark (not product code): what does the sm_100a tensor pipe accept per tcgen05.mma
// from shared-memory operands, with cta_group::1 (M = 128) and cta_group::2 (M = 256), aligned and
// row-shifted B descriptors (the match kernel's implicit-GEMM trick), with and without concurrent bulk
// copies into shared memory?  Also checks the cta_group::2 operand / accumulator layout numerically.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/_bin/umma_bench scripts/umma_bench.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// 64-byte swizzle: rows of 64 B (32 bf16), 8-row groups 512 B apart, layout type 4
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <int CG>
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc, uint32_t pred) {
  if (CG == 1)
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(pred) : "memory");
  else
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                 "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(pred) : "memory");
}
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar, uint32_t pred) {
  if (CG == 1)
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
                 ::"r"(bar), "r"(pred) : "memory");
  else
    asm volatile("{\n\t.reg .pred q;\n\t.reg .b16 m;\n\tsetp.ne.b32 q, %1, 0;\n\tmov.b16 m, 3;\n\t"
                 "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}"
                 ::"r"(bar), "r"(pred) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
        "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]),
        "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]),
        "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

// Layout of the K-major SW128 operand tile: row r = 128 bytes (64 bf16), 16-byte chunk c stored at chunk c ^ (r & 7).
__device__ __forceinline__ uint32_t sw128_off(int r, int kk) {
  return (uint32_t)r * 128u + (uint32_t)((((kk >> 3) ^ (r & 7)) << 4) + (kk & 7) * 2);
}

__device__ __forceinline__ uint32_t sw64_off(int r, int kk) {   // kk in [0, 32)
  return (uint32_t)r * 64u + (uint32_t)((((kk >> 3) ^ ((r >> 1) & 3)) << 4) + (kk & 7) * 2);
}

constexpr int kARows = 128, kAStages = 4;
constexpr int kBRows = 704;   // >= 256 + 3*128 + 3 + slack
constexpr int kABytes = kARows * 128, kBBytes = kBRows * 128;
constexpr int kSmem = kAStages * kABytes + kBBytes + 1024 + 256;

struct Args {
  int iters;        // groups of (16 shifts x 4 k-steps) MMAs
  int N;            // MMA N
  int shift_mode;   // 0: no shift, 1: 8-row aligned shifts, 2: real shifts dy*W+dx (W = 128)
  int bulk;         // 1: a second warp streams 16 KB bulk copies global -> shared per 4 MMAs' worth
  int verify;       // 1: integer test pattern, dump D
  int vshift;       // verification: B descriptor row shift
  int sw64;         // verification: 64-byte swizzle, K = 32 per tile (2 MMAs)
  const uint8_t* gsrc;
  float* dump;      // [CG*128][N]
  long long* cycles;  // per CTA
};

// mode of the issuing loop is fully warp-uniform: every lane runs the loop, one elected lane issues.
template <int CG>
__global__ void __launch_bounds__(192, 1) bench_kernel(const Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const uint32_t sA = base, sB = base + kAStages * kABytes;
  const uint32_t sBar = sB + kBBytes;       // [0] mma done, [1..4] bulk full
  const uint32_t sTmem = sBar + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;

  // ---- fill operands ----
  if (a.verify) {
    // A[m][kk] = ((m_glob * 3 + kk) % 7) - 3 ; B[n][kk] = ((n_glob + 2 * kk) % 5) - 2   (exact in bf16 / fp32)
    for (int i = threadIdx.x; i < kARows * 64; i += blockDim.x) {
      const int r = i >> 6, kk = i & 63;
      const int mg = (int)rank * 128 + r;
      if (a.sw64) { if (kk < 32) *reinterpret_cast<__nv_bfloat16*>(sm + sw64_off(r, kk)) = __float2bfloat16((float)((mg * 3 + kk) % 7 - 3)); }
      else *reinterpret_cast<__nv_bfloat16*>(sm + sw128_off(r, kk)) = __float2bfloat16((float)((mg * 3 + kk) % 7 - 3));
    }
    for (int i = threadIdx.x; i < kBRows * 64; i += blockDim.x) {
      const int r = i >> 6, kk = i & 63;
      // CTA `rank` holds the B rows of its half of N: global row = rank * (N / CG) + r
      const int ng = (int)rank * (a.N / CG) + r;
      if (a.sw64) { if (kk < 32) *reinterpret_cast<__nv_bfloat16*>(sm + kAStages * kABytes + sw64_off(r, kk)) = __float2bfloat16((float)((ng + 2 * kk) % 5 - 2)); }
      else *reinterpret_cast<__nv_bfloat16*>(sm + kAStages * kABytes + sw128_off(r, kk)) =
          __float2bfloat16((float)((ng + 2 * kk) % 5 - 2));
    }
  } else {
    uint32_t s = 1234567u + blockIdx.x * 7919u + threadIdx.x;
    for (int i = threadIdx.x; i < (kAStages * kABytes + kBBytes) / 4; i += blockDim.x) {
      s = s * 1664525u + 1013904223u;
      // two bf16 in [-2, 2): sign + exponent 0x3f/0x3e.. keep it simple: mantissa random, exponent 126..127
      const uint32_t lo = 0x3f00u | ((s >> 8) & 0xffu) | ((s & 1u) << 15);
      const uint32_t hi = 0x3f00u | ((s >> 16) & 0xffu) | ((s & 2u) << 14);
      reinterpret_cast<uint32_t*>(sm)[i] = lo | (hi << 16);
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

  if (warp == 0) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sTmem), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sTmem), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  } else if (warp == 1 && lane == 0) {
    mbar_init(sBar, 1);
    for (int i = 1; i <= 4; ++i) mbar_init(sBar + 8u * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(sm + (sTmem - base));

  long long t0 = 0, t1 = 0;
  if (warp == 1 && rank == 0) {
    // ===== MMA issuer: warp-uniform loop, elected lane issues =====
    const uint32_t pred = elect_one();
    const uint32_t idesc = make_idesc(128 * CG, a.N);
    const uint64_t adesc0 = make_desc_sw128(sA), bdesc0 = make_desc_sw128(sB);
    t0 = clock64();
    if (a.verify && a.sw64) {
      const uint64_t a64 = make_desc_sw64(sA), b64 = make_desc_sw64(sB);
#pragma unroll
      for (int k = 0; k < 2; ++k)
        umma<CG>(tmem_base, a64 + 2u * k, b64 + (uint64_t)(a.vshift * 4) + 2u * k, idesc, k ? 1u : 0u, pred);
    } else if (a.verify) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma<CG>(tmem_base, adesc0 + 2u * k, bdesc0 + (uint64_t)(a.vshift * 8) + 2u * k, idesc, k ? 1u : 0u, pred);
    } else {
      for (int it = 0; it < a.iters; ++it) {
        int dy = 0, dx = 0;
        for (int s = 0; s < 16; ++s) {
          int sh = 0;
          if (a.shift_mode == 1) sh = dy * 128 + dx * 8;
          else if (a.shift_mode == 2) sh = dy * 128 + dx;
          const uint64_t ad = adesc0 + (uint64_t)((s & (kAStages - 1)) * (kABytes >> 4));
          const uint64_t bd = bdesc0 + (uint64_t)(sh * 8);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma<CG>(tmem_base, ad + 2u * k, bd + 2u * k, idesc, (it | s | k) ? 1u : 0u, pred);
          if (++dx == 4) { dx = 0; ++dy; }
        }
      }
    }
    umma_commit<CG>(sBar, pred);
    __syncwarp();
  } else if (warp == 2 && a.bulk && !a.verify) {
    // ===== bulk-copy traffic: 16 KB per 4 MMAs into the A stages (contents irrelevant) =====
    if (lane == 0) {
      uint32_t ph[4] = {0, 0, 0, 0};
      const int total = a.iters * 16;
      for (int i = 0; i < total; ++i) {
        const int st = i & 3;
        const uint32_t bar = sBar + 8u * (1 + st);
        if (i >= 4) { mbar_wait(bar, ph[st]); ph[st] ^= 1u; }
        mbar_expect_tx(bar, kABytes);
        const uint8_t* src = a.gsrc + ((size_t)((blockIdx.x * 64 + (i & 63)) & 4095)) * kABytes;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(sA + st * kABytes), "l"(src), "r"((uint32_t)kABytes), "r"(bar) : "memory");
      }
      for (int st = 0; st < 4; ++st) { mbar_wait(sBar + 8u * (1 + st), ph[st]); }
    }
  }
  // everyone (both CTAs of a pair) waits for the MMAs
  mbar_wait(sBar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 1 && rank == 0 && lane == 0) {
    t1 = clock64();
    a.cycles[blockIdx.x] = t1 - t0;
  }
  if (a.verify && warp >= 2) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    for (int c0 = 0; c0 < a.N; c0 += 32) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      for (int j = 0; j < 32; ++j) a.dump[((size_t)rank * 128 + row) * a.N + c0 + j] = v[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 0) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <int CG>
static void launch(const Args& a, int grid, cudaStream_t st) {
  auto kern = bench_kernel<CG>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = kSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, kern, a));
}

template <int CG>
static bool verify(int N, int vshift, int sw64 = 0) {
  Args a = {};
  a.N = N; a.verify = 1; a.vshift = vshift; a.sw64 = sw64;
  const int KK = sw64 ? 32 : 64;
  const int M = 128 * CG;
  CK(cudaMalloc(&a.dump, sizeof(float) * M * N));
  CK(cudaMemset(a.dump, 0xff, sizeof(float) * M * N));
  CK(cudaMalloc(&a.cycles, sizeof(long long) * 8));
  launch<CG>(a, CG, 0);
  CK(cudaDeviceSynchronize());
  std::vector<float> d(M * N);
  CK(cudaMemcpy(d.data(), a.dump, sizeof(float) * M * N, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      // column n of the pair's tile: CTA (n / (N/CG)) row (n % (N/CG)) + vshift of that CTA's B buffer
      const int half = N / CG, owner = n / half, rl = n % half + vshift;
      const int ng = owner * half + rl;
      float ref = 0.f;
      for (int kk = 0; kk < KK; ++kk) ref += (float)((m * 3 + kk) % 7 - 3) * (float)((ng + 2 * kk) % 5 - 2);
      if (d[m * N + n] != ref) {
        if (bad < 5) printf("  mismatch CG=%d m=%d n=%d got %g want %g\n", CG, m, n, d[m * N + n], ref);
        ++bad;
      }
    }
  printf("verify cta_group::%d N=%d shift=%d sw64=%d : %s (%d mismatches)\n", CG, N, vshift, sw64, bad ? "FAIL" : "ok", bad);
  cudaFree(a.dump); cudaFree(a.cycles);
  return bad == 0;
}

template <int CG>
static void bench(int N, int shift_mode, int bulk, int iters, const uint8_t* gsrc) {
  Args a = {};
  a.N = N; a.iters = iters; a.shift_mode = shift_mode; a.bulk = bulk; a.gsrc = gsrc;
  CK(cudaMalloc(&a.cycles, sizeof(long long) * 148));
  CK(cudaMemset(a.cycles, 0, sizeof(long long) * 148));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0));
    launch<CG>(a, 148, 0);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  std::vector<long long> cyc(148);
  CK(cudaMemcpy(cyc.data(), a.cycles, sizeof(long long) * 148, cudaMemcpyDeviceToHost));
  double sum = 0; int cnt = 0;
  for (int i = 0; i < 148; ++i) if (cyc[i] > 0) { sum += (double)cyc[i]; ++cnt; }
  const double mmas = (double)iters * 64;                      // per issuing CTA
  const double flop = 2.0 * 128 * CG * N * 16 * mmas * cnt;    // cnt issuers, each M = 128*CG
  printf("{\"cta_group\": %d, \"M\": %d, \"N\": %d, \"shift_mode\": %d, \"bulk\": %d, \"cycles_per_mma\": %.1f, "
         "\"ms\": %.4f, \"tflops\": %.1f, \"issuers\": %d, \"eff_clock_mhz\": %.0f}\n",
         CG, 128 * CG, N, shift_mode, bulk, sum / cnt / mmas, best, flop / (best * 1e-3) / 1e12, cnt,
         sum / cnt / (best * 1e-3) / 1e6);
  cudaFree(a.cycles);
}

int main() {
  bool ok = true;
  ok &= verify<1>(256, 0);
  ok &= verify<1>(256, 5);
  ok &= verify<2>(256, 0);
  ok &= verify<2>(256, 5);
  ok &= verify<2>(256, 131);
  ok &= verify<2>(128, 3);
  ok &= verify<1>(256, 0, 1);
  ok &= verify<1>(256, 1, 1);
  ok &= verify<1>(256, 5, 1);
  ok &= verify<1>(256, 131, 1);
  ok &= verify<1>(256, 387, 1);
  uint8_t* gsrc;
  CK(cudaMalloc(&gsrc, (size_t)4096 * kABytes));
  CK(cudaMemset(gsrc, 0x3c, (size_t)4096 * kABytes));
  const int iters = 400;
  for (int bulk = 0; bulk < 2; ++bulk)
    for (int sm = 0; sm < 3; ++sm) {
      bench<1>(256, sm, bulk, iters, gsrc);
      bench<2>(256, sm, bulk, iters, gsrc);
    }
  bench<1>(128, 2, 1, iters, gsrc);
  bench<2>(128, 2, 1, iters, gsrc);
  bench<1>(64, 2, 0, iters, gsrc);
  printf("%s\n", ok ? "ALL VERIFY OK" : "VERIFY FAILED");
  return 0;
}
