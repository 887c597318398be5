// Stand-alone check of the CTA-pair (cta_group::2) mechanisms before they go into a production kernel:
// cluster launch (2,1,1), tcgen05.alloc.cta_group::2, TMA loads of both CTAs completing on the LEADER's mbarrier,
// one tcgen05.mma.cta_group::2 tile D[256 x 256] = A[256 x K] * B[256 x K]^T (bf16, K = 64), commit multicast to both
// CTAs, each CTA reading its 128 rows of D from its own TMEM.  Compared with a CPU matmul.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I kcl_ltss_bioatm_b200/csrc scripts/test_2cta.cu -o scripts/test_2cta -lcuda
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "ptx.cuh"
#include "tmap.cuh"
namespace plume {
void set_error(const std::string& m) { fprintf(stderr, "error: %s\n", m.c_str()); }
}  // namespace plume
using namespace plume;

__device__ int g_dbg = 0;

template <int KB, int N>   // K = 64 * KB; D is 256 x N
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
    pair_gemm(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, float* __restrict__ D,
              int iters, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  __shared__ uint64_t bar_full[KB], bar_tfull;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  constexpr int A_BYTES = 128 * 128, B_BYTES = (N / 2) * 128;   // per CTA: 128 rows of A, half of the N rows of B
  if (threadIdx.x == 0) {
    for (int k = 0; k < KB; ++k) mbar_init(smem_u32(&bar_full[k]), 1);
    mbar_init(smem_u32(&bar_tfull), 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(smem_u32(&tmem_ptr), 256);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tm = tmem_ptr;
  long long t0 = clock64();
  if (warp == 0 && lane == 0) {
    for (int k = 0; k < KB; ++k) {
      const uint32_t a = sbase + k * (A_BYTES + B_BYTES), b = a + A_BYTES;
      if (rank == 0) mbar_expect_tx(smem_u32(&bar_full[k]), 2 * (A_BYTES + B_BYTES));   // both CTAs' bytes
      tma_load_2d_2cta(a, &mapA, smem_u32(&bar_full[k]), k * 64, rank * 128);
      tma_load_2d_2cta(b, &mapB, smem_u32(&bar_full[k]), k * 64, rank * (N / 2));
    }
  }
  if (warp == 2 && lane == 0 && rank == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(256, N, 0, 0);
    constexpr uint32_t hi = umma_desc_hi_sw128(1024);
    for (int it = 0; it < iters; ++it) {
      for (int k = 0; k < KB; ++k) {
        if (it == 0) mbar_wait(smem_u32(&bar_full[k]), 0, 1, &g_dbg);
        tc_fence_after();
        const uint32_t a_lo = umma_desc_lo(sbase + k * (A_BYTES + B_BYTES), 16);
        const uint32_t b_lo = a_lo + (A_BYTES >> 4);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          umma_bf16_lohi_2cta(tm, a_lo + 2 * j, hi, b_lo + 2 * j, hi, idesc, (it | k | j) != 0 ? 1u : 0u);
      }
    }
    umma_commit_2cta(smem_u32(&bar_tfull), 3u);
  }
  if (warp < 4) {
    mbar_wait(smem_u32(&bar_tfull), 0, 2, &g_dbg);
    tc_fence_after();
    if (threadIdx.x == 0 && cycles) cycles[rank] = clock64() - t0;
    const int row = rank * 128 + warp * 32 + lane;
    for (int c = 0; c < N; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tm + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) D[row * 256 + c + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2cta(tm, 256);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

template <int N>
static int run(const std::vector<float>& Af, const std::vector<float>& Bf, __nv_bfloat16* dA, __nv_bfloat16* dB, float* dD,
               long long* dC) {
  constexpr int KB = 2, K = 64 * KB;
  CUtensorMap mapA, mapB;
  if (make_mat_map(&mapA, dA, 256, K, 64, 128) || make_mat_map(&mapB, dB, 256, K, 64, N / 2)) return 1;
  const int smem = KB * (16384 + (N / 2) * 128) + 1024;
  cudaFuncSetAttribute(pair_gemm<KB, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaMemset(dD, 0xff, 256 * 256 * 4);
  pair_gemm<KB, N><<<2, 192, smem>>>(mapA, mapB, dD, 1, dC);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    int dbg = 0;
    cudaMemcpyFromSymbol(&dbg, g_dbg, 4);
    printf("N=%d kernel failed: %s (watchdog tag 0x%x)\n", N, cudaGetErrorString(e), dbg);
    return 2;
  }
  std::vector<float> D(256 * 256);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double worst = 0;
  int bad = 0;
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < N; ++n) {
      double r = 0;
      for (int k = 0; k < K; ++k) r += double(Af[m * K + k]) * Bf[n * K + k];
      const double d = fabs(D[m * 256 + n] - r);
      if (!(d <= 1e-3)) ++bad;
      worst = d > worst ? d : worst;
    }
  printf("pair GEMM 256x%dx%d: max |err| %.3e, %d mismatching elements -> %s\n", N, K, worst, bad, bad ? "FAIL" : "OK");
  if (bad) return 3;
  long long h[2];
  pair_gemm<KB, N><<<148, 192, smem>>>(mapA, mapB, dD, 2000, dC);
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("timing run failed: %s\n", cudaGetErrorString(e)); return 4; }
  cudaMemcpy(h, dC, 16, cudaMemcpyDeviceToHost);
  printf("  74 pairs, 2000 x %d MMAs per pair: %.1f cycles per M=256,N=%d,K=16 pair MMA (single CTA, M=128: %d)\n", 4 * KB,
         double(h[0]) / (2000 * 4.0 * KB), N, N == 64 ? 48 : (N == 128 ? 64 : 128));
  return 0;
}

int main() {
  constexpr int K = 128;
  std::vector<__nv_bfloat16> A(256 * K), B(256 * K);
  std::vector<float> Af(256 * K), Bf(256 * K);
  srand(1);
  for (int i = 0; i < 256 * K; ++i) {
    Af[i] = bf((rand() % 2001 - 1000) / 1000.0f);
    Bf[i] = bf((rand() % 2001 - 1000) / 1000.0f);
    A[i] = __float2bfloat16(Af[i]);
    B[i] = __float2bfloat16(Bf[i]);
  }
  __nv_bfloat16 *dA, *dB;
  float* dD;
  long long* dC;
  cudaMalloc(&dA, A.size() * 2);
  cudaMalloc(&dB, B.size() * 2);
  cudaMalloc(&dD, 256 * 256 * 4);
  cudaMalloc(&dC, 16);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  int rc = run<256>(Af, Bf, dA, dB, dD, dC);
  if (!rc) rc = run<128>(Af, Bf, dA, dB, dD, dC);
  if (!rc) rc = run<64>(Af, Bf, dA, dB, dD, dC);
  return rc;
}
