// Micro-benchmark: issue rate of tcgen05.mma (kind::f16, bf16 -> fp32, M = 128) on sm_100a as a function
// of N and of the number of independent accumulators, with a lean, unrolled issue loop (descriptors are
// loop-invariant registers), so that the tensor pipe rather than the issuing thread is measured.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I kcl_ltss_bioatm_b200/csrc scripts/bench_mma.cu -o scripts/bench_mma
#include <cstdio>
#include <cstdlib>
#include "ptx.cuh"
using namespace plume;

// ACCS independent accumulators (column ranges), 8 distinct A slices (two 16 KB tiles), B fixed
template <int N, int ACCS, bool WARP_UNIFORM>
__global__ void __launch_bounds__(128, 1) bench(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (2 * 16384 + 32768) / 16; i += 128)
    reinterpret_cast<uint4*>(smem_raw + (sbase - raw))[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_ptr), 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_ptr;
  if (warp == 1) {
    const bool leader = WARP_UNIFORM ? (elect_one() != 0) : ((threadIdx.x & 31) == 0);
    if (WARP_UNIFORM || leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
      uint64_t ad[8], bd[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) ad[i] = umma_desc_sw128(sbase + (i >> 2) * 16384 + (i & 3) * 32, 16, 1024);
#pragma unroll
      for (int i = 0; i < 4; ++i) bd[i] = umma_desc_sw128(sbase + 2 * 16384 + i * 32, 16, 1024);
      long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (leader) umma_bf16(tm + (j % ACCS) * N, ad[j], bd[j & 3], idesc, 1);
        }
      }
      if (leader) umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), 0, 1, nullptr);
      long long t1 = clock64();
      if (leader) out[blockIdx.x] = t1 - t0;
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

template <int N, int ACCS, bool WU>
void run(int grid, int iters, long long* d_out) {
  const int smem = 2 * 16384 + 32768 + 1024;
  cudaFuncSetAttribute(bench<N, ACCS, WU>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  bench<N, ACCS, WU><<<grid, 128, smem>>>(iters, d_out);
  bench<N, ACCS, WU><<<grid, 128, smem>>>(iters, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N %d: %s\n", N, cudaGetErrorString(e)); exit(1); }
  long long h[148];
  cudaMemcpy(h, d_out, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  const double per = double(mx) / (8.0 * iters);
  printf("N=%3d accs=%d %s grid=%3d: %6.1f cycles/mma -> %5.0f MAC/clk/SM (%.0f%% of 4096)\n", N, ACCS,
         WU ? "warp-uniform" : "single-lane ", grid, per, 128.0 * N * 16 / per, 100.0 * 128.0 * N * 16 / per / 4096);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, sizeof(long long) * 148);
  run<64, 1, false>(148, 2000, d_out);
  run<64, 1, true>(148, 2000, d_out);
  run<64, 2, true>(148, 2000, d_out);
  run<64, 4, true>(148, 2000, d_out);
  run<64, 8, true>(148, 2000, d_out);
  run<128, 1, false>(148, 2000, d_out);
  run<128, 1, true>(148, 2000, d_out);
  run<128, 2, true>(148, 2000, d_out);
  run<128, 4, true>(148, 2000, d_out);
  run<256, 1, false>(148, 2000, d_out);
  run<256, 1, true>(148, 2000, d_out);
  run<256, 2, true>(148, 2000, d_out);
  return 0;
}
