// Two MMA-issuing threads (different warps) alternating groups of MMAs into the same accumulator, each with
// the per-group shared-memory counter poll + tcgen05.commit of the real kernel: does a second issuer hide
// the ~300-cycle sync cost that a single issuer exposes?
#include <cstdio>
#include <cstdlib>
#include "ptx.cuh"
using namespace plume;

__device__ __forceinline__ void wait_counter(const volatile uint32_t* ctr, uint32_t need) {
  if (*ctr >= need) return;
  while (*ctr < need) {}
}

template <int N, int GROUP, int ISSUERS>
__global__ void __launch_bounds__(256, 1) bench(int groups, long long* out) {
  constexpr int kHaloBytes = 18432;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  __shared__ uint64_t bars[40];
  __shared__ uint32_t tmem_ptr;
  __shared__ volatile uint32_t ctr[4];
  __shared__ long long tstart[2], tend[2];
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += 256)
    reinterpret_cast<uint4*>(smem_raw + (sbase - raw))[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 40; ++i) mbar_init(smem_u32(&bars[i]), 1);
    ctr[0] = ctr[1] = ctr[2] = 0x7fffffff;
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_ptr), 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_ptr;
  const int me = (warp == 7) ? 0 : ((warp == 6) ? 1 : -1);
  if (me >= 0 && me < ISSUERS && (threadIdx.x & 31) == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    constexpr uint32_t hi = umma_desc_hi_sw128(1024);
    const uint32_t a_lo_base = umma_desc_lo(sbase, 16);
    const uint32_t b_lo_base = umma_desc_lo(sbase + 6 * kHaloBytes, 16);
    tstart[me] = clock64();
    int sa = 0;
    for (int g = 0; g < groups; ++g) {
      if (++sa == 6) sa = 0;
      if ((g % ISSUERS) != me) continue;
      wait_counter(ctr + 0, g + 1);
      tc_fence_after();
      const uint32_t a_lo = a_lo_base + sa * (kHaloBytes >> 4);
      const uint32_t b_lo = b_lo_base + (g % 3) * (N * 128 >> 4);
#pragma unroll
      for (int j = 0; j < GROUP; ++j)
        umma_bf16_lohi(tm, a_lo + (j / 4) * 64 + 2 * (j % 4), hi, b_lo + 2 * (j % 4), hi, idesc, 1);
      umma_commit(smem_u32(&bars[sa]));
    }
    umma_commit(smem_u32(&bars[30 + me]));
    mbar_wait(smem_u32(&bars[30 + me]), 0, 1, nullptr);
    tend[me] = clock64();
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x == 0) {
    long long t0 = tstart[0], t1 = tend[0];
    if (ISSUERS == 2) { t0 = min(t0, tstart[1]); t1 = max(t1, tend[1]); }
    out[blockIdx.x] = t1 - t0;
  }
  if (warp == 0) tmem_dealloc(tm, 512);
}

template <int N, int GROUP, int ISSUERS>
void run(long long* d_out) {
  const int smem = 200 * 1024 + 1024;
  const int groups = 2400 / GROUP * 4;
  cudaFuncSetAttribute(bench<N, GROUP, ISSUERS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int r = 0; r < 2; ++r) bench<N, GROUP, ISSUERS><<<148, 256, smem>>>(groups, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); exit(1); }
  long long h[148];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("N=%3d group=%2d issuers=%d : %6.1f cycles/mma (ideal %d)\n", N, GROUP, ISSUERS,
         double(mx) / (double(groups) * GROUP), N == 64 ? 48 : (N == 128 ? 64 : 128));
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, sizeof(long long) * 148);
  run<64, 12, 1>(d_out);
  run<64, 12, 2>(d_out);
  run<128, 12, 1>(d_out);
  run<128, 12, 2>(d_out);
  run<128, 4, 1>(d_out);
  run<128, 4, 2>(d_out);
  run<256, 4, 1>(d_out);
  run<256, 4, 2>(d_out);
  return 0;
}
