// Which part of a real issue loop slows tcgen05.mma issue?  Variants add, one at a time, what the conv
// kernel does around its MMAs: commit per group, fence per group, a (satisfied) mbarrier wait per group,
// run-time descriptor arithmetic.  N = 64 (48 cycles/MMA when the tensor pipe is the limit).
#include <cstdio>
#include <cstdlib>
#include "ptx.cuh"
using namespace plume;

template <int N, int GROUP, int VARIANT>
__global__ void __launch_bounds__(128, 1) bench(int iters, int rt_zero, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  __shared__ uint64_t bar, bar_grp[4], bar_self;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (6 * 18432 + 73728) / 16; i += 128)
    reinterpret_cast<uint4*>(smem_raw + (sbase - raw))[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_init(smem_u32(&bar_self), 1);
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bar_grp[i]), 1);
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_ptr), 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_ptr;
  if (threadIdx.x == 32) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    constexpr uint32_t hi = umma_desc_hi_sw128(1024);
    const uint32_t a_base = sbase, b_base = sbase + 6 * 18432;
    int sa = rt_zero, self_phase = 0;
    uint32_t accumulate = rt_zero;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      // one "slot": GROUP MMAs
      uint32_t early_ok = 1;
      if (VARIANT & 16) {  // early non-blocking probe of the (already satisfied) barrier, consumed after the MMAs
        mbar_arrive(smem_u32(&bar_self));
        asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}"
                     : "=r"(early_ok) : "r"(smem_u32(&bar_self)), "r"(self_phase) : "memory");
      }
      if (VARIANT & 32) {  // same with try_wait
        mbar_arrive(smem_u32(&bar_self));
        early_ok = mbar_try_wait(smem_u32(&bar_self), self_phase);
      }
      if (VARIANT & 4) {  // a satisfied mbarrier wait (arrive by myself first)
        mbar_arrive(smem_u32(&bar_self));
        mbar_wait(smem_u32(&bar_self), self_phase, 1, nullptr);
        self_phase ^= 1;
      }
      if (VARIANT & 2) tc_fence_after();
      uint32_t a_lo, b_lo;
      if (VARIANT & 8) {  // run-time descriptor arithmetic as in the kernel
        a_lo = umma_desc_lo(a_base + sa * 18432, 16);
        b_lo = umma_desc_lo(b_base + (sa % 3) * N * 128, 16);
      } else {
        a_lo = umma_desc_lo(a_base, 16);
        b_lo = umma_desc_lo(b_base, 16);
      }
#pragma unroll
      for (int j = 0; j < GROUP; ++j) {
        umma_bf16_lohi(tm, a_lo + (j / 4) * 64 + 2 * (j % 4), hi, b_lo + 2 * (j % 4), hi, idesc, accumulate);
        accumulate = 1;
      }
      if (VARIANT & 1) umma_commit(smem_u32(&bar_grp[it & 3]));
      if (VARIANT & 48) {
        if (!early_ok) mbar_wait(smem_u32(&bar_self), self_phase, 1, nullptr);
        self_phase ^= 1;
      }
      if (++sa == 6) sa = 0;
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0, 1, nullptr);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

template <int N, int GROUP, int VARIANT>
void run(long long* d_out) {
  const int smem = 6 * 18432 + 73728 + 1024;
  const int iters = 4000 / GROUP * 4;
  cudaFuncSetAttribute(bench<N, GROUP, VARIANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  bench<N, GROUP, VARIANT><<<148, 128, smem>>>(iters, 0, d_out);
  bench<N, GROUP, VARIANT><<<148, 128, smem>>>(iters, 0, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); exit(1); }
  long long h[148];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("N=%3d group=%2d commit=%d fence=%d wait=%d rtdesc=%d early_test=%d early_try=%d : %6.1f cycles/mma\n", N, GROUP,
         VARIANT & 1, (VARIANT >> 1) & 1, (VARIANT >> 2) & 1, (VARIANT >> 3) & 1, (VARIANT >> 4) & 1, (VARIANT >> 5) & 1,
         double(mx) / (double(iters) * GROUP));
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, sizeof(long long) * 148);
  run<64, 12, 0>(d_out);
  run<64, 12, 15>(d_out);
  run<64, 12, 16 + 11>(d_out);
  run<64, 12, 32 + 11>(d_out);
  run<64, 4, 15>(d_out);
  run<64, 4, 16 + 11>(d_out);
  run<64, 4, 32 + 11>(d_out);
  run<128, 4, 15>(d_out);
  run<128, 4, 16 + 11>(d_out);
  run<128, 4, 32 + 11>(d_out);
  run<256, 4, 15>(d_out);
  run<256, 4, 16 + 11>(d_out);
  return 0;
}
