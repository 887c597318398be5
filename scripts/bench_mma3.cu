// Replica of the conv3 MMA-issuer loop (resident-weight mode and ring mode) with every counter pre-satisfied:
// measures what the issuing thread's own instruction stream costs per MMA, and tries leaner variants.
#include <cstdio>
#include <cstdlib>
#include "ptx.cuh"
using namespace plume;

__device__ int g_dbg_word_b;
#ifndef POLL
#define POLL 0
#endif
__device__ __forceinline__ uint32_t ld_ctr(const volatile uint32_t* ctr) {
#if POLL == 0
  return *ctr;                                   // volatile generic load
#elif POLL == 1
  uint32_t v;                                    // ld.shared (non-generic), relaxed
  asm volatile("ld.relaxed.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32((const void*)ctr)) : "memory");
  return v;
#elif POLL == 2
  uint32_t v;                                    // plain ld.shared.volatile
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32((const void*)ctr)) : "memory");
  return v;
#elif POLL == 3
  uint32_t v;                                    // weak ld.shared, asm not volatile (may be hoisted!)
  asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32((const void*)ctr)));
  return v;
#else
  uint32_t v;                                    // weak ld.shared, volatile asm + memory clobber (never hoisted)
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32((const void*)ctr)) : "memory");
  return v;
#endif
}
__device__ __forceinline__ void wait_counter(const volatile uint32_t* ctr, uint32_t need) {
  if (ld_ctr(ctr) >= need) return;
  while (ld_ctr(ctr) < need) {}
}

// VARIANT 0: loop as in the kernel.  1: counter value prefetched one group ahead.  2: no counter polls at all.
// 3: no commits.  4: neither.
template <int N, bool RESIDENT, int VARIANT>
__global__ void __launch_bounds__(128, 1) bench(int tiles, int kbs, int a_slots, int b_slots, long long* out) {
  constexpr int B_BYTES = N * 128;
  constexpr int kHaloBytes = 18432;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  __shared__ uint64_t bars[40];
  __shared__ uint32_t tmem_ptr;
  __shared__ volatile uint32_t ctr[4];
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += 128)
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
  const uint32_t off_a = 0, off_b = a_slots * kHaloBytes;
  auto a_empty = [&](int s) { return smem_u32(&bars[s]); };
  auto b_empty = [&](int s) { return smem_u32(&bars[8 + s]); };
  auto tfull = [&](int a) { return smem_u32(&bars[16 + a]); };
  if (threadIdx.x == 96) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    constexpr uint32_t hi = umma_desc_hi_sw128(1024);
    const uint32_t a_lo_base = umma_desc_lo(sbase + off_a, 16);
    const uint32_t b_lo_base = umma_desc_lo(sbase + off_b, 16);
    constexpr uint32_t kHaloUnits = kHaloBytes >> 4, kBUnits = B_BYTES >> 4;
    const uint32_t b_tap_step = 3 * kbs * kBUnits;
    uint32_t a_cnt = 0, b_cnt = 0;
    int sa = 0, sb = 0;
    long long n_mma = 0;
    uint32_t pre = (VARIANT == 1) ? ctr[0] : 0;
    long long t0 = clock64();
    for (int it = 0; it < tiles; ++it) {
      if (VARIANT != 2 && VARIANT != 4) wait_counter(ctr + 2, it + 1);
      tc_fence_after();
      const uint32_t d_tmem = tm + (it & 1) * N;
      uint32_t accumulate = 0;
      for (int kb = 0; kb < kbs; ++kb) {
#pragma unroll 1
        for (int dwi = 0; dwi < 3; ++dwi) {
          ++a_cnt;
          if (VARIANT == 0 || VARIANT == 3) wait_counter(ctr + 0, a_cnt);
          if (VARIANT == 1) { if (pre < a_cnt) wait_counter(ctr + 0, a_cnt); pre = ctr[0]; }
          if (RESIDENT) tc_fence_after();
          const uint32_t a_lo = a_lo_base + sa * kHaloUnits;
          uint32_t b_lo = b_lo_base + (dwi * kbs + kb) * kBUnits;
#pragma unroll
          for (int dhi = 0; dhi < 3; ++dhi) {
            if (!RESIDENT) {
              ++b_cnt;
              if (VARIANT == 0 || VARIANT == 3 || VARIANT == 1) wait_counter(ctr + 1, b_cnt);
              tc_fence_after();
              b_lo = b_lo_base + sb * kBUnits;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_bf16_lohi(d_tmem, a_lo + dhi * 64 + 2 * k, hi, b_lo + 2 * k, hi, idesc, accumulate);
              accumulate = 1;
            }
            n_mma += 4;
            if (!RESIDENT) {
              if (VARIANT < 3) umma_commit(b_empty(sb));
              if (++sb == b_slots) sb = 0;
            } else {
              b_lo += b_tap_step;
            }
          }
          if (VARIANT < 3) umma_commit(a_empty(sa));
          if (++sa == a_slots) sa = 0;
        }
      }
      if (VARIANT < 3) umma_commit(tfull(it & 1));
    }
    umma_commit(smem_u32(&bars[30]));
    mbar_wait(smem_u32(&bars[30]), 0, 1, nullptr);
    long long t1 = clock64();
    out[blockIdx.x * 2] = t1 - t0;
    out[blockIdx.x * 2 + 1] = n_mma;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

template <int N, bool RESIDENT, int VARIANT>
void run(int kbs, int a_slots, int b_slots, long long* d_out) {
  const int smem = 200 * 1024 + 1024;
  cudaFuncSetAttribute(bench<N, RESIDENT, VARIANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int r = 0; r < 2; ++r) bench<N, RESIDENT, VARIANT><<<148, 128, smem>>>(200, kbs, a_slots, b_slots, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); exit(1); }
  long long h[296];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("N=%3d %s kb=%d variant=%d : %6.1f cycles/mma (ideal %d)\n", N, RESIDENT ? "resident" : "ring    ", kbs, VARIANT,
         double(h[0]) / double(h[1]), N == 64 ? 48 : (N == 128 ? 64 : 128));
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, sizeof(long long) * 296);
  printf("POLL=%d\n", POLL);
  run<64, true, 0>(1, 6, 0, d_out);
  run<64, true, 2>(1, 6, 0, d_out);
  run<128, false, 0>(2, 4, 6, d_out);
  run<256, false, 0>(4, 3, 4, d_out);
  run<256, false, 2>(4, 3, 4, d_out);
  run<256, false, 4>(4, 3, 4, d_out);
  return 0;
}
