// Weight-gradient MMA pattern (M=128, N=128, K=16, both operands MN-major, 12 MMAs per pipeline stage =
// 3 vertical taps x 4 K steps): does sharing one operand across the three taps of a K step reduce the
// shared-memory operand traffic?  Variants:
//   0: tap-major (the kernel as of r1h): A = X(tap,k), B = dY(k)
//   1: k-major, same roles, no hints
//   2: roles swapped (A = dY(k), B = X(tap,k)), k-major, collector::a::fill / use / lastuse
//   3: roles swapped, k-major, no hints
//   4: roles swapped, k-major, every MMA ::fill (control: hint syntax without reuse)
// Usage: bench_mma5 [stages_per_run]
#include <cstdio>
#include <cstdlib>
#include "ptx.cuh"
using namespace plume;

#define MMA_ASM(SUFFIX)                                                                               \
  asm volatile(                                                                                       \
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"                                                   \
      "setp.ne.b32 p, %6, 0;\n\t"                                                                     \
      "mov.b64 da, {%1, %2};\n\t"                                                                     \
      "mov.b64 db, {%3, %4};\n\t"                                                                     \
      "tcgen05.mma.cta_group::1.kind::f16" SUFFIX " [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),          \
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)                         \
      : "memory")

template <int HINT>  // 0 none, 1 fill, 2 use, 3 lastuse
__device__ __forceinline__ void mma_hint(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                         uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  if constexpr (HINT == 0) MMA_ASM("");
  if constexpr (HINT == 1) MMA_ASM(".collector::a::fill");
  if constexpr (HINT == 2) MMA_ASM(".collector::a::use");
  if constexpr (HINT == 3) MMA_ASM(".collector::a::lastuse");
}

constexpr int kXBox = 10240, kStage = 2 * kXBox + 16384, kStages = 5;

template <int VARIANT>
__global__ void __launch_bounds__(128, 1) bench(int nstages, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  __shared__ uint64_t bars[8];
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kStages * kStage / 16; i += 128)
    reinterpret_cast<uint4*>(smem_raw + (sbase - raw))[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1);
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_ptr), 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_ptr;
  if (threadIdx.x == 96) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128, 1, 1);
    constexpr uint32_t hi = umma_desc_hi_sw128(1024);
    const long long t0 = clock64();
    int stage = 0;
    for (int i = 0; i < nstages; ++i) {
      const uint32_t x_addr = sbase + stage * kStage;
      const uint32_t dy_lo = umma_desc_lo(x_addr + 2 * kXBox, 8192);
      if (VARIANT == 0) {
#pragma unroll
        for (int blk = 0; blk < 3; ++blk) {
          const uint32_t x_lo = umma_desc_lo(x_addr + blk * 1024, kXBox);
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_hint<0>(tm + blk * 128, x_lo + 128 * k, hi, dy_lo + 128 * k, hi, idesc, 1);
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
          for (int blk = 0; blk < 3; ++blk) {
            const uint32_t x_lo = umma_desc_lo(x_addr + blk * 1024, kXBox) + 128 * k;
            const uint32_t d = tm + blk * 128;
            if (VARIANT == 1) mma_hint<0>(d, x_lo, hi, dy_lo + 128 * k, hi, idesc, 1);
            if (VARIANT == 2) {
              if (blk == 0) mma_hint<1>(d, dy_lo + 128 * k, hi, x_lo, hi, idesc, 1);
              if (blk == 1) mma_hint<2>(d, dy_lo + 128 * k, hi, x_lo, hi, idesc, 1);
              if (blk == 2) mma_hint<3>(d, dy_lo + 128 * k, hi, x_lo, hi, idesc, 1);
            }
            if (VARIANT == 3) mma_hint<0>(d, dy_lo + 128 * k, hi, x_lo, hi, idesc, 1);
            if (VARIANT == 4) mma_hint<1>(d, dy_lo + 128 * k, hi, x_lo, hi, idesc, 1);
          }
        }
      }
      if (++stage == kStages) stage = 0;
    }
    umma_commit(smem_u32(&bars[0]));
    mbar_wait(smem_u32(&bars[0]), 0, 1, nullptr);
    out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

template <int V>
static void run(int nstages, long long* d_out, int ctas) {
  const size_t smem = kStages * kStage + 1024;
  cudaFuncSetAttribute(bench<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  bench<V><<<ctas, 128, smem>>>(nstages, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d_out, sizeof(long long) * ctas, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < ctas; ++i) avg += h[i];
  avg /= ctas;
  printf("variant %d ctas %3d: %s  %.1f cycles per MMA (%.0f per 12-MMA stage)\n", V, ctas,
         cudaGetErrorString(e), avg / (12.0 * nstages), avg / nstages);
}

int main(int argc, char** argv) {
  const int nstages = argc > 1 ? atoi(argv[1]) : 2000;
  long long* d_out;
  cudaMalloc(&d_out, sizeof(long long) * 148);
  for (int ctas : {1, 148}) {
    run<0>(nstages, d_out, ctas);
    run<1>(nstages, d_out, ctas);
    run<2>(nstages, d_out, ctas);
    run<3>(nstages, d_out, ctas);
    run<4>(nstages, d_out, ctas);
  }
  return 0;
}
