// HBM-bandwidth kernels of the UNet hot path: BatchNorm apply / backward, 2x2 max pool (+argmax),
// 1x1 head + sigmoid + BCE/Dice, fused Adam, weight packing, tile cut / stitch.
// All activation traffic is 128-bit vectorised (8 bf16), coalesced along the NHWC channel dimension;
// grids are sized in multiples of the SM count and walk the data with a grid-stride loop in which a
// thread keeps the same 8 channels, so per-channel parameters live in registers.
//
// No reference counterpart: the reference repository has no model code.
#include "bandwidth.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace plume {

namespace {

constexpr int kThreads = 256;

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

int gcd_int(int a, int b) { return b == 0 ? a : gcd_int(b, a % b); }

// Grid for `work` 16-byte vectors where the loop step (grid*256) must be a multiple of `cv`, the number
// of vectors per pixel, so that a thread's channel group is loop-invariant.
int grid_for(long long work, int cv, int blocks_per_sm = 8, int min_per_thread = 1) {
  // min_per_thread > 1: kernels that end with a per-block channel reduction (shared + global atomics)
  // should not be split so finely that the reduction dominates on small tensors
  static const int env_mpt = getenv("PLUME_BW_MIN_PER_THREAD") ? atoi(getenv("PLUME_BW_MIN_PER_THREAD")) : 0;
  static const int env_bps = getenv("PLUME_BW_BLOCKS_PER_SM") ? atoi(getenv("PLUME_BW_BLOCKS_PER_SM")) : 0;
  if (env_mpt > 0 && min_per_thread > 1) min_per_thread = env_mpt;   // diagnostics: sweep the reduction kernels' grids
  if (env_bps > 0 && min_per_thread > 1) blocks_per_sm = env_bps;
  const int g = cv / gcd_int(cv, kThreads);
  long long need = (work + 1ll * kThreads * min_per_thread - 1) / (1ll * kThreads * min_per_thread);
  long long cap = 1ll * sm_count() * blocks_per_sm;
  long long grid = std::max(1ll, std::min(need, cap));
  grid = (grid + g - 1) / g * g;
  return static_cast<int>(grid);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    return -3;
  }
  return 0;
}

// ---- deterministic mode ---------------------------------------------------------------------------
// PLUME_DETERMINISTIC=1 (or plume_set_deterministic): every reduction that normally ends in floating-point atomics
// (no fixed order) instead writes per-block partial sums to a library-owned scratch buffer and a second small
// kernel adds them up in block order, one thread per output element.  Inside a block the per-thread partials are
// combined in thread order.  Two runs then produce bit-identical results.
int g_det = -1;
struct DetScratch {
  void* ptr = nullptr;
  size_t bytes = 0;
};
DetScratch g_scratch[2];
}  // namespace

bool deterministic() {
  if (g_det < 0) {
    const char* e = getenv("PLUME_DETERMINISTIC");
    g_det = (e && atoi(e) != 0) ? 1 : 0;
  }
  return g_det == 1;
}
void set_deterministic(int on) { g_det = on ? 1 : 0; }

// which: 0 = bandwidth-kernel and conv-statistics reductions (main / chain stream), 1 = weight gradients (side stream).
// Grown with cudaMalloc when too small -- never inside a stream capture (the un-captured warm-up step sizes it).
void* det_scratch(int which, size_t bytes) {
  DetScratch& sc = g_scratch[which];
  if (sc.bytes < bytes) {
    // an outgrown buffer is NOT freed: a captured CUDA graph may still hold its address
    sc.ptr = nullptr;
    sc.bytes = 0;
    const size_t want = std::max(bytes + bytes / 2, static_cast<size_t>(32) << 20);
    if (cudaMalloc(&sc.ptr, want) != cudaSuccess) {
      cudaGetLastError();
      set_error("deterministic mode: cannot allocate the partial-sum scratch (inside a stream capture?)");
      return nullptr;
    }
    sc.bytes = want;
  }
  return sc.ptr;
}

// out[c] += sum over b (in order) of part[b * n + c], c < n
__global__ void ordered_sum_kernel(const float* __restrict__ part, int blocks, int n, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  float s = 0.f;
  for (int b = 0; b < blocks; ++b) s += part[static_cast<size_t>(b) * n + c];
  out[c] += s;
}
__global__ void ordered_sum_f64_kernel(const double* __restrict__ part, int blocks, int row, int n,
                                       double* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  double s = 0.0;
  for (int b = 0; b < blocks; ++b) s += part[static_cast<size_t>(b) * row + c];
  out[c] += s;
}
int ordered_sum_f64(const double* part, int blocks, int row, int n, double* out, cudaStream_t s) {
  ordered_sum_f64_kernel<<<(n + 255) / 256, 256, 0, s>>>(part, blocks, row, n, out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error(std::string("ordered_sum_f64: ") + cudaGetErrorString(e));
    return -3;
  }
  return 0;
}
int ordered_sum(const float* part, int blocks, int n, float* out, cudaStream_t s) {
  ordered_sum_kernel<<<(n + 255) / 256, 256, 0, s>>>(part, blocks, n, out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error(std::string("ordered_sum: ") + cudaGetErrorString(e));
    return -3;
  }
  return 0;
}

namespace {

struct alignas(16) BF8 {
  uint32_t u[4];
};
__device__ __forceinline__ BF8 ld_bf8(const __nv_bfloat16* p) {
  BF8 r;
  const uint4 v = *reinterpret_cast<const uint4*>(p);
  r.u[0] = v.x; r.u[1] = v.y; r.u[2] = v.z; r.u[3] = v.w;
  return r;
}
__device__ __forceinline__ BF8 ld_bf8_stream(const __nv_bfloat16* p) {
  BF8 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.u[0]), "=r"(r.u[1]), "=r"(r.u[2]), "=r"(r.u[3])
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_bf8(__nv_bfloat16* p, const BF8& v) {
  *reinterpret_cast<uint4*>(p) = make_uint4(v.u[0], v.u[1], v.u[2], v.u[3]);
}
__device__ __forceinline__ void unpack8(const BF8& v, float (&f)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = unpack_bf16x2(v.u[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ BF8 pack8(const float (&f)[8]) {
  BF8 v;
#pragma unroll
  for (int i = 0; i < 4; ++i) v.u[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
  return v;
}
// Activation storage traits: a thread moves 8 channels of one pixel per access.
//   __nv_bfloat16 : one 16-byte vector (the default storage);
//   Split2        : the "bf16x3" high-precision mode.  A value is stored as hi + lo, two bf16 numbers
//                   (hi = bf16(v), lo = bf16(v - hi): 16 significant bits), in two channel PLANES per pixel:
//                   pixel stride ld (in bf16 elements) covers both planes, hi of channel c at [p*ld + c], lo at
//                   [p*ld + ld/2 + c].  Each plane is an ordinary NHWC bf16 tensor as far as TMA is concerned, so
//                   the GEMMs form x*w = x_hi*w_hi + x_hi*w_lo + x_lo*w_hi with three passes of the bf16 MMAs.
// `plane` (= ld / 2) is ignored by the bf16 traits.  `pack` is where the storage rounding happens.
struct Split2 {
  __nv_bfloat16 v;
};
struct SP8 {
  BF8 hi, lo;
};
template <typename T>
struct Act;
template <>
struct Act<__nv_bfloat16> {
  using V8 = BF8;
  static constexpr int kPlanes = 1;
  static __device__ __forceinline__ V8 ld(const __nv_bfloat16* p, long long) { return ld_bf8(p); }
  static __device__ __forceinline__ V8 ld_stream(const __nv_bfloat16* p, long long) { return ld_bf8_stream(p); }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, long long, const V8& v) { st_bf8(p, v); }
  static __device__ __forceinline__ void unpack(const V8& v, float (&f)[8]) { unpack8(v, f); }
  static __device__ __forceinline__ V8 pack(const float (&f)[8]) { return pack8(f); }
  static __device__ __forceinline__ V8 zero() {
    V8 v;
    v.u[0] = v.u[1] = v.u[2] = v.u[3] = 0u;
    return v;
  }
};
template <>
struct Act<Split2> {
  using V8 = SP8;
  static constexpr int kPlanes = 2;
  static __device__ __forceinline__ const __nv_bfloat16* b(const Split2* p) {
    return reinterpret_cast<const __nv_bfloat16*>(p);
  }
  static __device__ __forceinline__ V8 ld(const Split2* p, long long plane) {
    V8 v;
    v.hi = ld_bf8(b(p));
    v.lo = ld_bf8(b(p) + plane);
    return v;
  }
  static __device__ __forceinline__ V8 ld_stream(const Split2* p, long long plane) {
    V8 v;
    v.hi = ld_bf8_stream(b(p));
    v.lo = ld_bf8_stream(b(p) + plane);
    return v;
  }
  static __device__ __forceinline__ void st(Split2* p, long long plane, const V8& v) {
    st_bf8(reinterpret_cast<__nv_bfloat16*>(p), v.hi);
    st_bf8(reinterpret_cast<__nv_bfloat16*>(p) + plane, v.lo);
  }
  static __device__ __forceinline__ void unpack(const V8& v, float (&f)[8]) {
    float l[8];
    unpack8(v.hi, f);
    unpack8(v.lo, l);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] += l[i];
  }
  static __device__ __forceinline__ V8 pack(const float (&f)[8]) {
    V8 v;
    v.hi = pack8(f);
    float h[8], r[8];
    unpack8(v.hi, h);
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = f[i] - h[i];  // exact in fp32
    v.lo = pack8(r);
    return v;
  }
  static __device__ __forceinline__ V8 zero() {
    V8 v;
    v.hi.u[0] = v.hi.u[1] = v.hi.u[2] = v.hi.u[3] = 0u;
    v.lo = v.hi;
    return v;
  }
};
// Launches `call` with T bound to the activation storage type (dt: 0 = bf16, 1 = bf16x3 split planes).
#define PLUME_ACT_DISPATCH(dt, ...)       \
  do {                                    \
    if ((dt) == 1) {                      \
      using T = Split2;                   \
      __VA_ARGS__;                        \
    } else {                              \
      using T = __nv_bfloat16;            \
      __VA_ARGS__;                        \
    }                                     \
  } while (0)

__device__ __forceinline__ void ld_f8(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// Block-level accumulation of K per-thread partial sums per channel into global fp32 arrays.
// Threads with the same channel group first combine in shared memory, then one atomic per channel.
template <int K>
__device__ __forceinline__ void block_channel_reduce(float (&acc)[K][8], int cv_idx, int CV,
                                                     float* const (&out)[K], float* s_acc,
                                                     float* det_part = nullptr) {
  const int span = min(CV, kThreads);
  const int slot = cv_idx % span;
  const int cv_block_base = cv_idx - slot;  // identical for all threads of the block (see grid_for)
  if (det_part != nullptr) {
    // deterministic: s_acc holds every thread's partials (kThreads * K * 8 floats); output element (k, slot, e)
    // is summed by ONE thread over the contributing threads in thread order (slot(t) = (slot(0) + t) % span),
    // then stored to this block's row of the partial-sum scratch: part[block][k][C] (zeroed by the launcher)
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int e = 0; e < 8; ++e) s_acc[(threadIdx.x * K + k) * 8 + e] = acc[k][e];
    __shared__ int s_slot0;
    if (threadIdx.x == 0) s_slot0 = slot;
    __syncthreads();
    const int slot0 = s_slot0;
    const int C = CV * 8;
    for (int i = threadIdx.x; i < K * span * 8; i += kThreads) {
      const int k = i / (span * 8);
      const int r = i % (span * 8);
      const int sl = r >> 3, e = r & 7;
      float v = 0.f;
      for (int t = (sl - slot0 + span) % span; t < kThreads; t += span) v += s_acc[(t * K + k) * 8 + e];
      det_part[(static_cast<size_t>(blockIdx.x) * K + k) * C + cv_block_base * 8 + r] = v;
    }
    return;
  }
  // s_acc: K * min(CV,256) * 8 floats, zeroed here
  for (int i = threadIdx.x; i < K * span * 8; i += kThreads) s_acc[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(&s_acc[(k * span + slot) * 8 + e], acc[k][e]);
  __syncthreads();
  // which global channel does smem slot s map to?  slot = cv % span and cv = (tid + const) % CV with
  // the loop-invariant mapping, so slot s holds channel group (cv0 - cv0 % span) + s where cv0 is this
  // block's first channel group; when CV <= 256, span == CV and the group is simply s.
  for (int i = threadIdx.x; i < K * span * 8; i += kThreads) {
    const int k = i / (span * 8);
    const int r = i % (span * 8);
    const float v = s_acc[i];
    if (v != 0.f) atomicAdd(out[k] + cv_block_base * 8 + r, v);
  }
}

// Host side of a deterministic channel reduction: scratch for `grid` blocks x K x C partial sums (zeroed: a
// block only writes the channel groups it covers), and the ordered final sums into out[k].
struct DetReduce {
  float* part = nullptr;
  int grid = 0, K = 0, C = 0;
  int begin(int grid_, int K_, int C_, cudaStream_t s) {
    grid = grid_; K = K_; C = C_;
    if (!deterministic()) return 0;
    const size_t bytes = static_cast<size_t>(grid) * K * C * sizeof(float);
    part = static_cast<float*>(det_scratch(0, bytes));
    if (!part) return -2;
    if (cudaMemsetAsync(part, 0, bytes, s) != cudaSuccess) {
      set_error("deterministic reduction: cudaMemsetAsync failed");
      return -2;
    }
    return 0;
  }
  // dynamic shared memory a reducing kernel needs
  size_t smem(int CV) const {
    return part ? static_cast<size_t>(kThreads) * K * 8 * sizeof(float)
                : static_cast<size_t>(K) * std::min(CV, kThreads) * 8 * sizeof(float);
  }
  // out[k] += ordered sum; the K arrays are separate pointers
  int finish(float* const* outs, cudaStream_t s) const {
    if (!part) return 0;
    for (int k = 0; k < K; ++k) {
      if (!outs[k]) continue;
      // row stride between blocks is K*C: view plane k as `grid` rows of K*C with offset k*C
      ordered_sum_strided(k, outs[k], s);
    }
    return 0;
  }
  void ordered_sum_strided(int k, float* out, cudaStream_t s) const;
};

__global__ void ordered_sum_strided_kernel(const float* __restrict__ part, int blocks, int row, int n,
                                           float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  float s = 0.f;
  for (int b = 0; b < blocks; ++b) s += part[static_cast<size_t>(b) * row + c];
  out[c] += s;
}
void DetReduce::ordered_sum_strided(int k, float* out, cudaStream_t s) const {
  ordered_sum_strided_kernel<<<(C + 255) / 256, 256, 0, s>>>(part + static_cast<size_t>(k) * C, grid, K * C, C, out);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// pad_channels
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void pad_channels_kernel(const __nv_bfloat16* __restrict__ in, int cvs,
                                    T* __restrict__ out, int cvd, long long total) {
  // the source is always plain bf16 (the input bands); the destination is in the activation storage format
  const long long step = 1ll * gridDim.x * blockDim.x;
  const long long ldo = 1ll * Act<T>::kPlanes * cvd * 8;
  for (long long i = 1ll * blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
    const long long p = i / cvd;
    const int cv = static_cast<int>(i % cvd);
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = 0.f;
    if (cv < cvs) unpack8(ld_bf8_stream(in + (p * cvs + cv) * 8), f);
    Act<T>::st(out + p * ldo + cv * 8, ldo >> 1, Act<T>::pack(f));
  }
}

int pad_channels(const void* in, int Cs, void* out, int Cd, long long pixels, int dt, cudaStream_t s) {
  if (Cs % 8 || Cd % 8 || Cs > Cd || Cs <= 0) {
    set_error("pad_channels: channel counts must be multiples of 8 with Cs <= Cd");
    return -1;
  }
  if (pixels <= 0) return 0;
  const long long total = pixels * (Cd / 8);
  const int grid = grid_for(total, 1);
  PLUME_ACT_DISPATCH(dt, (pad_channels_kernel<T><<<grid, kThreads, 0, s>>>(
                             static_cast<const __nv_bfloat16*>(in), Cs / 8, static_cast<T*>(out), Cd / 8, total)));
  return check_launch("pad_channels");
}

// ------------------------------------------------------------------------------------------------
// BatchNorm finalize / fold
// ------------------------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(const double* sum, const double* sq, double inv_count,
                                   float unbias, const float* gamma, const float* beta, float eps,
                                   float momentum, float* running_mean, float* running_var,
                                   float* scale, float* shift, float* mean_out, float* invstd_out,
                                   int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  // mean and variance in double: E[y^2] - E[y]^2 cancels catastrophically in fp32 once |mean| >> std
  const double mean_d = sum[c] * inv_count;
  const float mean = static_cast<float>(mean_d);
  const float var = static_cast<float>(fmax(sq[c] * inv_count - mean_d * mean_d, 0.0));
  const float invstd = rsqrtf(var + eps);
  const float g = gamma ? gamma[c] : 1.f;
  const float b = beta ? beta[c] : 0.f;
  const float sc = g * invstd;
  scale[c] = sc;
  shift[c] = b - mean * sc;
  if (mean_out) mean_out[c] = mean;
  if (invstd_out) invstd_out[c] = invstd;
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
  if (running_var) running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * unbias;
}

int bn_finalize(const double* sum, const double* sq, long long count, const float* gamma,
                const float* beta, float eps, float momentum, float* running_mean,
                float* running_var, float* scale, float* shift, float* mean, float* invstd, int C,
                cudaStream_t s) {
  if (count <= 0 || C <= 0) {
    set_error("bn_finalize: empty batch");
    return -1;
  }
  const float unbias = count > 1 ? static_cast<float>(count) / static_cast<float>(count - 1) : 1.f;
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, s>>>(sum, sq, 1.0 / static_cast<double>(count), unbias,
                                                     gamma, beta, eps, momentum, running_mean,
                                                     running_var, scale, shift, mean, invstd, C);
  return check_launch("bn_finalize");
}

__global__ void bn_fold_kernel(const float* gamma, const float* beta, const float* rm,
                               const float* rv, const float* bias, float eps, float* scale,
                               float* shift, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = (gamma ? gamma[c] : 1.f) * rsqrtf(rv[c] + eps);
  scale[c] = sc;
  shift[c] = ((bias ? bias[c] : 0.f) - rm[c]) * sc + (beta ? beta[c] : 0.f);
}

int bn_fold_eval(const float* gamma, const float* beta, const float* rm, const float* rv,
                 const float* bias, float eps, float* scale, float* shift, int C, cudaStream_t s) {
  bn_fold_kernel<<<(C + 127) / 128, 128, 0, s>>>(gamma, beta, rm, rv, bias, eps, scale, shift, C);
  return check_launch("bn_fold_eval");
}

// ------------------------------------------------------------------------------------------------
// scale/shift/act (BN apply + ReLU)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
    scale_shift_act_kernel(const T* __restrict__ y, long long ldy,
                           const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                           T* __restrict__ a, long long lda, long long pixels, int CV) {
  const long long i0 = 1ll * blockIdx.x * kThreads + threadIdx.x;
  const int cv = static_cast<int>(i0 % CV);
  const long long pstep = (1ll * gridDim.x * kThreads) / CV;
  float sc[8], sh[8];
  ld_f8(scale + cv * 8, sc);
  ld_f8(shift + cv * 8, sh);
  long long p = i0 / CV;
  // two pixels in flight per iteration
  for (; p + pstep < pixels; p += 2 * pstep) {
    const typename Act<T>::V8 v0 = Act<T>::ld_stream(y + p * ldy + cv * 8, ldy >> 1);
    const typename Act<T>::V8 v1 = Act<T>::ld_stream(y + (p + pstep) * ldy + cv * 8, ldy >> 1);
    float f0[8], f1[8];
    Act<T>::unpack(v0, f0);
    Act<T>::unpack(v1, f1);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      f0[e] = fmaf(f0[e], sc[e], sh[e]);
      f1[e] = fmaf(f1[e], sc[e], sh[e]);
      if (relu) {
        f0[e] = fmaxf(f0[e], 0.f);
        f1[e] = fmaxf(f1[e], 0.f);
      }
    }
    Act<T>::st(a + p * lda + cv * 8, lda >> 1, Act<T>::pack(f0));
    Act<T>::st(a + (p + pstep) * lda + cv * 8, lda >> 1, Act<T>::pack(f1));
  }
  if (p < pixels) {
    const typename Act<T>::V8 v0 = Act<T>::ld_stream(y + p * ldy + cv * 8, ldy >> 1);
    float f0[8];
    Act<T>::unpack(v0, f0);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      f0[e] = fmaf(f0[e], sc[e], sh[e]);
      if (relu) f0[e] = fmaxf(f0[e], 0.f);
    }
    Act<T>::st(a + p * lda + cv * 8, lda >> 1, Act<T>::pack(f0));
  }
}

int scale_shift_act(const void* y, int ldy, const float* scale, const float* shift, int relu, void* a,
                    int lda, long long pixels, int C, int dt, cudaStream_t s) {
  if (C % 8 || ldy % 8 || lda % 8 || C <= 0) {
    set_error("scale_shift_act: channels/strides must be multiples of 8");
    return -1;
  }
  if (pixels <= 0) return 0;
  const int CV = C / 8;
  const int grid = grid_for(pixels * CV, CV);
  PLUME_ACT_DISPATCH(dt, (scale_shift_act_kernel<T><<<grid, kThreads, 0, s>>>(
                             static_cast<const T*>(y), ldy, scale, shift, relu, static_cast<T*>(a), lda, pixels, CV)));
  return check_launch("scale_shift_act");
}

// ------------------------------------------------------------------------------------------------
// scale/shift/act fused with 2x2 max pool (+argmax), and the plain pool
// ------------------------------------------------------------------------------------------------
template <typename T, bool kAct, bool kSkip>
__global__ void __launch_bounds__(kThreads)
    act_pool_kernel(const T* __restrict__ y, long long ldy,
                    const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                    T* __restrict__ skip, long long ldskip,
                    T* __restrict__ pooled, long long ldpooled,
                    uint8_t* __restrict__ argmax, int N, int Ho, int Wo, int CV) {
  // one thread per (pooled pixel, channel group)
  const long long i0 = 1ll * blockIdx.x * kThreads + threadIdx.x;
  const int cv = static_cast<int>(i0 % CV);
  const long long pstep = (1ll * gridDim.x * kThreads) / CV;
  const long long opix = 1ll * N * Ho * Wo;
  float sc[8], sh[8];
  if (kAct) {
    ld_f8(scale + cv * 8, sc);
    ld_f8(shift + cv * 8, sh);
  }
  const int W = 2 * Wo;
  for (long long op = i0 / CV; op < opix; op += pstep) {
    const int wo = static_cast<int>(op % Wo);
    const long long t = op / Wo;
    const int ho = static_cast<int>(t % Ho);
    const long long n = t / Ho;
    const long long ip = (n * (2 * Ho) + 2 * ho) * W + 2 * wo;  // top-left input pixel
    typename Act<T>::V8 v[4];
    v[0] = Act<T>::ld_stream(y + ip * ldy + cv * 8, ldy >> 1);
    v[1] = Act<T>::ld_stream(y + (ip + 1) * ldy + cv * 8, ldy >> 1);
    v[2] = Act<T>::ld_stream(y + (ip + W) * ldy + cv * 8, ldy >> 1);
    v[3] = Act<T>::ld_stream(y + (ip + W + 1) * ldy + cv * 8, ldy >> 1);
    float f[4][8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      Act<T>::unpack(v[k], f[k]);
      if (kAct) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          f[k][e] = fmaf(f[k][e], sc[e], sh[e]);
          if (relu) f[k][e] = fmaxf(f[k][e], 0.f);
        }
        v[k] = Act<T>::pack(f[k]);
        Act<T>::unpack(v[k], f[k]);  // pool the rounded values (what the next layer reads)
      }
    }
    if (kSkip) {
      Act<T>::st(skip + ip * ldskip + cv * 8, ldskip >> 1, v[0]);
      Act<T>::st(skip + (ip + 1) * ldskip + cv * 8, ldskip >> 1, v[1]);
      Act<T>::st(skip + (ip + W) * ldskip + cv * 8, ldskip >> 1, v[2]);
      Act<T>::st(skip + (ip + W + 1) * ldskip + cv * 8, ldskip >> 1, v[3]);
    }
    float best[8];
    uint32_t idx_lo = 0, idx_hi = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float b = f[0][e];
      uint32_t bi = 0;
      if (f[1][e] > b) { b = f[1][e]; bi = 1; }
      if (f[2][e] > b) { b = f[2][e]; bi = 2; }
      if (f[3][e] > b) { b = f[3][e]; bi = 3; }
      best[e] = b;
      if (e < 4) idx_lo |= bi << (8 * e);
      else idx_hi |= bi << (8 * (e - 4));
    }
    Act<T>::st(pooled + op * ldpooled + cv * 8, ldpooled >> 1, Act<T>::pack(best));
    *reinterpret_cast<uint2*>(argmax + (op * CV + cv) * 8) = make_uint2(idx_lo, idx_hi);
  }
}

static int pool_common(bool act, bool with_skip, const void* y, int ldy, const float* scale,
                       const float* shift, int relu, void* skip, int ldskip, void* pooled,
                       int ldpooled, uint8_t* argmax, int N, int H, int W, int C, int dt, cudaStream_t s) {
  if (C % 8 || ldy % 8 || ldpooled % 8 || (with_skip && ldskip % 8) || C <= 0) {
    set_error("maxpool: channels/strides must be multiples of 8");
    return -1;
  }
  if ((H & 1) || (W & 1)) {
    set_error("maxpool: H and W must be even");
    return -1;
  }
  if (N <= 0 || H <= 0 || W <= 0) return 0;
  const int CV = C / 8;
  const long long work = 1ll * N * (H / 2) * (W / 2) * CV;
  const int grid = grid_for(work, CV);
  PLUME_ACT_DISPATCH(dt, {
    auto yy = static_cast<const T*>(y);
    auto sk = static_cast<T*>(skip);
    auto po = static_cast<T*>(pooled);
    if (act && with_skip)
      act_pool_kernel<T, true, true><<<grid, kThreads, 0, s>>>(yy, ldy, scale, shift, relu, sk, ldskip, po,
                                                               ldpooled, argmax, N, H / 2, W / 2, CV);
    else if (act)
      act_pool_kernel<T, true, false><<<grid, kThreads, 0, s>>>(yy, ldy, scale, shift, relu, sk, ldskip, po,
                                                                ldpooled, argmax, N, H / 2, W / 2, CV);
    else
      act_pool_kernel<T, false, false><<<grid, kThreads, 0, s>>>(yy, ldy, scale, shift, relu, sk, ldskip, po,
                                                                 ldpooled, argmax, N, H / 2, W / 2, CV);
  });
  return check_launch("maxpool2x2");
}

int scale_shift_act_pool(const void* y, int ldy, const float* scale, const float* shift, int relu,
                         void* skip, int ldskip, void* pooled, int ldpooled, uint8_t* argmax, int N,
                         int H, int W, int C, int dt, cudaStream_t s) {
  return pool_common(true, skip != nullptr, y, ldy, scale, shift, relu, skip, ldskip, pooled, ldpooled,
                     argmax, N, H, W, C, dt, s);
}
int maxpool2x2_fwd(const void* x, int ldx, void* y, int ldy, uint8_t* argmax, int N, int H, int W,
                   int C, int dt, cudaStream_t s) {
  return pool_common(false, false, x, ldx, nullptr, nullptr, 0, nullptr, 0, y, ldy, argmax, N, H, W, C,
                     dt, s);
}

// Optional fused BatchNorm-backward reduction: the gradient a kernel has just produced (dx of the pool backward,
// dfeat of the head backward) is the `da` of the BatchNorm in front of it, so the kernel can accumulate
// sum(g) and sum(g * xhat) (g = da masked by ReLU) itself -- one extra read of y instead of a separate
// bn_bwd_reduce pass over da AND y.  The sums use the ROUNDED gradient, i.e. what bn_bwd_apply will read back.
template <typename T>
struct BnReduceState {
  float sc[8], sh[8], mu[8], is[8];
  float acc[2][8];
  __device__ __forceinline__ void init(const BnReduceArgs& a, int cv) {
    ld_f8(a.scale + cv * 8, sc);
    ld_f8(a.shift + cv * 8, sh);
    ld_f8(a.mean + cv * 8, mu);
    ld_f8(a.invstd + cv * 8, is);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[0][e] = acc[1][e] = 0.f;
  }
  __device__ __forceinline__ typename Act<T>::V8 load_y(const BnReduceArgs& a, long long p, int cv) const {
    return Act<T>::ld_stream(static_cast<const T*>(a.y) + p * a.ldy + cv * 8, a.ldy >> 1);
  }
  // v: the stored (rounded) gradient of a pixel, yv: the layer's raw conv output there (load_y, issued early so
  // that the loads of several pixels are in flight together)
  __device__ __forceinline__ void add(const BnReduceArgs& a, const typename Act<T>::V8& v,
                                      const typename Act<T>::V8& yv) {
    float g[8], yy[8];
    Act<T>::unpack(v, g);
    Act<T>::unpack(yv, yy);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float gg = (!a.relu || fmaf(yy[e], sc[e], sh[e]) > 0.f) ? g[e] : 0.f;
      acc[0][e] += gg;
      acc[1][e] = fmaf(gg, (yy[e] - mu[e]) * is[e], acc[1][e]);
    }
  }
  __device__ __forceinline__ void finish(const BnReduceArgs& a, int cv, int CV, float* s_acc) {
    float* const outs[2] = {a.sum_g, a.sum_gx};
    block_channel_reduce<2>(acc, cv, CV, outs, s_acc, a.det_part);
  }
};

template <typename T, bool kReduce>
__global__ void __launch_bounds__(kThreads)
    maxpool_bwd_kernel(const T* __restrict__ dy, long long lddy,
                       const uint8_t* __restrict__ argmax, const T* __restrict__ dskip,
                       long long lddskip, T* __restrict__ dx, long long lddx, int N, int Ho,
                       int Wo, int CV, const BnReduceArgs bn) {
  extern __shared__ float s_acc[];
  const long long i0 = 1ll * blockIdx.x * kThreads + threadIdx.x;
  const int cv = static_cast<int>(i0 % CV);
  const long long pstep = (1ll * gridDim.x * kThreads) / CV;
  const long long opix = 1ll * N * Ho * Wo;
  const int W = 2 * Wo;
  BnReduceState<T> red;
  if (kReduce) red.init(bn, cv);
  for (long long op = i0 / CV; op < opix; op += pstep) {
    const int wo = static_cast<int>(op % Wo);
    const long long t = op / Wo;
    const int ho = static_cast<int>(t % Ho);
    const long long n = t / Ho;
    const long long ip = (n * (2 * Ho) + 2 * ho) * W + 2 * wo;
    float g[8];
    Act<T>::unpack(Act<T>::ld_stream(dy + op * lddy + cv * 8, lddy >> 1), g);
    const uint2 am = *reinterpret_cast<const uint2*>(argmax + (op * CV + cv) * 8);
    const long long off[4] = {ip, ip + 1, ip + W, ip + W + 1};
    typename Act<T>::V8 sv[4], yv[4];   // all loads of the 2 x 2 window first: up to 8 vectors in flight
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (dskip) sv[k] = Act<T>::ld_stream(dskip + off[k] * lddskip + cv * 8, lddskip >> 1);
      if (kReduce) yv[k] = red.load_y(bn, off[k], cv);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float o[8];
      if (dskip) {
        Act<T>::unpack(sv[k], o);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const uint32_t bi = ((e < 4 ? am.x : am.y) >> (8 * (e & 3))) & 0xffu;
        if (bi == static_cast<uint32_t>(k)) o[e] += g[e];
      }
      const typename Act<T>::V8 ov = Act<T>::pack(o);
      Act<T>::st(dx + off[k] * lddx + cv * 8, lddx >> 1, ov);
      if (kReduce) red.add(bn, ov, yv[k]);
    }
  }
  if (kReduce) red.finish(bn, cv, CV, s_acc);
}

static int bn_reduce_check(const BnReduceArgs* bn, const char* who) {
  if (!bn) return 0;
  if (!bn->y || !bn->scale || !bn->shift || !bn->mean || !bn->invstd || !bn->sum_g || !bn->sum_gx || bn->ldy % 8) {
    set_error(std::string(who) + ": incomplete BatchNorm-reduction arguments");
    return -1;
  }
  return 0;
}

int maxpool2x2_bwd(const void* dy, int lddy, const uint8_t* argmax, const void* dskip, int lddskip,
                   void* dx, int lddx, int N, int H, int W, int C, int dt, cudaStream_t s,
                   const BnReduceArgs* bn_in) {
  if (C % 8 || lddy % 8 || lddx % 8 || (dskip && lddskip % 8) || C <= 0) {
    set_error("maxpool_bwd: channels/strides must be multiples of 8");
    return -1;
  }
  if (H % 2 || W % 2) {
    set_error("maxpool_bwd: H and W must be even");
    return -1;
  }
  if (bn_reduce_check(bn_in, "maxpool_bwd")) return -1;
  if (N <= 0 || H <= 0 || W <= 0) return 0;
  const int CV = C / 8;
  const long long work = 1ll * N * (H / 2) * (W / 2) * CV;
  BnReduceArgs bn{};
  DetReduce dr;
  int grid = grid_for(work, CV);
  size_t smem = 0;
  if (bn_in) {
    bn = *bn_in;
    grid = grid_for(work, CV, 4, 4);
    if (dr.begin(grid, 2, C, s)) return -2;
    bn.det_part = dr.part;
    smem = dr.smem(CV);
  }
  PLUME_ACT_DISPATCH(dt, {
    auto a0 = static_cast<const T*>(dy);
    auto a1 = static_cast<const T*>(dskip);
    auto a2 = static_cast<T*>(dx);
    if (bn_in)
      maxpool_bwd_kernel<T, true><<<grid, kThreads, smem, s>>>(a0, lddy, argmax, a1, lddskip, a2, lddx, N, H / 2,
                                                               W / 2, CV, bn);
    else
      maxpool_bwd_kernel<T, false><<<grid, kThreads, 0, s>>>(a0, lddy, argmax, a1, lddskip, a2, lddx, N, H / 2,
                                                             W / 2, CV, bn);
  });
  if (int r = check_launch("maxpool2x2_bwd")) return r;
  if (!bn_in) return 0;
  float* const outs[2] = {bn.sum_g, bn.sum_gx};
  return dr.finish(outs, s);
}

// ------------------------------------------------------------------------------------------------
// BatchNorm + ReLU backward
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
    bn_bwd_reduce_kernel(const T* __restrict__ da, long long ldda,
                         const T* __restrict__ y, long long ldy,
                         const float* __restrict__ scale, const float* __restrict__ shift,
                         const float* __restrict__ mean, const float* __restrict__ invstd, int relu,
                         float* sum_g, float* sum_gx, long long pixels, int CV, float* det_part) {
  extern __shared__ float s_acc[];
  const long long i0 = 1ll * blockIdx.x * kThreads + threadIdx.x;
  const int cv = static_cast<int>(i0 % CV);
  const long long pstep = (1ll * gridDim.x * kThreads) / CV;
  float sc[8], sh[8], mu[8], is[8];
  ld_f8(scale + cv * 8, sc);
  ld_f8(shift + cv * 8, sh);
  ld_f8(mean + cv * 8, mu);
  ld_f8(invstd + cv * 8, is);
  float acc[2][8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[0][e] = acc[1][e] = 0.f;
  constexpr int U = 4;  // pixels in flight per thread: 8 independent 16-byte loads cover the HBM latency
  long long p = i0 / CV;
  for (; p + (U - 1) * pstep < pixels; p += U * pstep) {
    typename Act<T>::V8 gv[U], yv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      gv[u] = Act<T>::ld_stream(da + (p + u * pstep) * ldda + cv * 8, ldda >> 1);
      yv[u] = Act<T>::ld_stream(y + (p + u * pstep) * ldy + cv * 8, ldy >> 1);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float g[8], yy[8];
      Act<T>::unpack(gv[u], g);
      Act<T>::unpack(yv[u], yy);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float gg = (!relu || fmaf(yy[e], sc[e], sh[e]) > 0.f) ? g[e] : 0.f;
        acc[0][e] += gg;
        acc[1][e] = fmaf(gg, (yy[e] - mu[e]) * is[e], acc[1][e]);
      }
    }
  }
  for (; p < pixels; p += pstep) {
    float g[8], yy[8];
    Act<T>::unpack(Act<T>::ld_stream(da + p * ldda + cv * 8, ldda >> 1), g);
    Act<T>::unpack(Act<T>::ld_stream(y + p * ldy + cv * 8, ldy >> 1), yy);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float gg = (!relu || fmaf(yy[e], sc[e], sh[e]) > 0.f) ? g[e] : 0.f;
      acc[0][e] += gg;
      acc[1][e] = fmaf(gg, (yy[e] - mu[e]) * is[e], acc[1][e]);
    }
  }
  float* const outs[2] = {sum_g, sum_gx};
  block_channel_reduce<2>(acc, cv, CV, outs, s_acc, det_part);
}

int bn_bwd_reduce(const void* da, int ldda, const void* y, int ldy, const float* scale,
                  const float* shift, const float* mean, const float* invstd, int relu, float* sum_g,
                  float* sum_gx, long long pixels, int C, int dt, cudaStream_t s) {
  if (C % 8 || ldda % 8 || ldy % 8 || C <= 0) {
    set_error("bn_bwd_reduce: channels/strides must be multiples of 8");
    return -1;
  }
  if (pixels <= 0) return 0;
  const int CV = C / 8;
  // two blocks per SM = what the 108 registers allow: one wave, no tail (244 -> 215 us over the five layer shapes)
  const int grid = grid_for(pixels * CV, CV, 2, 16);
  DetReduce dr;
  if (dr.begin(grid, 2, C, s)) return -2;
  const size_t smem = dr.smem(CV);
  PLUME_ACT_DISPATCH(dt, (bn_bwd_reduce_kernel<T><<<grid, kThreads, smem, s>>>(
                             static_cast<const T*>(da), ldda, static_cast<const T*>(y), ldy, scale, shift, mean,
                             invstd, relu, sum_g, sum_gx, pixels, CV, dr.part)));
  if (int r = check_launch("bn_bwd_reduce")) return r;
  float* const outs[2] = {sum_g, sum_gx};
  return dr.finish(outs, s);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
    bn_bwd_apply_kernel(const T* __restrict__ da, long long ldda,
                        const T* __restrict__ y, long long ldy,
                        const float* __restrict__ scale, const float* __restrict__ shift,
                        const float* __restrict__ mean, const float* __restrict__ invstd, int relu,
                        const float* __restrict__ sum_g, const float* __restrict__ sum_gx,
                        float inv_count, T* __restrict__ dy, long long lddy, float* sum_dy,
                        float* dgamma, float* dbeta, int accumulate, long long pixels, int CV,
                        float* det_part) {
  extern __shared__ float s_acc[];
  // the finished per-channel sums are the BatchNorm parameter gradients; one block hands them over (sum_g /
  // sum_gx are a per-backward scratch, so that micro-batch accumulation never feeds stale sums into this pass)
  if (blockIdx.x == 0 && dgamma != nullptr) {
    for (int c = threadIdx.x; c < CV * 8; c += kThreads) {
      dgamma[c] = (accumulate ? dgamma[c] : 0.f) + sum_gx[c];
      dbeta[c] = (accumulate ? dbeta[c] : 0.f) + sum_g[c];
    }
  }
  const long long i0 = 1ll * blockIdx.x * kThreads + threadIdx.x;
  const int cv = static_cast<int>(i0 % CV);
  const long long pstep = (1ll * gridDim.x * kThreads) / CV;
  float sc[8], sh[8], mu[8], is[8], mg[8], mgx[8];
  ld_f8(scale + cv * 8, sc);
  ld_f8(shift + cv * 8, sh);
  ld_f8(mean + cv * 8, mu);
  ld_f8(invstd + cv * 8, is);
  ld_f8(sum_g + cv * 8, mg);
  ld_f8(sum_gx + cv * 8, mgx);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    mg[e] *= inv_count;
    mgx[e] *= inv_count;
  }
  float acc[1][8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[0][e] = 0.f;
  auto one = [&](const typename Act<T>::V8& gvv, const typename Act<T>::V8& yvv, long long pp) {
    float g[8], yy[8], o[8];
    Act<T>::unpack(gvv, g);
    Act<T>::unpack(yvv, yy);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float gg = (!relu || fmaf(yy[e], sc[e], sh[e]) > 0.f) ? g[e] : 0.f;
      const float xhat = (yy[e] - mu[e]) * is[e];
      o[e] = sc[e] * (gg - mg[e] - xhat * mgx[e]);
    }
    const typename Act<T>::V8 ov = Act<T>::pack(o);
    Act<T>::st(dy + pp * lddy + cv * 8, lddy >> 1, ov);
    if (sum_dy) {
      float r[8];
      Act<T>::unpack(ov, r);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[0][e] += r[e];
    }
  };
  constexpr int U = 4;
  long long p = i0 / CV;
  for (; p + (U - 1) * pstep < pixels; p += U * pstep) {
    typename Act<T>::V8 gv[U], yv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      gv[u] = Act<T>::ld_stream(da + (p + u * pstep) * ldda + cv * 8, ldda >> 1);
      yv[u] = Act<T>::ld_stream(y + (p + u * pstep) * ldy + cv * 8, ldy >> 1);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) one(gv[u], yv[u], p + u * pstep);
  }
  for (; p < pixels; p += pstep)
    one(Act<T>::ld_stream(da + p * ldda + cv * 8, ldda >> 1), Act<T>::ld_stream(y + p * ldy + cv * 8, ldy >> 1), p);
  if (sum_dy) {
    float* const outs[1] = {sum_dy};
    block_channel_reduce<1>(acc, cv, CV, outs, s_acc, det_part);
  }
}

int bn_bwd_apply(const void* da, int ldda, const void* y, int ldy, const float* scale,
                 const float* shift, const float* mean, const float* invstd, int relu,
                 const float* sum_g, const float* sum_gx, void* dy, int lddy, float* sum_dy,
                 float* dgamma, float* dbeta, int accumulate, long long pixels, int C, int dt, cudaStream_t s) {
  if ((dgamma == nullptr) != (dbeta == nullptr)) {
    set_error("bn_bwd_apply: dgamma and dbeta go together");
    return -1;
  }
  if (C % 8 || ldda % 8 || ldy % 8 || lddy % 8 || C <= 0) {
    set_error("bn_bwd_apply: channels/strides must be multiples of 8");
    return -1;
  }
  if (pixels <= 0) return 0;
  const int CV = C / 8;
  const int grid = grid_for(pixels * CV, CV, 2, 16);  // one wave at the kernel's occupancy (see bn_bwd_reduce)
  DetReduce dr;
  if (sum_dy && dr.begin(grid, 1, C, s)) return -2;
  const size_t smem = dr.part ? dr.smem(CV) : 1ull * std::min(CV, kThreads) * 8 * sizeof(float);
  PLUME_ACT_DISPATCH(dt, (bn_bwd_apply_kernel<T><<<grid, kThreads, smem, s>>>(
                             static_cast<const T*>(da), ldda, static_cast<const T*>(y), ldy, scale, shift, mean,
                             invstd, relu, sum_g, sum_gx, 1.f / static_cast<float>(pixels), static_cast<T*>(dy),
                             lddy, sum_dy, dgamma, dbeta, accumulate, pixels, CV, dr.part)));
  if (int r = check_launch("bn_bwd_apply")) return r;
  float* const outs[1] = {sum_dy};
  return dr.finish(outs, s);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
    relu_bwd_kernel(const T* __restrict__ da, long long ldda,
                    const T* __restrict__ a, long long lda,
                    T* __restrict__ dy, long long lddy, float* sum_dy, long long pixels,
                    int CV, float* det_part) {
  extern __shared__ float s_acc[];
  const long long i0 = 1ll * blockIdx.x * kThreads + threadIdx.x;
  const int cv = static_cast<int>(i0 % CV);
  const long long pstep = (1ll * gridDim.x * kThreads) / CV;
  float acc[1][8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[0][e] = 0.f;
  for (long long p = i0 / CV; p < pixels; p += pstep) {
    float g[8], aa[8];
    Act<T>::unpack(Act<T>::ld_stream(da + p * ldda + cv * 8, ldda >> 1), g);
    Act<T>::unpack(Act<T>::ld_stream(a + p * lda + cv * 8, lda >> 1), aa);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      g[e] = aa[e] > 0.f ? g[e] : 0.f;
      acc[0][e] += g[e];
    }
    Act<T>::st(dy + p * lddy + cv * 8, lddy >> 1, Act<T>::pack(g));
  }
  if (sum_dy) {
    float* const outs[1] = {sum_dy};
    block_channel_reduce<1>(acc, cv, CV, outs, s_acc, det_part);
  }
}

int relu_bwd(const void* da, int ldda, const void* a, int lda, void* dy, int lddy, float* sum_dy,
             long long pixels, int C, int dt, cudaStream_t s) {
  if (C % 8 || ldda % 8 || lda % 8 || lddy % 8 || C <= 0) {
    set_error("relu_bwd: channels/strides must be multiples of 8");
    return -1;
  }
  if (pixels <= 0) return 0;
  const int CV = C / 8;
  const int grid = grid_for(pixels * CV, CV, 4);
  DetReduce dr;
  if (sum_dy && dr.begin(grid, 1, C, s)) return -2;
  const size_t smem = dr.part ? dr.smem(CV) : 1ull * std::min(CV, kThreads) * 8 * sizeof(float);
  PLUME_ACT_DISPATCH(dt, (relu_bwd_kernel<T><<<grid, kThreads, smem, s>>>(
                             static_cast<const T*>(da), ldda, static_cast<const T*>(a), lda, static_cast<T*>(dy),
                             lddy, sum_dy, pixels, CV, dr.part)));
  if (int r = check_launch("relu_bwd")) return r;
  float* const outs[1] = {sum_dy};
  return dr.finish(outs, s);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
    channel_sum_kernel(const T* __restrict__ x, long long ldx, float* out,
                       long long pixels, int CV, float* det_part) {
  extern __shared__ float s_acc[];
  const long long i0 = 1ll * blockIdx.x * kThreads + threadIdx.x;
  const int cv = static_cast<int>(i0 % CV);
  const long long pstep = (1ll * gridDim.x * kThreads) / CV;
  float acc[1][8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[0][e] = 0.f;
  for (long long p = i0 / CV; p < pixels; p += pstep) {
    float f[8];
    Act<T>::unpack(Act<T>::ld_stream(x + p * ldx + cv * 8, ldx >> 1), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[0][e] += f[e];
  }
  float* const outs[1] = {out};
  block_channel_reduce<1>(acc, cv, CV, outs, s_acc, det_part);
}

int channel_sum(const void* x, int ldx, float* out, long long pixels, int C, int dt, cudaStream_t s) {
  if (C % 8 || ldx % 8 || C <= 0) {
    set_error("channel_sum: channels/strides must be multiples of 8");
    return -1;
  }
  if (pixels <= 0) return 0;
  const int CV = C / 8;
  const int grid = grid_for(pixels * CV, CV, 4, 16);
  DetReduce dr;
  if (dr.begin(grid, 1, C, s)) return -2;
  const size_t smem = dr.smem(CV);
  PLUME_ACT_DISPATCH(dt, (channel_sum_kernel<T><<<grid, kThreads, smem, s>>>(static_cast<const T*>(x), ldx, out,
                                                                            pixels, CV, dr.part)));
  if (int r = check_launch("channel_sum")) return r;
  float* const outs[1] = {out};
  return dr.finish(outs, s);
}

// ------------------------------------------------------------------------------------------------
// 1x1 head + sigmoid + BCE / Dice
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float z) { return 1.f / (1.f + __expf(-z)); }

// CV (= C/8, a power of two <= 32) consecutive lanes share one pixel.
template <typename T>
__global__ void __launch_bounds__(kThreads)
    head_fwd_kernel(const T* __restrict__ feat, long long ldf,
                    const float* __restrict__ w, const float* __restrict__ b,
                    const uint8_t* __restrict__ target, float* __restrict__ logits, float* sums,
                    long long pixels, int CV, float* det_part) {
  __shared__ float s_red[4][kThreads / 32];
  const long long i0 = 1ll * blockIdx.x * kThreads + threadIdx.x;
  const int cv = static_cast<int>(i0 % CV);
  const long long pstep = (1ll * gridDim.x * kThreads) / CV;
  float wv[8];
  ld_f8(w + cv * 8, wv);
  const float bias = b ? __ldg(b) : 0.f;
  float a_bce = 0.f, a_pt = 0.f, a_p = 0.f, a_t = 0.f;
  // every lane of a pixel group runs the same trip count, so the shuffles below are convergent;
  // four pixels are in flight per thread to cover the HBM latency
  constexpr int U = 4;
  const long long p0 = i0 / CV;
  const long long iters = (pixels + pstep - 1) / pstep;
  for (long long it = 0; it < iters; it += U) {
    typename Act<T>::V8 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long p = p0 + (it + u) * pstep;
      if (it + u < iters && p < pixels) v[u] = Act<T>::ld_stream(feat + p * ldf + cv * 8, ldf >> 1);
      else v[u] = Act<T>::zero();
    }
    float dots[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float f[8];
      Act<T>::unpack(v[u], f);
      float dot = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) dot = fmaf(f[e], wv[e], dot);
      for (int o = CV >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      dots[u] = dot;  // every lane of the pixel group now holds the full dot product
    }
    // spread the per-pixel epilogue (sigmoid / log) over the lanes of the group: lane u takes pixel u
    // (when the group has fewer lanes than pixels in flight, lane 0 takes them in turn)
    const int per_lane = CV >= U ? 1 : U;
    for (int k = 0; k < per_lane; ++k) {
      const int u = CV >= U ? cv : k;
      if (u < U && (CV >= U || cv == 0)) {
        float zsel = dots[0];
#pragma unroll
        for (int uu = 1; uu < U; ++uu) zsel = (u == uu) ? dots[uu] : zsel;
        const long long p = p0 + (it + u) * pstep;
        if (it + u < iters && p < pixels) {
          const float z = zsel + bias;
          logits[p] = z;
          if (target) {
            const float t = target[p] ? 1.f : 0.f;
            const float pr = sigmoidf_(z);
            a_bce += fmaxf(z, 0.f) - z * t + log1pf(__expf(-fabsf(z)));
            a_pt += pr * t;
            a_p += pr;
            a_t += t;
          }
        }
      }
    }
  }
  if (!target) return;
  a_bce = warp_sum(a_bce);
  a_pt = warp_sum(a_pt);
  a_p = warp_sum(a_p);
  a_t = warp_sum(a_t);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    s_red[0][warp] = a_bce;
    s_red[1][warp] = a_pt;
    s_red[2][warp] = a_p;
    s_red[3][warp] = a_t;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    float v = 0.f;
    for (int k = 0; k < kThreads / 32; ++k) v += s_red[threadIdx.x][k];
    if (det_part) det_part[blockIdx.x * 4 + threadIdx.x] = v;   // deterministic: summed in block order afterwards
    else atomicAdd(sums + threadIdx.x, v);
  }
}

static bool head_cv_ok(int C) {
  const int CV = C / 8;
  return C % 8 == 0 && CV >= 1 && CV <= 32 && (CV & (CV - 1)) == 0;
}

int head_fwd(const void* feat, int ldf, const float* w, const float* b, const uint8_t* target,
             float* logits, float* sums, long long pixels, int C, int dt, cudaStream_t s) {
  if (!head_cv_ok(C) || ldf % 8) {
    set_error("head_fwd: C must be 8*2^k <= 256 and ld a multiple of 8");
    return -1;
  }
  if (pixels <= 0) return 0;
  const int CV = C / 8;
  const int grid = grid_for(pixels * CV, CV, 4);
  float* part = nullptr;
  if (target && deterministic()) {
    part = static_cast<float*>(det_scratch(0, static_cast<size_t>(grid) * 4 * sizeof(float)));
    if (!part) return -2;
  }
  PLUME_ACT_DISPATCH(dt, (head_fwd_kernel<T><<<grid, kThreads, 0, s>>>(static_cast<const T*>(feat), ldf, w, b,
                                                                      target, logits, sums, pixels, CV, part)));
  if (int r = check_launch("head_fwd")) return r;
  return part ? ordered_sum(part, grid, 4, sums, s) : 0;
}

__global__ void head_loss_kernel(const float* sums, float inv_pixels, float bce_w, float dice_w,
                                 float eps, float* out) {
  const float bce = sums[0] * inv_pixels;
  const float dice = 1.f - (2.f * sums[1] + eps) / (sums[2] + sums[3] + eps);
  out[0] = bce_w * bce + dice_w * dice;
  out[1] = bce;
  out[2] = dice;
}

int head_loss(const float* sums, long long pixels, float bce_w, float dice_w, float eps, float* out,
              cudaStream_t s) {
  if (pixels <= 0) {
    set_error("head_loss: empty batch");
    return -1;
  }
  head_loss_kernel<<<1, 1, 0, s>>>(sums, 1.f / static_cast<float>(pixels), bce_w, dice_w, eps, out);
  return check_launch("head_loss");
}

template <typename T, bool kReduce>
__global__ void __launch_bounds__(kThreads)
    head_bwd_kernel(const T* __restrict__ feat, long long ldf,
                    const float* __restrict__ w, const float* __restrict__ logits,
                    const uint8_t* __restrict__ target, const float* __restrict__ sums, float inv_pixels,
                    float bce_w, float dice_w, float eps, float grad_scale,
                    T* __restrict__ dfeat, long long lddf, float* dw, float* db,
                    long long pixels, int CV, float* det_part, float* det_db, const BnReduceArgs bn) {
  extern __shared__ float s_acc[];
  __shared__ float s_db[kThreads / 32];
  const long long i0 = 1ll * blockIdx.x * kThreads + threadIdx.x;
  const int cv = static_cast<int>(i0 % CV);
  const long long pstep = (1ll * gridDim.x * kThreads) / CV;
  float wv[8];
  ld_f8(w + cv * 8, wv);
  // d(dice)/dp_i = -(2 t_i (S+eps) - (2I+eps)) / (S+eps)^2,  S = sum p + sum t, I = sum p t
  const float I2 = 2.f * sums[1] + eps;
  const float S = sums[2] + sums[3] + eps;
  const float invS2 = 1.f / (S * S);
  float acc[1][8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[0][e] = 0.f;
  float a_db = 0.f;
  BnReduceState<T> red;
  if (kReduce) red.init(bn, cv);
  auto one = [&](long long p, const typename Act<T>::V8& fv, const typename Act<T>::V8& yv, float z, float t) {
    const float pr = sigmoidf_(z);
    const float ddice = -(2.f * t * S - I2) * invS2;
    const float dz = grad_scale * (bce_w * (pr - t) * inv_pixels + dice_w * ddice * pr * (1.f - pr));
    float f[8], o[8];
    Act<T>::unpack(fv, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      o[e] = dz * wv[e];
      acc[0][e] = fmaf(dz, f[e], acc[0][e]);
    }
    const typename Act<T>::V8 ov = Act<T>::pack(o);
    Act<T>::st(dfeat + p * lddf + cv * 8, lddf >> 1, ov);
    if (kReduce) red.add(bn, ov, yv);
    if (cv == 0) a_db += dz;
  };
  constexpr int U = 4;
  long long p = i0 / CV;
  for (; p + (U - 1) * pstep < pixels; p += U * pstep) {
    typename Act<T>::V8 fv[U], yv[U];
    float zz[U], tt[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      fv[u] = Act<T>::ld_stream(feat + (p + u * pstep) * ldf + cv * 8, ldf >> 1);
      if (kReduce) yv[u] = red.load_y(bn, p + u * pstep, cv);
      zz[u] = __ldg(logits + p + u * pstep);
      tt[u] = target[p + u * pstep] ? 1.f : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) one(p + u * pstep, fv[u], yv[u], zz[u], tt[u]);
  }
  for (; p < pixels; p += pstep) {
    typename Act<T>::V8 yv1;
    if (kReduce) yv1 = red.load_y(bn, p, cv);
    one(p, Act<T>::ld_stream(feat + p * ldf + cv * 8, ldf >> 1), yv1, __ldg(logits + p), target[p] ? 1.f : 0.f);
  }
  float* const outs[1] = {dw};
  block_channel_reduce<1>(acc, cv, CV, outs, s_acc, det_part);
  if (kReduce) {
    __syncthreads();   // s_acc is reused
    red.finish(bn, cv, CV, s_acc);
  }
  a_db = warp_sum(a_db);
  if ((threadIdx.x & 31) == 0) s_db[threadIdx.x >> 5] = a_db;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int k = 0; k < kThreads / 32; ++k) v += s_db[k];
    if (det_db) det_db[blockIdx.x] = v;
    else atomicAdd(db, v);
  }
}

int head_bwd(const void* feat, int ldf, const float* w, const float* logits, const uint8_t* target,
             const float* sums, float bce_w, float dice_w, float eps, float grad_scale, void* dfeat,
             int lddf, float* dw, float* db, long long pixels, int C, int dt, cudaStream_t s,
             const BnReduceArgs* bn_in) {
  if (!head_cv_ok(C) || ldf % 8 || lddf % 8) {
    set_error("head_bwd: C must be 8*2^k <= 256 and strides multiples of 8");
    return -1;
  }
  if (bn_reduce_check(bn_in, "head_bwd")) return -1;
  if (pixels <= 0) return 0;
  const int CV = C / 8;
  const int grid = grid_for(pixels * CV, CV, 4);
  BnReduceArgs bn{};
  if (bn_in) bn = *bn_in;
  // deterministic mode: scratch = [grid][C] partial dw rows | [grid] partial db | [grid][2][C] BatchNorm sums
  float *part_dw = nullptr, *part_db = nullptr;
  const bool det = deterministic();
  if (det) {
    const size_t n_dw = static_cast<size_t>(grid) * C, n_db = grid, n_bn = bn_in ? 2 * n_dw : 0;
    part_dw = static_cast<float*>(det_scratch(0, (n_dw + n_db + n_bn) * sizeof(float)));
    if (!part_dw) return -2;
    if (cudaMemsetAsync(part_dw, 0, (n_dw + n_db + n_bn) * sizeof(float), s) != cudaSuccess) {
      set_error("head_bwd: cudaMemsetAsync failed");
      return -2;
    }
    part_db = part_dw + n_dw;
    bn.det_part = bn_in ? part_db + n_db : nullptr;
  }
  const int K = bn_in ? 2 : 1;
  const size_t smem = det ? static_cast<size_t>(kThreads) * K * 8 * sizeof(float)
                          : static_cast<size_t>(K) * CV * 8 * sizeof(float);
  PLUME_ACT_DISPATCH(dt, {
    auto f0 = static_cast<const T*>(feat);
    auto f1 = static_cast<T*>(dfeat);
    const float ip = 1.f / static_cast<float>(pixels);
    if (bn_in)
      head_bwd_kernel<T, true><<<grid, kThreads, smem, s>>>(f0, ldf, w, logits, target, sums, ip, bce_w, dice_w, eps,
                                                            grad_scale, f1, lddf, dw, db, pixels, CV, part_dw, part_db,
                                                            bn);
    else
      head_bwd_kernel<T, false><<<grid, kThreads, smem, s>>>(f0, ldf, w, logits, target, sums, ip, bce_w, dice_w, eps,
                                                             grad_scale, f1, lddf, dw, db, pixels, CV, part_dw, part_db,
                                                             bn);
  });
  if (int r = check_launch("head_bwd")) return r;
  if (!det) return 0;
  if (int r = ordered_sum(part_dw, grid, C, dw, s)) return r;
  if (int r = ordered_sum(part_db, grid, 1, db, s)) return r;
  if (bn_in) {
    DetReduce dr;
    dr.part = bn.det_part; dr.grid = grid; dr.K = 2; dr.C = C;
    float* const outs[2] = {bn.sum_g, bn.sum_gx};
    return dr.finish(outs, s);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// fp32 <-> bf16 casts of flat buffers (gradient buckets compressed for the data-parallel all-reduce)
// ------------------------------------------------------------------------------------------------
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  const long long step = 1ll * gridDim.x * blockDim.x;
  const long long n8 = n / 8;
  for (long long i = 1ll * blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += step) {
    float f[8];
    ld_f8(in + i * 8, f);
    st_bf8(out + i * 8, pack8(f));
  }
  for (long long i = n8 * 8 + 1ll * blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step)
    out[i] = __float2bfloat16_rn(in[i]);
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, long long n) {
  const long long step = 1ll * gridDim.x * blockDim.x;
  const long long n8 = n / 8;
  for (long long i = 1ll * blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += step) {
    float f[8];
    unpack8(ld_bf8(in + i * 8), f);
    *reinterpret_cast<float4*>(out + i * 8) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(out + i * 8 + 4) = make_float4(f[4], f[5], f[6], f[7]);
  }
  for (long long i = n8 * 8 + 1ll * blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step)
    out[i] = __bfloat162float(in[i]);
}
int cast_f32_bf16(const float* in, void* out, long long n, cudaStream_t s) {
  if (n <= 0) return 0;
  if ((reinterpret_cast<uintptr_t>(in) & 31) || (reinterpret_cast<uintptr_t>(out) & 15)) {
    set_error("cast_f32_bf16: buffers must be 32- / 16-byte aligned");
    return -1;
  }
  cast_f32_bf16_kernel<<<grid_for(n / 8 + 1, 1), kThreads, 0, s>>>(in, static_cast<__nv_bfloat16*>(out), n);
  return check_launch("cast_f32_bf16");
}
int cast_bf16_f32(const void* in, float* out, long long n, cudaStream_t s) {
  if (n <= 0) return 0;
  if ((reinterpret_cast<uintptr_t>(out) & 31) || (reinterpret_cast<uintptr_t>(in) & 15)) {
    set_error("cast_bf16_f32: buffers must be 16- / 32-byte aligned");
    return -1;
  }
  cast_bf16_f32_kernel<<<grid_for(n / 8 + 1, 1), kThreads, 0, s>>>(static_cast<const __nv_bfloat16*>(in), out, n);
  return check_launch("cast_bf16_f32");
}

// ------------------------------------------------------------------------------------------------
// Adam
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
    adam_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ m,
                float* __restrict__ v, long long n4, long long n, float lr_t, float beta1, float beta2,
                float omb1, float omb2, float eps, float inv_bc2_sqrt, float gscale) {
  const long long step = 1ll * gridDim.x * blockDim.x;
  for (long long i = 1ll * blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += step) {
    float4 p = reinterpret_cast<float4*>(param)[i];
    const float4 g = reinterpret_cast<const float4*>(grad)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pp = reinterpret_cast<float*>(&p);
    const float* gp = reinterpret_cast<const float*>(&g);
    float* mp = reinterpret_cast<float*>(&mm);
    float* vp = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float ge = gp[e] * gscale;
      mp[e] = beta1 * mp[e] + omb1 * ge;
      vp[e] = beta2 * vp[e] + omb2 * ge * ge;
      pp[e] -= lr_t * mp[e] / (sqrtf(vp[e]) * inv_bc2_sqrt + eps);
    }
    reinterpret_cast<float4*>(param)[i] = p;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  // tail (n not a multiple of 4)
  const long long tail0 = n4 * 4;
  const long long ti = tail0 + 1ll * blockIdx.x * blockDim.x + threadIdx.x;
  if (ti < n) {
    const float ge = grad[ti] * gscale;
    const float me = beta1 * m[ti] + omb1 * ge;
    const float ve = beta2 * v[ti] + omb2 * ge * ge;
    m[ti] = me;
    v[ti] = ve;
    param[ti] -= lr_t * me / (sqrtf(ve) * inv_bc2_sqrt + eps);
  }
}

// Same update with the step-dependent coefficients read from device memory, so that a captured CUDA
// graph of the whole training step can be replayed (the host refreshes the 8 floats before each replay):
// coef = {lr/(1-b1^t), beta1, beta2, 1-beta1, 1-beta2, eps, 1/sqrt(1-b2^t), grad_scale}
__global__ void __launch_bounds__(kThreads)
    adam_dev_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ m,
                    float* __restrict__ v, long long n4, long long n, const float* __restrict__ coef) {
  const float lr_t = coef[0], beta1 = coef[1], beta2 = coef[2], omb1 = coef[3], omb2 = coef[4], eps = coef[5],
              inv_bc2_sqrt = coef[6], gscale = coef[7];
  const long long step = 1ll * gridDim.x * blockDim.x;
  for (long long i = 1ll * blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += step) {
    float4 p = reinterpret_cast<float4*>(param)[i];
    const float4 g = reinterpret_cast<const float4*>(grad)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pp = reinterpret_cast<float*>(&p);
    const float* gp = reinterpret_cast<const float*>(&g);
    float* mp = reinterpret_cast<float*>(&mm);
    float* vp = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float ge = gp[e] * gscale;
      mp[e] = beta1 * mp[e] + omb1 * ge;
      vp[e] = beta2 * vp[e] + omb2 * ge * ge;
      pp[e] -= lr_t * mp[e] / (sqrtf(vp[e]) * inv_bc2_sqrt + eps);
    }
    reinterpret_cast<float4*>(param)[i] = p;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  const long long ti = n4 * 4 + 1ll * blockIdx.x * blockDim.x + threadIdx.x;
  if (ti < n) {
    const float ge = grad[ti] * gscale;
    const float me = beta1 * m[ti] + omb1 * ge;
    const float ve = beta2 * v[ti] + omb2 * ge * ge;
    m[ti] = me;
    v[ti] = ve;
    param[ti] -= lr_t * me / (sqrtf(ve) * inv_bc2_sqrt + eps);
  }
}

int adam_dev(float* param, const float* grad, float* m, float* v, long long n, const float* coef,
             cudaStream_t s) {
  if (n <= 0) return 0;
  const long long n4 = n / 4;
  const int grid = grid_for(std::max(n4, 1ll), 1);
  adam_dev_kernel<<<grid, kThreads, 0, s>>>(param, grad, m, v, n4, n, coef);
  return check_launch("adam_dev");
}

int adam(float* param, const float* grad, float* m, float* v, long long n, double lr, double beta1,
         double beta2, double eps, int step, float grad_scale, cudaStream_t s) {
  if (n <= 0) return 0;
  if (step < 1) {
    set_error("adam: step must be >= 1");
    return -1;
  }
  const double bc1 = 1.0 - std::pow(beta1, step);
  const double bc2 = 1.0 - std::pow(beta2, step);
  const float lr_t = static_cast<float>(lr / bc1);
  const float inv_bc2_sqrt = static_cast<float>(1.0 / std::sqrt(bc2));
  const long long n4 = n / 4;
  const int grid = grid_for(std::max(n4, 1ll), 1);
  adam_kernel<<<grid, kThreads, 0, s>>>(param, grad, m, v, n4, n, lr_t, static_cast<float>(beta1),
                                        static_cast<float>(beta2), static_cast<float>(1.0 - beta1),
                                        static_cast<float>(1.0 - beta2), static_cast<float>(eps),
                                        inv_bc2_sqrt, grad_scale);
  return check_launch("adam");
}

// ------------------------------------------------------------------------------------------------
// weight packing
// ------------------------------------------------------------------------------------------------
// w, wf: [co][t][ci] (ci contiguous);  wd: [ci][8-t][co] (co contiguous).  One block transposes a
// 32 (co) x 32 (ci) tile of one tap through shared memory so that both the fp32 reads and the two bf16
// writes are coalesced.
__global__ void __launch_bounds__(256)
    pack_conv3x3_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf,
                        __nv_bfloat16* __restrict__ wd, int Cout, int Cin) {
  __shared__ float tile[32][33];
  const int t = blockIdx.z;
  const int co0 = blockIdx.y * 32, ci0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int co = co0 + ty + 8 * k, ci = ci0 + tx;
    float v = 0.f;
    if (co < Cout && ci < Cin) {
      const long long i = (1ll * co * 9 + t) * Cin + ci;
      v = w[i];
      if (wf) wf[i] = __float2bfloat16_rn(v);
    }
    tile[ty + 8 * k][tx] = v;
  }
  if (!wd) return;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ci = ci0 + ty + 8 * k, co = co0 + tx;
    if (co < Cout && ci < Cin)
      wd[(1ll * ci * 9 + (8 - t)) * Cout + co] = __float2bfloat16_rn(tile[tx][ty + 8 * k]);
  }
}

int pack_conv3x3(const float* w, void* wf, void* wd, int Cout, int Cin, cudaStream_t s) {
  if (Cout <= 0 || Cin <= 0) return 0;
  dim3 grid((Cin + 31) / 32, (Cout + 31) / 32, 9);
  pack_conv3x3_kernel<<<grid, 256, 0, s>>>(w, static_cast<__nv_bfloat16*>(wf),
                                           static_cast<__nv_bfloat16*>(wd), Cout, Cin);
  return check_launch("pack_conv3x3");
}

__global__ void pack_convT_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf,
                                  __nv_bfloat16* __restrict__ wd, int Cout, int Cin) {
  // w, wf: [ij][co][ci];  wd: [ci][ij][co]
  const long long total = 4ll * Cout * Cin;
  const long long step = 1ll * gridDim.x * blockDim.x;
  for (long long i = 1ll * blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
    const int ci = static_cast<int>(i % Cin);
    const int co = static_cast<int>((i / Cin) % Cout);
    const int ij = static_cast<int>(i / (1ll * Cin * Cout));
    const __nv_bfloat16 v = __float2bfloat16_rn(w[i]);
    if (wf) wf[i] = v;
    if (wd) wd[(1ll * ci * 4 + ij) * Cout + co] = v;
  }
}

int pack_convT2x2(const float* w, void* wf, void* wd, int Cout, int Cin, cudaStream_t s) {
  const long long total = 4ll * Cout * Cin;
  if (total <= 0) return 0;
  const int grid = grid_for(total, 1);
  pack_convT_kernel<<<grid, kThreads, 0, s>>>(w, static_cast<__nv_bfloat16*>(wf),
                                              static_cast<__nv_bfloat16*>(wd), Cout, Cin);
  return check_launch("pack_convT2x2");
}

// Table-driven version: one launch packs every layer.  A block handles one 32 x 32 (co, ci) tile of one tap
// of one layer; both layouts are a transpose of the master:
//   conv3x3 : w[(co*9 + t)*Cin + ci]     -> wd[(ci*9 + 8-t)*Cout + co]
//   convT2x2: w[(t*Cout + co)*Cin + ci]  -> wd[(ci*4 + t)*Cout + co]
__global__ void __launch_bounds__(256)
    pack_batch_kernel(const plume_pack_desc* __restrict__ descs, int n) {
  __shared__ float tile[32][33];
  __shared__ int s_entry;
  // which layer owns this block: one coalesced read of the first_block column, then a ballot
  if (threadIdx.x < 32) {
    int e = 0;
    for (int base = 0; base < n; base += 32) {
      const int i = base + threadIdx.x;
      const bool le = i < n && descs[i].first_block <= static_cast<int>(blockIdx.x);
      e += __popc(__ballot_sync(0xffffffffu, le));
    }
    if (threadIdx.x == 0) s_entry = e - 1;
  }
  __syncthreads();
  const plume_pack_desc d = descs[s_entry];
  const int local = blockIdx.x - d.first_block;
  const int tiles_ci = (d.Cin + 31) / 32, tiles_co = (d.Cout + 31) / 32;
  const int ci0 = (local % tiles_ci) * 32;
  const int co0 = ((local / tiles_ci) % tiles_co) * 32;
  const int t = local / (tiles_ci * tiles_co);
  const int conv3 = (d.kind & 1) == 0;
  const bool split = (d.kind & 2) != 0;   // bf16x3 mode: the hi matrix is followed by the lo matrix (same layout)
  const int taps = conv3 ? 9 : 4;
  const long long numel = 1ll * taps * d.Cout * d.Cin;
  const float* __restrict__ w = d.w;
  __nv_bfloat16* __restrict__ wf = static_cast<__nv_bfloat16*>(d.w_fwd_bf16);
  __nv_bfloat16* __restrict__ wd = static_cast<__nv_bfloat16*>(d.w_dgrad_bf16);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int co = co0 + ty + 8 * k, ci = ci0 + tx;
    float v = 0.f;
    if (co < d.Cout && ci < d.Cin) {
      const long long i = conv3 ? (1ll * co * 9 + t) * d.Cin + ci : (1ll * t * d.Cout + co) * d.Cin + ci;
      v = w[i];
      if (wf) {
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        wf[i] = h;
        if (split) wf[numel + i] = __float2bfloat16_rn(v - __bfloat162float(h));
      }
    }
    tile[ty + 8 * k][tx] = v;
  }
  if (!wd) return;
  __syncthreads();
  const int td = conv3 ? 8 - t : t;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ci = ci0 + ty + 8 * k, co = co0 + tx;
    if (co < d.Cout && ci < d.Cin) {
      const float v = tile[tx][ty + 8 * k];
      const long long j = (1ll * ci * taps + td) * d.Cout + co;
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      wd[j] = h;
      if (split) wd[numel + j] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}

int pack_blocks(int kind, int Cout, int Cin) {
  if (Cout <= 0 || Cin <= 0) return 0;
  return ((Cin + 31) / 32) * ((Cout + 31) / 32) * ((kind & 1) == 0 ? 9 : 4);
}

int pack_batch(const plume_pack_desc* descs, int n, int total_blocks, cudaStream_t s) {
  if (n <= 0 || total_blocks <= 0) return 0;
  pack_batch_kernel<<<total_blocks, 256, 0, s>>>(descs, n);
  return check_launch("pack_batch");
}

// ------------------------------------------------------------------------------------------------
// tiled inference: cut tiles out of a scene, stitch logits back by centre crop + threshold
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void extract_tiles_kernel(const __nv_bfloat16* __restrict__ scene, int Hs, int Ws, int cvs,
                                     const int* __restrict__ ys, const int* __restrict__ xs, int tile,
                                     T* __restrict__ tiles, int cvd, long long total) {
  // the scene is always plain bf16; the tiles are in the activation storage format
  const long long step = 1ll * gridDim.x * blockDim.x;
  const long long ldo = 1ll * Act<T>::kPlanes * cvd * 8;
  for (long long i = 1ll * blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
    const int cv = static_cast<int>(i % cvd);
    long long t = i / cvd;
    const long long pix = t;
    const int x = static_cast<int>(t % tile);
    t /= tile;
    const int y = static_cast<int>(t % tile);
    const int k = static_cast<int>(t / tile);
    const int sy = ys[k] + y, sx = xs[k] + x;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = 0.f;
    if (cv < cvs && sy >= 0 && sy < Hs && sx >= 0 && sx < Ws)
      unpack8(ld_bf8(scene + ((1ll * sy * Ws + sx) * cvs + cv) * 8), f);
    Act<T>::st(tiles + pix * ldo + cv * 8, ldo >> 1, Act<T>::pack(f));
  }
}

int extract_tiles(const void* scene, int Hs, int Ws, int Cs, const int* ys, const int* xs, int count,
                  int tile, void* tiles, int Cd, int dt, cudaStream_t s) {
  if (Cs % 8 || Cd % 8 || Cs > Cd || Cs <= 0) {
    set_error("extract_tiles: channel counts must be multiples of 8 with Cs <= Cd");
    return -1;
  }
  if (count <= 0) return 0;
  const long long total = 1ll * count * tile * tile * (Cd / 8);
  const int grid = grid_for(total, 1);
  PLUME_ACT_DISPATCH(dt, (extract_tiles_kernel<T><<<grid, kThreads, 0, s>>>(
                             static_cast<const __nv_bfloat16*>(scene), Hs, Ws, Cs / 8, ys, xs, tile, static_cast<T*>(tiles),
                             Cd / 8,
                             total)));
  return check_launch("extract_tiles");
}

__global__ void stitch_kernel(const float* __restrict__ logits, const int* __restrict__ ys,
                              const int* __restrict__ xs, int T, int margin, float thr,
                              uint8_t* __restrict__ mask, float* __restrict__ prob, int Hs, int Ws,
                              long long total) {
  const long long step = 1ll * gridDim.x * blockDim.x;
  for (long long i = 1ll * blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
    const int x = static_cast<int>(i % T);
    const long long t = i / T;
    const int y = static_cast<int>(t % T);
    const int k = static_cast<int>(t / T);
    const int y0 = ys[k], x0 = xs[k];
    const int sy = y0 + y, sx = x0 + x;
    if (sy < 0 || sy >= Hs || sx < 0 || sx >= Ws) continue;
    // a tile owns its interior; a border strip is owned only where the tile touches the scene edge
    const bool own_y = (y >= margin || y0 <= 0) && (y < T - margin || y0 + T >= Hs);
    const bool own_x = (x >= margin || x0 <= 0) && (x < T - margin || x0 + T >= Ws);
    if (!(own_y && own_x)) continue;
    const float z = logits[i];
    mask[1ll * sy * Ws + sx] = z >= thr ? 1 : 0;
    if (prob) prob[1ll * sy * Ws + sx] = 1.f / (1.f + __expf(-z));
  }
}

int stitch_threshold(const float* logits, const int* ys, const int* xs, int count, int T, int margin,
                     float thr, uint8_t* mask, float* prob, int Hs, int Ws, cudaStream_t s) {
  if (count <= 0) return 0;
  if (margin < 0 || 2 * margin >= T) {
    set_error("stitch_threshold: margin must satisfy 0 <= 2*margin < T");
    return -1;
  }
  const long long total = 1ll * count * T * T;
  const int grid = grid_for(total, 1);
  stitch_kernel<<<grid, kThreads, 0, s>>>(logits, ys, xs, T, margin, thr, mask, prob, Hs, Ws, total);
  return check_launch("stitch_threshold");
}

}  // namespace plume
