// Implicit-GEMM convolution kernels for sm_100a: TMA-fed tcgen05.mma with TMEM accumulators.
//
//  * igemm_fwd_kernel  : pixels x out-channels GEMM, K = taps x in-channels, both operands K-major.
//                        A tiles are 4-D TMA boxes of the NHWC activation (halo/padding comes from
//                        TMA out-of-bounds zero fill, one load per filter tap); B tiles are 2-D boxes
//                        of the packed weight matrix.  Persistent CTAs, warp-specialised:
//                        warp0 = TMA producer, warp1 = MMA issuer, warp2 = TMEM allocator,
//                        warps4-7 = epilogue (TMEM -> regs -> scale/shift/ReLU -> bf16 -> swizzled smem
//                        -> TMA store; optional per-channel sum / sum-of-squares for BatchNorm).
//                        Two TMEM accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
//  * igemm_wgrad_kernel: in-channels x out-channels GEMM per filter tap, K = pixels.  Both operands are
//                        MN-major (channels are the contiguous dimension of NHWC), so the same TMA boxes
//                        feed the MMA without a transpose.  Split-K over pixel tiles; the partial sums are
//                        added into dW with fp32 reductions (RED.ADD) from the epilogue.
//
// No reference counterpart: the reference repository has no model code.
#include "igemm.cuh"
#include "bandwidth.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <cstdlib>
#include <mutex>

#ifndef PLUME_PAIR_ISSUERS
#define PLUME_PAIR_ISSUERS 2   // MMA issuer threads of the CTA-pair conv3 kernels in the resident / triple modes
#endif

namespace plume {

__device__ int g_dbg_word = 0;

static long long* g_prof_buf = nullptr;
void set_prof_buffer(long long* buf) { g_prof_buf = buf; }

int read_debug_word() {
  int v = 0;
  cudaMemcpyFromSymbol(&v, g_dbg_word, sizeof(int));
  return v;
}

struct TmapPack4 {
  CUtensorMap m[4];
};
// views 0-3, and in the bf16x3 mode their lo planes at 4-7
struct TmapPack8 {
  CUtensorMap m[8];
};
struct TmapPack2 {
  CUtensorMap m[2];
};

// SMs left to the persistent GEMM kernels.  Under data parallelism NCCL's all-reduce kernels hold a few SMs for
// milliseconds; a persistent kernel with one CTA per SM and statically assigned tiles then runs its last CTAs as a
// second wave and takes up to twice as long.  plume_set_sm_margin(k) (PLUME_SM_MARGIN) sizes the persistent grids
// and the split-K wave arithmetic for `SMs - k`, so every CTA is resident from the start.
static int g_sm_margin = -1;
void set_sm_margin(int k) { g_sm_margin = k < 0 ? 0 : k; }
int get_sm_margin() {
  if (g_sm_margin < 0) {
    const char* e = getenv("PLUME_SM_MARGIN");
    g_sm_margin = e ? std::max(0, atoi(e)) : 0;
  }
  return g_sm_margin;
}
static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return std::max(8, n - get_sm_margin());
}

// Deterministic BatchNorm statistics: `rows` rows of [sum | sq] (c doubles each) in the scratch, zeroed; after the
// launch the rows are added into stat_sum / stat_sq in row order.
static double* det_stats_begin(int rows, int c, cudaStream_t s) {
  const size_t bytes = static_cast<size_t>(rows) * 2 * c * sizeof(double);
  double* part = static_cast<double*>(det_scratch(0, bytes));
  if (part && cudaMemsetAsync(part, 0, bytes, s) != cudaSuccess) {
    set_error("deterministic statistics: cudaMemsetAsync failed");
    return nullptr;
  }
  return part;
}
static int det_stats_finish(const double* part, int rows, int c, double* stat_sum, double* stat_sq, cudaStream_t s) {
  if (int r = ordered_sum_f64(part, rows, 2 * c, c, stat_sum, s)) return r;
  return ordered_sum_f64(part + c, rows, 2 * c, c, stat_sq, s);
}

static int pow2_floor(int v) {
  int p = 1;
  while (p * 2 <= v) p *= 2;
  return p;
}

// Spatial box of `target` pixels (a power of two) for an image of W x H: as square as the image allows.
static void pick_box(int W, int H, int target, int max_w, int* bw, int* bh, int* bn) {
  int w = std::min(pow2_floor(W), max_w);
  w = std::min(w, target);
  int h = std::min(pow2_floor(H), target / w);
  *bw = w;
  *bh = h;
  *bn = target / (w * h);
}

// =================================================================================================
// Forward kernel
// =================================================================================================
struct FwdParams {
  int tiles_w, tiles_h, tiles_n;  // spatial tiling of the GEMM M dimension
  int TW, TH, TN;                 // box extents, TW*TH*TN == 128
  int W, H, N;                    // output extents (row-valid mask)
  int n_tiles;                    // tiles along GEMM N
  int num_taps, per_tap_view;     // per_tap_view: tap t reads input view t (no spatial shift)
  int kb_per_tap;                 // Cin / 64
  int k_per_tap;                  // Cin
  int cout_per_view;
  int n_total;                    // rows of the weight matrix (bf16x3: the lo matrix follows at row n_total)
  const float* scale;
  const float* shift;
  int relu;
  double* stat_sum;
  double* stat_sq;
  double* det_part;               // deterministic mode: per-(CTA, warpgroup) rows [sum | sq] of det_c doubles each
  int det_c;
};

template <int BLOCK_N, int STAGES, int STAGING, int EPI_WG = 1>
struct FwdSmem {
  static constexpr int A_BYTES = 128 * 128;      // 128 pixels x 64 bf16 (one 128-B swizzle row each)
  static constexpr int B_BYTES = BLOCK_N * 128;  // BLOCK_N out-channels x 64 bf16
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STG_BYTES = 128 * 128;    // 128 pixels x 64 bf16 outputs
  static constexpr int OFF_STG = STAGES * STAGE_BYTES;              // STAGING buffers per epilogue warpgroup
  static constexpr int OFF_PARAM = OFF_STG + EPI_WG * STAGING * STG_BYTES;  // per warpgroup: scale|shift|sum|sq
  static constexpr int OFF_BAR = OFF_PARAM + EPI_WG * 4 * BLOCK_N * 4;
  static constexpr int NUM_BARS = 2 * STAGES + 4;
  static constexpr int OFF_TMEMPTR = OFF_BAR + NUM_BARS * 8;
  static constexpr int TOTAL = OFF_TMEMPTR + 16 + 1024;  // + slack for the 1024-B alignment
  static_assert(TOTAL <= 232448, "shared memory budget exceeded");
};

// 32 accumulator columns of one pixel row -> scale/shift/ReLU -> bf16 -> four 16-byte chunks of the
// 128-byte-swizzled staging row (chunk index XOR row%8, the layout the TMA store expects).
__device__ __forceinline__ void epi_store_half(const uint32_t (&v)[32], int jbase, const float* sc,
                                               const float* sh, int relu, bool valid, uint8_t* stg,
                                               int row) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float y[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float t = fmaf(__uint_as_float(v[j * 8 + e]), sc[j * 8 + e], sh[j * 8 + e]);
      if (relu) t = fmaxf(t, 0.0f);
      y[e] = valid ? t : 0.0f;
    }
    uint4 o;
    o.x = pack_bf16x2(y[0], y[1]);
    o.y = pack_bf16x2(y[2], y[3]);
    o.z = pack_bf16x2(y[4], y[5]);
    o.w = pack_bf16x2(y[6], y[7]);
    *reinterpret_cast<uint4*>(stg + row * 128 + (((jbase + j) ^ (row & 7)) << 4)) = o;
  }
}

// bf16x3 mode: the same, with the value split into hi = bf16(t) and lo = bf16(t - hi), staged as two tiles.
__device__ __forceinline__ void epi_store_half_split(const uint32_t (&v)[32], int jbase, const float* sc,
                                                     const float* sh, int relu, bool valid, uint8_t* stg_hi,
                                                     uint8_t* stg_lo, int row) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float y[8], r[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float t = fmaf(__uint_as_float(v[j * 8 + e]), sc[j * 8 + e], sh[j * 8 + e]);
      if (relu) t = fmaxf(t, 0.0f);
      y[e] = valid ? t : 0.0f;
    }
    uint4 o;
    o.x = pack_bf16x2(y[0], y[1]);
    o.y = pack_bf16x2(y[2], y[3]);
    o.z = pack_bf16x2(y[4], y[5]);
    o.w = pack_bf16x2(y[6], y[7]);
    const uint32_t hw[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 h = unpack_bf16x2(hw[e]);
      r[2 * e] = y[2 * e] - h.x;          // exact in fp32
      r[2 * e + 1] = y[2 * e + 1] - h.y;
    }
    uint4 l;
    l.x = pack_bf16x2(r[0], r[1]);
    l.y = pack_bf16x2(r[2], r[3]);
    l.z = pack_bf16x2(r[4], r[5]);
    l.w = pack_bf16x2(r[6], r[7]);
    const int off = row * 128 + (((jbase + j) ^ (row & 7)) << 4);
    *reinterpret_cast<uint4*>(stg_hi + off) = o;
    *reinterpret_cast<uint4*>(stg_lo + off) = l;
  }
}

// SPLIT = the bf16x3 high-precision mode: every activation is two bf16 planes (hi, lo; maps 0-3 / 4-7), the
// weight matrix is followed by its lo matrix at row n_total, and the K loop runs three passes into the same
// accumulator: x_hi*w_hi + x_hi*w_lo + x_lo*w_hi (the dropped lo*lo term is 2^-18 relative).  The epilogue
// splits the fp32 result again (STAGING == 2: one staging tile per plane).
// EPI_WG = 2: two epilogue warpgroups (warps 4-7 and 8-11); warpgroup g drains TMEM accumulator g, i.e. every
// other tile of the CTA, with its own staging buffers.  The transposed convolutions (K = Cin, one or four taps)
// are bound by the epilogue, not by the MMAs: one warpgroup needs ~2.5 k cycles per 128 x 64 chunk.
template <int BLOCK_N, int STAGES, int STAGING, bool SPLIT, int EPI_WG>
__global__ void __launch_bounds__(128 + 128 * EPI_WG, 1)
    igemm_fwd_kernel(const __grid_constant__ TmapPack8 amaps,
                     const __grid_constant__ CUtensorMap bmap,
                     const __grid_constant__ TmapPack8 omaps, const FwdParams p) {
  static_assert(!SPLIT || STAGING == 2, "bf16x3 mode stages a hi and a lo tile");
  using L = FwdSmem<BLOCK_N, STAGES, STAGING, EPI_WG>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t sbase = (raw_addr + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (sbase - raw_addr);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  const uint32_t bar0 = sbase + L::OFF_BAR;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + 2 + a); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(gbase + L::OFF_TMEMPTR);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&amaps.m[0]);
    tma_prefetch_desc(&bmap);
    tma_prefetch_desc(&omaps.m[0]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 128);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(sbase + L::OFF_TMEMPTR, 2 * BLOCK_N);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int total_tiles = m_tiles * p.n_tiles;
  constexpr int PASSES = SPLIT ? 3 : 1;
  const int kiters = PASSES * p.num_taps * p.kb_per_tap;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one lane)
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles;
        const int m_tile = tile / p.n_tiles;
        const int w0 = (m_tile % p.tiles_w) * p.TW;
        const int h0 = ((m_tile / p.tiles_w) % p.tiles_h) * p.TH;
        const int n0 = (m_tile / (p.tiles_w * p.tiles_h)) * p.TN;
        for (int pt = 0; pt < PASSES * p.num_taps; ++pt) {
          const int pass = SPLIT ? pt / p.num_taps : 0;
          const int tap = SPLIT ? pt - pass * p.num_taps : pt;
          const int a_plane = pass == 2 ? 4 : 0;                 // x_lo in the third pass
          const int b_row0 = pass == 1 ? p.n_total : 0;          // w_lo in the second pass
          int dw = 0, dh = 0, av = 0;
          if (p.per_tap_view) {
            av = tap;
          } else if (p.num_taps == 9) {
            dh = tap / 3 - 1;
            dw = tap % 3 - 1;
          }
          for (int kb = 0; kb < p.kb_per_tap; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1, 1, &g_dbg_word);
            const uint32_t a_addr = sbase + stage * L::STAGE_BYTES;
            const uint32_t b_addr = a_addr + L::A_BYTES;
            mbar_expect_tx(full_bar(stage), L::STAGE_BYTES);
            tma_load_4d(a_addr, &amaps.m[av + a_plane], full_bar(stage), kb * 64, w0 + dw, h0 + dh, n0);
            tma_load_2d(b_addr, &bmap, full_bar(stage), tap * p.k_per_tap + kb * 64,
                        b_row0 + n_tile * BLOCK_N);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one lane)
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BLOCK_N, 0, 0);
      int stage = 0, phase = 0, it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const int acc_phase = (it >> 1) & 1;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1, 2, &g_dbg_word);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        uint32_t accumulate = 0;
        for (int kit = 0; kit < kiters; ++kit) {
          mbar_wait(full_bar(stage), phase, 3, &g_dbg_word);
          tc_fence_after();
          const uint32_t a_lo = umma_desc_lo(sbase + stage * L::STAGE_BYTES, 16);
          const uint32_t b_lo = a_lo + (L::A_BYTES >> 4);
          constexpr uint32_t hi = umma_desc_hi_sw128(1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_bf16_lohi(d_tmem, a_lo + 2 * k, hi, b_lo + 2 * k, hi, idesc, accumulate);
            accumulate = 1;
          }
          umma_commit(empty_bar(stage));  // slot reusable once these MMAs have read it
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (4 warps per warpgroup)
    const int wg = (warp - 4) >> 2;  // warpgroup: drains the tiles it = wg, wg + EPI_WG, ...
    const int q = warp & 3;          // the TMEM lane quadrant this warp may read
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 128 - wg * 128;
    const uint32_t bar_a = 1 + 2 * wg, bar_b = 2 + 2 * wg;
    const uint32_t off_stg = L::OFF_STG + wg * STAGING * L::STG_BYTES;
    float* s_scale = reinterpret_cast<float*>(gbase + L::OFF_PARAM) + wg * 4 * BLOCK_N;
    float* s_shift = s_scale + BLOCK_N;
    float* s_sum = s_shift + BLOCK_N;
    float* s_sq = s_sum + BLOCK_N;
    const bool do_stats = p.stat_sum != nullptr;
    int it = wg;
    int chunk_ctr = 0;
    for (int tile = blockIdx.x + wg * gridDim.x; tile < total_tiles; tile += EPI_WG * gridDim.x, it += EPI_WG) {
      const int n_tile = tile % p.n_tiles;
      const int m_tile = tile / p.n_tiles;
      const int w0 = (m_tile % p.tiles_w) * p.TW;
      const int h0 = ((m_tile / p.tiles_w) % p.tiles_h) * p.TH;
      const int n0 = (m_tile / (p.tiles_w * p.tiles_h)) * p.TN;
      const int acc = it & 1;
      const int acc_phase = (it >> 1) & 1;
      // A tile may span several output views (transposed conv: the four (i,j) phases are the four 64-column
      // chunks of one 256-wide tile, so the input tile is read once for all of them); a 64-column chunk never
      // straddles views (cout_per_view % 64 == 0).  Per-channel vectors are indexed by n % cout_per_view.
      const int col0 = n_tile * BLOCK_N;

      for (int c = et; c < BLOCK_N; c += 128) {
        const int ch = (col0 + c) % p.cout_per_view;
        s_scale[c] = p.scale ? p.scale[ch] : 1.0f;
        s_shift[c] = p.shift ? p.shift[ch] : 0.0f;
        s_sum[c] = 0.0f;
        s_sq[c] = 0.0f;
      }
      const int pw = row % p.TW;
      const int ph = (row / p.TW) % p.TH;
      const int pn = row / (p.TW * p.TH);
      const bool valid = (w0 + pw < p.W) && (h0 + ph < p.H) && (n0 + pn < p.N);

      mbar_wait(tfull_bar(acc), acc_phase, 4, &g_dbg_word);
      tc_fence_after();

#pragma unroll 1
      for (int chunk = 0; chunk < BLOCK_N / 64; ++chunk, ++chunk_ctr) {
        uint32_t v0[32], v1[32];
        {
          const uint32_t taddr =
              tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N + chunk * 64;
          tmem_ld_32x32b_x32(taddr, v0);
          tmem_ld_32x32b_x32(taddr + 32, v1);
          tmem_ld_wait();
        }
        if (chunk == BLOCK_N / 64 - 1) {
          tc_fence_before();
          mbar_arrive(tempty_bar(acc));  // accumulator drained: MMA may overwrite it
        }
        const int sb = (STAGING == 2 && !SPLIT) ? (chunk_ctr & 1) : 0;
        // staging buffer `sb` (bf16x3: both buffers) no longer being read by an earlier store
        if (et == 0) tma_store_wait_read<SPLIT ? 0 : STAGING - 1>();
        named_bar_sync(bar_a, 128);

        uint8_t* stg = gbase + off_stg + sb * L::STG_BYTES;
        if (SPLIT) {
          epi_store_half_split(v0, 0, s_scale + chunk * 64, s_shift + chunk * 64, p.relu, valid, stg,
                               stg + L::STG_BYTES, row);
          epi_store_half_split(v1, 4, s_scale + chunk * 64 + 32, s_shift + chunk * 64 + 32, p.relu, valid, stg,
                               stg + L::STG_BYTES, row);
        } else {
          epi_store_half(v0, 0, s_scale + chunk * 64, s_shift + chunk * 64, p.relu, valid, stg, row);
          epi_store_half(v1, 4, s_scale + chunk * 64 + 32, s_shift + chunk * 64 + 32, p.relu, valid,
                         stg, row);
        }
        fence_proxy_async_smem();
        named_bar_sync(bar_b, 128);
        if (et == 0) {
          const int col = col0 + chunk * 64;
          tma_store_4d(&omaps.m[col / p.cout_per_view], sbase + off_stg + sb * L::STG_BYTES,
                       col % p.cout_per_view, w0, h0, n0);
          if (SPLIT)
            tma_store_4d(&omaps.m[4 + col / p.cout_per_view], sbase + off_stg + L::STG_BYTES,
                         col % p.cout_per_view, w0, h0, n0);
          tma_store_commit();
        }
        if (do_stats) {
          // column sums of the rounded outputs: lane <-> channel pair, warp <-> 32-row group
          const int j = lane >> 2, wsub = lane & 3;
          float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll 8
          for (int rr = 0; rr < 32; ++rr) {
            const int r = q * 32 + rr;
            const int off = r * 128 + ((j ^ (r & 7)) << 4) + wsub * 4;
            float2 f = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(stg + off));
            if (SPLIT) {  // the stored value is hi + lo
              const float2 l = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(stg + L::STG_BYTES + off));
              f.x += l.x;
              f.y += l.y;
            }
            s0 += f.x;
            s1 += f.y;
            q0 = fmaf(f.x, f.x, q0);
            q1 = fmaf(f.y, f.y, q1);
          }
          if (p.det_part) {
            for (int turn = 0; turn < 4; ++turn) {   // deterministic: the four warps take turns
              if (q == turn) {
                s_sum[chunk * 64 + 2 * lane] += s0;
                s_sum[chunk * 64 + 2 * lane + 1] += s1;
                s_sq[chunk * 64 + 2 * lane] += q0;
                s_sq[chunk * 64 + 2 * lane + 1] += q1;
              }
              named_bar_sync(bar_a, 128);
            }
          } else {
            atomicAdd(&s_sum[chunk * 64 + 2 * lane], s0);
            atomicAdd(&s_sum[chunk * 64 + 2 * lane + 1], s1);
            atomicAdd(&s_sq[chunk * 64 + 2 * lane], q0);
            atomicAdd(&s_sq[chunk * 64 + 2 * lane + 1], q1);
          }
        }
      }
      if (do_stats) {
        named_bar_sync(bar_a, 128);
        for (int c = et; c < BLOCK_N; c += 128) {
          const int ch = (col0 + c) % p.cout_per_view;
          // fp64 accumulators: a channel with |mean| >> std loses its variance to cancellation in fp32
          if (p.det_part) {
            // deterministic: this warpgroup's row of the scratch; successive tiles of a warpgroup are ordered by
            // the named barriers above, and within a tile every channel has one writer
            double* rowp = p.det_part + (EPI_WG * blockIdx.x + wg) * 2ll * p.det_c;
            rowp[ch] += static_cast<double>(s_sum[c]);
            rowp[p.det_c + ch] += static_cast<double>(s_sq[c]);
          } else {
            atomicAdd(p.stat_sum + ch, static_cast<double>(s_sum[c]));
            atomicAdd(p.stat_sq + ch, static_cast<double>(s_sq[c]));
          }
        }
      }
    }
    if (et == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 2 * BLOCK_N);
  static_assert(EPI_WG == 1 || EPI_WG == 2, "one or two epilogue warpgroups");
}

template <int BLOCK_N, int STAGES, int STAGING, bool SPLIT = false, int EPI_WG = 1>
static int launch_fwd_inst(const TmapPack8& amaps, const CUtensorMap& bmap, const TmapPack8& omaps,
                           const FwdParams& p, int total_tiles, cudaStream_t stream) {
  using L = FwdSmem<BLOCK_N, STAGES, STAGING, EPI_WG>;
  auto kern = igemm_fwd_kernel<BLOCK_N, STAGES, STAGING, SPLIT, EPI_WG>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] {
    attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
  });
  if (attr_err != cudaSuccess) {
    set_error(std::string("cudaFuncSetAttribute(igemm_fwd): ") + cudaGetErrorString(attr_err));
    return -2;
  }
  const int grid = std::min(total_tiles, num_sms());
  kern<<<grid, 128 + 128 * EPI_WG, L::TOTAL, stream>>>(amaps, bmap, omaps, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error(std::string("igemm_fwd launch: ") + cudaGetErrorString(e));
    return -3;
  }
  return 0;
}

// =================================================================================================
// 3x3 convolution kernel with halo reuse ("conv3"): the forward kernel above reloads the activation
// tile once per filter tap, which makes the narrow layers (64 / 128 output channels) L2-bound
// (profiles/: 3.6 GB of L2->SM traffic for a 268 MB input).  Here an M tile is 8 (w) x 16 (h) pixels,
// so each group of 8 consecutive GEMM rows (one 1024-byte swizzle atom) is one image row of the tile.
// One TMA box of 8 x 18 pixels (one halo row above and below) then serves the three vertical taps:
// tap dh simply starts the UMMA descriptor (1 + dh) atoms further.  Horizontal taps still need their
// own boxes (a one-pixel shift moves pixels across atoms).  A traffic drops from 9 x 16 KB to
// 3 x 18 KB per 64 input channels.  When the whole weight slice [BLOCK_N][9*Cin] fits in shared memory
// it is loaded once per CTA and stays resident; otherwise weight tiles flow through their own ring.
// =================================================================================================
// Division by a runtime constant without the ~20-instruction integer divide (tile decoding runs in
// every epilogue thread once per tile).  Granlund-Montgomery round-up method, exact for all 32-bit n.
struct FastDiv {
  uint32_t mul, sh1, sh2, d;
};
static FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;
  f.mul = static_cast<uint32_t>(((1ull << 32) * ((1ull << l) - d)) / d + 1);
  f.sh1 = l < 1 ? l : 1;
  f.sh2 = l > 1 ? l - 1 : 0;
  return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) {
  const uint32_t t = __umulhi(f.mul, n);
  return (t + ((n - t) >> f.sh1)) >> f.sh2;
}

struct Conv3Params {
  int tiles_w, tiles_h;           // per image; tile = 8 x 16 pixels
  int W, H, N;
  int n_tiles;                    // tiles along GEMM N
  int kb;                         // Cin / 64
  int cin;
  int a_slots, b_slots;           // ring depths (b_slots == 0: weights resident)
  FastDiv div_ntiles, div_tiles_img, div_tiles_w;
  const float* scale;
  const float* shift;
  int relu;
  double* stat_sum;
  double* stat_sq;
  double* det_part;               // deterministic mode: per-(CTA, warpgroup) rows [sum | sq] of `cout` doubles each
  int cout;
  long long* prof;                // optional [grid][8] cycle counters (plume_debug_set_prof)
};

constexpr int kHaloBytes = 18 * 1024;  // 8 w x 18 h pixels x 64 channels bf16
constexpr int kMaxSlots = 8;
constexpr int kConv3Threads = 448;     // 8 epilogue warps + 6 single-thread role warps

// One 32-column half of a 64-channel chunk: scale/shift (+ReLU) -> bf16 -> swizzled staging row.
template <bool RELU, bool MASK>
__device__ __forceinline__ void epi_half(const uint32_t (&v)[32], int jbase, const float* sc,
                                         const float* sh, bool valid, uint8_t* stg, int row) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float y[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float t = fmaf(__uint_as_float(v[j * 8 + e]), sc[j * 8 + e], sh[j * 8 + e]);
      if (RELU) t = fmaxf(t, 0.0f);
      y[e] = t;
    }
    uint4 o;
    o.x = pack_bf16x2(y[0], y[1]);
    o.y = pack_bf16x2(y[2], y[3]);
    o.z = pack_bf16x2(y[4], y[5]);
    o.w = pack_bf16x2(y[6], y[7]);
    if (MASK && !valid) o = make_uint4(0u, 0u, 0u, 0u);  // rows outside the image must not reach the stats
    *reinterpret_cast<uint4*>(stg + row * 128 + (((jbase + j) ^ (row & 7)) << 4)) = o;
  }
}

// Spin on a shared-memory counter published by a relay thread.  The load is deliberately a WEAK
// ld.shared: a volatile / relaxed (strong) load issued behind tcgen05.mma instructions stalls the thread
// for ~300 cycles (it appears to wait for the asynchronous MMAs in flight), which drains the tensor pipe;
// the weak load returns in ~30 cycles (scripts/bench_mma3.cu: 83 -> 59 cycles per N=64 MMA, 108 -> 71 at
// N=128, 133 -> 128 at N=256).  Shared memory is not cached, so a weak load in a loop still observes the
// relay's store; what follows the wait is ordered by the control dependency on the loaded value.
__device__ __forceinline__ uint32_t ld_counter(const volatile uint32_t* ctr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(const_cast<const uint32_t*>(ctr))) : "memory");
  return v;
}
__device__ __forceinline__ void wait_counter(const volatile uint32_t* ctr, uint32_t need, int tag) {
  if (ld_counter(ctr) >= need) return;  // common case: the data is already there (weak, cheap probe)
  // not there yet: the pipe is starving anyway, so spin with strong loads (a weak load inside a loop may be
  // kept in a register by the assembler and never observe the relay's store)
  long long t0 = clock64();
  while (*ctr < need) {
    if (clock64() - t0 > PLUME_WATCHDOG_CYCLES) {
      atomicExch(&g_dbg_word, 0x7100 | tag);
      __threadfence_system();
      __trap();
    }
  }
}

// Warp roles (the warp arbiter prefers higher warp ids within a scheduler, so the single-thread roles sit
// above the epilogue warps):
//   0-7   epilogue, two warpgroups; warpgroup g drains TMEM accumulator g
//   8     TMEM allocator, then relay for "accumulator drained" barriers
//   9     TMA producer
//   10    barrier init, then relay for "halo slot full" barriers
//   11,13 MMA issuers (two threads that alternate MMA groups)
//   12    relay for "weight slot full" barriers (ring mode)
//
// What limits small-N MMAs is the issuing thread, not the tensor pipe (scripts/bench_mma*.cu): the pipe
// buffers about one MMA beyond the running one, and every observation of shared state by the issuing
// thread -- an mbarrier probe, even a plain shared-memory load -- costs it ~300 cycles, a tcgen05.commit
// ~80.  Measured cycles per M=128 MMA at N = 64 / 128: 48 / 64 with nothing in between, 83 / 108 with one
// poll + commit per 12 (4) MMAs.  Hence:
//   * relay threads do the mbarrier waits and publish monotonically increasing counters in shared
//     memory (the data was written by TMA before the mbarrier completed, which the relay observed before
//     bumping the counter, which the issuer reads before issuing);
//   * MMAs are issued in groups of TPG taps (12 MMAs, or 4 for 256-wide tiles whose weight tiles are too
//     big to be grouped) and TWO issuer threads alternate groups: while one polls / commits, the other's
//     MMAs keep the pipe full (55.6 / 64.1 cycles per MMA in the micro-benchmark).  Each issuer
//     accumulates into its OWN TMEM tile (its groups in a fixed order) and the epilogue adds the two
//     partial accumulators, so results do not depend on how the two instruction streams interleave.
//
// MODE 0: weights resident in shared memory (whole [BLOCK_N][9*Cin] slice), group = one halo slot
// MODE 1: weight ring of "triples" (the three vertical taps of one halo slot share a barrier)
// MODE 2: weight ring of single tiles (BLOCK_N = 256), group = one tap
// PAIR (MODE 2 only, launched as clusters of two CTAs): tcgen05.mma.cta_group::2.  The two CTAs of a pair work on two
// adjacent M tiles with the same N tile; each loads its own halo boxes and HALF of every weight tile (128 of the 256
// rows), both CTAs' TMA loads complete on the LEADER's "full" barriers (rank 0), the leader's relays and its single
// issuer thread see them and issue M = 256 MMAs that read A from both CTAs and the two halves of B, commits are
// multicast to both CTAs' "empty" / "accumulator full" barriers, each CTA's epilogue drains its own TMEM and
// arrives on the leader's "accumulator drained" barrier.  Per SM the operand traffic of an N = 256 MMA drops from
// 21.5 KB to 13.5 KB per 128 cycles, i.e. below what shared memory delivers (scripts/test_2cta.cu: 128.1 cycles per
// pair MMA, correct against a CPU matmul).
template <int BLOCK_N, int MODE, bool HALF_STAGE, bool PAIR = false>
__global__ void __launch_bounds__(kConv3Threads, 1)
    igemm_conv3_kernel(const __grid_constant__ CUtensorMap amap,
                       const __grid_constant__ CUtensorMap bmap,
                       const __grid_constant__ CUtensorMap omap, const Conv3Params p) {
  constexpr bool RESIDENT = MODE == 0;
  constexpr int TPG = MODE == 2 ? 1 : 3;   // taps per MMA group
  constexpr int GPS = 3 / TPG;             // groups per halo slot
  // 512-cycle groups of N=256 MMAs hide a single issuer's sync.  The pair kernels of the other modes keep two issuers:
  // with one (-DPLUME_PAIR_ISSUERS=1, A/B of two builds on one box, gpurun_out/r2ab) the 64-wide layers slow down from
  // 190 to 240 us and 128 -> 64 from 318 to 431 us -- they run at the pair instruction's floor of ~80 cycles
  // (scripts/test_2cta.cu) only when the issue overhead of one thread hides behind the other's MMAs.
  constexpr uint32_t ISSUERS = MODE == 2 ? 1 : (PAIR ? PLUME_PAIR_ISSUERS : 2);
  constexpr int B_BYTES = (PAIR ? BLOCK_N / 2 : BLOCK_N) * 128;   // PAIR: this CTA's half of the weight tile
  constexpr int BSLOT_BYTES = TPG * B_BYTES;
  // HALF_STAGE: stage 64 rows (half a tile) at a time.  128-wide tiles in triple mode need a third weight slot more
  // than a full staging tile; resident weight slices of 144 KB (64 -> 128 and 128 -> 64 channels) fit beside three
  // halo slots only with the smaller staging buffers.
  constexpr int STG_ROWS = HALF_STAGE ? 64 : 128;
  constexpr int STG_BYTES = STG_ROWS * 128;
  constexpr int kWarpAlloc = 8, kWarpProducer = 9, kWarpInit = 10, kWarpMma0 = 11, kWarpRelayB = 12,
                kWarpMma1 = 13;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t sbase = (raw_addr + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (sbase - raw_addr);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;      // position in the CTA pair; rank 0 is the leader
  const bool leader = rank == 0;
  // virtual CTA index / count over which work items are distributed (PAIR: the pair)
  const int vcta = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int vgrid = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  const uint32_t b_region = RESIDENT ? 9u * p.kb * B_BYTES : static_cast<uint32_t>(p.b_slots) * BSLOT_BYTES;
  const uint32_t off_a = 0;
  const uint32_t off_b = off_a + p.a_slots * kHaloBytes;
  const uint32_t off_stg = off_b + b_region;                // one staging buffer per epilogue warpgroup
  const uint32_t off_param = off_stg + 2 * STG_BYTES;       // per warpgroup: scale|shift|sum|sq
  const uint32_t off_bar = off_param + 2 * 4 * BLOCK_N * 4;
  const uint32_t bar0 = sbase + off_bar;
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (kMaxSlots + s); };
  auto b_full = [&](int s) { return bar0 + 8u * (2 * kMaxSlots + s); };
  auto b_empty = [&](int s) { return bar0 + 8u * (3 * kMaxSlots + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (4 * kMaxSlots + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (4 * kMaxSlots + 2 + a); };
  const uint32_t bres_full = bar0 + 8u * (4 * kMaxSlots + 4);
  volatile uint32_t* tmem_ptr_smem =
      reinterpret_cast<volatile uint32_t*>(gbase + off_bar + 8 * (4 * kMaxSlots + 5));
  // relay counters: [0] halo slots landed, [1] weight slots landed, [2] accumulators drained
  volatile uint32_t* ctr = reinterpret_cast<volatile uint32_t*>(gbase + off_bar + 8 * (4 * kMaxSlots + 6));

  if (warp == kWarpProducer && lane == 0) {
    tma_prefetch_desc(&amap);
    tma_prefetch_desc(&bmap);
    tma_prefetch_desc(&omap);
  }
  if (warp == kWarpInit && lane == 0) {
    for (int s = 0; s < kMaxSlots; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), GPS);  // every MMA group that read the slot commits once
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), ISSUERS);  // every issuer commits after its last group of the tile
      mbar_init(tempty_bar(a), PAIR ? 256 : 128);   // PAIR: both CTAs' epilogue warpgroups arrive on the leader's
    }
    mbar_init(bres_full, 1);
    ctr[0] = 0;
    ctr[1] = 0;
    ctr[2] = 2;  // both accumulators start drained
    fence_mbar_init();
  }
  if (warp == kWarpAlloc) {
    if (PAIR) {
      tmem_alloc_2cta(sbase + off_bar + 8 * (4 * kMaxSlots + 5), 2 * ISSUERS * BLOCK_N);
      tmem_relinquish_2cta();
    } else {
      tmem_alloc(sbase + off_bar + 8 * (4 * kMaxSlots + 5), 2 * ISSUERS * BLOCK_N);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int tiles_img = p.tiles_w * p.tiles_h;
  // work items: tiles, or (PAIR) pairs of adjacent M tiles with the same N tile (the host guarantees an even count)
  const int total_tiles = PAIR ? (tiles_img * p.N / 2) * p.n_tiles : tiles_img * p.N * p.n_tiles;
  const int my_tiles = (total_tiles - vcta + vgrid - 1) / vgrid;
  const int slots_per_tile = 3 * p.kb;  // halo slots per tile

  // work item -> (n tile, image, w0, h0) of THIS CTA's tile, without integer divides
  auto decode = [&](int tile, int& n_tile, int& img, int& w0, int& h0) {
    uint32_t m_tile = fdiv(tile, p.div_ntiles);
    n_tile = tile - m_tile * p.n_tiles;
    if (PAIR) m_tile = 2 * m_tile + rank;
    img = fdiv(m_tile, p.div_tiles_img);
    const uint32_t rem = m_tile - img * tiles_img;
    const uint32_t ty = fdiv(rem, p.div_tiles_w);
    w0 = (rem - ty * p.tiles_w) * 8;
    h0 = ty * 16;
  };

  if (warp == kWarpProducer) {
    // ------------------------------------------------------------------ TMA producer (one lane)
    if (lane == 0) {
      if (RESIDENT) {
        // n_tiles == 1 in this mode: the CTA's weight slice never changes (PAIR: this CTA's half of its rows;
        // the leader's barrier counts both halves)
        if (!PAIR) mbar_expect_tx(bres_full, b_region);
        else if (leader) mbar_expect_tx(bres_full, 2 * b_region);
        for (int t = 0; t < 9; ++t)
          for (int kb = 0; kb < p.kb; ++kb) {
            if (PAIR)
              tma_load_2d_2cta(sbase + off_b + (t * p.kb + kb) * B_BYTES, &bmap, bres_full, t * p.cin + kb * 64,
                               rank * (BLOCK_N / 2));
            else
              tma_load_2d(sbase + off_b + (t * p.kb + kb) * B_BYTES, &bmap, bres_full,
                          t * p.cin + kb * 64, 0);
          }
      }
      int sa = 0, pa = 0, sb = 0, pb = 0;
      long long t_wait = 0, t_begin = clock64();
      for (int tile = vcta; tile < total_tiles; tile += vgrid) {
        int n_tile, img, w0, h0;
        decode(tile, n_tile, img, w0, h0);
        for (int kb = 0; kb < p.kb; ++kb) {
          for (int dwi = 0; dwi < 3; ++dwi) {
            const long long tw0 = p.prof ? clock64() : 0;
            mbar_wait_relaxed(a_empty(sa), pa ^ 1, 1, &g_dbg_word);
            if (p.prof) t_wait += clock64() - tw0;
            if (PAIR) {
              // the leader's barrier counts both CTAs' bytes; the peer only issues its load
              if (leader) mbar_expect_tx(a_full(sa), 2 * kHaloBytes);
              tma_load_4d_2cta(sbase + off_a + sa * kHaloBytes, &amap, a_full(sa), kb * 64, w0 + dwi - 1, h0 - 1, img);
            } else {
              mbar_expect_tx(a_full(sa), kHaloBytes);
              tma_load_4d(sbase + off_a + sa * kHaloBytes, &amap, a_full(sa), kb * 64, w0 + dwi - 1,
                          h0 - 1, img);
            }
            if (++sa == p.a_slots) { sa = 0; pa ^= 1; }
            if (!RESIDENT) {
#pragma unroll
              for (int g = 0; g < GPS; ++g) {
                mbar_wait_relaxed(b_empty(sb), pb ^ 1, 8, &g_dbg_word);
                if (PAIR) {
                  if (leader) mbar_expect_tx(b_full(sb), 2 * BSLOT_BYTES);
                } else {
                  mbar_expect_tx(b_full(sb), BSLOT_BYTES);
                }
#pragma unroll
                for (int t = 0; t < TPG; ++t) {
                  const int dhi = g * TPG + t;
                  if (PAIR)   // this CTA's half of the weight tile: rows [rank * 128, +128) of the 256
                    tma_load_2d_2cta(sbase + off_b + sb * BSLOT_BYTES + t * B_BYTES, &bmap, b_full(sb),
                                     (dhi * 3 + dwi) * p.cin + kb * 64, n_tile * BLOCK_N + rank * (BLOCK_N / 2));
                  else
                    tma_load_2d(sbase + off_b + sb * BSLOT_BYTES + t * B_BYTES, &bmap, b_full(sb),
                                (dhi * 3 + dwi) * p.cin + kb * 64, n_tile * BLOCK_N);
                }
                if (++sb == p.b_slots) { sb = 0; pb ^= 1; }
              }
            }
          }
        }
      }
      if (p.prof) {
        p.prof[blockIdx.x * 8 + 0] = clock64() - t_begin;  // producer: total
        p.prof[blockIdx.x * 8 + 1] = t_wait;               // producer: waiting for a free halo slot
      }
    }
  } else if (warp == kWarpInit) {
    // ------------------------------------------------------------------ relay: halo slots full
    if (lane == 0 && leader) {
      int sa = 0, pa = 0;
      const uint32_t n = static_cast<uint32_t>(my_tiles) * slots_per_tile;
      for (uint32_t i = 0; i < n; ++i) {
        mbar_wait(a_full(sa), pa, 11, &g_dbg_word);
        ctr[0] = i + 1;
        if (++sa == p.a_slots) { sa = 0; pa ^= 1; }
      }
    }
  } else if (warp == kWarpRelayB) {
    // ------------------------------------------------------------------ relay: weight slots full
    if (lane == 0 && leader) {
      if (RESIDENT) {
        mbar_wait(bres_full, 0, 9, &g_dbg_word);
        ctr[1] = 1;
      } else {
        int sb = 0, pb = 0;
        const uint32_t n = static_cast<uint32_t>(my_tiles) * slots_per_tile * GPS;
        for (uint32_t i = 0; i < n; ++i) {
          mbar_wait(b_full(sb), pb, 12, &g_dbg_word);
          ctr[1] = i + 1;
          if (++sb == p.b_slots) { sb = 0; pb ^= 1; }
        }
      }
    }
  } else if (warp == kWarpAlloc) {
    // ------------------------------------------------------------------ relay: accumulators drained
    if (lane == 0 && leader) {
      for (int it = 0; it < my_tiles; ++it) {
        // tile `it` may start once the epilogue drained the previous user of accumulator it & 1
        mbar_wait(tempty_bar(it & 1), ((it >> 1) & 1) ^ 1, 13, &g_dbg_word);
        if (it >= 2) ctr[2] = it + 1;
      }
    }
  } else if (warp == kWarpMma0 || warp == kWarpMma1) {
    // ------------------------------------------------------------------ MMA issuers (one lane each)
    if (lane == 0 && leader && (warp == kWarpMma0 || ISSUERS == 2)) {
      const uint32_t me = warp == kWarpMma0 ? 0u : 1u;
      constexpr uint32_t idesc = umma_idesc_bf16(PAIR ? 256 : 128, BLOCK_N, 0, 0);
      constexpr uint32_t hi = umma_desc_hi_sw128(1024);
      const long long t_begin = clock64();
      const uint32_t a_lo_base = umma_desc_lo(sbase + off_a, 16);
      const uint32_t b_lo_base = umma_desc_lo(sbase + off_b, 16);
      constexpr uint32_t kHaloUnits = kHaloBytes >> 4, kBUnits = B_BYTES >> 4;
      const uint32_t b_tap_step = 3 * p.kb * kBUnits;  // resident: next vertical tap
      if (RESIDENT) wait_counter(ctr + 1, 1, 1);
      uint32_t G = 0;   // global group index; this thread owns the groups with (G & 1) == me
      uint32_t A = 0;   // global halo-slot index
      int sa = 0, sb = 0;
      for (int it = 0; it < my_tiles; ++it) {
        // accumulator (tile parity, issuer): each issuer owns one TMEM tile per tile parity
        const uint32_t d_tmem = tmem_base + ((it & 1) * ISSUERS + me) * BLOCK_N;
        bool entered = false;  // has this thread started accumulating tile `it` yet?
        for (int kb = 0; kb < p.kb; ++kb) {
#pragma unroll 1
          for (int dwi = 0; dwi < 3; ++dwi, ++A) {
            const uint32_t a_lo = a_lo_base + sa * kHaloUnits;
#pragma unroll
            for (int g = 0; g < GPS; ++g, ++G) {
              if (ISSUERS == 1 || (G & 1u) == me) {
                uint32_t accumulate = 1;
                if (!entered) {
                  wait_counter(ctr + 2, it + 1, 2);   // accumulators of this parity drained by the epilogue
                  accumulate = 0;                      // this thread's first group zeroes its accumulator
                  entered = true;
                }
                wait_counter(ctr + 0, A + 1, 3);
                uint32_t b_lo;
                if (RESIDENT) {
                  b_lo = b_lo_base + (dwi * p.kb + kb) * kBUnits;
                } else {
                  wait_counter(ctr + 1, G + 1, 4);
                  b_lo = b_lo_base + sb * (BSLOT_BYTES >> 4);
                }
                tc_fence_after();
#pragma unroll
                for (int t = 0; t < TPG; ++t) {
                  const int dhi = g * TPG + t;   // vertical tap = one 1024-byte atom further into the halo
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    if (PAIR)
                      umma_bf16_lohi_2cta(d_tmem, a_lo + dhi * 64 + 2 * k, hi, b_lo + 2 * k, hi, idesc, accumulate);
                    else
                      umma_bf16_lohi(d_tmem, a_lo + dhi * 64 + 2 * k, hi, b_lo + 2 * k, hi, idesc, accumulate);
                    accumulate = 1;
                  }
                  b_lo += RESIDENT ? b_tap_step : kBUnits;
                }
                if (PAIR) {   // both CTAs' producers get their slots back
                  umma_commit_2cta(b_empty(sb), 3u);
                  umma_commit_2cta(a_empty(sa), 3u);
                } else {
                  if (!RESIDENT) umma_commit(b_empty(sb));
                  umma_commit(a_empty(sa));
                }
              }
              if (!RESIDENT) {
                if (++sb == p.b_slots) sb = 0;
              }
            }
            if (++sa == p.a_slots) sa = 0;
          }
        }
        // arrives once this thread's MMAs of the tile are complete (PAIR: in both CTAs)
        if (PAIR) umma_commit_2cta(tfull_bar(it & 1), 3u);
        else umma_commit(tfull_bar(it & 1));
      }
      if (p.prof && me == 0) p.prof[blockIdx.x * 8 + 2] = clock64() - t_begin;  // MMA issuer: total
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------------ epilogue: two warpgroups,
    // warpgroup g drains accumulator g (tiles it = g, g+2, ...), so each has two tile times per tile.
    const int g = warp >> 2;
    const int q = warp & 3;  // TMEM lane quadrant this warp may read
    const int row = q * 32 + lane;
    const int et = threadIdx.x - g * 128;
    const uint32_t bar_a = 1 + 2 * g, bar_b = 2 + 2 * g;
    float* s_scale = reinterpret_cast<float*>(gbase + off_param) + g * 4 * BLOCK_N;
    float* s_shift = s_scale + BLOCK_N;
    float* s_sum = s_shift + BLOCK_N;
    float* s_sq = s_sum + BLOCK_N;
    uint8_t* stg = gbase + off_stg + g * STG_BYTES;
    const uint32_t stg_s = sbase + off_stg + g * STG_BYTES;
    const bool do_stats = p.stat_sum != nullptr;
    const bool relu = p.relu != 0;
    int last_n_tile = -1;
    int it = g;
    long long t_epi_wait = 0;
    const long long t_epi_begin = clock64();
    // BatchNorm statistics: a tile's column sums are formed in fp32 (shared memory), then carried across this
    // warpgroup's tiles in fp64 REGISTERS (thread et owns channels et, et + 128) and added to the global fp64
    // accumulators once per run of tiles with the same channel slice -- normally once per CTA.  One global
    // atomic per channel per TILE (16 K tiles x 128 addresses on the 64-wide layers) doubled those layers'
    // time when the accumulators became fp64 (340 vs 170 us).
    constexpr int ACC_PER_THREAD = (BLOCK_N + 127) / 128;
    double acc_s[ACC_PER_THREAD], acc_q[ACC_PER_THREAD];
#pragma unroll
    for (int i = 0; i < ACC_PER_THREAD; ++i) acc_s[i] = acc_q[i] = 0.0;
    int acc_ch0 = -1;
    auto flush_stats = [&]() {
      if (acc_ch0 < 0) return;
#pragma unroll
      for (int i = 0; i < ACC_PER_THREAD; ++i) {
        const int c = et + 128 * i;
        if (c < BLOCK_N) {
          if (p.det_part) {
            // deterministic: this warpgroup's own row of the scratch (zeroed by the launcher); thread et is the
            // only writer of its channels, flushes of one warpgroup are sequential
            double* rowp = p.det_part + (2 * blockIdx.x + g) * 2ll * p.cout;
            rowp[acc_ch0 + c] += acc_s[i];
            rowp[p.cout + acc_ch0 + c] += acc_q[i];
          } else {
            atomicAdd(p.stat_sum + acc_ch0 + c, acc_s[i]);
            atomicAdd(p.stat_sq + acc_ch0 + c, acc_q[i]);
          }
        }
        acc_s[i] = acc_q[i] = 0.0;
      }
    };
    for (int tile = vcta + g * vgrid; tile < total_tiles; tile += 2 * vgrid, it += 2) {
      int n_tile, img, w0, h0;
      decode(tile, n_tile, img, w0, h0);
      const int acc_phase = (it >> 1) & 1;
      const int ch0 = n_tile * BLOCK_N;
      if (do_stats && ch0 != acc_ch0) {
        flush_stats();
        acc_ch0 = ch0;
      }
      if (n_tile != last_n_tile) {  // per-channel parameters change only with the N tile
        for (int c = et; c < BLOCK_N; c += 128) {
          s_scale[c] = p.scale ? p.scale[ch0 + c] : 1.0f;
          s_shift[c] = p.shift ? p.shift[ch0 + c] : 0.0f;
        }
        last_n_tile = n_tile;
      }
      if (do_stats)
        for (int c = et; c < BLOCK_N; c += 128) {
          s_sum[c] = 0.0f;
          s_sq[c] = 0.0f;
        }
      const bool full_tile = (w0 + 8 <= p.W) && (h0 + 16 <= p.H);
      const bool valid = (w0 + (row & 7) < p.W) && (h0 + (row >> 3) < p.H);
      {
        const long long tw0 = p.prof ? clock64() : 0;
        mbar_wait_relaxed(tfull_bar(g), acc_phase, 4, &g_dbg_word);
        if (p.prof && et == 0) t_epi_wait += clock64() - tw0;
      }
      tc_fence_after();
#pragma unroll 1
      for (int chunk = 0; chunk < BLOCK_N / 64; ++chunk) {
        uint32_t v0[32], v1[32];
        {
          // the issuers' partial accumulators of this tile parity sit side by side; add them in a fixed order
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                 g * ISSUERS * BLOCK_N + chunk * 64;
          tmem_ld_32x32b_x32(taddr, v0);
          tmem_ld_32x32b_x32(taddr + 32, v1);
          if (ISSUERS == 2) {
            uint32_t w0r[32];
            tmem_ld_32x32b_x32(taddr + BLOCK_N, w0r);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) v0[e] = __float_as_uint(__uint_as_float(v0[e]) + __uint_as_float(w0r[e]));
            tmem_ld_32x32b_x32(taddr + BLOCK_N + 32, w0r);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) v1[e] = __float_as_uint(__uint_as_float(v1[e]) + __uint_as_float(w0r[e]));
          } else {
            tmem_ld_wait();
          }
        }
        if (chunk == BLOCK_N / 64 - 1) {
          tc_fence_before();
          if (PAIR) mbar_arrive_leader(tempty_bar(g));   // the issuer lives in the leader CTA
          else mbar_arrive(tempty_bar(g));
        }
        const float* sc = s_scale + chunk * 64;
        const float* sh = s_shift + chunk * 64;
#pragma unroll 1
        for (int half = 0; half < (HALF_STAGE ? 2 : 1); ++half) {
          if (et == 0) tma_store_wait_read<0>();  // the previous store has finished reading the buffer
          named_bar_sync(bar_a, 128);
          const bool mine = !HALF_STAGE || (row >> 6) == half;
          const int srow = HALF_STAGE ? (row & 63) : row;
          if (mine) {
            if (full_tile) {
              if (relu) {
                epi_half<true, false>(v0, 0, sc, sh, true, stg, srow);
                epi_half<true, false>(v1, 4, sc + 32, sh + 32, true, stg, srow);
              } else {
                epi_half<false, false>(v0, 0, sc, sh, true, stg, srow);
                epi_half<false, false>(v1, 4, sc + 32, sh + 32, true, stg, srow);
              }
            } else {
              if (relu) {
                epi_half<true, true>(v0, 0, sc, sh, valid, stg, srow);
                epi_half<true, true>(v1, 4, sc + 32, sh + 32, valid, stg, srow);
              } else {
                epi_half<false, true>(v0, 0, sc, sh, valid, stg, srow);
                epi_half<false, true>(v1, 4, sc + 32, sh + 32, valid, stg, srow);
              }
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(bar_b, 128);
          if (et == 0) {
            tma_store_4d(&omap, stg_s, ch0 + chunk * 64, w0, h0 + (HALF_STAGE ? 8 * half : 0), img);
            tma_store_commit();
          }
          if (do_stats) {
            // column sums of the rounded outputs: lane <-> channel pair, warp <-> row group
            constexpr int RPW = STG_ROWS / 4;  // staged rows per warp
            const int j = lane >> 2, wsub = lane & 3;
            float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll 8
            for (int rr = 0; rr < RPW; ++rr) {
              const int r = q * RPW + rr;
              const uint32_t u =
                  *reinterpret_cast<const uint32_t*>(stg + r * 128 + ((j ^ (r & 7)) << 4) + wsub * 4);
              const float2 f = unpack_bf16x2(u);
              s0 += f.x;
              s1 += f.y;
              q0 = fmaf(f.x, f.x, q0);
              q1 = fmaf(f.y, f.y, q1);
            }
            if (p.det_part) {
              // deterministic: the four warps add their row groups' sums one after the other
              for (int turn = 0; turn < 4; ++turn) {
                if (q == turn) {
                  s_sum[chunk * 64 + 2 * lane] += s0;
                  s_sum[chunk * 64 + 2 * lane + 1] += s1;
                  s_sq[chunk * 64 + 2 * lane] += q0;
                  s_sq[chunk * 64 + 2 * lane + 1] += q1;
                }
                named_bar_sync(bar_a, 128);
              }
            } else {
              atomicAdd(&s_sum[chunk * 64 + 2 * lane], s0);
              atomicAdd(&s_sum[chunk * 64 + 2 * lane + 1], s1);
              atomicAdd(&s_sq[chunk * 64 + 2 * lane], q0);
              atomicAdd(&s_sq[chunk * 64 + 2 * lane + 1], q1);
            }
          }
        }
      }
      if (do_stats) {
        named_bar_sync(bar_a, 128);
#pragma unroll
        for (int i = 0; i < ACC_PER_THREAD; ++i) {
          const int c = et + 128 * i;
          if (c < BLOCK_N) {   // the same thread zeroes s_sum[c] at the top of the next tile
            acc_s[i] += static_cast<double>(s_sum[c]);
            acc_q[i] += static_cast<double>(s_sq[c]);
          }
        }
      }
    }
    if (do_stats) flush_stats();
    if (et == 0) tma_store_wait_all<0>();
    if (p.prof && et == 0 && g == 0) {
      p.prof[blockIdx.x * 8 + 6] = clock64() - t_epi_begin;  // epilogue warpgroup 0: total
      p.prof[blockIdx.x * 8 + 7] = t_epi_wait;               // waiting for a finished accumulator
    }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) {
    cluster_sync_all();   // the leader's MMAs read the peer's shared memory and commit to its barriers until the end
    if (warp == kWarpAlloc) tmem_dealloc_2cta(tmem_base, 2 * ISSUERS * BLOCK_N);
  } else {
    if (warp == kWarpAlloc) tmem_dealloc(tmem_base, 2 * ISSUERS * BLOCK_N);
  }
}

static size_t conv3_smem_bytes(int block_n, int a_slots, int b_tiles, bool half_stage = false, bool pair = false) {
  return static_cast<size_t>(a_slots) * kHaloBytes + static_cast<size_t>(b_tiles) * (pair ? block_n / 2 : block_n) * 128 +
         2 * (half_stage ? 8192 : 16384) + 2 * 4 * block_n * 4 + 8 * (4 * kMaxSlots + 8) + 16 + 1024;
}

// CTA-pair launch: clusters of two CTAs, an even grid, work items = tile pairs.
template <int BLOCK_N, int MODE, bool HALF_STAGE>
static int launch_conv3_pair(const CUtensorMap& amap, const CUtensorMap& bmap, const CUtensorMap& omap,
                             const Conv3Params& p, size_t smem, int total_pairs, cudaStream_t stream) {
  auto kern = igemm_conv3_kernel<BLOCK_N, MODE, HALF_STAGE, true>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] {
    attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  });
  if (attr_err != cudaSuccess) {
    set_error(std::string("cudaFuncSetAttribute(igemm_conv3 pair): ") + cudaGetErrorString(attr_err));
    return -2;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2u * static_cast<unsigned>(std::min(total_pairs, num_sms() / 2)), 1, 1);
  cfg.blockDim = dim3(kConv3Threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, amap, bmap, omap, p);
  if (e != cudaSuccess) {
    set_error(std::string("igemm_conv3 pair launch: ") + cudaGetErrorString(e));
    return -3;
  }
  return 0;
}

template <int BLOCK_N, int MODE, bool HALF_STAGE>
static int launch_conv3_inst(const CUtensorMap& amap, const CUtensorMap& bmap, const CUtensorMap& omap,
                             const Conv3Params& p, size_t smem, int total_tiles, cudaStream_t stream) {
  auto kern = igemm_conv3_kernel<BLOCK_N, MODE, HALF_STAGE>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] {
    attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  });
  if (attr_err != cudaSuccess) {
    set_error(std::string("cudaFuncSetAttribute(igemm_conv3): ") + cudaGetErrorString(attr_err));
    return -2;
  }
  const int grid = std::min(total_tiles, num_sms());
  kern<<<grid, kConv3Threads, smem, stream>>>(amap, bmap, omap, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error(std::string("igemm_conv3 launch: ") + cudaGetErrorString(e));
    return -3;
  }
  return 0;
}

// conv3x3 fwd / dgrad through the halo-reuse kernel.  Returns 1 if the shape is not eligible (caller
// falls back to the generic forward kernel), 0 on success, < 0 on error.
static int try_launch_conv3(const FwdDesc& d) {
  static const bool disabled = getenv("PLUME_DISABLE_CONV3") != nullptr;
  if (disabled) return 1;
  if (d.num_taps != 9 || d.num_in_views != 1 || d.num_out_views != 1) return 1;
  const int W = d.in[0].W, H = d.in[0].H, N = d.in[0].N;
  if (H < 16 || W < 8) return 1;
  const int cout = d.cout_per_view;
  const int block_n = (cout % 256 == 0) ? 256 : (cout % 128 == 0 ? 128 : 64);
  const int n_tiles = cout / block_n;
  const int kb = d.Cin / 64;
  const size_t limit = 232448;

  Conv3Params p;
  p.tiles_w = (W + 7) / 8;
  p.tiles_h = (H + 15) / 16;
  p.W = W; p.H = H; p.N = N;
  p.n_tiles = n_tiles;
  p.kb = kb;
  p.cin = d.Cin;
  p.div_ntiles = make_fastdiv(n_tiles);
  p.div_tiles_img = make_fastdiv(p.tiles_w * p.tiles_h);
  p.div_tiles_w = make_fastdiv(p.tiles_w);
  p.scale = d.scale; p.shift = d.shift; p.relu = d.relu;
  p.stat_sum = d.stat_sum; p.stat_sq = d.stat_sq;
  p.det_part = nullptr;
  p.cout = cout;
  p.prof = g_prof_buf;

  // mode 0: weights resident when the whole [block_n][9*Cin] slice fits beside >= 4 halo slots (a slot is
  //         12 MMAs of work; fewer cannot cover the TMA latency)
  // mode 1: weight triples (block_n <= 128), mode 2: single weight tiles (block_n == 256)
  // (PLUME_CONV3_RESIDENT_HALF=0 disables the second form: A/B switch)
  static const bool resident_half = !(getenv("PLUME_CONV3_RESIDENT_HALF") && atoi(getenv("PLUME_CONV3_RESIDENT_HALF")) == 0);
  // CTA pairs (tcgen05.mma.cta_group::2) when the M tiles pair up: each CTA holds HALF of every weight tile, so weight
  // slices that are resident halve, ring slots halve, and an MMA reads half of B per SM.  PLUME_CONV3_PAIR: 0 = never,
  // 1 = only the 256-wide ring mode, 2 (default) = every mode.
  static const int pair_level = getenv("PLUME_CONV3_PAIR") ? atoi(getenv("PLUME_CONV3_PAIR")) : 2;
  const long long m_tiles = 1ll * p.tiles_w * p.tiles_h * N;
  const bool can_pair = m_tiles % 2 == 0 && m_tiles * n_tiles >= 4;
  bool pair = can_pair && ((pair_level >= 2) || (pair_level == 1 && block_n == 256));
  auto fits = [&](int a_slots, int b_tiles_, bool half) {
    return conv3_smem_bytes(block_n, a_slots, b_tiles_, half, pair) <= limit;
  };
  int mode, b_tiles;
  bool half_stage = false;
  if (n_tiles == 1 && block_n <= 128 && fits(4, 9 * kb, false)) {
    mode = 0;
    p.b_slots = 0;
    b_tiles = 9 * kb;
    p.a_slots = 4;
  } else if (resident_half && !pair && n_tiles == 1 && block_n == 128 && fits(3, 9 * kb, true)) {
    // single CTAs, the 144 KB slice of a 64 -> 128 layer: resident beside three halo slots when the staging buffers
    // are halved; streaming it through the ring instead re-writes the weights into shared memory for every tile.
    // Measured on one box (gpurun_out/r2u): 64 -> 128 @128x128 125 -> 117 us, @256x256 (a dgrad) 317 -> 290 us.  The
    // same for 128 -> 64 (64-wide tiles, kb = 2) is SLOWER, 349 -> 402 us: three halo slots are 1.7 k cycles of
    // 48-cycle MMAs, too little look-ahead, where the ring configuration affords six -- so 128-wide tiles only.  As a
    // CTA pair the 64 -> 128 slice fits with full staging (branch above); for the 128 -> 128 slice (144 KB per CTA of
    // a pair) the ring with six halo slots wins again: 146 -> 133 us forward, 139 -> 116 us dgrad (gpurun_out/r2z).
    mode = 0;
    half_stage = true;
    p.b_slots = 0;
    b_tiles = 9 * kb;
    p.a_slots = 3;
  } else if (block_n <= 128) {
    mode = 1;
    p.b_slots = block_n == 128 ? 3 : 4;   // triples of 48 KB / 24 KB (half of that per CTA of a pair)
    b_tiles = 3 * p.b_slots;
    p.a_slots = 3;
  } else {
    mode = 2;
    p.b_slots = pair ? 6 : 4;
    b_tiles = p.b_slots;
    p.a_slots = pair ? 4 : 3;
  }
  if (mode == 1 && block_n == 128) half_stage = true;   // selects the HALF_STAGE instantiation below
  if (mode == 2 && pair) {
    while (p.b_slots < kMaxSlots && fits(p.a_slots, b_tiles + 1, false)) { ++p.b_slots; ++b_tiles; }
  } else {
    while (p.a_slots < 6 && fits(p.a_slots + 1, b_tiles, half_stage)) ++p.a_slots;
    if (mode == 1 && pair)   // the halved triples leave room for a deeper weight ring
      while (p.b_slots < kMaxSlots && fits(p.a_slots, b_tiles + 3, half_stage)) { ++p.b_slots; b_tiles += 3; }
  }
  const size_t smem = conv3_smem_bytes(block_n, p.a_slots, b_tiles, half_stage, pair);
  if (smem > limit) return 1;

  CUtensorMap amap, bmap, omap;
  if (d.in[0].C > d.Cin || d.in[0].C % 8 || d.out[0].C != cout) {  // fewer input channels: TMA zero-fills
    set_error("igemm_conv3: view channel counts do not match");
    return -1;
  }
  if (make_act_map(&amap, d.in[0], 64, 8, 18, 1)) return -1;
  if (make_act_map(&omap, d.out[0], 64, 8, half_stage ? 8 : 16, 1)) return -1;
  if (make_mat_map(&bmap, d.wmat, cout, 9ll * d.Cin, 64, pair ? block_n / 2 : block_n)) return -1;
  const long long total = 1ll * p.tiles_w * p.tiles_h * N * n_tiles;
  if (total > 0x7fffffffll) {
    set_error("igemm_conv3: too many tiles");
    return -1;
  }
  const int tt = static_cast<int>(total);
  const int det_rows = 2 * std::min(tt, num_sms()) + 2;   // one row per (CTA, epilogue warpgroup)
  if (d.stat_sum && deterministic()) {
    p.det_part = det_stats_begin(det_rows, cout, d.stream);
    if (!p.det_part) return -2;
  }
  int r;
  if (pair) {
    const int tp = tt / 2;
    if (mode == 0 && half_stage)
      r = block_n == 128 ? launch_conv3_pair<128, 0, true>(amap, bmap, omap, p, smem, tp, d.stream)
                         : launch_conv3_pair<64, 0, true>(amap, bmap, omap, p, smem, tp, d.stream);
    else if (mode == 0)
      r = block_n == 128 ? launch_conv3_pair<128, 0, false>(amap, bmap, omap, p, smem, tp, d.stream)
                         : launch_conv3_pair<64, 0, false>(amap, bmap, omap, p, smem, tp, d.stream);
    else if (mode == 1)
      r = block_n == 128 ? launch_conv3_pair<128, 1, true>(amap, bmap, omap, p, smem, tp, d.stream)
                         : launch_conv3_pair<64, 1, false>(amap, bmap, omap, p, smem, tp, d.stream);
    else
      r = launch_conv3_pair<256, 2, false>(amap, bmap, omap, p, smem, tp, d.stream);
  } else if (mode == 0 && half_stage) {
    r = block_n == 128 ? launch_conv3_inst<128, 0, true>(amap, bmap, omap, p, smem, tt, d.stream)
                       : launch_conv3_inst<64, 0, true>(amap, bmap, omap, p, smem, tt, d.stream);
  } else if (mode == 0) {
    r = block_n == 128 ? launch_conv3_inst<128, 0, false>(amap, bmap, omap, p, smem, tt, d.stream)
                       : launch_conv3_inst<64, 0, false>(amap, bmap, omap, p, smem, tt, d.stream);
  } else if (mode == 1) {
    r = block_n == 128 ? launch_conv3_inst<128, 1, true>(amap, bmap, omap, p, smem, tt, d.stream)
                       : launch_conv3_inst<64, 1, false>(amap, bmap, omap, p, smem, tt, d.stream);
  } else {
    r = launch_conv3_inst<256, 2, false>(amap, bmap, omap, p, smem, tt, d.stream);
  }
  if (r == 0 && p.det_part) r = det_stats_finish(p.det_part, det_rows, cout, d.stat_sum, d.stat_sq, d.stream);
  return r;
}

int launch_igemm_fwd(const FwdDesc& d) {
  if (d.Cin % 64 != 0 || d.Cin <= 0) {
    set_error("igemm_fwd: input channels must be a positive multiple of 64");
    return -1;
  }
  if (d.cout_per_view % 64 != 0 || d.cout_per_view <= 0) {
    set_error("igemm_fwd: output channels must be a positive multiple of 64");
    return -1;
  }
  const int W = d.in[0].W, H = d.in[0].H, N = d.in[0].N;
  if (W <= 0 || H <= 0 || N <= 0) {
    set_error("igemm_fwd: empty input");
    return -1;
  }
  for (int i = 0; i < d.num_out_views; ++i)
    if (d.out[i].W != W || d.out[i].H != H || d.out[i].N != N) {
      set_error("igemm_fwd: output view extents differ from the input extents");
      return -1;
    }
  if (!d.split) {
    const int r = try_launch_conv3(d);
    if (r <= 0) return r;
  }
  int bw, bh, bn;
  pick_box(W, H, 128, 16, &bw, &bh, &bn);
  // 64-channel views (the first decoder level's transposed conv): one 256-wide tile covers all four views, so
  // the input tile is read once instead of four times (153 -> 137 us); wider views keep one view per tile
  // (spanning measured slower there: 72 -> 76 us)
  const int block_n = (d.cout_per_view == 64 && d.num_out_views == 4)
                          ? 256
                          : ((d.cout_per_view % 256 == 0) ? 256 : (d.cout_per_view % 128 == 0 ? 128 : 64));

  TmapPack8 amaps, omaps;
  CUtensorMap bmap;
  for (int i = 0; i < 4; ++i) {
    const ActView& v = d.in[i < d.num_in_views ? i : 0];
    if (v.C > d.Cin || v.C % 8) {  // fewer channels than the weights' K per tap: TMA zero-fills the rest
      set_error("igemm_fwd: input view channel count > Cin");
      return -1;
    }
    if (make_act_map(&amaps.m[i], v, 64, bw, bh, bn)) return -1;
    const ActView& o = d.out[i < d.num_out_views ? i : 0];
    if (o.C != d.cout_per_view) {
      set_error("igemm_fwd: output view channel count != cout_per_view");
      return -1;
    }
    if (make_act_map(&omaps.m[i], o, 64, bw, bh, bn)) return -1;
    if (d.split) {
      if (v.plane <= 0 || o.plane <= 0) {
        set_error("igemm_fwd: bf16x3 views need their lo-plane offset");
        return -1;
      }
      if (make_act_map(&amaps.m[4 + i], lo_plane(v), 64, bw, bh, bn)) return -1;
      if (make_act_map(&omaps.m[4 + i], lo_plane(o), 64, bw, bh, bn)) return -1;
    } else {
      amaps.m[4 + i] = amaps.m[i];
      omaps.m[4 + i] = omaps.m[i];
    }
  }
  const long long n_total = static_cast<long long>(d.num_out_views) * d.cout_per_view;
  const long long k_total = static_cast<long long>(d.num_taps) * d.Cin;
  if (make_mat_map(&bmap, d.wmat, d.split ? 2 * n_total : n_total, k_total, 64, block_n)) return -1;

  FwdParams p;
  p.TW = bw; p.TH = bh; p.TN = bn;
  p.W = W; p.H = H; p.N = N;
  p.tiles_w = (W + bw - 1) / bw;
  p.tiles_h = (H + bh - 1) / bh;
  p.tiles_n = (N + bn - 1) / bn;
  p.n_tiles = static_cast<int>(n_total / block_n);
  p.num_taps = d.num_taps;
  p.per_tap_view = d.num_in_views > 1 ? 1 : 0;
  p.kb_per_tap = d.Cin / 64;
  p.k_per_tap = d.Cin;
  p.cout_per_view = d.cout_per_view;
  p.n_total = static_cast<int>(n_total);
  p.scale = d.scale;
  p.shift = d.shift;
  p.relu = d.relu;
  p.stat_sum = d.stat_sum;
  p.stat_sq = d.stat_sq;
  p.det_part = nullptr;
  p.det_c = d.cout_per_view;
  const long long total = 1ll * p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles;
  if (total > 0x7fffffffll) {
    set_error("igemm_fwd: too many tiles");
    return -1;
  }
  const int det_rows = 2 * std::min(static_cast<int>(total), num_sms());   // (CTA, epilogue warpgroup), at most two
  if (d.stat_sum && deterministic()) {
    p.det_part = det_stats_begin(det_rows, p.det_c, d.stream);
    if (!p.det_part) return -2;
  }
  static const bool one_wg = getenv("PLUME_FWD_ONE_EPILOGUE_WG") != nullptr;   // A/B switch (scripts/time_convT.py)
  int r;
  if (d.split) {
    r = block_n == 256   ? launch_fwd_inst<256, 3, 2, true>(amaps, bmap, omaps, p, (int)total, d.stream)
        : block_n == 128 ? launch_fwd_inst<128, 4, 2, true>(amaps, bmap, omaps, p, (int)total, d.stream)
                         : launch_fwd_inst<64, 4, 2, true>(amaps, bmap, omaps, p, (int)total, d.stream);
  } else if (one_wg) {
    r = block_n == 256   ? launch_fwd_inst<256, 4, 1>(amaps, bmap, omaps, p, (int)total, d.stream)
        : block_n == 128 ? launch_fwd_inst<128, 5, 2>(amaps, bmap, omaps, p, (int)total, d.stream)
                         : launch_fwd_inst<64, 6, 2>(amaps, bmap, omaps, p, (int)total, d.stream);
  } else {
    r = block_n == 256   ? launch_fwd_inst<256, 3, 2, false, 2>(amaps, bmap, omaps, p, (int)total, d.stream)
        : block_n == 128 ? launch_fwd_inst<128, 4, 2, false, 2>(amaps, bmap, omaps, p, (int)total, d.stream)
                         : launch_fwd_inst<64, 6, 2, false, 2>(amaps, bmap, omaps, p, (int)total, d.stream);
  }
  if (r == 0 && p.det_part) r = det_stats_finish(p.det_part, det_rows, p.det_c, d.stat_sum, d.stat_sq, d.stream);
  return r;
}

// =================================================================================================
// Weight-gradient kernel
// =================================================================================================
struct WgradParams {
  int tiles_w, tiles_h, tiles_n;  // pixel tiling (64 pixels per K block)
  int PW, PH, PN;
  int num_taps, per_tap_dy;       // per_tap_dy: tap t reads dY view t and X is unshifted (convT)
  int cin, cout;
  int pair_taps;                  // cin == 64: an M tile stacks two taps
  int ci_tiles;                   // cin / 128 when !pair_taps
  int m_units, n_tiles;
  int splits, ktiles_per_split;
  float* dw;                      // fp32 weight gradient, element (tap t, ci, co) at co*s_co + t*s_t + ci
  long long s_co, s_t;
  float* det_part;                // deterministic mode: [splits][dw_numel] partial sums (stored, not added)
  long long dw_numel;
};

template <int BLOCK_N, int STAGES>
struct WgSmem {
  static constexpr int A_BYTES = 2 * 64 * 128;              // two boxes of 64 pixels x 64 channels
  static constexpr int B_BYTES = (BLOCK_N / 64) * 64 * 128; // BLOCK_N/64 such boxes
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int OFF_BAR = STAGES * STAGE_BYTES;
  static constexpr int NUM_BARS = 2 * STAGES + 1;
  static constexpr int OFF_TMEMPTR = OFF_BAR + NUM_BARS * 8;
  static constexpr int TOTAL = OFF_TMEMPTR + 16 + 1024;
  static_assert(TOTAL <= 232448, "shared memory budget exceeded");
};

// SPLIT (bf16x3 mode): X and dY are hi/lo plane pairs (xmaps.m[0/1], dymaps.m[0-3 / 4-7]); the K loop runs over
// the CTA's pixel tiles three times: x_hi*dy_hi + x_hi*dy_lo + x_lo*dy_hi, all into the same accumulator.
template <int BLOCK_N, int STAGES, bool SPLIT>
__global__ void __launch_bounds__(256, 1)
    igemm_wgrad_kernel(const __grid_constant__ TmapPack2 xmaps,
                       const __grid_constant__ TmapPack8 dymaps, const WgradParams p) {
  using L = WgSmem<BLOCK_N, STAGES>;
  constexpr int PASSES = SPLIT ? 3 : 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t sbase = (raw_addr + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (sbase - raw_addr);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t bar0 = sbase + L::OFF_BAR;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
  const uint32_t tfull_bar = bar0 + 8u * (2 * STAGES);
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(gbase + L::OFF_TMEMPTR);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&xmaps.m[0]);
    tma_prefetch_desc(&dymaps.m[0]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(sbase + L::OFF_TMEMPTR, BLOCK_N < 32 ? 32 : BLOCK_N);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // ---- work decomposition: blockIdx.x = (m unit, n tile), blockIdx.y = K split
  const int n_tile = blockIdx.x % p.n_tiles;
  const int unit = blockIdx.x / p.n_tiles;
  const int split = blockIdx.y;
  int tapA, tapB, cA0, cA1, row_base;
  if (p.pair_taps) {
    tapA = 2 * unit;
    tapB = min(2 * unit + 1, p.num_taps - 1);
    cA0 = 0;
    cA1 = 0;
    row_base = tapA * 64;
  } else {
    tapA = tapB = unit / p.ci_tiles;
    cA0 = (unit % p.ci_tiles) * 128;
    cA1 = cA0 + 64;
    row_base = tapA * p.cin + cA0;
  }
  const int total_ktiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int kt_begin = split * p.ktiles_per_split;
  const int kt_end = min(kt_begin + p.ktiles_per_split, total_ktiles);
  const int nk = kt_end - kt_begin;  // host guarantees nk >= 1

  if (warp == 0) {
    if (lane == 0) {
      int dwA = 0, dhA = 0, dwB = 0, dhB = 0;
      if (!p.per_tap_dy && p.num_taps == 9) {
        dhA = tapA / 3 - 1; dwA = tapA % 3 - 1;
        dhB = tapB / 3 - 1; dwB = tapB % 3 - 1;
      }
      const int dyv0 = p.per_tap_dy ? tapA : 0;
      int stage = 0, phase = 0;
      for (int pk = 0; pk < PASSES * nk; ++pk) {
        const int pass = SPLIT ? pk / nk : 0;
        const int kt = kt_begin + (SPLIT ? pk - pass * nk : pk);
        const CUtensorMap& xmap = xmaps.m[pass == 2 ? 1 : 0];   // x_lo in the third pass
        const int dyv = dyv0 + (pass == 1 ? 4 : 0);             // dy_lo in the second pass
        const int w0 = (kt % p.tiles_w) * p.PW;
        const int h0 = ((kt / p.tiles_w) % p.tiles_h) * p.PH;
        const int n0 = (kt / (p.tiles_w * p.tiles_h)) * p.PN;
        mbar_wait(empty_bar(stage), phase ^ 1, 5, &g_dbg_word);
        const uint32_t a_addr = sbase + stage * L::STAGE_BYTES;
        const uint32_t b_addr = a_addr + L::A_BYTES;
        mbar_expect_tx(full_bar(stage), L::STAGE_BYTES);
        tma_load_4d(a_addr, &xmap, full_bar(stage), cA0, w0 + dwA, h0 + dhA, n0);
        tma_load_4d(a_addr + 8192, &xmap, full_bar(stage), cA1, w0 + dwB, h0 + dhB, n0);
#pragma unroll
        for (int b = 0; b < BLOCK_N / 64; ++b)
          tma_load_4d(b_addr + b * 8192, &dymaps.m[dyv], full_bar(stage),
                      n_tile * BLOCK_N + b * 64, w0, h0, n0);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BLOCK_N, 1, 1);
      int stage = 0, phase = 0;
      for (int i = 0; i < PASSES * nk; ++i) {
        mbar_wait(full_bar(stage), phase, 6, &g_dbg_word);
        tc_fence_after();
        const uint32_t a_addr = sbase + stage * L::STAGE_BYTES;
        const uint32_t b_addr = a_addr + L::A_BYTES;
        // MN-major: 16 pixels (two 8-pixel atoms, SBO = 1024 B apart) per MMA; 64-channel groups
        // are LBO = 8192 B apart (one TMA box each); a K step is 2048 B = 128 descriptor units.
        constexpr uint32_t hi = umma_desc_hi_sw128(1024);
        const uint32_t a_lo = umma_desc_lo(a_addr, 8192), b_lo = umma_desc_lo(b_addr, 8192);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_lohi(tmem_base, a_lo + 128 * k, hi, b_lo + 128 * k, hi, idesc, (i | k) != 0 ? 1u : 0u);
        umma_commit(empty_bar(stage));
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(tfull_bar);
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    const int row = q * 32 + lane;
    const int grow = row_base + row;
    const bool row_ok = grow < p.num_taps * p.cin && (!p.pair_taps || row < 64 || tapB != tapA);
    mbar_wait_relaxed(tfull_bar, 0, 7, &g_dbg_word);
    tc_fence_after();
    // split-K partial sums are added straight into dW with fp32 reductions (RED.ADD): consecutive lanes are
    // consecutive input channels = consecutive addresses, so every warp instruction is one 128-byte line
    const int t = grow / p.cin, ci = grow - t * p.cin;
    float* base = p.det_part ? p.det_part + split * p.dw_numel : p.dw;
    float* dst = base + t * p.s_t + ci + static_cast<long long>(n_tile * BLOCK_N) * p.s_co;
#pragma unroll 1
    for (int c = 0; c < BLOCK_N; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, v);
      tmem_ld_wait();
      if (row_ok) {
        if (p.det_part) {   // deterministic: this split's own copy, summed over the splits in order afterwards
#pragma unroll
          for (int j = 0; j < 32; ++j) dst[(c + j) * p.s_co] = __uint_as_float(v[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(dst + (c + j) * p.s_co, __uint_as_float(v[j]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, BLOCK_N < 32 ? 32 : BLOCK_N);
}

static void wgrad_geometry(int N, int H, int W, int* bw, int* bh, int* bn, int* ktiles) {
  pick_box(W, H, 64, 16, bw, bh, bn);
  const int tw = (W + *bw - 1) / *bw, th = (H + *bh - 1) / *bh, tn = (N + *bn - 1) / *bn;
  *ktiles = tw * th * tn;
}

static int wgrad_block_n(int Cout) { return Cout % 256 == 0 ? 256 : (Cout % 128 == 0 ? 128 : 64); }

// One place decides which weight-gradient kernel runs and how K is split, so that what the C ABI reports
// (plume_wgrad_splits) always matches what the launcher does.
struct WgradConfig {
  bool halo;        // igemm_wgrad3_kernel (3x3, halo reuse) vs the generic igemm_wgrad_kernel
  int block_n;
  int ctas_mn;      // CTAs per split
  int ktiles;       // 64-pixel K tiles
  int splits, ktiles_per_split;
};

static int split_for(int ctas, int ktiles, int min_ktiles, int* per_out) {
  // Split K so that the CTAs fill whole waves of the 148 SMs: one CTA past a wave boundary costs a whole
  // extra wave (measured: 99 splits x 3 CTAs = 297 CTAs ran 401 us, 148 x 3 = 444 CTAs 299 us).  Cost of a
  // candidate in K-tile units = waves * (K tiles per CTA + kEpilogue), kEpilogue ~ pipeline fill + the
  // fp32 reductions of one CTA (about 6 us per extra wave measured on the 64 -> 64 layer).
  const int kSms = num_sms();
  constexpr int kEpilogue = 10, kMaxWaves = 4;
  static const int force_waves = [] {  // diagnostics: PLUME_WGRAD_WAVES=n -> the most splits that fit n waves
    const char* e = getenv("PLUME_WGRAD_WAVES");
    const int v = e ? atoi(e) : 0;
    return (v >= 1 && v <= 8) ? v : 0;
  }();
  const int s_cap = std::max(1, ktiles / min_ktiles);
  int best_s = 1, best_per = ktiles;
  long long best_cost = -1;
  if (force_waves) {
    best_s = std::min(s_cap, std::max(1, (force_waves * kSms) / ctas));
    best_per = (ktiles + best_s - 1) / best_s;
  } else {
    const int s_max = std::min(s_cap, std::max(1, (kMaxWaves * kSms) / ctas));
    for (int s = 1; s <= s_max; ++s) {
      const int per = (ktiles + s - 1) / s;
      const int s_real = (ktiles + per - 1) / per;
      const long long waves = (1ll * ctas * s_real + kSms - 1) / kSms;
      const long long cost = waves * (per + kEpilogue);
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost;
        best_s = s_real;
        best_per = per;
      }
    }
  }
  *per_out = best_per;
  return (ktiles + best_per - 1) / best_per;
}

static WgradConfig wgrad_config(int N, int H, int W, int num_taps, int Cin, int Cout, int dy_views) {
  static const bool no_halo = getenv("PLUME_DISABLE_WGRAD3") != nullptr;
  WgradConfig c;
  c.halo = !no_halo && num_taps == 9 && dy_views == 1 && H >= 8 && W >= 8 &&
           (Cin == 64 || Cin % 128 == 0);
  if (c.halo) {
    // Cin == 64: one CTA covers all nine taps (five 128-row blocks of TMEM, 64 columns each)
    // otherwise : one CTA covers the three vertical taps of one horizontal tap for 128 input channels
    c.block_n = (Cin == 64) ? 64 : (Cout % 128 == 0 ? 128 : 64);
    const int units = (Cin == 64) ? 1 : 3 * (Cin / 128);
    c.ctas_mn = units * (Cout / c.block_n);
    c.ktiles = ((W + 7) / 8) * ((H + 7) / 8) * N;
    c.splits = split_for(c.ctas_mn, c.ktiles, 16, &c.ktiles_per_split);
  } else {
    int bw, bh, bn;
    wgrad_geometry(N, H, W, &bw, &bh, &bn, &c.ktiles);
    c.block_n = wgrad_block_n(Cout);
    const int m_units = (Cin == 64) ? (num_taps + 1) / 2 : num_taps * (Cin / 128);
    c.ctas_mn = m_units * (Cout / c.block_n);
    c.splits = split_for(c.ctas_mn, c.ktiles, 8, &c.ktiles_per_split);
  }
  return c;
}

int wgrad_plan(int N, int H, int W, int num_taps, int Cin, int Cout) {
  return wgrad_config(N, H, W, num_taps, Cin, Cout, num_taps == 9 ? 1 : 4).splits;
}
static int wgrad_plan_views(int N, int H, int W, int num_taps, int Cin, int Cout, int dy_views) {
  return wgrad_config(N, H, W, num_taps, Cin, Cout, dy_views).splits;
}

// =================================================================================================
// 3x3 weight gradient with halo reuse ("wgrad3").  K tile = 8 (w) x 8 (h) pixels; with MN-major
// operands an 8-pixel image row of the tile is one 1024-byte atom, so one X box of 8 x 10 pixels
// feeds the three vertical taps (descriptor start moved by (1+dh) atoms) and dY is loaded once per
// K tile instead of once per tap.  Accumulators for all taps of the CTA live side by side in TMEM:
//   Cin == 64 : 9 taps x 64 rows = five 128-row blocks (tap pairs), 3 X boxes (dw = -1,0,1) per K tile
//   Cin >= 128: 3 vertical taps x 128 rows = three blocks for one dw and one 128-channel slice
// =================================================================================================
struct Wgrad3Params {
  int tiles_w, tiles_h, N;        // K tiling: 8 x 8 pixel tiles per image
  int cin, cout;
  int mode9;                      // 1: Cin == 64, all nine taps in one CTA
  int ci_blocks;                  // Cin / 128 (mode9 == 0)
  int n_tiles;
  int splits, ktiles_per_split;
  int stages;
  float* dw;                      // fp32, element (tap t, ci, co) at co*s_co + t*s_t + ci
  long long s_co, s_t;
  float* det_part;                // deterministic mode: [splits][dw_numel] partial sums, one MMA issuer
  long long dw_numel;
};

constexpr int kXBox = 10 * 1024;  // 8 w x 10 h pixels x 64 channels

// Warp roles: 0-3 epilogue (TMEM lane quadrant = warp), 4 TMEM allocator, 5 TMA producer,
// 6 barrier init then relay ("stage full" barriers -> shared-memory counter), 7 and 8 MMA issuers.
// The two issuers alternate pipeline stages (12 or 20 MMAs each) so that one's counter poll / commit
// overlaps the other's MMAs (see igemm_conv3_kernel).  They accumulate into the same TMEM blocks: the
// split-K sums are combined with atomics anyway, so the order of the stages' contributions is free; only
// the zeroing stage 0 must come first, which issuer 1 waits for (ctr[1]).
constexpr int kWgrad3Threads = 288;

// PAIR (Cin >= 256, 128 output channels per CTA pair, launched as clusters of two CTAs along x): the two CTAs take two
// units (horizontal tap x 128-channel slice) of the same N tile and K split, so they share dY -- each loads HALF of
// the dY tile (64 of its 128 channels) and the leader issues M = 256 MMAs (cta_group::2); see igemm_conv3_kernel.
template <int BLOCK_N, bool PAIR = false>
__global__ void __launch_bounds__(kWgrad3Threads, 1)
    igemm_wgrad3_kernel(const __grid_constant__ CUtensorMap xmap,
                        const __grid_constant__ CUtensorMap dymap, const Wgrad3Params p) {
  static_assert(!PAIR || BLOCK_N == 128, "CTA pairs: 128-wide tiles");
  constexpr int DY_BYTES = (PAIR ? BLOCK_N / 128 : BLOCK_N / 64) * 8192;   // PAIR: this CTA's half of the dY tile
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  constexpr int kWarpAlloc = 4, kWarpProducer = 5, kWarpRelay = 6, kWarpMma0 = 7, kWarpMma1 = 8;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t sbase = (raw_addr + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (sbase - raw_addr);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  const int x_bytes = p.mode9 ? 3 * kXBox : 2 * kXBox;
  const int stage_bytes = x_bytes + DY_BYTES;
  const uint32_t off_bar = p.stages * stage_bytes;
  const uint32_t bar0 = sbase + off_bar;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kMaxSlots + s); };
  const uint32_t tfull_bar = bar0 + 8u * (2 * kMaxSlots);
  volatile uint32_t* tmem_ptr_smem =
      reinterpret_cast<volatile uint32_t*>(gbase + off_bar + 8 * (2 * kMaxSlots + 1));
  volatile uint32_t* ctr = reinterpret_cast<volatile uint32_t*>(gbase + off_bar + 8 * (2 * kMaxSlots + 2));
  const int nblocks = p.mode9 ? 5 : 3;
  const uint32_t tmem_cols = (nblocks * BLOCK_N <= 256) ? 256u : 512u;

  // work item: blockIdx.x = (unit, n tile) -- PAIR: ((unit pair, n tile), rank) --, blockIdx.y = K split
  const int item = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int n_tile = item % p.n_tiles;
  const int unit = PAIR ? 2 * (item / p.n_tiles) + static_cast<int>(rank) : item / p.n_tiles;
  const int split = blockIdx.y;
  const int dwi = p.mode9 ? 0 : unit / p.ci_blocks;   // horizontal tap of this CTA (mode 3)
  const int cb = p.mode9 ? 0 : unit % p.ci_blocks;    // 128-channel slice of this CTA (mode 3)
  const int tiles_img = p.tiles_w * p.tiles_h;
  const int total_ktiles = tiles_img * p.N;
  const int kt_begin = split * p.ktiles_per_split;
  const int kt_end = min(kt_begin + p.ktiles_per_split, total_ktiles);
  const int nk = kt_end - kt_begin;

  if (warp == kWarpProducer && lane == 0) {
    tma_prefetch_desc(&xmap);
    tma_prefetch_desc(&dymap);
  }
  if (warp == kWarpRelay && lane == 0) {
    for (int s = 0; s < kMaxSlots; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, (nk > 1 && !p.det_part) ? 2 : 1);  // every issuer that has a stage commits once
    ctr[0] = 0;
    ctr[1] = 0;
    fence_mbar_init();
  }
  if (warp == kWarpAlloc) {
    if (PAIR) {
      tmem_alloc_2cta(sbase + off_bar + 8 * (2 * kMaxSlots + 1), tmem_cols);
      tmem_relinquish_2cta();
    } else {
      tmem_alloc(sbase + off_bar + 8 * (2 * kMaxSlots + 1), tmem_cols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == kWarpProducer) {
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int kt = kt_begin; kt < kt_end; ++kt) {
        const int img = kt / tiles_img;
        const int rem = kt % tiles_img;
        const int w0 = (rem % p.tiles_w) * 8;
        const int h0 = (rem / p.tiles_w) * 8;
        mbar_wait_relaxed(empty_bar(stage), phase ^ 1, 5, &g_dbg_word);
        const uint32_t x_addr = sbase + stage * stage_bytes;
        const uint32_t dy_addr = x_addr + x_bytes;
        if (PAIR) {
          // both CTAs' bytes are counted on the leader's barrier; this CTA loads its own X boxes and its half of dY
          if (leader) mbar_expect_tx(full_bar(stage), 2 * stage_bytes);
          tma_load_4d_2cta(x_addr, &xmap, full_bar(stage), cb * 128, w0 + dwi - 1, h0 - 1, img);
          tma_load_4d_2cta(x_addr + kXBox, &xmap, full_bar(stage), cb * 128 + 64, w0 + dwi - 1, h0 - 1, img);
          tma_load_4d_2cta(dy_addr, &dymap, full_bar(stage), n_tile * BLOCK_N + static_cast<int>(rank) * 64, w0, h0,
                           img);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
          continue;
        }
        mbar_expect_tx(full_bar(stage), stage_bytes);
        if (p.mode9) {
#pragma unroll
          for (int d = 0; d < 3; ++d)
            tma_load_4d(x_addr + d * kXBox, &xmap, full_bar(stage), 0, w0 + d - 1, h0 - 1, img);
        } else {
          tma_load_4d(x_addr, &xmap, full_bar(stage), cb * 128, w0 + dwi - 1, h0 - 1, img);
          tma_load_4d(x_addr + kXBox, &xmap, full_bar(stage), cb * 128 + 64, w0 + dwi - 1, h0 - 1,
                      img);
        }
#pragma unroll
        for (int b = 0; b < BLOCK_N / 64; ++b)
          tma_load_4d(dy_addr + b * 8192, &dymap, full_bar(stage), n_tile * BLOCK_N + b * 64, w0, h0,
                      img);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == kWarpRelay) {
    if (lane == 0 && leader) {
      int stage = 0, phase = 0;
      for (int i = 0; i < nk; ++i) {
        mbar_wait(full_bar(stage), phase, 6, &g_dbg_word);
        ctr[0] = i + 1;
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == kWarpMma0 || warp == kWarpMma1) {
    const int me = warp == kWarpMma0 ? 0 : 1;
    // deterministic mode: one issuer takes every stage, so the accumulation order inside the CTA is fixed
    const int istep = p.det_part ? 1 : 2;
    if (lane == 0 && leader && me < nk && (me == 0 || !p.det_part)) {
      constexpr uint32_t idesc = umma_idesc_bf16(PAIR ? 256 : 128, BLOCK_N, 1, 1);
      constexpr uint32_t hi = umma_desc_hi_sw128(1024);
      if (me == 1) wait_counter(ctr + 1, 1, 8);  // the zeroing stage has been issued
      int stage = me % p.stages;
      for (int i = me; i < nk; i += istep) {
        wait_counter(ctr, i + 1, 6);
        tc_fence_after();
        const uint32_t x_addr = sbase + stage * stage_bytes;
        const uint32_t b_lo = umma_desc_lo(x_addr + x_bytes, 8192);
        const uint32_t accflag = i != 0 ? 1u : 0u;
        if (p.mode9) {
          // rows 0-63: tap 2*blk, rows 64-127: tap 2*blk+1, taps ordered (dw major, dh minor):
          // tap t lives at box (t/3) of the stage, atom (t%3); LBO = distance between the two taps
#pragma unroll
          for (int blk = 0; blk < 5; ++blk) {
            const int t0 = 2 * blk;
            const uint32_t a_lo = umma_desc_lo(x_addr + (t0 / 3) * kXBox + (t0 % 3) * 1024,
                                               blk == 1 ? (kXBox - 2048) : 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_lohi(tmem_base + blk * BLOCK_N, a_lo + 128 * k, hi, b_lo + 128 * k, hi, idesc,
                             k == 0 ? accflag : 1u);
          }
        } else {
          // blk == vertical tap; the two 64-channel boxes are kXBox apart
#pragma unroll
          for (int blk = 0; blk < 3; ++blk) {
            const uint32_t a_lo = umma_desc_lo(x_addr + blk * 1024, kXBox);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (PAIR)
                umma_bf16_lohi_2cta(tmem_base + blk * BLOCK_N, a_lo + 128 * k, hi, b_lo + 128 * k, hi, idesc,
                                    k == 0 ? accflag : 1u);
              else
                umma_bf16_lohi(tmem_base + blk * BLOCK_N, a_lo + 128 * k, hi, b_lo + 128 * k, hi, idesc,
                               k == 0 ? accflag : 1u);
            }
          }
        }
        if (PAIR) umma_commit_2cta(empty_bar(stage), 3u);
        else umma_commit(empty_bar(stage));
        if (i == 0) ctr[1] = 1;
        stage += istep;
        while (stage >= p.stages) stage -= p.stages;
      }
      if (PAIR) umma_commit_2cta(tfull_bar, 3u);
      else umma_commit(tfull_bar);
    }
  } else if (warp < 4) {
    const int q = warp;
    const int row = q * 32 + lane;
    mbar_wait_relaxed(tfull_bar, 0, 7, &g_dbg_word);
    tc_fence_after();
    for (int blk = 0; blk < nblocks; ++blk) {
      int wt, ci;  // weight tap index (r*3+s) and input channel of this TMEM lane
      bool ok = true;
      if (p.mode9) {
        const int t = 2 * blk + (row >> 6);  // (dw major, dh minor)
        ok = t < 9;
        wt = (t % 3) * 3 + (t / 3);
        ci = row & 63;
      } else {
        wt = blk * 3 + dwi;
        ci = cb * 128 + row;
      }
      // fp32 reductions straight into dW: lanes = consecutive ci = one 128-byte line per instruction
      float* base = p.det_part ? p.det_part + split * p.dw_numel : p.dw;
      float* dst = base + wt * p.s_t + ci + static_cast<long long>(n_tile * BLOCK_N) * p.s_co;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + blk * BLOCK_N + c, v);
        tmem_ld_wait();
        if (ok) {
          if (p.det_part) {
#pragma unroll
            for (int j = 0; j < 32; ++j) dst[(c + j) * p.s_co] = __uint_as_float(v[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(dst + (c + j) * p.s_co, __uint_as_float(v[j]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) {
    cluster_sync_all();
    if (warp == kWarpAlloc) tmem_dealloc_2cta(tmem_base, tmem_cols);
  } else {
    if (warp == kWarpAlloc) tmem_dealloc(tmem_base, tmem_cols);
  }
}

static int launch_wgrad3_pair(const CUtensorMap& xmap, const CUtensorMap& dymap, const Wgrad3Params& p, size_t smem,
                              int ctas_mn, cudaStream_t stream) {
  auto kern = igemm_wgrad3_kernel<128, true>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] {
    attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                      cudaSharedmemCarveoutMaxShared);
  });
  if (attr_err != cudaSuccess) {
    set_error(std::string("cudaFuncSetAttribute(igemm_wgrad3 pair): ") + cudaGetErrorString(attr_err));
    return -2;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas_mn, p.splits, 1);     // ctas_mn is even: clusters of two along x
  cfg.blockDim = dim3(kWgrad3Threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, xmap, dymap, p);
  if (e != cudaSuccess) {
    set_error(std::string("igemm_wgrad3 pair launch: ") + cudaGetErrorString(e));
    return -3;
  }
  return 0;
}

template <int BLOCK_N>
static int launch_wgrad3_inst(const CUtensorMap& xmap, const CUtensorMap& dymap, const Wgrad3Params& p,
                              size_t smem, int ctas_mn, cudaStream_t stream) {
  auto kern = igemm_wgrad3_kernel<BLOCK_N>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] {
    attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    // full shared-memory carve-out, so that a block of another stream's kernel fits beside this CTA
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                      cudaSharedmemCarveoutMaxShared);
  });
  if (attr_err != cudaSuccess) {
    set_error(std::string("cudaFuncSetAttribute(igemm_wgrad3): ") + cudaGetErrorString(attr_err));
    return -2;
  }
  dim3 grid(ctas_mn, p.splits);
  kern<<<grid, kWgrad3Threads, smem, stream>>>(xmap, dymap, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error(std::string("igemm_wgrad3 launch: ") + cudaGetErrorString(e));
    return -3;
  }
  return 0;
}

static int launch_wgrad3(const WgradDesc& d, const WgradConfig& c) {
  const int W = d.x.W, H = d.x.H, N = d.x.N;
  Wgrad3Params p;
  p.tiles_w = (W + 7) / 8;
  p.tiles_h = (H + 7) / 8;
  p.N = N;
  p.cin = d.Cin;
  p.cout = d.Cout;
  p.mode9 = d.Cin == 64 ? 1 : 0;
  p.ci_blocks = d.Cin / 128;
  p.n_tiles = d.Cout / c.block_n;
  p.splits = c.splits;
  p.ktiles_per_split = c.ktiles_per_split;
  p.dw = d.dw;
  p.s_co = d.s_co;
  p.s_t = d.s_t;
  p.det_part = d.det_part;
  p.dw_numel = 9ll * d.Cin * d.Cout;
  // CTA pairs when two units of the same N tile exist for every pair (an even number of 128-channel slices).  OFF by
  // default (PLUME_WGRAD3_PAIR=1 enables it): correct (tests/test_gpu_ops.py) but measured no faster on any layer
  // (gpurun_out/r2aa: 202.1 vs 201.8, 107.6 vs 107.6, 205.4 vs 205.5 us ...) -- the 128-wide weight-gradient tiles
  // already run at 80-86 % of the tensor pipe and are not bound by the dY reads the pairing halves.
  static const bool pair_on = getenv("PLUME_WGRAD3_PAIR") && atoi(getenv("PLUME_WGRAD3_PAIR")) != 0;
  const bool pair = pair_on && !p.mode9 && c.block_n == 128 && p.ci_blocks % 2 == 0;
  const int stage_bytes = (p.mode9 ? 3 : 2) * kXBox + (pair ? 1 : c.block_n / 64) * 8192;
  const int overhead = 8 * (2 * kMaxSlots + 3) + 16 + 1024;
  // Leave room beside a CTA for one block of a bandwidth kernel on another stream (see unet.py).
  static const int smem_budget = [] {
    const char* e = getenv("PLUME_WGRAD3_SMEM");
    const int v = e ? atoi(e) : 0;
    return (v >= 65536 && v <= 232448) ? v : 212992;
  }();
  p.stages = std::min(kMaxSlots, (smem_budget - overhead) / stage_bytes);
  const size_t smem = static_cast<size_t>(p.stages) * stage_bytes + overhead;
  CUtensorMap xmap, dymap;
  if (make_act_map(&xmap, d.x, 64, 8, 10, 1)) return -1;
  if (make_act_map(&dymap, d.dy[0], 64, 8, 8, 1)) return -1;
  if (pair) return launch_wgrad3_pair(xmap, dymap, p, smem, c.ctas_mn, d.stream);
  if (c.block_n == 128) return launch_wgrad3_inst<128>(xmap, dymap, p, smem, c.ctas_mn, d.stream);
  return launch_wgrad3_inst<64>(xmap, dymap, p, smem, c.ctas_mn, d.stream);
}

template <int BLOCK_N, int STAGES, bool SPLIT = false>
static int launch_wgrad_inst(const TmapPack2& xmap, const TmapPack8& dymaps, const WgradParams& p,
                             cudaStream_t stream) {
  using L = WgSmem<BLOCK_N, STAGES>;
  auto kern = igemm_wgrad_kernel<BLOCK_N, STAGES, SPLIT>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] {
    attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
  });
  if (attr_err != cudaSuccess) {
    set_error(std::string("cudaFuncSetAttribute(igemm_wgrad): ") + cudaGetErrorString(attr_err));
    return -2;
  }
  dim3 grid(p.m_units * p.n_tiles, p.splits);
  kern<<<grid, 256, L::TOTAL, stream>>>(xmap, dymaps, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error(std::string("igemm_wgrad launch: ") + cudaGetErrorString(e));
    return -3;
  }
  return 0;
}

static int launch_igemm_wgrad_impl(const WgradDesc& d);

// Deterministic mode: every K split stores its partial sums to its own slice of the scratch (zeroed first: a CTA
// writes only the rows it owns) and `ordered_sum` adds the slices to dW in split order.
int launch_igemm_wgrad(const WgradDesc& d0) {
  if (!deterministic()) return launch_igemm_wgrad_impl(d0);
  WgradDesc d = d0;
  const int splits = wgrad_plan_views(d.x.N, d.x.H, d.x.W, d.num_taps, d.Cin, d.Cout, d.split ? 4 : d.num_dy_views);
  const long long numel = 1ll * d.num_taps * d.Cin * d.Cout;
  const size_t bytes = static_cast<size_t>(splits) * numel * sizeof(float);
  d.det_part = static_cast<float*>(det_scratch(1, bytes));
  if (!d.det_part) return -2;
  if (cudaMemsetAsync(d.det_part, 0, bytes, d.stream) != cudaSuccess) {
    set_error("deterministic wgrad: cudaMemsetAsync failed");
    return -2;
  }
  if (int r = launch_igemm_wgrad_impl(d)) return r;
  return ordered_sum(d.det_part, splits, static_cast<int>(numel), d.dw, d.stream);
}

static int launch_igemm_wgrad_impl(const WgradDesc& d) {
  if (d.Cin == 64 && d.num_dy_views > 1) {
    set_error("igemm_wgrad: per-tap dY views need Cin to be a multiple of 128");
    return -1;
  }
  if (!(d.Cin == 64 || d.Cin % 128 == 0) || d.Cin <= 0) {
    set_error("igemm_wgrad: input channels must be 64 or a multiple of 128");
    return -1;
  }
  if (d.Cout % 64 != 0 || d.Cout <= 0) {
    set_error("igemm_wgrad: output channels must be a positive multiple of 64");
    return -1;
  }
  const int W = d.x.W, H = d.x.H, N = d.x.N;
  if (W <= 0 || H <= 0 || N <= 0) {
    set_error("igemm_wgrad: empty input");
    return -1;
  }
  if (d.x.C > d.Cin || d.x.C % 8) {  // fewer channels: TMA zero-fills, those dW columns receive zeros
    set_error("igemm_wgrad: x view channel count > Cin");
    return -1;
  }
  for (int i = 0; i < d.num_dy_views; ++i) {
    const ActView& v = d.dy[i];
    if (v.C != d.Cout || v.W != W || v.H != H || v.N != N) {
      set_error("igemm_wgrad: dy view extents do not match");
      return -1;
    }
  }
  // bf16x3 mode always takes the generic kernel (dy_views = 4 rules the halo kernel out)
  const WgradConfig c = wgrad_config(N, H, W, d.num_taps, d.Cin, d.Cout, d.split ? 4 : d.num_dy_views);
  if (c.halo) return launch_wgrad3(d, c);

  int bw, bh, bn, ktiles;
  wgrad_geometry(N, H, W, &bw, &bh, &bn, &ktiles);
  const int block_n = c.block_n;
  TmapPack2 xmap;
  TmapPack8 dymaps;
  if (make_act_map(&xmap.m[0], d.x, 64, bw, bh, bn)) return -1;
  if (d.split) {
    if (d.x.plane <= 0) {
      set_error("igemm_wgrad: bf16x3 views need their lo-plane offset");
      return -1;
    }
    if (make_act_map(&xmap.m[1], lo_plane(d.x), 64, bw, bh, bn)) return -1;
  } else {
    xmap.m[1] = xmap.m[0];
  }
  for (int i = 0; i < 4; ++i) {
    const ActView& v = d.dy[i < d.num_dy_views ? i : 0];
    if (make_act_map(&dymaps.m[i], v, 64, bw, bh, bn)) return -1;
    if (d.split) {
      if (v.plane <= 0) {
        set_error("igemm_wgrad: bf16x3 views need their lo-plane offset");
        return -1;
      }
      if (make_act_map(&dymaps.m[4 + i], lo_plane(v), 64, bw, bh, bn)) return -1;
    } else {
      dymaps.m[4 + i] = dymaps.m[i];
    }
  }
  WgradParams p;
  p.PW = bw; p.PH = bh; p.PN = bn;
  p.tiles_w = (W + bw - 1) / bw;
  p.tiles_h = (H + bh - 1) / bh;
  p.tiles_n = (N + bn - 1) / bn;
  p.num_taps = d.num_taps;
  p.per_tap_dy = d.num_dy_views > 1 ? 1 : 0;
  p.cin = d.Cin;
  p.cout = d.Cout;
  p.pair_taps = d.Cin == 64 ? 1 : 0;
  p.ci_tiles = d.Cin / 128;
  p.m_units = p.pair_taps ? (d.num_taps + 1) / 2 : d.num_taps * p.ci_tiles;
  p.n_tiles = d.Cout / block_n;
  p.splits = c.splits;
  p.ktiles_per_split = c.ktiles_per_split;
  p.dw = d.dw;
  p.s_co = d.s_co;
  p.s_t = d.s_t;
  p.det_part = d.det_part;
  p.dw_numel = 1ll * d.num_taps * d.Cin * d.Cout;
  if (d.split) {
    switch (block_n) {
      case 256: return launch_wgrad_inst<256, 4, true>(xmap, dymaps, p, d.stream);
      case 128: return launch_wgrad_inst<128, 6, true>(xmap, dymaps, p, d.stream);
      default:  return launch_wgrad_inst<64, 8, true>(xmap, dymaps, p, d.stream);
    }
  }
  switch (block_n) {
    case 256: return launch_wgrad_inst<256, 4>(xmap, dymaps, p, d.stream);
    case 128: return launch_wgrad_inst<128, 6>(xmap, dymaps, p, d.stream);
    default:  return launch_wgrad_inst<64, 8>(xmap, dymaps, p, d.stream);
  }
}

}  // namespace plume
