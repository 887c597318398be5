// Threshold-sweep kernels: the reference's plume-extent search
// (plume_identifier_gaussian_profile.py:142-202) for all thresholds of a sweep at once.
//   threshold_masks_kernel : aod > t -> erosion -> dilation (cross footprint), every threshold from one read
//   ccl_*_kernel           : 8-connected component labelling by union-find over each mask plane
//   fire_extents_kernel    : per (threshold, fire) the size of the component nearest to the fire in its window
// Integer / boolean work; results are exact against oracle/sweep_ref.py.
#include "bandwidth.cuh"
#include "sweep_bits.cuh"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <string>

namespace plume {

namespace {
int check_launch_sweep(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    return -3;
  }
  return 0;
}
constexpr int kMaxThresholds = 64;
}  // namespace

// The image is float32 or float64 (the reference's MAIAC AOD is int16 * 0.001 = float64, tools.py:88); thresholds are
// float64.  A float64 image is compared in float64.  For a float32 v and a double t,  v > t  <=>  v > rd(t)  with
// rd(t) the largest float32 <= t: if t is a float32 itself nothing changes; otherwise t lies strictly between two
// consecutive float32 a < t < b and v > t <=> v >= b <=> v > a = rd(t).  So for float32 images the thresholds are
// rounded DOWN once (__double2float_rd) and the comparisons run in fp32.
template <typename V>
__device__ __forceinline__ V image_threshold(double t);
template <>
__device__ __forceinline__ float image_threshold<float>(double t) {
  return __double2float_rd(t);
}
template <>
__device__ __forceinline__ double image_threshold<double>(double t) {
  return t;
}

// ------------------------------------------------------------------------------------------------
// masks[t][y][x] = dilate(erode(aod > thr[t])).  With the cross footprint that is
//   OR over q in cross(p), q inside the image, of  AND over r in cross(q) of (r outside the image or aod[r] > t)
// (erosion sees set pixels beyond the border, dilation unset ones).  A thread loads the 13 values of the
// diamond around its pixel once and evaluates every threshold from registers.
// Comparison: image_threshold<V> above.
// ------------------------------------------------------------------------------------------------
template <typename V>
__global__ void __launch_bounds__(256)
    threshold_masks_kernel(const V* __restrict__ aod, int H, int W, const double* __restrict__ thr, int T,
                           uint8_t* __restrict__ masks) {
  __shared__ V s_thr[kMaxThresholds];
  if (threadIdx.x < T) s_thr[threadIdx.x] = image_threshold<V>(thr[threadIdx.x]);
  __syncthreads();
  const long long pixels = 1ll * H * W;
  const long long i = 1ll * blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pixels) return;
  const int y = static_cast<int>(i / W), x = static_cast<int>(i % W);
  // the 13 diamond positions, bit index = (dy + 2) * 5 + (dx + 2); only |dy| + |dx| <= 2 are used
  V v[25];
  uint32_t outside = 0;   // bit set: position outside the image (counts as set for the erosion)
#pragma unroll
  for (int dy = -2; dy <= 2; ++dy)
#pragma unroll
    for (int dx = -2; dx <= 2; ++dx) {
      if ((dy < 0 ? -dy : dy) + (dx < 0 ? -dx : dx) > 2) continue;
      const int b = (dy + 2) * 5 + (dx + 2);
      const int yy = y + dy, xx = x + dx;
      const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
      if (!in) outside |= 1u << b;
      v[b] = in ? aod[1ll * yy * W + xx] : V(0);
    }
  for (int t = 0; t < T; ++t) {
    const V th = s_thr[t];
    uint32_t set = outside;   // bit b: position counts as set for the erosion
#pragma unroll
    for (int dy = -2; dy <= 2; ++dy)
#pragma unroll
      for (int dx = -2; dx <= 2; ++dx) {
        if ((dy < 0 ? -dy : dy) + (dx < 0 ? -dx : dx) > 2) continue;
        const int b = (dy + 2) * 5 + (dx + 2);
        if (v[b] > th) set |= 1u << b;
      }
    // eroded(q): q inside the image and the cross around q all set; q = p + (dy, dx), |dy| + |dx| <= 1
    auto eroded = [&](int dy, int dx) {
      const int b = (dy + 2) * 5 + (dx + 2);
      const uint32_t cross = (1u << b) | (1u << (b - 5)) | (1u << (b + 5)) | (1u << (b - 1)) | (1u << (b + 1));
      return ((outside >> b) & 1u) == 0 && (set & cross) == cross;
    };
    const bool m = eroded(0, 0) || eroded(-1, 0) || eroded(1, 0) || eroded(0, -1) || eroded(0, 1);
    masks[1ll * t * pixels + i] = m ? 1 : 0;
  }
}

int threshold_masks(const void* aod, int f64, int H, int W, const double* thr, int T, uint8_t* masks, cudaStream_t s) {
  if (H <= 0 || W <= 0 || T <= 0) return 0;
  if (T > kMaxThresholds) {
    set_error("threshold_masks: at most 64 thresholds per call");
    return -1;
  }
  const long long pixels = 1ll * H * W;
  const unsigned grid = static_cast<unsigned>((pixels + 255) / 256);
  if (f64) threshold_masks_kernel<double><<<grid, 256, 0, s>>>(static_cast<const double*>(aod), H, W, thr, T, masks);
  else threshold_masks_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(aod), H, W, thr, T, masks);
  return check_launch_sweep("threshold_masks");
}

// ------------------------------------------------------------------------------------------------
// Connected components, 8-connectivity.  parent[] holds, per plane, a union-find forest over pixel indices;
// links always point to the smaller index, so a component's root is its smallest row-major index (the
// canonical label).  Run based: a warp owns 32 consecutive pixels of a row.
//   init  : every set pixel points straight at the first pixel of its horizontal run inside the warp's
//           segment (ballot + count-leading-zeros, no atomics).
//   merge : unions only where a connection is new -- a segment's first pixel with the pixel to its left, and
//           with the row above only at the leftmost pixel of a run that touches an upper run: at a run start
//           north (or, if north is unset, north-west and north-east); inside a run only north-east when north
//           is unset (everything else was already joined by the pixel to the left).
//   flatten: every pixel looks up its root; the size counters are bumped once per (warp, root) with
//           __match_any_sync instead of once per pixel (large components would serialise on one address).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int uf_find(const int* parent, int x) {
  int p = parent[x];
  while (p != x) {
    x = p;
    p = parent[x];
  }
  return x;
}

__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) {
      const int t = a;
      a = b;
      b = t;
    }
    // a > b: hang root a under b unless someone re-rooted a meanwhile
    const int old = atomicMin(&parent[a], b);
    if (old == a) return;
    a = old;
  }
}

struct RowSeg {
  int t, y, x;       // plane, row, column of this lane's pixel
  bool valid;        // inside the image
};
__device__ __forceinline__ RowSeg row_segment(int H, int W, int T) {
  const int segs = (W + 31) / 32;
  const long long warp = (1ll * blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  RowSeg r;
  r.t = static_cast<int>(warp / (1ll * H * segs));
  const int rem = static_cast<int>(warp % (1ll * H * segs));
  r.y = rem / segs;
  r.x = (rem % segs) * 32 + (threadIdx.x & 31);
  r.valid = r.t < T && r.x < W;
  return r;
}

__global__ void __launch_bounds__(256)
    ccl_init_kernel(const uint8_t* __restrict__ masks, int T, int H, int W, int* __restrict__ parent,
                    int* __restrict__ sizes) {
  const RowSeg r = row_segment(H, W, T);
  const int lane = threadIdx.x & 31;
  const long long i = (1ll * r.t * H + r.y) * W + r.x;
  const bool m = r.valid && masks[i];
  const uint32_t bits = __ballot_sync(0xffffffffu, m);
  if (!r.valid) return;
  int p = -1;
  if (m) {
    const uint32_t lower = (1u << lane) - 1u;
    const uint32_t zeros = ~bits & lower;                    // unset pixels to the left, inside the segment
    const int start = zeros ? 32 - __clz(zeros) : 0;         // first lane of this lane's run
    p = r.y * W + (r.x - lane) + start;
  }
  parent[i] = p;
  sizes[i] = 0;
}

__global__ void __launch_bounds__(256)
    ccl_merge_kernel(const uint8_t* __restrict__ masks, int T, int H, int W, int* __restrict__ parent_all) {
  const RowSeg r = row_segment(H, W, T);
  if (!r.valid) return;
  const int lane = threadIdx.x & 31;
  const long long plane = 1ll * H * W;
  const uint8_t* m = masks + r.t * plane;
  int* parent = parent_all + r.t * plane;
  const int i = r.y * W + r.x;
  if (!m[i]) return;
  const bool w = r.x > 0 && m[i - 1];
  if (lane == 0 && w) uf_union(parent, i, i - 1);            // the run continues from the segment to the left
  if (r.y == 0) return;
  const bool n = m[i - W];
  const bool ne = r.x + 1 < W && m[i - W + 1];
  if (w) {
    if (!n && ne) uf_union(parent, i, i - W + 1);
  } else if (n) {
    uf_union(parent, i, i - W);
  } else {
    if (r.x > 0 && m[i - W - 1]) uf_union(parent, i, i - W - 1);
    if (ne) uf_union(parent, i, i - W + 1);
  }
}

__global__ void __launch_bounds__(256)
    ccl_flatten_kernel(int T, int H, int W, int* __restrict__ parent_all, int* __restrict__ sizes_all) {
  const RowSeg r = row_segment(H, W, T);
  const int lane = threadIdx.x & 31;
  const long long plane = 1ll * H * W;
  int root = -1;
  if (r.valid) {
    int* parent = parent_all + r.t * plane;
    const int i = r.y * W + r.x;
    if (parent[i] >= 0) {
      root = uf_find(parent, i);
      parent[i] = root;   // readers that race with this see the old parent or the root: both are ancestors
    }
  }
  const uint32_t active = __ballot_sync(0xffffffffu, root >= 0);
  if (root < 0) return;
  const uint32_t same = __match_any_sync(active, root);
  if (lane == __ffs(same) - 1) atomicAdd(&sizes_all[r.t * plane + root], __popc(same));
}

int label_components(const uint8_t* masks, int T, int H, int W, int* labels, int* sizes, cudaStream_t s) {
  if (T <= 0 || H <= 0 || W <= 0) return 0;
  const long long plane = 1ll * H * W;
  if (plane >= 0x7FFFFFFFll) {
    set_error("label_components: plane too large");
    return -1;
  }
  const long long warps = 1ll * T * H * ((W + 31) / 32);
  const long long blocks = (warps + 7) / 8;
  if (blocks >= 0x7FFFFFFFll) {
    set_error("label_components: too many pixels for one call");
    return -1;
  }
  const unsigned grid = static_cast<unsigned>(blocks);
  ccl_init_kernel<<<grid, 256, 0, s>>>(masks, T, H, W, labels, sizes);
  ccl_merge_kernel<<<grid, 256, 0, s>>>(masks, T, H, W, labels);
  ccl_flatten_kernel<<<grid, 256, 0, s>>>(T, H, W, labels, sizes);
  return check_launch_sweep("label_components");
}

// ------------------------------------------------------------------------------------------------
// extents[t][f] = size of the component whose pixel is nearest to fire f inside the (2 win + 1)^2 window
// (Euclidean distance; first pixel in row-major window order on ties, as numpy's argmin over the window),
// 0 if the window holds no component.  One warp per (threshold, fire).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    fire_extents_kernel(const int* __restrict__ labels, const int* __restrict__ sizes, int T, int H, int W,
                        const int* __restrict__ fire_rc, int n_fires, int win, int* __restrict__ extents) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= T * n_fires) return;
  const int t = warp / n_fires, f = warp % n_fires;
  const int r = fire_rc[2 * f], c = fire_rc[2 * f + 1];
  const int side = 2 * win + 1;
  const int* lab = labels + 1ll * t * H * W;
  unsigned long long best = ~0ull;   // (d2 << 32) | window index
  for (int k = lane; k < side * side; k += 32) {
    const int dy = k / side - win, dx = k % side - win;
    const int y = r + dy, x = c + dx;
    if (y < 0 || y >= H || x < 0 || x >= W) continue;
    if (lab[y * W + x] >= 0) {
      const unsigned long long key = (static_cast<unsigned long long>(dy * dy + dx * dx) << 32) | static_cast<unsigned>(k);
      best = key < best ? key : best;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other < best ? other : best;
  }
  if (lane == 0) {
    int out = 0;
    if (best != ~0ull) {
      const int k = static_cast<int>(best & 0xFFFFFFFFu);
      const int y = r + k / side - win, x = c + k % side - win;
      out = sizes[1ll * t * H * W + lab[y * W + x]];
    }
    extents[t * n_fires + f] = out;
  }
}

int fire_extents(const int* labels, const int* sizes, int T, int H, int W, const int* fire_rc, int n_fires, int win,
                 int* extents, cudaStream_t s) {
  if (T <= 0 || n_fires <= 0) return 0;
  if (win < 0 || win > 1000) {
    set_error("fire_extents: bad window");
    return -1;
  }
  const long long warps = 1ll * T * n_fires;
  fire_extents_kernel<<<static_cast<unsigned>((warps * 32 + 255) / 256), 256, 0, s>>>(labels, sizes, T, H, W, fire_rc,
                                                                                    n_fires, win, extents);
  return check_launch_sweep("fire_extents");
}

// ================================================================================================
// Bit-plane sweep: the same three steps on masks packed 32 pixels per word (sweep_bits.cuh holds the per-thread
// bodies, shared with the CPU emulation test).  The byte-mask / int32-label kernels above stay for callers that
// want dense planes; ThresholdSweep.extents and find_plume_extents use this path.
//   mask_bits_kernel    : a warp walks down a 64-column strip of the image.  Per row coalesced loads; per pixel a
//                         binary search gives its comparison bits against all thresholds of the chunk, a warp-wide
//                         bit transpose turns them into row words per threshold; the lane that owns a threshold
//                         runs erosion and dilation for it on 36-bit row windows (bit-parallel over the 32
//                         columns of a word), a sliding window of two rows of B and of E in registers.
//   bits_init / merge / flatten : union-find over word-local runs (<= 16 per word, typically 1), one thread per
//                         word; parent / size entries only exist at run starts.
//   bits_extents_kernel : one warp per (threshold, fire), one window row per lane, nearest set bit per row by clz / ffs.
// ================================================================================================
using namespace sweepbits;

constexpr int kStripRows = 16;      // output rows per warp of mask_bits_kernel (+ 4 halo rows); 8 when the grid is small
constexpr int kMaskWarps = 4;

// A warp takes two adjacent 32-column strips and kStripRows rows (the two extra window columns between the strips are
// ordinary pixels of the other strip; lanes 0..3 load the four outer ones).  Every lane finds how many thresholds of
// the chunk lie below its pixel (binary search in the chunk's ascending threshold table: the comparison results
// against ALL thresholds are the low bits of one word), the warp transposes the 32 x 32 bit matrix with five shuffles,
// and lane r then holds the row words of the r-th smallest threshold and writes to that threshold's plane.
// (First version: one ballot per threshold and operand, 75 ballots per row instead of 14 shuffles: 69 vs 54 us.)
template <typename V>
__global__ void __launch_bounds__(kMaskWarps * 32)
    mask_bits_kernel(const V* __restrict__ aod, int H, int W, const double* __restrict__ thr, int T, int strip_rows,
                     uint32_t* __restrict__ bits, int2* __restrict__ ent_all) {
  __shared__ V s_thr[32];        // the chunk's thresholds, ascending, padded with +inf
  __shared__ V s_raw[32];
  __shared__ int s_plane[32];    // threshold index (inside the chunk) of the r-th smallest threshold
  const int chunk = blockIdx.y;
  const int Tc = min(32, T - 32 * chunk);
  if (threadIdx.x < 32) {
    V mine = threadIdx.x < Tc ? image_threshold<V>(thr[32 * chunk + threadIdx.x]) : V(INFINITY);
    if (mine != mine) mine = V(INFINITY);                                 // v > NaN is false for every v, like v > +inf
    s_raw[threadIdx.x] = mine;
    s_thr[threadIdx.x] = V(INFINITY);
  }
  __syncthreads();
  if (threadIdx.x < Tc) {
    const int r = rank_of(s_raw, Tc, static_cast<int>(threadIdx.x));
    s_thr[r] = s_raw[threadIdx.x];
    s_plane[r] = threadIdx.x;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int segs = (W + 31) / 32, pairs = (segs + 1) / 2;
  const int strips = (H + strip_rows - 1) / strip_rows;
  const long long wg = 1ll * blockIdx.x * kMaskWarps + (threadIdx.x >> 5);
  if (wg >= 1ll * pairs * strips) return;
  const int sp = static_cast<int>(wg % pairs), y0 = static_cast<int>(wg / pairs) * strip_rows;
  const int y_end = min(y0 + strip_rows, H);
  const int seg0 = 2 * sp, seg1 = 2 * sp + 1;
  const uint64_t colmask0 = window_colmask(seg0, W), colmask1 = window_colmask(seg1, W);
  const uint64_t outcols0 = ~colmask0 & kWin36, outcols1 = ~colmask1 & kWin36;
  const int x0 = 64 * sp + lane, x1 = x0 + 32;
  const int xe = lane < 2 ? 64 * sp - 2 + lane : 64 * sp + 62 + lane;          // lanes 0..3: columns -2, -1, +64, +65
  const bool in0 = x0 < W, in1 = x1 < W, ine = lane < 4 && xe >= 0 && xe < W;
  OpenState st0, st1;
  st0.b1 = st0.b2 = st0.e1 = st0.e2 = 0;
  st1 = st0;
  V n0 = 0, n1 = 0, ne = 0;
  if (y0 - 2 >= 0) {
    const V* row = aod + 1ll * (y0 - 2) * W;
    if (in0) n0 = __ldg(row + x0);
    if (in1) n1 = __ldg(row + x1);
    if (ine) ne = __ldg(row + xe);
  }
  const int plane = 32 * chunk + (lane < Tc ? s_plane[lane] : lane);     // the threshold this lane owns
  uint32_t* out = bits + (1ll * plane * H) * segs + seg0;
  // fused call: the lane that owns a threshold also creates the union-find entries of the runs it emits (bits_init)
  const Geom g = make_geom(H, W);
  int2* ent = ent_all ? ent_all + plane * g.ent_per_plane : nullptr;
  for (int yy = y0 - 2; yy <= y_end + 1; ++yy) {
    const V v0 = n0, v1 = n1, ve = ne;
    const bool row_in = yy >= 0 && yy < H;
    if (yy + 1 >= 0 && yy + 1 < H && yy + 1 <= y_end + 1) {
      const V* row = aod + 1ll * (yy + 1) * W;
      n0 = in0 ? __ldg(row + x0) : V(0);
      n1 = in1 ? __ldg(row + x1) : V(0);
      ne = ine ? __ldg(row + xe) : V(0);
    }
    uint64_t b0 = kWin36, b1 = kWin36;
    if (row_in) {                                                        // warp-uniform
      uint32_t c0 = low_mask32(count_below(s_thr, v0));                   // bit r: pixel > r-th smallest threshold
      uint32_t c1 = low_mask32(count_below(s_thr, v1));
      const uint32_t me = low_mask32(count_below(s_thr, ve));
#pragma unroll
      for (int j = 16; j > 0; j >>= 1) {
        c0 = transpose32_step(c0, __shfl_xor_sync(0xffffffffu, c0, j), lane, j);
        c1 = transpose32_step(c1, __shfl_xor_sync(0xffffffffu, c1, j), lane, j);
      }
      uint32_t e = 0;                                                     // the four outer columns live in lanes 0..3
#pragma unroll
      for (int j = 0; j < 4; ++j) e |= ((__shfl_sync(0xffffffffu, me, j) >> lane) & 1u) << j;
      b0 = (static_cast<uint64_t>(c0) << 2) | (e & 3u) | (static_cast<uint64_t>(c1 & 3u) << 34) | outcols0;
      b1 = (static_cast<uint64_t>(c1) << 2) | (c0 >> 30) | (static_cast<uint64_t>((e >> 2) & 3u) << 34) | outcols1;
    }
    const bool prev_in = yy - 1 >= 0 && yy - 1 < H;
    const uint32_t w0 = open_step(st0, b0, prev_in, colmask0);
    const uint32_t w1 = open_step(st1, b1, prev_in, colmask1);
    const int r = yy - 2;
    if (r >= y0 && lane < Tc) {
      out[1ll * r * segs] = w0;
      if (seg1 < segs) out[1ll * r * segs + 1] = w1;
      if (ent) {
        init_word(w0, ent, g, r, seg0);
        init_word(w1, ent, g, r, seg1);                                 // zero when seg1 is beyond the image
      }
    }
  }
}

namespace {
int mask_bits_launch(const void* aod, int f64, int H, int W, const double* thr, int T, uint32_t* bits, int2* ent,
                     cudaStream_t s) {
  if (H <= 0 || W <= 0 || T <= 0) return 0;
  const Geom g = make_geom(H, W);
  if (g.ent_per_plane >= 0x7FFFFFFFll) {
    set_error("threshold_mask_bits: plane too large");
    return -1;
  }
  // 16 output rows per warp cost 4 halo rows (25 % extra work); a call with few thresholds does not fill the GPU with
  // that (1200 x 1200 x 25: 9.6 warps per SM) and is latency bound, so it takes 8-row strips (twice the warps)
  const long long chunks = (T + 31) / 32;
  const long long warps16 = 1ll * ((g.segs + 1) / 2) * ((H + kStripRows - 1) / kStripRows);
  static const int force_rows = [] {
    const char* v = std::getenv("PLUME_SWEEP_STRIP_ROWS");
    return v ? std::atoi(v) : 0;
  }();
  static const int sms = plume_num_sms() > 0 ? plume_num_sms() : 148;
  const int strip_rows = force_rows > 0 ? force_rows : (warps16 * chunks < 16ll * sms ? kStripRows / 2 : kStripRows);
  const long long warps = 1ll * ((g.segs + 1) / 2) * ((H + strip_rows - 1) / strip_rows);
  const dim3 grid(static_cast<unsigned>((warps + kMaskWarps - 1) / kMaskWarps), static_cast<unsigned>(chunks));
  if (grid.y > 65535u) {
    set_error("threshold_mask_bits: too many thresholds");
    return -1;
  }
  if (f64) mask_bits_kernel<double><<<grid, kMaskWarps * 32, 0, s>>>(static_cast<const double*>(aod), H, W, thr, T, strip_rows, bits, ent);
  else mask_bits_kernel<float><<<grid, kMaskWarps * 32, 0, s>>>(static_cast<const float*>(aod), H, W, thr, T, strip_rows, bits, ent);
  return check_launch_sweep("threshold_mask_bits");
}
}  // namespace

int threshold_mask_bits(const void* aod, int f64, int H, int W, const double* thr, int T, uint32_t* bits,
                        cudaStream_t s) {
  return mask_bits_launch(aod, f64, H, W, thr, T, bits, nullptr, s);
}

// byte masks [T][H][W] -> bit planes (one warp per word)
__global__ void __launch_bounds__(256)
    pack_mask_bits_kernel(const uint8_t* __restrict__ masks, long long words, int W, int segs,
                          uint32_t* __restrict__ bits) {
  const long long word = (1ll * blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (word >= words) return;
  const int lane = threadIdx.x & 31;
  const long long row = word / segs;                       // t * H + y
  const int x = static_cast<int>(word % segs) * 32 + lane;
  const uint32_t w = __ballot_sync(0xffffffffu, x < W && masks[row * W + x] != 0);
  if (lane == 0) bits[word] = w;
}

int pack_mask_bits(const uint8_t* masks, int T, int H, int W, uint32_t* bits, cudaStream_t s) {
  if (T <= 0 || H <= 0 || W <= 0) return 0;
  const Geom g = make_geom(H, W);
  const long long words = g.words_per_plane * T;
  const long long blocks = (words * 32 + 255) / 256;
  if (blocks >= 0x7FFFFFFFll) {
    set_error("pack_mask_bits: too many pixels for one call");
    return -1;
  }
  pack_mask_bits_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(masks, words, W, g.segs, bits);
  return check_launch_sweep("pack_mask_bits");
}

struct WordPos {
  long long idx;   // word index over all planes
  int t, y, seg;
  bool valid;
};
__device__ __forceinline__ WordPos word_pos(const Geom& g, int T) {
  WordPos p;
  p.idx = 1ll * blockIdx.x * blockDim.x + threadIdx.x;
  p.valid = p.idx < g.words_per_plane * T;
  p.t = static_cast<int>(p.idx / g.words_per_plane);
  const int rem = static_cast<int>(p.idx % g.words_per_plane);
  p.y = rem / g.segs;
  p.seg = rem % g.segs;
  return p;
}

__global__ void __launch_bounds__(256)
    bits_init_kernel(const uint32_t* __restrict__ bits, int T, Geom g, int2* __restrict__ ent) {
  const WordPos p = word_pos(g, T);
  if (!p.valid) return;
  const uint32_t w = bits[p.idx];
  if (w) init_word(w, ent + p.t * g.ent_per_plane, g, p.y, p.seg);
}

__global__ void __launch_bounds__(256)
    bits_merge_kernel(const uint32_t* __restrict__ bits, int T, Geom g, int2* __restrict__ ent) {
  const WordPos p = word_pos(g, T);
  if (!p.valid) return;
  const uint32_t w = bits[p.idx];
  if (!w) return;
  const uint32_t left = p.seg > 0 ? bits[p.idx - 1] : 0u;
  uint32_t up_l = 0, up_c = 0, up_r = 0;
  if (p.y > 0) {
    const uint32_t* up = bits + p.idx - g.segs;
    up_c = up[0];
    if (p.seg > 0) up_l = up[-1];
    if (p.seg + 1 < g.segs) up_r = up[1];
  }
  merge_word(w, left, up_l, up_c, up_r, ent + p.t * g.ent_per_plane, g, p.y, p.seg);
}

// Every run looks up its root and stores it as its parent; the size counters are bumped once per (warp, root)
// and round: a large component would otherwise serialise tens of thousands of atomics on one address.
__global__ void __launch_bounds__(256)
    bits_flatten_kernel(const uint32_t* __restrict__ bits, int T, Geom g, int2* __restrict__ ent_all) {
  const WordPos p = word_pos(g, T);
  const int lane = threadIdx.x & 31;
  uint32_t rest = p.valid ? bits[p.idx] : 0u;
  int2* ent = ent_all + p.t * g.ent_per_plane;
  while (true) {
    const uint32_t active = __ballot_sync(0xffffffffu, rest != 0u);
    if (!active) break;
    if (rest != 0u) {
      int len;
      const int e = pop_run(rest, g, p.y, p.seg, len);
      const int root = uf_root(ent, e);
      if (root != e) ent[e].x = root;   // readers racing with this see the old parent or the root: both ancestors
      const unsigned long long key = (static_cast<unsigned long long>(p.t) << 32) | static_cast<unsigned>(root);
      const uint32_t same = __match_any_sync(active, key);
      int total = 0;
      for (uint32_t m = same; m; m &= m - 1u) total += __shfl_sync(same, len, __ffs(m) - 1);
      if (lane == __ffs(same) - 1) atomicAdd(&ent[root].y, total);
    }
  }
}

__global__ void __launch_bounds__(256)
    bits_extents_kernel(const uint32_t* __restrict__ bits, const int2* __restrict__ ent_all, int T, Geom g,
                        const int* __restrict__ fire_rc, int n_fires, int win, int* __restrict__ extents) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= T * n_fires) return;
  const int t = warp / n_fires, f = warp % n_fires;
  const int r = fire_rc[2 * f], c = fire_rc[2 * f + 1];
  const uint32_t* plane = bits + t * g.words_per_plane;
  unsigned long long best = kNoKey;
  for (int dy = -win + lane; dy <= win; dy += 32) {
    const int y = r + dy;
    if (y < 0 || y >= g.H) continue;
    const unsigned long long k = best_in_row(plane + 1ll * y * g.segs, g, c, dy, win);
    best = k < best ? k : best;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other < best ? other : best;
  }
  if (lane == 0) extents[t * n_fires + f] = extent_of_key(best, plane, ent_all + t * g.ent_per_plane, g, r, c, win);
}

// The component nearest to each fire in the plane chosen for it (find_plume_mask, gaussian_profile.py:306-331:
// label(mask) -> extract_label -> labelled_mask == label), from planes that bits_extents has labelled: one block per
// fire; warp 0 finds the nearest set pixel and its root, then every thread walks words of the plane and keeps the
// runs that hang under that root.  stats[f] = {area, min_row, min_col, max_row + 1, max_col + 1, root, 0, 0}
// (the bounding box as regionprops reports it; all zero and root = -1 when the fire has no component / no plane).
__global__ void __launch_bounds__(256)
    fire_components_kernel(const uint32_t* __restrict__ bits, const int2* __restrict__ ent_all, int T, Geom g,
                           const int* __restrict__ fire_rc, const int* __restrict__ plane_of_fire, int win,
                           uint32_t* __restrict__ comp, int* __restrict__ stats) {
  __shared__ int s_root;
  __shared__ int s_stat[5];
  const int f = blockIdx.x, lane = threadIdx.x & 31;
  const int p = plane_of_fire[f];
  const bool has_plane = p >= 0 && p < T;
  if (threadIdx.x == 0) {
    s_root = -1;
    s_stat[0] = 0, s_stat[1] = g.H, s_stat[2] = g.W, s_stat[3] = 0, s_stat[4] = 0;
  }
  __syncthreads();
  const uint32_t* plane = bits + (has_plane ? p : 0) * g.words_per_plane;
  const int2* ent = ent_all + (has_plane ? p : 0) * g.ent_per_plane;
  if (has_plane && threadIdx.x < 32) {
    const int r = fire_rc[2 * f], c = fire_rc[2 * f + 1];
    unsigned long long best = kNoKey;
    for (int dy = -win + lane; dy <= win; dy += 32) {
      const int y = r + dy;
      if (y < 0 || y >= g.H) continue;
      const unsigned long long k = best_in_row(plane + 1ll * y * g.segs, g, c, dy, win);
      best = k < best ? k : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other < best ? other : best;
    }
    if (lane == 0) s_root = root_of_key(best, plane, ent, g, r, c, win);
  }
  __syncthreads();
  const int root = s_root;
  uint32_t* out = comp + f * g.words_per_plane;
  int area = 0, y_min = g.H, x_min = g.W, y_max = 0, x_max = 0;
  for (long long i = threadIdx.x; i < g.words_per_plane; i += blockDim.x) {
    uint32_t o = 0;
    if (root >= 0) {
      const uint32_t w = plane[i];
      if (w) {
        const int y = static_cast<int>(i / g.segs), seg = static_cast<int>(i % g.segs);
        o = component_word(w, ent, g, y, seg, root);
        if (o) {
          area += __popc(o);
          y_min = min(y_min, y), y_max = max(y_max, y + 1);
          x_min = min(x_min, 32 * seg + ctz32(o)), x_max = max(x_max, 32 * seg + 32 - clz32(o));
        }
      }
    }
    out[i] = o;
  }
  if (area) {
    atomicAdd(&s_stat[0], area);
    atomicMin(&s_stat[1], y_min);
    atomicMin(&s_stat[2], x_min);
    atomicMax(&s_stat[3], y_max);
    atomicMax(&s_stat[4], x_max);
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    int v = 0;
    if (threadIdx.x < 5) v = s_stat[0] ? s_stat[threadIdx.x] : 0;
    else if (threadIdx.x == 5) v = root;
    stats[8 * f + threadIdx.x] = v;
  }
}

namespace {
size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }
}  // namespace

size_t sweep_workspace_bytes(int H, int W, int T) {
  if (H <= 0 || W <= 0 || T <= 0) return 0;
  const Geom g = make_geom(H, W);
  return align256(static_cast<size_t>(g.ent_per_plane) * T * sizeof(int2)) +
         align256(static_cast<size_t>(g.words_per_plane) * T * sizeof(uint32_t));
}

namespace {
int bits_extents_launch(const uint32_t* bits, int T, int H, int W, const int* fire_rc, int n_fires, int win,
                        void* workspace, size_t workspace_bytes, int* extents, bool entries_ready, cudaStream_t s) {
  if (T <= 0 || H <= 0 || W <= 0 || n_fires <= 0) return 0;
  const Geom g = make_geom(H, W);
  if (g.ent_per_plane >= 0x7FFFFFFFll) {
    set_error("bits_extents: plane too large");
    return -1;
  }
  if (win < 0 || win > 1000) {
    set_error("bits_extents: bad window");
    return -1;
  }
  if (workspace_bytes < sweep_workspace_bytes(H, W, T)) {
    set_error("bits_extents: workspace smaller than plume_sweep_workspace_bytes(H, W, T)");
    return -1;
  }
  const long long blocks = (g.words_per_plane * T + 255) / 256;
  if (blocks >= 0x7FFFFFFFll || 1ll * T * n_fires * 32 >= 0x7FFFFFFFll) {
    set_error("bits_extents: too many pixels / fires for one call");
    return -1;
  }
  int2* ent = static_cast<int2*>(workspace);
  const unsigned grid = static_cast<unsigned>(blocks);
  if (!entries_ready) bits_init_kernel<<<grid, 256, 0, s>>>(bits, T, g, ent);
  bits_merge_kernel<<<grid, 256, 0, s>>>(bits, T, g, ent);
  bits_flatten_kernel<<<grid, 256, 0, s>>>(bits, T, g, ent);
  bits_extents_kernel<<<static_cast<unsigned>((1ll * T * n_fires * 32 + 255) / 256), 256, 0, s>>>(
      bits, ent, T, g, fire_rc, n_fires, win, extents);
  return check_launch_sweep("bits_extents");
}
}  // namespace

int bits_extents(const uint32_t* bits, int T, int H, int W, const int* fire_rc, int n_fires, int win, void* workspace,
                 size_t workspace_bytes, int* extents, cudaStream_t s) {
  return bits_extents_launch(bits, T, H, W, fire_rc, n_fires, win, workspace, workspace_bytes, extents, false, s);
}

int fire_components(const uint32_t* bits, int T, int H, int W, const int* fire_rc, const int* plane_of_fire, int n_fires,
                    int win, const void* workspace, size_t workspace_bytes, uint32_t* comp, int* stats, cudaStream_t s) {
  if (T <= 0 || H <= 0 || W <= 0 || n_fires <= 0) return 0;
  const Geom g = make_geom(H, W);
  if (g.ent_per_plane >= 0x7FFFFFFFll || win < 0 || win > 1000) {
    set_error("fire_components: plane too large or bad window");
    return -1;
  }
  if (workspace_bytes < sweep_workspace_bytes(H, W, T)) {
    set_error("fire_components: workspace smaller than plume_sweep_workspace_bytes(H, W, T)");
    return -1;
  }
  fire_components_kernel<<<static_cast<unsigned>(n_fires), 256, 0, s>>>(bits, static_cast<const int2*>(workspace), T, g,
                                                                        fire_rc, plane_of_fire, win, comp, stats);
  return check_launch_sweep("fire_components");
}

int sweep_extents(const void* aod, int f64, int H, int W, const double* thr, int T, const int* fire_rc, int n_fires, int win,
                  void* workspace, size_t workspace_bytes, int* extents, cudaStream_t s) {
  if (T <= 0 || H <= 0 || W <= 0 || n_fires <= 0) return 0;
  if (workspace_bytes < sweep_workspace_bytes(H, W, T)) {
    set_error("sweep_extents: workspace smaller than plume_sweep_workspace_bytes(H, W, T)");
    return -1;
  }
  const Geom g = make_geom(H, W);
  uint32_t* bits = reinterpret_cast<uint32_t*>(static_cast<char*>(workspace) +
                                               align256(static_cast<size_t>(g.ent_per_plane) * T * sizeof(int2)));
  const int rc = mask_bits_launch(aod, f64, H, W, thr, T, bits, static_cast<int2*>(workspace), s);
  if (rc) return rc;
  return bits_extents_launch(bits, T, H, W, fire_rc, n_fires, win, workspace, workspace_bytes, extents, true, s);
}

}  // namespace plume
