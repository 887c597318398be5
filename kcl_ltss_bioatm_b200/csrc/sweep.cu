// Threshold-sweep kernels: the reference's plume-extent search
// (plume_identifier_gaussian_profile.py:142-202) for all thresholds of a sweep at once.
//   threshold_masks_kernel : aod > t -> erosion -> dilation (cross footprint), every threshold from one read
//   ccl_*_kernel           : 8-connected component labelling by union-find over each mask plane
//   fire_extents_kernel    : per (threshold, fire) the size of the component nearest to the fire in its window
// Integer / boolean work; results are exact against oracle/sweep_ref.py.
#include "bandwidth.cuh"

#include <algorithm>
#include <cstdint>
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

// ------------------------------------------------------------------------------------------------
// masks[t][y][x] = dilate(erode(aod > thr[t])).  With the cross footprint that is
//   OR over q in cross(p), q inside the image, of  AND over r in cross(q) of (r outside the image or aod[r] > t)
// (erosion sees set pixels beyond the border, dilation unset ones).  A thread loads the 13 values of the
// diamond around its pixel once and evaluates every threshold from registers; the comparison is done in
// float64 because the reference compares a float32 image with float64 thresholds.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    threshold_masks_kernel(const float* __restrict__ aod, int H, int W, const double* __restrict__ thr, int T,
                           uint8_t* __restrict__ masks) {
  __shared__ double s_thr[kMaxThresholds];
  if (threadIdx.x < T) s_thr[threadIdx.x] = thr[threadIdx.x];
  __syncthreads();
  const long long pixels = 1ll * H * W;
  const long long i = 1ll * blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pixels) return;
  const int y = static_cast<int>(i / W), x = static_cast<int>(i % W);
  // diamond offsets: index = (dy + 2) * 5 + (dx + 2); only |dy| + |dx| <= 2 are used
  double v[5][5];
  bool inb[5][5];
#pragma unroll
  for (int dy = -2; dy <= 2; ++dy)
#pragma unroll
    for (int dx = -2; dx <= 2; ++dx) {
      if ((dy < 0 ? -dy : dy) + (dx < 0 ? -dx : dx) > 2) continue;
      const int yy = y + dy, xx = x + dx;
      const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
      inb[dy + 2][dx + 2] = in;
      v[dy + 2][dx + 2] = in ? static_cast<double>(aod[1ll * yy * W + xx]) : 0.0;
    }
  for (int t = 0; t < T; ++t) {
    const double th = s_thr[t];
    // set[dy][dx]: pixel counts as set for the erosion (outside the image = set)
    auto set = [&](int dy, int dx) { return !inb[dy + 2][dx + 2] || v[dy + 2][dx + 2] > th; };
    auto eroded = [&](int dy, int dx) {   // q = p + (dy, dx), |dy| + |dx| <= 1; outside the image = unset
      if (!inb[dy + 2][dx + 2]) return false;
      return set(dy, dx) && set(dy - 1, dx) && set(dy + 1, dx) && set(dy, dx - 1) && set(dy, dx + 1);
    };
    const bool m = eroded(0, 0) || eroded(-1, 0) || eroded(1, 0) || eroded(0, -1) || eroded(0, 1);
    masks[1ll * t * pixels + i] = m ? 1 : 0;
  }
}

int threshold_masks(const float* aod, int H, int W, const double* thr, int T, uint8_t* masks, cudaStream_t s) {
  if (H <= 0 || W <= 0 || T <= 0) return 0;
  if (T > kMaxThresholds) {
    set_error("threshold_masks: at most 64 thresholds per call");
    return -1;
  }
  const long long pixels = 1ll * H * W;
  threshold_masks_kernel<<<static_cast<unsigned>((pixels + 255) / 256), 256, 0, s>>>(aod, H, W, thr, T, masks);
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

}  // namespace plume
