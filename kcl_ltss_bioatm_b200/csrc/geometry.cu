// Label-geometry kernels: plume hulls -> masks (rank 1) and fire -> pixel geolocation
// (rank 3).  Integer / index work: results are bit-exact against the oracles in oracle/hull_ref.py and
// oracle/fire_ref.py, which are pinned to the reference's own functions (plume_selector.py:88-116,
// plume_identifier_gaussian_profile.py:65-123).
#include "bandwidth.cuh"

#include <algorithm>
#include <cstdint>
#include <string>

namespace plume {

namespace {
int check_launch_geo(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    return -3;
  }
  return 0;
}
}  // namespace

// ------------------------------------------------------------------------------------------------
// plume hulls -> masks.  A block owns a 32 (rows) x 128 (cols) pixel region of one window; polygons are
// taken 256 at a time: every thread bounding-box-tests one of them against the region, the survivors are
// compacted into a list and all threads walk that list.  A thread owns 16 consecutive pixels of one row (one 16-byte store); per polygon it
// first clips its span against the box, then evaluates the edge functions (int64, exact) per pixel.
// ------------------------------------------------------------------------------------------------
constexpr int kRasterRows = 32, kRasterCols = 128;

__global__ void __launch_bounds__(256)
    rasterize_hulls_kernel(const int* __restrict__ verts, const int* __restrict__ offs,
                           const int* __restrict__ bbox, int n_polys, const int* __restrict__ ys,
                           const int* __restrict__ xs, int Hm, int Wm, uint8_t* __restrict__ masks) {
  __shared__ int s_list[256];   // polygons (of the current 256) whose box overlaps this block's region
  __shared__ int s_warp[8];
  const int win = blockIdx.z;
  const int oy = ys ? ys[win] : 0, ox = xs ? xs[win] : 0;
  const int r0 = blockIdx.y * kRasterRows, c0 = blockIdx.x * kRasterCols;
  const int row = r0 + (threadIdx.x >> 3), col = c0 + (threadIdx.x & 7) * 16;
  // region in scene coordinates
  const int rx0 = ox + c0, rx1 = ox + min(c0 + kRasterCols, Wm) - 1;
  const int ry0 = oy + r0, ry1 = oy + min(r0 + kRasterRows, Hm) - 1;
  const int py = oy + row, px0 = ox + col;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t bits = 0;  // bit i: pixel col + i is inside some polygon
  for (int base = 0; base < n_polys; base += 256) {
    const int i = base + threadIdx.x;
    bool hit = false;
    if (i < n_polys) {
      const int4 b = *reinterpret_cast<const int4*>(bbox + 4 * i);
      hit = !(b.x > rx1 || b.z < rx0 || b.y > ry1 || b.w < ry0);
    }
    // compact the survivors (ballot + per-warp counts): the walk below costs one iteration per survivor
    const uint32_t ballot = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) s_warp[warp] = __popc(ballot);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = s_warp[q];
      if (q < warp) before += c;
      total += c;
    }
    if (hit) s_list[before + __popc(ballot & ((1u << lane) - 1u))] = i;
    __syncthreads();
    for (int j = 0; j < total; ++j) {
      const int p = s_list[j];
      const int4 b = *reinterpret_cast<const int4*>(bbox + 4 * p);
      if (py < b.y || py > b.w || px0 > b.z || px0 + 15 < b.x || bits == 0xFFFFu) continue;
      const int v0 = offs[p], v1 = offs[p + 1];
      uint32_t in = 0xFFFFu;
      int ax = verts[2 * (v1 - 1)], ay = verts[2 * (v1 - 1) + 1];
      for (int v = v0; v < v1 && in; ++v) {
        const int bx = verts[2 * v], by = verts[2 * v + 1];
        // edge function at pixel (px0 + k, py): e0 - k * dy, inside (or on the edge) when >= 0.  It is linear
        // along the span, so its signs at the two ends decide whole spans; only straddling edges go per pixel.
        const long long dx = bx - ax, dy = by - ay;
        const long long e0 = dx * (py - ay) - dy * (px0 - ax);
        const long long e15 = e0 - 15 * dy;
        if (e0 < 0 && e15 < 0) {
          in = 0;
        } else if (e0 < 0 || e15 < 0) {
          long long e = e0;
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            if (e < 0) in &= ~(1u << k);
            e -= dy;
          }
        }
        ax = bx;
        ay = by;
      }
      bits |= in;
    }
    __syncthreads();
  }
  if (row >= Hm || col >= Wm) return;
  uint8_t* dst = masks + (static_cast<long long>(win) * Hm + row) * Wm + col;
  if (col + 16 <= Wm && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t n = (bits >> (4 * q)) & 0xFu;
      w[q] = (n & 1u) | ((n & 2u) << 7) | ((n & 4u) << 14) | ((n & 8u) << 21);
    }
    *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
  } else {
    for (int k = 0; k < 16 && col + k < Wm; ++k) dst[k] = (bits >> k) & 1u;
  }
}

int rasterize_hulls(const int* verts, const int* offs, const int* bbox, int n_polys, const int* ys,
                    const int* xs, int count, int Hm, int Wm, uint8_t* masks, cudaStream_t s) {
  if (count <= 0 || Hm <= 0 || Wm <= 0) return 0;
  if ((ys == nullptr) != (xs == nullptr)) {
    set_error("rasterize_hulls: ys and xs must both be given or both be null");
    return -1;
  }
  if (ys == nullptr && count != 1) {
    set_error("rasterize_hulls: without window origins there is exactly one window");
    return -1;
  }
  if (count > 65535) {
    set_error("rasterize_hulls: at most 65535 windows per call");
    return -1;
  }
  dim3 grid((Wm + kRasterCols - 1) / kRasterCols, (Hm + kRasterRows - 1) / kRasterRows, count);
  rasterize_hulls_kernel<<<grid, 256, 0, s>>>(verts, offs, bbox, n_polys, ys, xs, Hm, Wm, masks);
  return check_launch_geo("rasterize_hulls");
}

// ------------------------------------------------------------------------------------------------
// fire -> pixel geolocation.  For every fire: among the pixels whose (lat, lon) lie strictly inside the
// +-half_box box around the fire, the first one in row-major order with the smallest haversine distance
// (float64, the reference's operation order, no fused multiply-adds).  Two passes over the lat/lon grids:
// pass 1 takes the minimum of the distance bit patterns (non-negative doubles order like unsigned integers),
// pass 2 the minimum linear index among the pixels that attain it.
// A block owns chunks of 1024 consecutive pixels (4 per thread, in registers).  Per chunk it reduces the
// chunk's lat/lon bounding box, tests every fire's box against it (one fire per thread, 256 at a time, bounds
// precomputed), compacts the few survivors with a ballot, and only those are tested per pixel -- the grids
// are read once per pass and a pixel meets a handful of fires instead of all of them.
// ------------------------------------------------------------------------------------------------
constexpr int kPixPerThread = 4;
constexpr int kChunk = 256 * kPixPerThread;

__device__ __forceinline__ double haversine_km(double lon1, double lat1, double lon2, double lat2) {
  constexpr double kRad = 3.141592653589793238462643383279502884 / 180.0;
  lon1 = __dmul_rn(lon1, kRad);
  lat1 = __dmul_rn(lat1, kRad);
  lon2 = __dmul_rn(lon2, kRad);
  lat2 = __dmul_rn(lat2, kRad);
  const double sdlat = sin(__dmul_rn(__dsub_rn(lat2, lat1), 0.5));   // x / 2.0 == x * 0.5 exactly
  const double sdlon = sin(__dmul_rn(__dsub_rn(lon2, lon1), 0.5));
  const double a = __dadd_rn(__dmul_rn(sdlat, sdlat),
                             __dmul_rn(__dmul_rn(cos(lat1), cos(lat2)), __dmul_rn(sdlon, sdlon)));
  return __dmul_rn(6367.0, __dmul_rn(2.0, asin(sqrt(a))));
}

// bounds[f] = (lat - h, lat + h, lon - h, lon + h), the exact float64 values the reference compares against
__global__ void fire_bounds_kernel(const double* __restrict__ fire_lat, const double* __restrict__ fire_lon,
                                   int n_fires, double half_box, double4* __restrict__ bounds) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_fires) return;
  const double la = fire_lat[f], lo = fire_lon[f];
  bounds[f] = make_double4(__dsub_rn(la, half_box), __dadd_rn(la, half_box), __dsub_rn(lo, half_box),
                           __dadd_rn(lo, half_box));
}

template <int PASS>
__global__ void __launch_bounds__(256)
    locate_fires_kernel(const double* __restrict__ lats, const double* __restrict__ lons, long long pixels,
                        const double* __restrict__ fire_lat, const double* __restrict__ fire_lon,
                        const double4* __restrict__ bounds, int n_fires,
                        unsigned long long* __restrict__ best_dist, unsigned int* __restrict__ best_idx) {
  __shared__ double s_red[4][8];   // per-warp min lat, max lat, min lon, max lon
  __shared__ int s_list[256];
  __shared__ int s_warp[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long n_chunks = (pixels + kChunk - 1) / kChunk;
  for (long long chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    const long long i0 = chunk * kChunk + threadIdx.x;
    double la[kPixPerThread], lo[kPixPerThread];
    double mn_la = 1e300, mx_la = -1e300, mn_lo = 1e300, mx_lo = -1e300;
#pragma unroll
    for (int k = 0; k < kPixPerThread; ++k) {
      const long long i = i0 + 256 * k;
      const bool in = i < pixels;
      la[k] = in ? lats[i] : 1e300;      // +huge: fails every "la < hi" test below
      lo[k] = in ? lons[i] : 1e300;
      if (in) {
        mn_la = fmin(mn_la, la[k]); mx_la = fmax(mx_la, la[k]);
        mn_lo = fmin(mn_lo, lo[k]); mx_lo = fmax(mx_lo, lo[k]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn_la = fmin(mn_la, __shfl_xor_sync(0xffffffffu, mn_la, o));
      mx_la = fmax(mx_la, __shfl_xor_sync(0xffffffffu, mx_la, o));
      mn_lo = fmin(mn_lo, __shfl_xor_sync(0xffffffffu, mn_lo, o));
      mx_lo = fmax(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, o));
    }
    __syncthreads();   // the previous chunk's readers of s_red / s_list are done
    if (lane == 0) {
      s_red[0][warp] = mn_la; s_red[1][warp] = mx_la; s_red[2][warp] = mn_lo; s_red[3][warp] = mx_lo;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      mn_la = fmin(mn_la, s_red[0][q]); mx_la = fmax(mx_la, s_red[1][q]);
      mn_lo = fmin(mn_lo, s_red[2][q]); mx_lo = fmax(mx_lo, s_red[3][q]);
    }
    for (int base = 0; base < n_fires; base += 256) {
      const int f = base + threadIdx.x;
      bool hit = false;
      if (f < n_fires) {
        const double4 b = bounds[f];
        // some pixel of the chunk may satisfy lo_lat < la < hi_lat and lo_lon < lo < hi_lon
        hit = mx_la > b.x && mn_la < b.y && mx_lo > b.z && mn_lo < b.w;
      }
      const uint32_t ballot = __ballot_sync(0xffffffffu, hit);
      __syncthreads();   // s_list / s_warp of the previous round consumed
      if (lane == 0) s_warp[warp] = __popc(ballot);
      __syncthreads();
      int before = 0, total = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int c = s_warp[q];
        if (q < warp) before += c;
        total += c;
      }
      if (hit) s_list[before + __popc(ballot & ((1u << lane) - 1u))] = f;
      __syncthreads();
      for (int j = 0; j < total; ++j) {
        const int t = s_list[j];
        const double4 b = bounds[t];
        unsigned long long target = 0;
        if (PASS == 2) target = best_dist[t];
#pragma unroll
        for (int k = 0; k < kPixPerThread; ++k) {
          if (la[k] > b.x && la[k] < b.y && lo[k] > b.z && lo[k] < b.w) {
            const unsigned long long d = static_cast<unsigned long long>(
                __double_as_longlong(haversine_km(fire_lon[t], fire_lat[t], lo[k], la[k])));
            if (PASS == 1) {
              atomicMin(&best_dist[t], d);
            } else if (d == target) {
              atomicMin(&best_idx[t], static_cast<unsigned int>(i0 + 256 * k));
            }
          }
        }
      }
    }
  }
}

__global__ void locate_fires_finish_kernel(const unsigned int* __restrict__ best_idx, int n_fires, int W,
                                           int* __restrict__ out_rc) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_fires) return;
  const unsigned int i = best_idx[f];
  out_rc[2 * f] = i == 0xFFFFFFFFu ? -1 : static_cast<int>(i / W);
  out_rc[2 * f + 1] = i == 0xFFFFFFFFu ? -1 : static_cast<int>(i % W);
}

// workspace: bounds double4[n] | best_dist u64[n] | best_idx u32[n]
size_t locate_fires_workspace_bytes(int n_fires) { return n_fires > 0 ? 48ull * n_fires : 0; }

int locate_fires(const double* lats, const double* lons, int H, int W, const double* fire_lat,
                 const double* fire_lon, int n_fires, double half_box, void* workspace, size_t workspace_bytes,
                 int* out_rc, cudaStream_t s) {
  if (n_fires <= 0) return 0;
  const long long pixels = 1ll * H * W;
  if (H <= 0 || W <= 0 || pixels >= 0xFFFFFFFFll) {
    set_error("locate_fires: image must hold between 1 and 2^32-2 pixels");
    return -1;
  }
  if (workspace_bytes < locate_fires_workspace_bytes(n_fires) || (reinterpret_cast<uintptr_t>(workspace) & 31)) {
    set_error("locate_fires: workspace too small or not 32-byte aligned (plume_locate_fires_workspace_bytes)");
    return -1;
  }
  double4* bounds = static_cast<double4*>(workspace);
  unsigned long long* best_dist = reinterpret_cast<unsigned long long*>(bounds + n_fires);
  unsigned int* best_idx = reinterpret_cast<unsigned int*>(best_dist + n_fires);
  if (cudaMemsetAsync(best_dist, 0xFF, 12ull * n_fires, s) != cudaSuccess) {
    set_error("locate_fires: cudaMemsetAsync failed");
    return -2;
  }
  int sms = 148;
  {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  fire_bounds_kernel<<<(n_fires + 255) / 256, 256, 0, s>>>(fire_lat, fire_lon, n_fires, half_box, bounds);
  const long long n_chunks = (pixels + kChunk - 1) / kChunk;
  const int grid = static_cast<int>(std::max(1ll, std::min(n_chunks, 1ll * sms * 8)));
  locate_fires_kernel<1><<<grid, 256, 0, s>>>(lats, lons, pixels, fire_lat, fire_lon, bounds, n_fires, best_dist,
                                              best_idx);
  locate_fires_kernel<2><<<grid, 256, 0, s>>>(lats, lons, pixels, fire_lat, fire_lon, bounds, n_fires, best_dist,
                                              best_idx);
  locate_fires_finish_kernel<<<(n_fires + 255) / 256, 256, 0, s>>>(best_idx, n_fires, W, out_rc);
  return check_launch_geo("locate_fires");
}

}  // namespace plume
