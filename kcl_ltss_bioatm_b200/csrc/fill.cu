// Nearest-valid fill of the AOD grid: the reference's interpolate_aod_nearest
// (plume_identifier_gaussian_profile.py:451-461 -- scipy NearestNDInterpolator over the pixels != NULL_VALUE,
// evaluated at every pixel), as an exact two-pass nearest-feature transform:
//   valid_bits_kernel     : validity as bit planes (32 pixels per word)
//   nearest_in_row_kernel : per pixel the signed offset to the nearest valid pixel of its ROW (clz / ffs over the row's
//                           words; the left one on ties)
//   nearest_valid_kernel  : per null pixel the rows y - k, y + k for growing k, candidate (k^2 + dx^2, row, column) from
//                           the row pass, until k^2 exceeds the best squared distance; the smallest
//                           (distance, row, column) wins, i.e. among equidistant valid pixels the first in row-major
//                           order (scipy's kd-tree picks one of them in traversal order: see oracle/sweep_ref.py).
// Values are copied, never computed: exact for float32 and float64 images.
#include "bandwidth.cuh"
#include "sweep_bits.cuh"

#include <cstdint>
#include <string>

namespace plume {

using namespace sweepbits;

namespace {
constexpr int kNoValid = 0x7FFFFFFF;

int check_launch_fill(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    return -3;
  }
  return 0;
}
size_t align256f(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }
}  // namespace

template <typename V>
__global__ void __launch_bounds__(256)
    valid_bits_kernel(const V* __restrict__ aod, long long words, int W, int segs, V null_value,
                      uint32_t* __restrict__ bits) {
  const long long word = (1ll * blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (word >= words) return;
  const int lane = threadIdx.x & 31;
  const long long y = word / segs;
  const int x = static_cast<int>(word % segs) * 32 + lane;
  const uint32_t w = __ballot_sync(0xffffffffu, x < W && aod[y * W + x] != null_value);   // NaN != null: valid, as numpy
  if (lane == 0) bits[word] = w;
}

__global__ void __launch_bounds__(256)
    nearest_in_row_kernel(const uint32_t* __restrict__ bits, int H, int W, int segs, int* __restrict__ dx_row) {
  const long long i = 1ll * blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 1ll * H * W) return;
  const int y = static_cast<int>(i / W), x = static_cast<int>(i % W);
  const uint32_t* row = bits + 1ll * y * segs;
  const int sg = x >> 5, b = x & 31;
  int left = -1, right = -1;                                   // columns of the nearest valid pixel at / left of x, right of x
  for (int s = sg; s >= 0; --s) {
    const uint32_t m = s == sg ? (row[s] & low_mask32(b + 1)) : row[s];
    if (m) {
      left = 32 * s + 31 - clz32(m);
      break;
    }
  }
  for (int s = sg; s < segs; ++s) {
    const uint32_t m = s == sg ? (row[s] & ~low_mask32(b + 1)) : row[s];
    if (m) {
      right = 32 * s + ctz32(m);
      break;
    }
  }
  int dx = kNoValid;
  if (left >= 0 && (right < 0 || x - left <= right - x)) dx = left - x;      // the left one on ties
  else if (right >= 0) dx = right - x;
  dx_row[i] = dx;
}

template <typename V>
__global__ void __launch_bounds__(256)
    nearest_valid_kernel(const V* __restrict__ aod, const int* __restrict__ dx_row, int H, int W, V* __restrict__ out) {
  const long long i = 1ll * blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 1ll * H * W) return;
  const int y = static_cast<int>(i / W), x = static_cast<int>(i % W);
  if (dx_row[i] == 0) {                                        // valid pixel: itself
    out[i] = aod[i];
    return;
  }
  long long best = 0x7FFFFFFFFFFFFFFFll;
  int best_y = -1, best_x = -1;
  for (int k = 0; 1ll * k * k <= best; ++k) {
    const int ya = y - k, yb = y + k;
    if (ya < 0 && yb >= H) break;
    if (ya >= 0) {
      const int dx = dx_row[1ll * ya * W + x];
      if (dx != kNoValid) {
        const long long d = 1ll * k * k + 1ll * dx * dx;
        if (d <= best) {                                       // row ya is above every candidate so far: it wins ties
          best = d;
          best_y = ya;
          best_x = x + dx;
        }
      }
    }
    if (k > 0 && yb < H) {
      const int dx = dx_row[1ll * yb * W + x];
      if (dx != kNoValid) {
        const long long d = 1ll * k * k + 1ll * dx * dx;
        if (d < best) {                                        // row yb is below every candidate so far: it loses ties
          best = d;
          best_y = yb;
          best_x = x + dx;
        }
      }
    }
  }
  out[i] = best_y >= 0 ? aod[1ll * best_y * W + best_x] : aod[i];   // no valid pixel at all: unchanged
}

size_t fill_nearest_workspace_bytes(int H, int W) {
  if (H <= 0 || W <= 0) return 0;
  const size_t segs = (static_cast<size_t>(W) + 31) / 32;
  return align256f(static_cast<size_t>(H) * segs * sizeof(uint32_t)) + align256f(static_cast<size_t>(H) * W * sizeof(int));
}

template <typename V>
static int fill_nearest_t(const V* aod, int H, int W, V null_value, void* workspace, size_t workspace_bytes, V* out,
                          cudaStream_t s) {
  if (H <= 0 || W <= 0) return 0;
  if (workspace_bytes < fill_nearest_workspace_bytes(H, W)) {
    set_error("fill_nearest: workspace smaller than plume_fill_nearest_workspace_bytes(H, W)");
    return -1;
  }
  const int segs = (W + 31) / 32;
  const long long words = 1ll * H * segs, pixels = 1ll * H * W;
  if ((pixels + 255) / 256 >= 0x7FFFFFFFll || (words * 32 + 255) / 256 >= 0x7FFFFFFFll) {
    set_error("fill_nearest: image too large");
    return -1;
  }
  uint32_t* bits = static_cast<uint32_t*>(workspace);
  int* dx_row = reinterpret_cast<int*>(static_cast<char*>(workspace) + align256f(static_cast<size_t>(words) * sizeof(uint32_t)));
  valid_bits_kernel<V><<<static_cast<unsigned>((words * 32 + 255) / 256), 256, 0, s>>>(aod, words, W, segs, null_value, bits);
  nearest_in_row_kernel<<<static_cast<unsigned>((pixels + 255) / 256), 256, 0, s>>>(bits, H, W, segs, dx_row);
  nearest_valid_kernel<V><<<static_cast<unsigned>((pixels + 255) / 256), 256, 0, s>>>(aod, dx_row, H, W, out);
  return check_launch_fill("fill_nearest");
}

int fill_nearest(const void* aod, int f64, int H, int W, double null_value, void* workspace, size_t workspace_bytes,
                 void* out, cudaStream_t s) {
  if (f64)
    return fill_nearest_t<double>(static_cast<const double*>(aod), H, W, null_value, workspace, workspace_bytes,
                                  static_cast<double*>(out), s);
  return fill_nearest_t<float>(static_cast<const float*>(aod), H, W, static_cast<float>(null_value), workspace,
                               workspace_bytes, static_cast<float*>(out), s);
}

}  // namespace plume
