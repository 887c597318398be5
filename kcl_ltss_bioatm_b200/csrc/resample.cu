// UTM projection and nearest-neighbour swath -> grid resampling: the GPU side of
// the reference's `utm_resampler` (/root/reference/src/features/tools.py:9-64), which delegates to pyproj (UTM) and
// pyresample (kd_tree.resample_nearest, radius of influence 10 km).  Neither library is available here, so their
// published behaviour is restated (see oracle/resample_ref.py for the statement and its anchors):
//   * UTM: Krueger series in the third flattening (Karney 2011), WGS84, k0 = 0.9996, false easting 500 km,
//     no false northing; fp64 throughout;
//   * nearest neighbour: 3-D Cartesian distance on the sphere R = 6 370 997 m between the target cell CENTRE and the
//     swath pixels, strictly below the radius; ties go to the smallest flat swath index.
// Instead of a kd-tree the swath pixels are counting-sorted into square buckets of the target plane (bucket edge
// >= 1.1 x radius: a pixel within `radius` metres of a cell centre lies in the 3 x 3 buckets around the centre's
// bucket as long as the projection's linear scale stays within 10 %, i.e. anywhere within ~25 degrees of the central
// meridian).  One thread per target cell then scans those buckets.  The result is an INDEX map, so one neighbour
// search serves every image on the same swath geometry (what pyresample calls get_neighbour_info / get_sample).
#include "bandwidth.cuh"

#include <cstdint>
#include <string>

namespace plume {

namespace {
int check_launch_rs(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    return -3;
  }
  return 0;
}

constexpr double kA = 6378137.0, kF = 1.0 / 298.257223563, kK0 = 0.9996, kE0 = 500000.0;
constexpr double kRSphere = 6370997.0;
constexpr double kPi = 3.14159265358979323846, kDeg = kPi / 180.0;

struct Krueger {
  double A, e, e2, alpha[6], beta[6];
};

Krueger make_krueger() {
  Krueger k;
  const double n = kF / (2.0 - kF);
  const double n2 = n * n, n3 = n2 * n, n4 = n3 * n, n5 = n4 * n, n6 = n5 * n;
  k.A = kA / (1 + n) * (1 + n2 / 4 + n4 / 64 + n6 / 256);
  k.e2 = kF * (2 - kF);
  k.e = sqrt(k.e2);
  k.alpha[0] = n / 2 - 2 * n2 / 3 + 5 * n3 / 16 + 41 * n4 / 180 - 127 * n5 / 288 + 7891 * n6 / 37800;
  k.alpha[1] = 13 * n2 / 48 - 3 * n3 / 5 + 557 * n4 / 1440 + 281 * n5 / 630 - 1983433 * n6 / 1935360;
  k.alpha[2] = 61 * n3 / 240 - 103 * n4 / 140 + 15061 * n5 / 26880 + 167603 * n6 / 181440;
  k.alpha[3] = 49561 * n4 / 161280 - 179 * n5 / 168 + 6601661 * n6 / 7257600;
  k.alpha[4] = 34729 * n5 / 80640 - 3418889 * n6 / 1995840;
  k.alpha[5] = 212378941 * n6 / 319334400;
  k.beta[0] = n / 2 - 2 * n2 / 3 + 37 * n3 / 96 - n4 / 360 - 81 * n5 / 512 + 96199 * n6 / 604800;
  k.beta[1] = n2 / 48 + n3 / 15 - 437 * n4 / 1440 + 46 * n5 / 105 - 1118711 * n6 / 3870720;
  k.beta[2] = 17 * n3 / 480 - 37 * n4 / 840 - 209 * n5 / 4480 + 5569 * n6 / 90720;
  k.beta[3] = 4397 * n4 / 161280 - 11 * n5 / 504 - 830251 * n6 / 7257600;
  k.beta[4] = 4583 * n5 / 161280 - 108847 * n6 / 3991680;
  k.beta[5] = 20648693 * n6 / 638668800;
  return k;
}

__device__ __forceinline__ double tau_prime(double tau, const Krueger& k) {
  const double s = sinh(k.e * atanh(k.e * tau / sqrt(1 + tau * tau)));
  return tau * sqrt(1 + s * s) - s * sqrt(1 + tau * tau);
}

// (lat, lon) in degrees -> UTM metres in `zone` (central meridian 6*zone - 183)
__device__ __forceinline__ void utm_fwd(double lat, double lon, double lon0, const Krueger& k, double* x, double* y) {
  double lam = (lon - lon0) * kDeg;
  lam -= 2 * kPi * floor((lam + kPi) / (2 * kPi));
  const double tp = tau_prime(tan(lat * kDeg), k);
  double sl, cl;
  sincos(lam, &sl, &cl);
  const double xi0 = atan2(tp, cl), eta0 = asinh(sl / sqrt(tp * tp + cl * cl));
  double xi = xi0, eta = eta0;
#pragma unroll
  for (int j = 1; j <= 6; ++j) {
    double s, c;
    sincos(2 * j * xi0, &s, &c);
    xi += k.alpha[j - 1] * s * cosh(2 * j * eta0);
    eta += k.alpha[j - 1] * c * sinh(2 * j * eta0);
  }
  *x = kE0 + kK0 * k.A * eta;
  *y = kK0 * k.A * xi;
}

__device__ __forceinline__ void utm_inv(double x, double y, double lon0, const Krueger& k, double* lat, double* lon) {
  const double xi = y / (kK0 * k.A), eta = (x - kE0) / (kK0 * k.A);
  double xip = xi, etap = eta;
#pragma unroll
  for (int j = 1; j <= 6; ++j) {
    double s, c;
    sincos(2 * j * xi, &s, &c);
    xip -= k.beta[j - 1] * s * cosh(2 * j * eta);
    etap -= k.beta[j - 1] * c * sinh(2 * j * eta);
  }
  const double sh = sinh(etap), cx = cos(xip);
  const double taup = sin(xip) / sqrt(sh * sh + cx * cx);
  double tau = taup;
#pragma unroll 1
  for (int it = 0; it < 5; ++it) {  // Newton on tau'(tau) = taup
    const double tpi = tau_prime(tau, k);
    tau += (taup - tpi) / sqrt(1 + tpi * tpi) * (1 + (1 - k.e2) * tau * tau) / ((1 - k.e2) * sqrt(1 + tau * tau));
  }
  *lat = atan(tau) / kDeg;
  *lon = atan2(sh, cx) / kDeg + lon0;
}

__device__ __forceinline__ void to_cartesian(double lat, double lon, double* c) {
  double sa, ca, so, co;
  sincos(lat * kDeg, &sa, &ca);
  sincos(lon * kDeg, &so, &co);
  c[0] = kRSphere * ca * co;
  c[1] = kRSphere * ca * so;
  c[2] = kRSphere * sa;
}

__global__ void utm_zone_hist_kernel(const double* __restrict__ lon, long long n, int* __restrict__ hist) {
  __shared__ int s_hist[64];
  if (threadIdx.x < 64) s_hist[threadIdx.x] = 0;
  __syncthreads();
  const long long step = 1ll * gridDim.x * blockDim.x;
  for (long long i = 1ll * blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
    const double l = lon[i];
    const double w = (l + 180) - floor((l + 180) / 360) * 360 - 180;   // tools.py:27
    const int z = static_cast<int>(floor((w + 180) / 6) + 1);          // tools.py:28
    if (z >= 0 && z < 64) atomicAdd(&s_hist[z], 1);
  }
  __syncthreads();
  if (threadIdx.x < 64 && s_hist[threadIdx.x]) atomicAdd(&hist[threadIdx.x], s_hist[threadIdx.x]);
}

__global__ void utm_forward_kernel(const double* __restrict__ lat, const double* __restrict__ lon, long long n,
                                   double lon0, Krueger k, double* __restrict__ x, double* __restrict__ y) {
  const long long step = 1ll * gridDim.x * blockDim.x;
  for (long long i = 1ll * blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step)
    utm_fwd(lat[i], lon[i], lon0, k, x + i, y + i);
}

__global__ void utm_inverse_kernel(const double* __restrict__ x, const double* __restrict__ y, long long n,
                                   double lon0, Krueger k, double* __restrict__ lat, double* __restrict__ lon) {
  const long long step = 1ll * gridDim.x * blockDim.x;
  for (long long i = 1ll * blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step)
    utm_inv(x[i], y[i], lon0, k, lat + i, lon + i);
}

struct Buckets {
  double x0, y0, edge;   // lower-left corner of bucket (0, 0), bucket edge in metres
  int nx, ny;
};

// pass 1: bucket id per swath pixel (-1: invalid coordinates or too far from the target area to matter) + counts
__global__ void bucket_assign_kernel(const double* __restrict__ lat, const double* __restrict__ lon, int n,
                                     double lon0, Krueger k, Buckets b, int* __restrict__ bucket_of,
                                     int* __restrict__ counts) {
  const int step = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
    const double la = lat[i], lo = lon[i];
    int id = -1;
    if (fabs(la) <= 90.0 && fabs(lo) <= 180.0) {   // pyresample drops invalid swath coordinates
      double x, y;
      utm_fwd(la, lo, lon0, k, &x, &y);
      const double fx = floor((x - b.x0) / b.edge), fy = floor((y - b.y0) / b.edge);
      if (fx >= 0 && fx < b.nx && fy >= 0 && fy < b.ny) {
        id = static_cast<int>(fy) * b.nx + static_cast<int>(fx);
        atomicAdd(&counts[id], 1);
      }
    }
    bucket_of[i] = id;
  }
}

// exclusive scan of the bucket counts (one block; a few thousand buckets)
__global__ void bucket_scan_kernel(const int* __restrict__ counts, int nb, int* __restrict__ start,
                                   int* __restrict__ cursor) {
  __shared__ int s_part[1024];
  const int per = (nb + blockDim.x - 1) / blockDim.x;
  const int lo = threadIdx.x * per, hi = min(lo + per, nb);
  int sum = 0;
  for (int i = lo; i < hi; ++i) sum += counts[i];
  s_part[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int t = 0; t < static_cast<int>(blockDim.x); ++t) {
      const int v = s_part[t];
      s_part[t] = run;
      run += v;
    }
  }
  __syncthreads();
  int run = s_part[threadIdx.x];
  for (int i = lo; i < hi; ++i) {
    start[i] = run;
    cursor[i] = run;
    run += counts[i];
  }
  if (threadIdx.x == blockDim.x - 1) start[nb] = run;
}

// pass 2: scatter (index, Cartesian coordinates) into bucket order
__global__ void bucket_scatter_kernel(const double* __restrict__ lat, const double* __restrict__ lon, int n,
                                      const int* __restrict__ bucket_of, int* __restrict__ cursor,
                                      int* __restrict__ sorted_idx, double* __restrict__ sorted_xyz) {
  const int step = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
    const int id = bucket_of[i];
    if (id < 0) continue;
    const int slot = atomicAdd(&cursor[id], 1);
    double c[3];
    to_cartesian(lat[i], lon[i], c);
    sorted_idx[slot] = i;
    sorted_xyz[3ll * slot + 0] = c[0];
    sorted_xyz[3ll * slot + 1] = c[1];
    sorted_xyz[3ll * slot + 2] = c[2];
  }
}

// one thread per target cell: centre -> lon/lat -> Cartesian, scan the 3 x 3 buckets around it
__global__ void __launch_bounds__(256)
    nearest_kernel(Buckets b, double min_x, double max_y, double psx, double psy, int x_size, int y_size,
                   double lon0, Krueger k, double radius, const int* __restrict__ start,
                   const int* __restrict__ sorted_idx, const double* __restrict__ sorted_xyz,
                   int* __restrict__ out_idx) {
  const long long cell = 1ll * blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= 1ll * x_size * y_size) return;
  const int col = static_cast<int>(cell % x_size), row = static_cast<int>(cell / x_size);
  const double x = min_x + (col + 0.5) * psx, y = max_y - (row + 0.5) * psy;   // row 0 = top
  double lat, lon, t[3];
  utm_inv(x, y, lon0, k, &lat, &lon);
  to_cartesian(lat, lon, t);
  const int bx = static_cast<int>(floor((x - b.x0) / b.edge)), by = static_cast<int>(floor((y - b.y0) / b.edge));
  double best = radius * radius;   // strict upper bound, like pykdtree's distance_upper_bound
  int best_i = -1;
  for (int dy = -1; dy <= 1; ++dy) {
    const int yy = by + dy;
    if (yy < 0 || yy >= b.ny) continue;
    const int xa = max(bx - 1, 0), xb = min(bx + 1, b.nx - 1);
    if (xa > xb) continue;
    // the three buckets of a row are contiguous in the sorted order
    const int s0 = start[yy * b.nx + xa], s1 = start[yy * b.nx + xb + 1];
    for (int s = s0; s < s1; ++s) {
      const double dx = sorted_xyz[3ll * s] - t[0], dyy = sorted_xyz[3ll * s + 1] - t[1],
                   dz = sorted_xyz[3ll * s + 2] - t[2];
      const double d2 = dx * dx + dyy * dyy + dz * dz;
      const int i = sorted_idx[s];
      if (d2 < best || (d2 == best && best_i >= 0 && i < best_i)) {
        best = d2;
        best_i = i;
      }
    }
  }
  out_idx[cell] = best_i;
}

template <typename T>
__global__ void gather_fill_kernel(const T* __restrict__ src, const int* __restrict__ idx, long long n, T fill,
                                   T* __restrict__ out) {
  const long long step = 1ll * gridDim.x * blockDim.x;
  for (long long i = 1ll * blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
    const int j = idx[i];
    out[i] = j >= 0 ? src[j] : fill;
  }
}

int grid_1d(long long n, int threads) {
  const long long b = (n + threads - 1) / threads;
  return static_cast<int>(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

Buckets make_buckets(double min_x, double min_y, double max_x, double max_y, double radius) {
  Buckets b;
  b.edge = 1.1 * radius;
  b.x0 = min_x - b.edge;
  b.y0 = min_y - b.edge;
  b.nx = static_cast<int>(ceil((max_x - min_x) / b.edge)) + 2;
  b.ny = static_cast<int>(ceil((max_y - min_y) / b.edge)) + 2;
  return b;
}
size_t align256(size_t v) { return (v + 255) / 256 * 256; }
}  // namespace

int utm_zone_histogram(const double* lon, long long n, int* hist64, cudaStream_t s) {
  if (cudaMemsetAsync(hist64, 0, 64 * sizeof(int), s) != cudaSuccess) {
    set_error("utm_zone_histogram: cudaMemsetAsync failed");
    return -2;
  }
  if (n <= 0) return 0;
  utm_zone_hist_kernel<<<grid_1d(n, 256), 256, 0, s>>>(lon, n, hist64);
  return check_launch_rs("utm_zone_histogram");
}

int utm_forward(const double* lat, const double* lon, long long n, int zone, double* x, double* y, cudaStream_t s) {
  if (zone < 1 || zone > 60) {
    set_error("utm_forward: zone must be 1..60");
    return -1;
  }
  if (n <= 0) return 0;
  utm_forward_kernel<<<grid_1d(n, 256), 256, 0, s>>>(lat, lon, n, 6.0 * zone - 183.0, make_krueger(), x, y);
  return check_launch_rs("utm_forward");
}

int utm_inverse(const double* x, const double* y, long long n, int zone, double* lat, double* lon, cudaStream_t s) {
  if (zone < 1 || zone > 60) {
    set_error("utm_inverse: zone must be 1..60");
    return -1;
  }
  if (n <= 0) return 0;
  utm_inverse_kernel<<<grid_1d(n, 256), 256, 0, s>>>(x, y, n, 6.0 * zone - 183.0, make_krueger(), lat, lon);
  return check_launch_rs("utm_inverse");
}

// ------------------------------------------------------------------------------------------------
// Geolocation of a MODIS sinusoidal grid (read_modis_aod, tools.py:97-128): x = linspace(x_start, x_stop, nx),
// y = linspace(y_start, y_stop, ny) in metres, meshgrid, inverse spherical sinusoidal projection
// (+proj=sinu +R=6371007.181 +nadgrids=@null -> EPSG:4326: the null grid shift makes the sphere's latitude /
// longitude the output, no datum step).  PROJ's operation order for the sphere: phi = y * (1 / R),
// lam = (x * (1 / R)) / cos(phi), longitude wrapped into [-pi, pi], radians * (180 / pi).  numpy's linspace:
// i * step + start with the last sample set to stop (both products and sums rounded separately: no FMA).
// ------------------------------------------------------------------------------------------------
__global__ void sinu_grid_kernel(double x_start, double x_stop, double x_step, double y_start, double y_stop,
                                 double y_step, int ny, int nx, double inv_r, double* __restrict__ lat,
                                 double* __restrict__ lon) {
  const long long n = 1ll * ny * nx, step = 1ll * gridDim.x * blockDim.x;
  constexpr double kPi = 3.14159265358979323846, kTwoPi = 6.28318530717958647693, kSpi = 3.14159265359;
  constexpr double kRadToDeg = 57.295779513082321;
  for (long long i = 1ll * blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
    const int r = static_cast<int>(i / nx), c = static_cast<int>(i % nx);
    const double x = (nx > 1 && c == nx - 1) ? x_stop : __dadd_rn(__dmul_rn(static_cast<double>(c), x_step), x_start);
    const double y = (ny > 1 && r == ny - 1) ? y_stop : __dadd_rn(__dmul_rn(static_cast<double>(r), y_step), y_start);
    const double phi = __dmul_rn(y, inv_r);
    double lam = __ddiv_rn(__dmul_rn(x, inv_r), cos(phi));
    if (fabs(lam) > kSpi) {                                  // PROJ adjlon
      lam += kPi;
      lam -= kTwoPi * floor(lam / kTwoPi);
      lam -= kPi;
    }
    lat[i] = __dmul_rn(phi, kRadToDeg);
    lon[i] = __dmul_rn(lam, kRadToDeg);
  }
}

int sinusoidal_grid_latlon(double x_start, double x_stop, double y_start, double y_stop, int ny, int nx, double radius,
                           double* lat, double* lon, cudaStream_t s) {
  if (ny <= 0 || nx <= 0) return 0;
  if (!(radius > 0)) {
    set_error("sinusoidal_grid_latlon: radius must be positive");
    return -1;
  }
  const double x_step = nx > 1 ? (x_stop - x_start) / (nx - 1) : 0.0, y_step = ny > 1 ? (y_stop - y_start) / (ny - 1) : 0.0;
  sinu_grid_kernel<<<grid_1d(1ll * ny * nx, 256), 256, 0, s>>>(x_start, x_stop, x_step, y_start, y_stop, y_step, ny, nx,
                                                               1.0 / radius, lat, lon);
  return check_launch_rs("sinusoidal_grid_latlon");
}

size_t resample_workspace_bytes(int n_src, double min_x, double min_y, double max_x, double max_y, double radius) {
  if (n_src <= 0 || !(radius > 0) || !(max_x >= min_x) || !(max_y >= min_y)) return 0;
  const Buckets b = make_buckets(min_x, min_y, max_x, max_y, radius);
  const size_t nb = static_cast<size_t>(b.nx) * b.ny;
  return align256((nb + 1) * 4) * 3 + align256(static_cast<size_t>(n_src) * 4) * 2 +
         align256(static_cast<size_t>(n_src) * 24);
}

int resample_nearest_index(const double* src_lat, const double* src_lon, int n_src, int zone, double min_x,
                           double min_y, double max_x, double max_y, int x_size, int y_size, double radius,
                           void* workspace, size_t workspace_bytes, int* out_idx, cudaStream_t s) {
  if (zone < 1 || zone > 60 || x_size <= 0 || y_size <= 0 || !(radius > 0) || !(max_x > min_x) || !(max_y > min_y)) {
    set_error("resample_nearest_index: bad area definition");
    return -1;
  }
  const long long cells = 1ll * x_size * y_size;
  if (n_src <= 0) {
    cudaMemsetAsync(out_idx, 0xff, cells * sizeof(int), s);   // every cell unfilled (-1)
    return 0;
  }
  const Buckets b = make_buckets(min_x, min_y, max_x, max_y, radius);
  const long long nb = 1ll * b.nx * b.ny;
  if (nb > (1 << 24)) {
    set_error("resample_nearest_index: area too large for the bucket grid");
    return -1;
  }
  const size_t need = resample_workspace_bytes(n_src, min_x, min_y, max_x, max_y, radius);
  if (!workspace || workspace_bytes < need) {
    set_error("resample_nearest_index: workspace too small (plume_resample_workspace_bytes)");
    return -1;
  }
  char* w = static_cast<char*>(workspace);
  int* counts = reinterpret_cast<int*>(w);      w += align256((nb + 1) * 4);
  int* start = reinterpret_cast<int*>(w);       w += align256((nb + 1) * 4);
  int* cursor = reinterpret_cast<int*>(w);      w += align256((nb + 1) * 4);
  int* bucket_of = reinterpret_cast<int*>(w);   w += align256(static_cast<size_t>(n_src) * 4);
  int* sorted_idx = reinterpret_cast<int*>(w);  w += align256(static_cast<size_t>(n_src) * 4);
  double* sorted_xyz = reinterpret_cast<double*>(w);
  if (cudaMemsetAsync(counts, 0, (nb + 1) * 4, s) != cudaSuccess) {
    set_error("resample_nearest_index: cudaMemsetAsync failed");
    return -2;
  }
  const Krueger k = make_krueger();
  const double lon0 = 6.0 * zone - 183.0;
  bucket_assign_kernel<<<grid_1d(n_src, 256), 256, 0, s>>>(src_lat, src_lon, n_src, lon0, k, b, bucket_of, counts);
  bucket_scan_kernel<<<1, 1024, 0, s>>>(counts, static_cast<int>(nb), start, cursor);
  bucket_scatter_kernel<<<grid_1d(n_src, 256), 256, 0, s>>>(src_lat, src_lon, n_src, bucket_of, cursor, sorted_idx,
                                                            sorted_xyz);
  const double psx = (max_x - min_x) / x_size, psy = (max_y - min_y) / y_size;
  nearest_kernel<<<static_cast<unsigned>((cells + 255) / 256), 256, 0, s>>>(
      b, min_x, max_y, psx, psy, x_size, y_size, lon0, k, radius, start, sorted_idx, sorted_xyz, out_idx);
  return check_launch_rs("resample_nearest_index");
}

int gather_fill(const void* src, int elem_bytes, const int* idx, long long n, double fill, void* out,
                cudaStream_t s) {
  if (n <= 0) return 0;
  if (elem_bytes == 4)
    gather_fill_kernel<float><<<grid_1d(n, 256), 256, 0, s>>>(static_cast<const float*>(src), idx, n,
                                                              static_cast<float>(fill), static_cast<float*>(out));
  else if (elem_bytes == 8)
    gather_fill_kernel<double><<<grid_1d(n, 256), 256, 0, s>>>(static_cast<const double*>(src), idx, n, fill,
                                                               static_cast<double*>(out));
  else {
    set_error("gather_fill: element size must be 4 (float32) or 8 (float64)");
    return -1;
  }
  return check_launch_rs("gather_fill");
}

}  // namespace plume
