/* TEST INFRASTRUCTURE -- plain C restatement of the label-generation rows, the same
 * algorithms as oracle/hull_ref.py, oracle/fire_ref.py and oracle/sweep_ref.py, used (a) as a second, independent
 * checker in tests/test_c_oracle.py (C == numpy == golden vectors recorded from the reference's functions) and (b) as
 * the cpu_baseline of scripts/bench_{rasterize,locate_fires,sweep}.py, where the numpy versions are too slow to be a
 * fair CPU number.  Built by oracle/Makefile into oracle/_build/libgeo_ref.so (gcc -O2 -ffp-contract=off: no fused
 * multiply-adds, so the float64 haversine rounds like numpy's).  Nothing under kcl_ltss_bioatm_b200/ or src/ may
 * load this library.
 *
 * Reference lines restated: plume_selector.py:88-116 (in_hull / find_plume_aod), plume_identifier_gaussian_profile.py
 * :65-82 (haversine), :85-106 (locate_fire_in_image search), :142-154 (generate_mask_dict), :157-202
 * (find_plume_extents / extract_label). */
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* mask[r][c] |= 1 where pixel (x, y) = (ox + c, oy + r) is inside or on the boundary of the counter-clockwise convex
 * polygon (vx[i], vy[i]), i < m.  Exact: int64 edge functions. */
void geo_rasterize_hull(const long long* vx, const long long* vy, int m, int H, int W, int oy, int ox,
                        unsigned char* mask) {
  long long xmin = vx[0], xmax = vx[0], ymin = vy[0], ymax = vy[0];
  for (int i = 1; i < m; ++i) {
    if (vx[i] < xmin) xmin = vx[i];
    if (vx[i] > xmax) xmax = vx[i];
    if (vy[i] < ymin) ymin = vy[i];
    if (vy[i] > ymax) ymax = vy[i];
  }
  for (int r = 0; r < H; ++r) {
    const long long py = oy + r;
    if (py < ymin || py > ymax) continue;
    for (int c = 0; c < W; ++c) {
      const long long px = ox + c;
      if (px < xmin || px > xmax) continue;
      int in = 1;
      for (int i = 0; i < m && in; ++i) {
        const int j = (i + 1) % m;
        const long long e = (vx[j] - vx[i]) * (py - vy[i]) - (vy[j] - vy[i]) * (px - vx[i]);
        if (e < 0) in = 0;
      }
      if (in) mask[(size_t)r * W + c] = 1;
    }
  }
}

static double haversine_km(double lon1, double lat1, double lon2, double lat2) {
  const double k = 3.141592653589793238462643383279502884 / 180.0;
  lon1 *= k; lat1 *= k; lon2 *= k; lat2 *= k;
  const double sdlat = sin((lat2 - lat1) / 2.0), sdlon = sin((lon2 - lon1) / 2.0);
  const double a = sdlat * sdlat + cos(lat1) * cos(lat2) * (sdlon * sdlon);
  return 6367 * (2 * asin(sqrt(a)));
}

/* out_rc[2f], out_rc[2f+1] = row, col of the first (row-major) pixel with the smallest haversine distance among the
 * pixels strictly inside the +-half box around fire f, or -1, -1. */
void geo_locate_fires(const double* lats, const double* lons, int H, int W, const double* flat, const double* flon,
                      int n, double half, long long* out_rc) {
  const size_t pixels = (size_t)H * W;
  for (int f = 0; f < n; ++f) {
    const double la0 = flat[f] - half, la1 = flat[f] + half, lo0 = flon[f] - half, lo1 = flon[f] + half;
    double best = 0;
    long long arg = -1;
    for (size_t i = 0; i < pixels; ++i) {
      if (lats[i] > la0 && lats[i] < la1 && lons[i] > lo0 && lons[i] < lo1) {
        const double d = haversine_km(flon[f], flat[f], lons[i], lats[i]);
        if (arg < 0 || d < best) { best = d; arg = (long long)i; }
      }
    }
    out_rc[2 * f] = arg < 0 ? -1 : arg / W;
    out_rc[2 * f + 1] = arg < 0 ? -1 : arg % W;
  }
}

/* out = dilate(erode(aod > t)), cross footprint; erosion sees set pixels beyond the border, dilation unset ones.
 * tmp: H*W scratch. */
void geo_threshold_mask(const float* aod, int H, int W, double t, unsigned char* tmp, unsigned char* out) {
#define SET(y, x) ((y) < 0 || (y) >= H || (x) < 0 || (x) >= W || (double)aod[(size_t)(y) * W + (x)] > t)
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x)
      tmp[(size_t)y * W + x] = SET(y, x) && SET(y - 1, x) && SET(y + 1, x) && SET(y, x - 1) && SET(y, x + 1);
#undef SET
#define ER(y, x) ((y) >= 0 && (y) < H && (x) >= 0 && (x) < W && tmp[(size_t)(y) * W + (x)])
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x)
      out[(size_t)y * W + x] = ER(y, x) || ER(y - 1, x) || ER(y + 1, x) || ER(y, x - 1) || ER(y, x + 1);
#undef ER
}

static int uf_find(int* p, int x) {
  while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; }
  return x;
}
static void uf_union(int* p, int a, int b) {
  a = uf_find(p, a); b = uf_find(p, b);
  if (a == b) return;
  if (a < b) p[b] = a; else p[a] = b;   /* the root is always the smaller index */
}

/* 8-connected components: labels[i] = -1 for background, else the smallest row-major index of i's component;
 * sizes[i] = component size at canonical indices, 0 elsewhere. */
void geo_label8(const unsigned char* mask, int H, int W, int* labels, int* sizes) {
  const int n = H * W;
  for (int i = 0; i < n; ++i) { labels[i] = mask[i] ? i : -1; sizes[i] = 0; }
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const int i = y * W + x;
      if (!mask[i]) continue;
      if (x > 0 && mask[i - 1]) uf_union(labels, i, i - 1);
      if (y > 0) {
        if (x > 0 && mask[i - W - 1]) uf_union(labels, i, i - W - 1);
        if (mask[i - W]) uf_union(labels, i, i - W);
        if (x + 1 < W && mask[i - W + 1]) uf_union(labels, i, i - W + 1);
      }
    }
  for (int i = 0; i < n; ++i)
    if (labels[i] >= 0) { labels[i] = uf_find(labels, i); sizes[labels[i]]++; }
}

/* extents[f] = size of the component nearest to fire f inside its (2 win + 1)^2 window (squared Euclidean pixel
 * distance, first in row-major window order on ties), 0 if none. */
void geo_fire_extents(const int* labels, const int* sizes, int H, int W, const long long* fire_rc, int n, int win,
                      int* extents) {
  for (int f = 0; f < n; ++f) {
    const int r = (int)fire_rc[2 * f], c = (int)fire_rc[2 * f + 1];
    long long best = -1;
    int lab = -1;
    for (int dy = -win; dy <= win; ++dy)
      for (int dx = -win; dx <= win; ++dx) {
        const int y = r + dy, x = c + dx;
        if (y < 0 || y >= H || x < 0 || x >= W) continue;
        const int l = labels[y * W + x];
        if (l < 0) continue;
        const long long d2 = (long long)dy * dy + (long long)dx * dx;
        if (best < 0 || d2 < best) { best = d2; lab = l; }
      }
    extents[f] = lab < 0 ? 0 : sizes[lab];
  }
}
