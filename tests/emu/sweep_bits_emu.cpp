// TEST INFRASTRUCTURE.  Sequential CPU run of the per-thread bodies of the bit-plane sweep kernels
// (kcl_ltss_bioatm_b200/csrc/sweep_bits.cuh, the same source the CUDA kernels compile) so that the bit logic is
// checked against the oracle without a GPU (tests/test_sweep_bits_emu.py builds this with g++).  The warp-level
// parts of the kernels (ballots, shuffles) are replaced by loops over 32 lanes with the kernel's own index
// arithmetic; threads run one after the other in an order the caller can permute.  Nothing in the package links this.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../kcl_ltss_bioatm_b200/csrc/sweep_bits.cuh"

using namespace plume::sweepbits;

namespace {
float round_down_to_float(double t) {          // __double2float_rd
  float f = static_cast<float>(t);
  if (static_cast<double>(f) > t) f = std::nextafterf(f, -INFINITY);
  return f;
}

// mask_bits_kernel<float, true>: two adjacent strips per warp; per lane the count of thresholds below its pixel, a
// 32 x 32 bit transpose across the "warp", lane r owns the r-th smallest threshold of the chunk
void mask_bits(const float* aod, int H, int W, const double* thr, int T, uint32_t* bits, int kStripRows) {
  const int segs = (W + 31) / 32, pairs = (segs + 1) / 2;
  const int strips = (H + kStripRows - 1) / kStripRows;
  for (int chunk = 0; chunk < (T + 31) / 32; ++chunk) {
    const int Tc = T - 32 * chunk < 32 ? T - 32 * chunk : 32;
    float raw[32], sorted[32];
    int plane_of[32];
    for (int i = 0; i < 32; ++i) {
      float mine = i < Tc ? round_down_to_float(thr[32 * chunk + i]) : INFINITY;
      if (mine != mine) mine = INFINITY;
      raw[i] = mine;
      sorted[i] = INFINITY;
    }
    for (int i = 0; i < Tc; ++i) {
      const int r = rank_of(raw, Tc, i);
      sorted[r] = raw[i];
      plane_of[r] = i;
    }
    for (int wg = 0; wg < pairs * strips; ++wg) {
      const int sp = wg % pairs, y0 = (wg / pairs) * kStripRows;
      const int y_end = y0 + kStripRows < H ? y0 + kStripRows : H;
      const int seg0 = 2 * sp, seg1 = 2 * sp + 1;
      const uint64_t colmask0 = window_colmask(seg0, W), colmask1 = window_colmask(seg1, W);
      const uint64_t outcols0 = ~colmask0 & kWin36, outcols1 = ~colmask1 & kWin36;
      OpenState st0[32], st1[32];
      std::memset(st0, 0, sizeof(st0));
      std::memset(st1, 0, sizeof(st1));
      for (int yy = y0 - 2; yy <= y_end + 1; ++yy) {
        uint32_t c0[32], c1[32], me[32];
        const bool row_in = yy >= 0 && yy < H;
        if (row_in) {
          for (int lane = 0; lane < 32; ++lane) {
            const int x0 = 64 * sp + lane, x1 = x0 + 32;
            const int xe = lane < 2 ? 64 * sp - 2 + lane : 64 * sp + 62 + lane;
            const bool in0 = x0 < W, in1 = x1 < W, ine = lane < 4 && xe >= 0 && xe < W;
            c0[lane] = low_mask32(count_below(sorted, in0 ? aod[1ll * yy * W + x0] : 0.f));
            c1[lane] = low_mask32(count_below(sorted, in1 ? aod[1ll * yy * W + x1] : 0.f));
            me[lane] = low_mask32(count_below(sorted, ine ? aod[1ll * yy * W + xe] : 0.f));
          }
          for (int j = 16; j > 0; j >>= 1) {
            uint32_t n0[32], n1[32];
            for (int lane = 0; lane < 32; ++lane) {
              n0[lane] = transpose32_step(c0[lane], c0[lane ^ j], lane, j);
              n1[lane] = transpose32_step(c1[lane], c1[lane ^ j], lane, j);
            }
            std::memcpy(c0, n0, sizeof(c0));
            std::memcpy(c1, n1, sizeof(c1));
          }
        }
        for (int lane = 0; lane < 32; ++lane) {
          uint64_t b0 = kWin36, b1 = kWin36;
          if (row_in) {
            uint32_t e = 0;
            for (int j = 0; j < 4; ++j) e |= ((me[j] >> lane) & 1u) << j;
            b0 = (static_cast<uint64_t>(c0[lane]) << 2) | (e & 3u) | (static_cast<uint64_t>(c1[lane] & 3u) << 34) | outcols0;
            b1 = (static_cast<uint64_t>(c1[lane]) << 2) | (c0[lane] >> 30) | (static_cast<uint64_t>((e >> 2) & 3u) << 34) | outcols1;
          }
          const bool prev_in = yy - 1 >= 0 && yy - 1 < H;
          const uint32_t w0 = open_step(st0[lane], b0, prev_in, colmask0);
          const uint32_t w1 = open_step(st1[lane], b1, prev_in, colmask1);
          const int r = yy - 2;
          if (r >= y0 && lane < Tc) {
            const long long plane = 32 * chunk + plane_of[lane];
            bits[(plane * H + r) * segs + seg0] = w0;
            if (seg1 < segs) bits[(plane * H + r) * segs + seg1] = w1;
          }
        }
      }
    }
  }
}

// a fixed pseudo-random permutation of 0..n-1 (seed 0: identity) -- thread order must not matter
std::vector<long long> order(long long n, unsigned seed) {
  std::vector<long long> o(n);
  for (long long i = 0; i < n; ++i) o[i] = i;
  if (seed) {
    uint64_t s = seed * 0x9E3779B97F4A7C15ull + 1;
    for (long long i = n - 1; i > 0; --i) {
      s = s * 6364136223846793005ull + 1442695040888963407ull;
      const long long j = static_cast<long long>((s >> 33) % static_cast<uint64_t>(i + 1));
      std::swap(o[i], o[j]);
    }
  }
  return o;
}
}  // namespace

extern "C" {

void emu_mask_bits(const float* aod, int H, int W, const double* thr, int T, uint32_t* bits, int strip_rows) {
  mask_bits(aod, H, W, thr, T, bits, strip_rows);
}

void emu_pack_bits(const uint8_t* masks, int T, int H, int W, uint32_t* bits) {
  const int segs = (W + 31) / 32;
  for (long long row = 0; row < 1ll * T * H; ++row)
    for (int sg = 0; sg < segs; ++sg) {
      uint32_t w = 0;
      for (int l = 0; l < 32; ++l)
        if (32 * sg + l < W && masks[row * W + 32 * sg + l]) w |= 1u << l;
      bits[row * segs + sg] = w;
    }
}

// init + merge + flatten + extents; `ent` is scratch of T * H * 16 * segs int2 (garbage-filled by the caller on purpose)
void emu_bits_extents(const uint32_t* bits, int T, int H, int W, const int* fire_rc, int n_fires, int win, int2* ent_all,
                      int* extents, unsigned seed) {
  const Geom g = make_geom(H, W);
  const long long words = g.words_per_plane * T;
  auto pos = [&](long long idx, int& t, int& y, int& seg) {
    t = static_cast<int>(idx / g.words_per_plane);
    const int rem = static_cast<int>(idx % g.words_per_plane);
    y = rem / g.segs;
    seg = rem % g.segs;
  };
  for (long long idx : order(words, seed)) {
    int t, y, seg;
    pos(idx, t, y, seg);
    if (bits[idx]) init_word(bits[idx], ent_all + t * g.ent_per_plane, g, y, seg);
  }
  for (long long idx : order(words, seed ? seed + 1 : 0)) {
    int t, y, seg;
    pos(idx, t, y, seg);
    const uint32_t w = bits[idx];
    if (!w) continue;
    const uint32_t left = seg > 0 ? bits[idx - 1] : 0u;
    uint32_t up_l = 0, up_c = 0, up_r = 0;
    if (y > 0) {
      const uint32_t* up = bits + idx - g.segs;
      up_c = up[0];
      if (seg > 0) up_l = up[-1];
      if (seg + 1 < g.segs) up_r = up[1];
    }
    merge_word(w, left, up_l, up_c, up_r, ent_all + t * g.ent_per_plane, g, y, seg);
  }
  for (long long idx : order(words, seed ? seed + 2 : 0)) {
    int t, y, seg;
    pos(idx, t, y, seg);
    uint32_t rest = bits[idx];
    int2* ent = ent_all + t * g.ent_per_plane;
    while (rest) {
      int len;
      const int e = pop_run(rest, g, y, seg, len);
      const int root = uf_root(ent, e);
      if (root != e) ent[e].x = root;
      ent[root].y += len;
    }
  }
  for (int t = 0; t < T; ++t)
    for (int f = 0; f < n_fires; ++f) {
      const int r = fire_rc[2 * f], c = fire_rc[2 * f + 1];
      const uint32_t* plane = bits + t * g.words_per_plane;
      unsigned long long best = kNoKey;
      for (int lane = 0; lane < 32; ++lane)
        for (int dy = -win + lane; dy <= win; dy += 32) {
          const int y = r + dy;
          if (y < 0 || y >= H) continue;
          const unsigned long long k = best_in_row(plane + 1ll * y * g.segs, g, c, dy, win);
          best = k < best ? k : best;
        }
      extents[t * n_fires + f] = extent_of_key(best, plane, ent_all + t * g.ent_per_plane, g, r, c, win);
    }
}

// fire_components_kernel after emu_bits_extents left the flattened entries in ent_all
void emu_fire_components(const uint32_t* bits, int T, int H, int W, const int* fire_rc, const int* plane_of_fire,
                         int n_fires, int win, const int2* ent_all, uint32_t* comp, int* stats) {
  const Geom g = make_geom(H, W);
  for (int f = 0; f < n_fires; ++f) {
    const int p = plane_of_fire[f];
    const bool has_plane = p >= 0 && p < T;
    const uint32_t* plane = bits + (has_plane ? p : 0) * g.words_per_plane;
    const int2* ent = ent_all + (has_plane ? p : 0) * g.ent_per_plane;
    int root = -1;
    if (has_plane) {
      const int r = fire_rc[2 * f], c = fire_rc[2 * f + 1];
      unsigned long long best = kNoKey;
      for (int dy = -win; dy <= win; ++dy) {
        const int y = r + dy;
        if (y < 0 || y >= H) continue;
        const unsigned long long k = best_in_row(plane + 1ll * y * g.segs, g, c, dy, win);
        best = k < best ? k : best;
      }
      root = root_of_key(best, plane, ent, g, r, c, win);
    }
    int area = 0, y_min = H, x_min = W, y_max = 0, x_max = 0;
    for (long long i = 0; i < g.words_per_plane; ++i) {
      uint32_t o = 0;
      if (root >= 0 && plane[i]) {
        const int y = static_cast<int>(i / g.segs), seg = static_cast<int>(i % g.segs);
        o = component_word(plane[i], ent, g, y, seg, root);
        if (o) {
          area += __builtin_popcount(o);
          y_min = y < y_min ? y : y_min;
          y_max = y + 1 > y_max ? y + 1 : y_max;
          const int lo = 32 * seg + ctz32(o), hi = 32 * seg + 32 - clz32(o);
          x_min = lo < x_min ? lo : x_min;
          x_max = hi > x_max ? hi : x_max;
        }
      }
      comp[f * g.words_per_plane + i] = o;
    }
    const int st[8] = {area, area ? y_min : 0, area ? x_min : 0, area ? y_max : 0, area ? x_max : 0, root, 0, 0};
    std::memcpy(stats + 8 * f, st, sizeof(st));
  }
}

long long emu_ent_count(int H, int W, int T) { return make_geom(H, W).ent_per_plane * T; }
}
