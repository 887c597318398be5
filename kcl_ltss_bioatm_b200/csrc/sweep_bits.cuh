// Bit-plane form of the threshold sweep (sweep.cu): the per-thread bodies of the kernels, written as
// __host__ __device__ functions so that tests/emu/sweep_bits_emu.cpp can run the very same code sequentially on
// the CPU against the oracle (test infrastructure only: nothing in the package calls the host instantiation).
//
// Data:
//   bits[t][y][seg]   uint32, bit i of word seg <-> pixel x = 32 seg + i of row y (segs = ceil(W / 32); bits
//                     beyond W are zero).  32 x smaller than an int32 label plane, 8 x smaller than a byte mask.
//   ent[t][y * pitch + (x >> 1)]   int2 {parent, size}, only touched where x is the first pixel of a WORD-LOCAL run
//                     (a maximal run of set bits inside one word).  Two run starts of a row are at least two
//                     pixels apart, so x >> 1 is unique per run; pitch = 16 segs.  A run is named by this entry
//                     index, which grows in row-major order, and union-find links always point to the smaller
//                     index: a component's root is its first run in row-major order.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define PLUME_HD __host__ __device__ __forceinline__
#else
#define PLUME_HD inline
struct int2 {
  int x, y;
};
#endif

namespace plume {
namespace sweepbits {

PLUME_HD int clz32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __clz(static_cast<int>(v));
#else
  return v ? __builtin_clz(v) : 32;
#endif
}
PLUME_HD int ctz32(uint32_t v) {   // v != 0
#if defined(__CUDA_ARCH__)
  return __ffs(static_cast<int>(v)) - 1;
#else
  return __builtin_ctz(v);
#endif
}
PLUME_HD int ctz64(uint64_t v) {   // v != 0
#if defined(__CUDA_ARCH__)
  return __ffsll(static_cast<long long>(v)) - 1;
#else
  return __builtin_ctzll(v);
#endif
}
PLUME_HD uint32_t low_mask32(int n) { return n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1u); }   // n lowest bits, 0 <= n

struct Geom {
  int H, W, segs, pitch;          // pitch = entries per row = 16 segs
  long long words_per_plane;      // H * segs
  long long ent_per_plane;        // H * pitch
};
PLUME_HD Geom make_geom(int H, int W) {
  Geom g;
  g.H = H;
  g.W = W;
  g.segs = (W + 31) / 32;
  g.pitch = 16 * g.segs;
  g.words_per_plane = 1ll * H * g.segs;
  g.ent_per_plane = 1ll * H * g.pitch;
  return g;
}
PLUME_HD int ent_index(const Geom& g, int y, int x) { return y * g.pitch + (x >> 1); }

// first bit of the word-local run of w that contains bit b (bit b is set)
PLUME_HD int run_start(uint32_t w, int b) {
  const uint32_t zeros = ~w & low_mask32(b);
  return zeros ? 32 - clz32(zeros) : 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Opening (binary_dilation(binary_erosion(.)), cross footprint) on 36-bit row windows.  Bit k of a window <-> column
// x = 32 seg - 2 + k.  A warp walks down a 32-column strip; the lane that owns a threshold keeps the two previous
// rows of B = (aod > t) and of E = erosion(B) and emits the finished word two rows behind the row just read.
//   B: columns / rows outside the image count as set (the erosion's border rule);
//   E: zero outside the image (the dilation's border rule).
// ---------------------------------------------------------------------------------------------------------------
constexpr uint64_t kWin36 = (1ull << 36) - 1ull;

// columns of the window that lie inside the image
PLUME_HD uint64_t window_colmask(int seg, int W) {
  const int lo = 2 - 32 * seg > 0 ? 2 - 32 * seg : 0;                 // first k with x >= 0
  int hi = W - 1 - 32 * seg + 2;                                     // last k with x < W
  if (hi > 35) hi = 35;
  if (hi < lo) return 0;
  const uint64_t upto_hi = hi >= 63 ? ~0ull : ((1ull << (hi + 1)) - 1ull);
  return upto_hi & ~((1ull << lo) - 1ull);
}

// Thresholds of a chunk as an ascending 32-entry table (padded with +inf): the number of entries below v, by a
// branch-free binary search.  A pixel's comparison results against all thresholds are then the low `count` bits.
template <typename V>
PLUME_HD int count_below(const V* s, V v) {
  int k = 0;
  if (v > s[15]) k = 16;
  if (v > s[k + 7]) k += 8;
  if (v > s[k + 3]) k += 4;
  if (v > s[k + 1]) k += 2;
  if (v > s[k]) k += 1;
  if (k == 31 && v > s[31]) k = 32;
  return k;
}
// position of threshold i in the ascending order of th[0..n) (equal values keep their order)
template <typename V>
PLUME_HD int rank_of(const V* th, int n, int i) {
  int r = 0;
  for (int j = 0; j < n; ++j) r += (th[j] < th[i] || (th[j] == th[i] && j < i)) ? 1 : 0;
  return r;
}
// One stage of the 32 x 32 bit-matrix transpose across a warp (lane l holds row l; stages j = 16, 8, 4, 2, 1;
// y = the word of lane l ^ j).  Afterwards lane t holds column t: bit l = bit t of lane l's original word.
PLUME_HD uint32_t transpose32_step(uint32_t x, uint32_t y, int lane, int j) {
  const uint32_t mask = j == 16 ? 0x0000FFFFu : j == 8 ? 0x00FF00FFu : j == 4 ? 0x0F0F0F0Fu : j == 2 ? 0x33333333u : 0x55555555u;
  return (lane & j) ? ((x & ~mask) | ((y >> j) & mask)) : ((x & mask) | ((y << j) & ~mask));
}

struct OpenState {
  uint64_t b1, b2;   // B of the previous row and of the one before
  uint64_t e1, e2;   // E of (row - 2) and (row - 3)
};

// Feed B of row `yy` (`prev_row_inside`: row yy - 1 lies inside the image); returns the opened word of row yy - 2,
// columns 32 seg .. 32 seg + 31.  Meaningful from the fifth call on a strip (the caller skips the first four).
PLUME_HD uint32_t open_step(OpenState& s, uint64_t bcur, bool prev_row_inside, uint64_t colmask) {
  uint64_t e0 = s.b1 & (s.b1 << 1) & (s.b1 >> 1) & s.b2 & bcur;        // E(yy - 1)
  e0 = prev_row_inside ? (e0 & colmask) : 0ull;
  const uint64_t d = s.e1 | (s.e1 << 1) | (s.e1 >> 1) | s.e2 | e0;     // D(yy - 2)
  s.b2 = s.b1;
  s.b1 = bcur;
  s.e2 = s.e1;
  s.e1 = e0;
  return static_cast<uint32_t>((d & colmask) >> 2);
}

// ---------------------------------------------------------------------------------------------------------------
// Union-find over run entries of one plane.
// ---------------------------------------------------------------------------------------------------------------
PLUME_HD int uf_root(const int2* ent, int x) {
  int p = ent[x].x;
  while (p != x) {
    x = p;
    p = ent[x].x;
  }
  return x;
}
// the same with path halving: every visited entry is re-pointed at its grandparent by a plain store.  Safe next to
// the atomicMin of uf_unite: only non-roots are written, always with an ancestor (a smaller index of the same set);
// a lost atomicMin result is harmless because the thread that issued it goes on uniting the previous parent with its
// target (uf_unite).  Measured on 1200 x 1200 x 75 planes: merge + flatten 163 -> 80 us; a band-local phase in shared
// memory and skipping unions the left neighbour also makes gained nothing on top (gpurun_out/r2ab).
PLUME_HD int uf_root_halving(int2* ent, int x) {
  int p = ent[x].x;
  while (p != x) {
    const int gp = ent[p].x;
    if (gp != p) ent[x].x = gp;
    x = p;
    p = gp;
  }
  return x;
}
PLUME_HD int uf_atomic_min(int2* ent, int a, int b) {
#if defined(__CUDA_ARCH__)
  return atomicMin(&ent[a].x, b);
#else
  const int old = ent[a].x;
  if (b < old) ent[a].x = b;
  return old;
#endif
}
PLUME_HD void uf_unite(int2* ent, int a, int b) {
  while (true) {
    a = uf_root_halving(ent, a);
    b = uf_root_halving(ent, b);
    if (a == b) return;
    if (a < b) {
      const int t = a;
      a = b;
      b = t;
    }
    // a > b: hang root a under b; if a stopped being a root meanwhile, carry on from where it points now
    // (the atomicMin may have re-hung a under b; uniting its previous parent with b keeps everything joined)
    const int old = uf_atomic_min(ent, a, b);
    if (old == a) return;
    a = old;
  }
}

// init: every word-local run of word (y, seg) becomes a singleton with size 0
PLUME_HD void init_word(uint32_t w, int2* ent, const Geom& g, int y, int seg) {
  uint32_t starts = w & ~(w << 1);
  while (starts) {
    const int s = ctz32(starts);
    starts &= starts - 1u;
    const int e = ent_index(g, y, 32 * seg + s);
    int2 v;
    v.x = e;
    v.y = 0;
    ent[e] = v;
  }
}

// merge: join every run of word (y, seg) with the run it continues from in the word to its left and with every
// run of the row above that touches it (8-connectivity: columns start - 1 .. end + 1).
//   left: word to the left in the same row (0 if seg == 0); up_l / up_c / up_r: the three words above (0 where absent)
PLUME_HD void merge_word(uint32_t w, uint32_t left, uint32_t up_l, uint32_t up_c, uint32_t up_r, int2* ent,
                         const Geom& g, int y, int seg) {
  // 34-bit window of the row above: bit j <-> column 32 seg - 1 + j
  const uint64_t up = static_cast<uint64_t>(up_l >> 31) | (static_cast<uint64_t>(up_c) << 1) |
                      (static_cast<uint64_t>(up_r & 1u) << 33);
  uint32_t rest = w;
  while (rest) {
    const int s = ctz32(rest);
    const uint32_t from_s = rest >> s;                                  // the run starts at bit 0 of this
    const int len = (~from_s) ? ctz32(~from_s) : 32;                    // s == 0 and the word is full: 32
    rest &= ~(low_mask32(len) << s);
    const int me = ent_index(g, y, 32 * seg + s);
    if (s == 0 && (left >> 31)) uf_unite(ent, me, ent_index(g, y, 32 * (seg - 1) + run_start(left, 31)));
    // columns s - 1 .. s + len of this word  <->  window bits s .. s + len + 1
    uint64_t touch = up & (((1ull << (len + 2)) - 1ull) << s);
    while (touch) {
      const int j = ctz64(touch);
      const uint64_t from_j = touch >> j;
      const int l2 = (~from_j) ? ctz64(~from_j) : 64;
      touch &= ~(((l2 >= 64) ? ~0ull : ((1ull << l2) - 1ull)) << j);
      // window bit j is column 32 seg - 1 + j of the row above: find the start of its word-local run
      int useg, ubit;
      uint32_t uw;
      if (j == 0) {
        useg = seg - 1, ubit = 31, uw = up_l;
      } else if (j == 33) {
        useg = seg + 1, ubit = 0, uw = up_r;
      } else {
        useg = seg, ubit = j - 1, uw = up_c;
      }
      // the rest of the stretch is joined to its first pixel by the row above itself (its own `left` rule)
      uf_unite(ent, me, ent_index(g, y - 1, 32 * useg + run_start(uw, ubit)));
    }
  }
}

// flatten: pops the lowest run of `rest`; returns its entry index and length
PLUME_HD int pop_run(uint32_t& rest, const Geom& g, int y, int seg, int& len) {
  const int s = ctz32(rest);
  const uint32_t from_s = rest >> s;
  len = (~from_s) ? ctz32(~from_s) : 32;
  rest &= ~(low_mask32(len) << s);
  return ent_index(g, y, 32 * seg + s);
}

// component extraction (after flatten: every run's parent is its root): the bits of word (y, seg) that belong to the
// component `root`
PLUME_HD uint32_t component_word(uint32_t w, const int2* ent, const Geom& g, int y, int seg, int root) {
  uint32_t out = 0, rest = w;
  while (rest) {
    const int s = ctz32(rest);
    const uint32_t from_s = rest >> s;
    const int len = (~from_s) ? ctz32(~from_s) : 32;
    const uint32_t run = low_mask32(len) << s;
    rest &= ~run;
    if (ent[ent_index(g, y, 32 * seg + s)].x == root) out |= run;
  }
  return out;
}

// ---------------------------------------------------------------------------------------------------------------
// extents: nearest set pixel to (r, c) in the (2 win + 1)^2 window; key = (squared distance << 32) | window index,
// the minimum key is numpy's first minimum over the row-major window.
// ---------------------------------------------------------------------------------------------------------------
constexpr unsigned long long kNoKey = ~0ull;

PLUME_HD unsigned long long window_key(int dy, int dx, int win) {
  const int side = 2 * win + 1;
  return (static_cast<unsigned long long>(dy * dy + dx * dx) << 32) |
         static_cast<unsigned>((dy + win) * side + (dx + win));
}

// best key of window row dy (image row r + dy, inside the image), reading the row's words from `row_bits`
PLUME_HD unsigned long long best_in_row(const uint32_t* row_bits, const Geom& g, int c, int dy, int win) {
  const int x0 = c - win > 0 ? c - win : 0;
  const int x1 = c + win < g.W - 1 ? c + win : g.W - 1;
  unsigned long long best = kNoKey;
  for (int sg = x0 >> 5; sg <= (x1 >> 5); ++sg) {
    const int lo = x0 - 32 * sg > 0 ? x0 - 32 * sg : 0;
    const int hi = x1 - 32 * sg < 31 ? x1 - 32 * sg : 31;
    const uint32_t m = row_bits[sg] & low_mask32(hi + 1) & ~low_mask32(lo);
    if (!m) continue;
    const int cl = c - 32 * sg;                                         // the fire's column inside this word
    const uint32_t at_or_left = cl < 0 ? 0u : (m & low_mask32(cl + 1));
    const uint32_t right = m & ~at_or_left;
    if (at_or_left) {
      const unsigned long long k = window_key(dy, 32 * sg + (31 - clz32(at_or_left)) - c, win);
      best = k < best ? k : best;
    }
    if (right) {
      const unsigned long long k = window_key(dy, 32 * sg + ctz32(right) - c, win);
      best = k < best ? k : best;
    }
  }
  return best;
}

// root entry of the component of the pixel a key names (after flatten: a run's parent is its root), -1 for no key
PLUME_HD int root_of_key(unsigned long long key, const uint32_t* plane_bits, const int2* ent, const Geom& g, int r, int c,
                         int win) {
  if (key == kNoKey) return -1;
  const int side = 2 * win + 1;
  const int k = static_cast<int>(key & 0xFFFFFFFFull);
  const int y = r + k / side - win, x = c + k % side - win;
  const uint32_t w = plane_bits[1ll * y * g.segs + (x >> 5)];
  return ent[ent_index(g, y, (x & ~31) + run_start(w, x & 31))].x;
}

// size of the component of the pixel a key names (after flatten: a run's parent is its root)
PLUME_HD int extent_of_key(unsigned long long key, const uint32_t* plane_bits, const int2* ent, const Geom& g, int r,
                           int c, int win) {
  if (key == kNoKey) return 0;
  const int side = 2 * win + 1;
  const int k = static_cast<int>(key & 0xFFFFFFFFull);
  const int y = r + k / side - win, x = c + k % side - win;
  const uint32_t w = plane_bits[1ll * y * g.segs + (x >> 5)];
  const int e = ent_index(g, y, (x & ~31) + run_start(w, x & 31));
  return ent[ent[e].x].y;
}

}  // namespace sweepbits
}  // namespace plume
