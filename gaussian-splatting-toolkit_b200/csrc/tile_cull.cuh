// tile_cull.cuh — exact (conservative) tile culling shared by binning.cu and binning_fast.cu.
//
// A (Gaussian, tile) pair of the reference's bounding-box list (helpers.cuh:11-34) can only change a pixel if some
// pixel p of the tile has alpha = min(0.999, o exp(-sigma(p))) >= 1/255 (forward.cu:360-363), i.e.
// q(p) = 2 sigma(p) = a dx^2 + 2 b dx dy + c dy^2 <= thr = 2 ln(255 o).  The reference `continue`s on every pixel of
// the other pairs, so dropping them changes nothing.  The sub-level set {q <= thr} is an ellipse; per TILE ROW its
// intersection with the row's band of pixel rows is bounded left / right by
//     l(dy) = (-b dy - sqrt(a thr - det dy^2)) / a   (convex)      r(dy) = (-b dy + sqrt(a thr - det dy^2)) / a   (concave)
// whose extrema over the band are attained at the ellipse's leftmost / rightmost point (dy = +- b ext_x / c) clamped into
// the band, so one row costs two square roots and yields a contiguous range of tiles — instead of one 45-instruction
// minimisation per tile.  Margins: thr carries +0.1 % + 1e-3 PLUS a term that scales with the magnitude of the
// quadratic form's terms at the far end of the bounding box (`extent` pixels): for long thin slanted Gaussians
// q = a dx^2 + 2 b dx dy + c dy^2 is a difference of terms ~1e5 that cancel to ~10, and the blend kernels / the
// reference evaluate it in a different order (0.5 (a dx^2 + c dy^2) + b dx dy), so their FP32 rounding differs by up to a
// few ulp of the LARGEST term; tile rectangles are taken unclipped (conservative as well).
// Degenerate inputs (non-positive-definite conic, NaN) are never culled.
#pragma once
#include "common.cuh"

namespace gsr {

struct CullEllipse {
  float a, b, inv_a, det, thr, ext_y, dy_left;  // dy_left = dy of the leftmost point; rightmost is -dy_left
  bool never_cull;  // keep the whole bounding box
  bool empty;       // opacity < 1/255: no pixel can ever pass the alpha test
};

GSR_HD CullEllipse make_cull_ellipse(float a, float b, float c, float opac, float extent = 0.f) {
  CullEllipse e;
  e.a = a;
  e.b = b;
  e.det = a * c - b * b;
  e.never_cull = !(a > 0.f && c > 0.f && e.det > 0.f) || !(opac == opac);
  e.empty = !e.never_cull && (255.f * opac < 0.999f);
  // 8 ulp (2^-23 each) of the sum of the absolute terms at distance `extent` in both coordinates
  e.thr = 2.f * gsr_logf(255.f * opac) * 1.001f + 1e-3f + 9.6e-7f * (a + 2.f * fabsf(b) + c) * extent * extent;
  e.inv_a = 1.f / a;
  const float inv_det = 1.f / e.det;
  e.ext_y = sqrtf(a * e.thr * inv_det);
  const float ext_x = sqrtf(c * e.thr * inv_det);
  e.dy_left = b * ext_x / c;
  return e;
}

// Tiles [j0, j1) of tile row `i` (clipped to the bounding box [x0, x1)) that the ellipse can reach; j0 >= j1 if none.
// inv_bw = 1.f / (float)block_width, hoisted out of the caller's row loop (the compiler does not hoist the MUFU.RCP of
// the fast-math division out of a divergent loop)
GSR_HD void cull_row_range(const CullEllipse &e, float mx, float my, int i, int x0, int x1,
                                               int block_width, float inv_bw, int &j0, int &j1) {
  if (e.never_cull) {
    j0 = x0;
    j1 = x1;
    return;
  }
  const float bw = (float)block_width;
  const float lo = fmaxf((float)(i * block_width) - my, -e.ext_y);
  const float hi = fminf((float)(i * block_width + block_width - 1) - my, e.ext_y);
  if (!(lo <= hi)) {
    j0 = j1 = x0;
    return;
  }
  const float dyl = fminf(fmaxf(e.dy_left, lo), hi), dyr = fminf(fmaxf(-e.dy_left, lo), hi);
  const float L = (-e.b * dyl - sqrtf(fmaxf(e.a * e.thr - e.det * dyl * dyl, 0.f))) * e.inv_a;
  const float R = (-e.b * dyr + sqrtf(fmaxf(e.a * e.thr - e.det * dyr * dyr, 0.f))) * e.inv_a;
  // tile j spans x in [j bw, j bw + bw - 1]
  const int lo_j = (int)ceilf((mx + L - (bw - 1.f)) * inv_bw);
  const int hi_j = (int)floorf((mx + R) * inv_bw);
  j0 = max(x0, lo_j);
  j1 = min(x1, hi_j + 1);
  if (j1 < j0) j1 = j0;
}

GSR_HD void cull_row_range(const CullEllipse &e, float mx, float my, int i, int x0, int x1,
                                               int block_width, int &j0, int &j1) {
  cull_row_range(e, mx, my, i, x0, x1, block_width, 1.f / (float)block_width, j0, j1);
}

typedef unsigned long long u64;

GSR_HD int gsr_clz(unsigned x) {
#ifdef __CUDA_ARCH__
  return __clz((int)x);
#else
  return x ? __builtin_clz(x) : 32;
#endif
}

// The tiles kept for a Gaussian (exact tile culling, above) inside its bounding box [x0, x1) x [y0, y1): calls
// f(tile_id) for each and returns the count.  For boxes of at most 64 tiles the decision is also returned as a mask
// (bit k = k-th tile of the box, row-major), which the fill pass walks instead of evaluating the ellipse again.
template <typename F>
GSR_HD int cull_tiles(float2 ctr, int r, float ca, float cb, float cc, float opac, int x0, int y0, int x1,
                                          int y1, int tiles_x, int block_width, u64 &mask, F f) {
  const int bw_tiles = x1 - x0, area = bw_tiles * (y1 - y0);
  mask = 0ull;
  const CullEllipse e = make_cull_ellipse(ca, cb, cc, opac, (float)(r + block_width));
  if (e.empty) return 0;
  int count = 0;
  u64 mk = 0ull;
  const float inv_bw = 1.f / (float)block_width;
  for (int i = y0; i < y1; ++i) {
    int j0, j1;
    cull_row_range(e, ctr.x, ctr.y, i, x0, x1, block_width, inv_bw, j0, j1);
    for (int j = j0; j < j1; ++j) f(i * tiles_x + j);
    const int cnt = j1 - j0;
    if (cnt > 0 && area <= 64) mk |= ((cnt >= 64) ? ~0ull : ((1ull << cnt) - 1ull)) << ((i - y0) * bw_tiles + (j0 - x0));
    count += cnt;
  }
  mask = mk;
  return count;
}

// Walk of a cached mask: tile of bit k = (y0 * tiles_x + x0) + k + (k / bw_tiles) * (tiles_x - bw_tiles).
// k / bw_tiles without an integer division per pair (25 of the 50 instructions of the loop it replaces):
// (k * magic) >> 16 is exact for k < 64 and every divisor up to 64 for any magic in [ceil(65536 / d), ceil(65536 / d) + 16]
// — the approximate reciprocal of fast-math cannot leave that interval
// (tests/test_abi.py::test_tile_mask_division_magic_is_exact).
template <typename F>
GSR_HD void walk_tile_mask(u64 mask, int x0, int y0, int bw_tiles, int tiles_x, F f) {
  const unsigned magic = (unsigned)(65536.f / (float)bw_tiles) + 1u;
  const int base = y0 * tiles_x + x0, skip = tiles_x - bw_tiles;
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    unsigned w = h ? (unsigned)(mask >> 32) : (unsigned)mask;
    const int k0 = 32 * h;
    while (w) {
      const unsigned low = w & (0u - w);
      const int k = k0 + 31 - gsr_clz(low);
      w ^= low;
      f(base + k + (int)(((unsigned)k * magic) >> 16) * skip);
    }
  }
}

}  // namespace gsr
