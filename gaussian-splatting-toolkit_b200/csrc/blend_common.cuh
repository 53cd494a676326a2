// blend_common.cuh — pieces shared by the 3-channel blend kernels (forward and adjoint).
#pragma once
#include <stdlib.h>

#include "common.cuh"

namespace gsr {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int BLEND_THREADS = 256;  // 16x16 tile, 8 warps; smaller block widths use fewer threads

// Packed per-Gaussian record staged in shared memory (3 x 16 bytes):
//   r0 = {x, y, ext_x, ext_y}        centre and conservative half-extents of the alpha >= 1/255 region
//   r1 = {A, B, C, opacity}          A = -0.5 a log2e, B = -b log2e, C = -0.5 c log2e  so that
//                                    exp(-sigma) = exp2(dx (A dx + B dy) + C dy dy)
//   r2 = {r, g, b, bits(gaussian id)}
struct BlendRecord {
  float4 r0, r1, r2;
};

// Conservative half-extents of the region where alpha = o * exp(-sigma) can reach 1/255:
// |dx| > ext_x = sqrt(2 ln(255 o) Sigma_xx) implies sigma >= dx^2 / (2 Sigma_xx) > ln(255 o)  (marginal bound of
// the Mahalanobis form), likewise in y.  NaN (never rejects) when the conic is not positive definite; negative
// (always rejects) when opacity < 1/255.  0.1 % + 0.01 px safety margin against rounding.
GSR_HD void alpha_extents(float a, float b, float c, float opac, float &ex, float &ey) {
  const float det = a * c - b * b;  // det(conic) = 1 / det(cov2d)
  const float tau2 = 2.f * gsr_logf(255.f * opac);
  if (!(tau2 >= 0.f)) {
    ex = (opac == opac) ? -1e30f : gsr_int_as_float(0x7fc00000);
    ey = ex;
    return;
  }
  ex = sqrtf(tau2 * c / det) * 1.001f + 0.01f;  // Sigma_xx = c / det
  ey = sqrtf(tau2 * a / det) * 1.001f + 0.01f;  // Sigma_yy = a / det
}

// The gathers go through L1 on purpose: the three scalar loads of a 12-byte conic / colour triple hit the sector the first
// one fetched.  Measured (gpurun_out/r2_run53_*): ld.global.nc.L1::no_allocate on these loads 0.369 -> 0.399 ms forward,
// 0.678 -> 0.707 ms adjoint; L1::evict_first 0.388 / 0.693 ms.
__device__ __forceinline__ BlendRecord gather_record(int g, const float2 *__restrict__ xys,
                                                     const float *__restrict__ conics,
                                                     const float *__restrict__ colors,
                                                     const float *__restrict__ opacities) {
  BlendRecord rec;
  const float2 xy = xys[g];
  const float opac = opacities[g];
  const float a = conics[3 * (size_t)g], b = conics[3 * (size_t)g + 1], c = conics[3 * (size_t)g + 2];
  float ex, ey;
  alpha_extents(a, b, c, opac, ex, ey);
  rec.r0 = make_float4(xy.x, xy.y, ex, ey);
  rec.r1 = make_float4(-0.5f * kLog2e * a, -kLog2e * b, -0.5f * kLog2e * c, opac);
  rec.r2 = make_float4(colors[3 * (size_t)g], colors[3 * (size_t)g + 1], colors[3 * (size_t)g + 2],
                       __int_as_float(g));
  return rec;
}

// thread -> pixel: a warp covers an 8x4 pixel sub-tile for 16x16 tiles, row-major order otherwise
__device__ __forceinline__ void map_pixel(int block_width, int &lx, int &ly) {
  const int tr = threadIdx.x;
  if (block_width == 16) {
    const int w = tr >> 5, l = tr & 31;
    lx = ((w & 1) << 3) + (l & 7);
    ly = ((w >> 1) << 2) + (l >> 3);
  } else {
    lx = tr % block_width;
    ly = tr / block_width;
  }
}

// Per-warp compaction: of the staged records [t_begin, t_end) keep those that can reach alpha >= 1/255 on one of
// the warp's pixels (rectangle [fx0,fx1]x[fy0,fy1], inside pixels only); their slot numbers are written
// (ascending) to `list`.  Returns the number kept (warp-uniform).  Lanes test 32 records per round, one ballot
// per round.  Two conservative tests per record:
//   1. the alpha-extent box (rec0.zw) against the rectangle;
//   2. for box survivors, per pixel ROW of the rectangle the maximum over x in [fx0,fx1] of the exponent
//      power(dx,dy) = A dx^2 + B dx dy + C dy^2 (concave in dx when A < 0: vertex dx* = -B dy / 2A clamped to the
//      row segment) against -log2(255 o), with a 0.1 % + 0.01 margin.  Exact in y, continuous in x: on the
//      BASELINE scene it removes another 16 % of the (warp, Gaussian) pairs the box lets through (the box
//      keeps the corners of the bounding box of a slanted ellipse), leaving 0.2 % false positives.
// Anything not provably below the threshold is kept (NaNs, A >= 0): comparisons are written so NaN keeps.
// PAD > 0: the PAD entries after the last survivor are set to `pad_slot` (a record that never contributes), so that a
// consumer may unroll / read ahead past the end of the list without a remainder loop.
// SHIFT: the list holds slot << SHIFT (SHIFT = 4: the byte offset of the slot's float4 in a record plane, which saves
// the consumer one address instruction per visit).
template <typename ListT = unsigned char, int PAD = 0, int SHIFT = 0>
__device__ __forceinline__ int compact_survivors(const float4 *__restrict__ rec0, const float4 *__restrict__ rec1,
                                                 int t_begin, int t_end, float fx0, float fx1, float fy0, float fy1,
                                                 ListT *__restrict__ list, int lane, int pad_slot = 0) {
  int n = 0;
  const unsigned lt_mask = (1u << lane) - 1u;
  for (int r = t_begin; r < t_end; r += 32) {
    const int t = r + lane;
    bool hit = false;
    if (t < t_end) {
      const float4 c = rec0[t];
      hit = !(c.x + c.z < fx0 || c.x - c.z > fx1 || c.y + c.w < fy0 || c.y - c.w > fy1);
      if (hit) {
        const float4 q = rec1[t];
        if (q.x < 0.f) {
          const float thr = -1.001f * __log2f(255.f * q.w) - 0.01f;
          const float k = -0.5f * q.y / q.x;
          const float lo = fx0 - c.x, hi = fx1 - c.x;
          bool keep = false;
          for (float fy = fy0; fy <= fy1; fy += 4.f) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float dy = fminf(fy + (float)i, fy1) - c.y;
              const float dx = fminf(fmaxf(k * dy, lo), hi);
              const float power = dx * (q.x * dx + q.y * dy) + q.z * dy * dy;
              keep = keep || !(power < thr);
            }
          }
          hit = keep;
        }
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (hit) list[n + __popc(m & lt_mask)] = (ListT)(t << SHIFT);
    n += __popc(m);
  }
  if (PAD > 0 && lane < PAD) list[n + lane] = (ListT)(pad_slot << SHIFT);
  __syncwarp();
  return n;
}

// ---- CTA-cooperative form of the same two tests for 16x16 tiles -----------------------------------------------------
// With compact_survivors every one of the 8 warps tests every staged record (8 x 107 warp-instructions per 32 records:
// 19 % of the forward's and 11 % of the adjoint's instructions, profiles/r02/ncu_blend_v9.txt).  Here the thread that
// STAGES a record evaluates it once, from its registers, against all eight 8x4 pixel blocks of the tile (warp w owns
// block column w & 1, block row w >> 1) and stores the 8-bit result next to the record; a warp then only gathers its
// bit (compact_from_masks).  Block rectangles are taken unclipped by the image (a superset of the clipped ones), the
// tests and their margins are those of compact_survivors: box test, then per pixel row the maximum of the exponent
// over the block's x segment against -log2(255 o); anything not provably below the threshold is kept (NaNs, A >= 0).
// Written with margins (>= 0 keeps) combined by min and read off the sign bit: the predicate form of the same logic
// compiles to ~130 instructions of predicate spilling (P2R / LOP3) on top of the ~300 of arithmetic.
GSR_HD unsigned block_mask_16(const float4 c, const float4 q, float tile_x0, float tile_y0) {
  // NaN anywhere (non-positive-definite conic, NaN opacity) or inf - inf: keep every block (fmin / fmax drop NaNs)
  const float probe = (c.x + c.y) + (c.z + c.w) + (q.x + q.y) + (q.z + q.w);
  const bool weird = !(probe == probe);
  const bool any_shape = !(q.x < 0.f);  // not concave in dx: only the box test applies
  const float thr = -1.001f * gsr_log2f(255.f * q.w) - 0.01f;
  const float k = -0.5f * q.y / q.x;
  const float lo0 = tile_x0 - c.x, hi0 = lo0 + 7.f, lo1 = lo0 + 8.f, hi1 = lo0 + 15.f;
  const float dy0 = tile_y0 - c.y;
  // box test:  !(x + ext < x_min || x - ext > x_max)  <=>  ext - max(x_min - x, x - x_max) >= 0
  const float box_l = c.z - fmaxf(lo0, -hi0), box_r = c.z - fmaxf(lo1, -hi1);
  unsigned reject = 0u;
#pragma unroll
  for (int g = 3; g >= 0; --g) {
    const float ylo = dy0 + (float)(4 * g);
    const float box_y = c.w - fmaxf(ylo, -(ylo + 3.f));
    float pm0 = -3.0e38f, pm1 = -3.0e38f;  // max of the exponent over the block's rows and its x segment
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float dy = dy0 + (float)(4 * g + i);
      const float v = k * dy, bdy = q.y * dy, cdy = q.z * dy * dy;
      const float dx0 = fminf(fmaxf(v, lo0), hi0), dx1 = fminf(fmaxf(v, lo1), hi1);
      pm0 = fmaxf(pm0, dx0 * (q.x * dx0 + bdy) + cdy);
      pm1 = fmaxf(pm1, dx1 * (q.x * dx1 + bdy) + cdy);
    }
    const float row0 = any_shape ? 1.f : pm0 - thr, row1 = any_shape ? 1.f : pm1 - thr;  // !(pm < thr)
    const float m1 = fminf(fminf(box_y, box_r), row1), m0 = fminf(fminf(box_y, box_l), row0);
    reject = (reject << 1) | (gsr_float_as_uint(m1) >> 31);  // bit 2 g + 1
    reject = (reject << 1) | (gsr_float_as_uint(m0) >> 31);  // bit 2 g
  }
  return weird ? 0xffu : (~reject & 0xffu);
}

// Per-warp list of the staged records [t_begin, t_end) whose block mask has this warp's bit (ascending slot numbers,
// padded like compact_survivors).  Returns the number kept (warp-uniform).  `masks` holds BLEND_THREADS bytes, 8-byte
// aligned; bytes at or beyond t_end may be stale (they are masked off).  Lane l owns slots 8 l .. 8 l + 7: one 8-byte
// load, the warp's bit of the eight bytes gathered with a multiply, an exclusive prefix over the lanes and a short
// store loop — ~75 instructions per (warp, batch) instead of 8 ballot rounds.
template <typename ListT = unsigned char, int PAD = 0, int SHIFT = 0>
__device__ __forceinline__ int compact_from_masks(const unsigned char *__restrict__ masks, int warp, int t_begin, int t_end,
                                                  ListT *__restrict__ list, int lane, int pad_slot = 0) {
  const uint2 m = reinterpret_cast<const uint2 *>(masks)[lane];
  const unsigned lo = (m.x >> warp) & 0x01010101u, hi = (m.y >> warp) & 0x01010101u;
  // bits 0, 8, 16, 24 -> bits 24..27 of the product (all partial products land on distinct bits: no carries)
  unsigned f = ((lo * 0x01020408u) >> 24) | (((hi * 0x01020408u) >> 24) << 4);
  const int s0 = 8 * lane;
  const int nb = min(max(t_begin - s0, 0), 8), ne = min(max(t_end - s0, 0), 8);
  f &= ((1u << ne) - 1u) & ~((1u << nb) - 1u);
  const int cnt = __popc(f);
  int incl = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int up = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += up;
  }
  const int n = __shfl_sync(0xffffffffu, incl, 31);
  ListT *dst = list + (incl - cnt);
  while (f) {
    const unsigned low = f & (0u - f);
    *dst++ = (ListT)((s0 + 31 - __clz((int)low)) << SHIFT);
    f ^= low;
  }
  if (PAD > 0 && lane < PAD) list[n + lane] = (ListT)(pad_slot << SHIFT);
  __syncwarp();
  return n;
}

// GSR_BLOCK_MASK = 0 keeps the per-warp tests (compact_survivors) in the 16x16 kernels; read once
inline bool blend_block_masks() {
  static const bool v = [] {
    const char *e = getenv("GSR_BLOCK_MASK");
    return !(e && e[0] == '0');
  }();
  return v;
}

}  // namespace gsr
