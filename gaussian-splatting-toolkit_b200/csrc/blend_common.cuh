// blend_common.cuh — pieces shared by the 3-channel blend kernels (forward and adjoint).
#pragma once
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
__device__ __forceinline__ void alpha_extents(float a, float b, float c, float opac, float &ex, float &ey) {
  const float det = a * c - b * b;  // det(conic) = 1 / det(cov2d)
  const float tau2 = 2.f * __logf(255.f * opac);
  if (!(tau2 >= 0.f)) {
    ex = (opac == opac) ? -1e30f : __int_as_float(0x7fc00000);
    ey = ex;
    return;
  }
  ex = sqrtf(tau2 * c / det) * 1.001f + 0.01f;  // Sigma_xx = c / det
  ey = sqrtf(tau2 * a / det) * 1.001f + 0.01f;  // Sigma_yy = a / det
}

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
template <typename ListT = unsigned char, int PAD = 0>
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
    if (hit) list[n + __popc(m & lt_mask)] = (ListT)t;
    n += __popc(m);
  }
  if (PAD > 0 && lane < PAD) list[n + lane] = (ListT)pad_slot;
  __syncwarp();
  return n;
}

}  // namespace gsr
