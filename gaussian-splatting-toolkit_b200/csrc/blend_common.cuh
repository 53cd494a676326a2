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

// Per-warp compaction: of the staged records [t_begin, t_end) keep those whose alpha-extent box overlaps the
// warp's pixel rectangle [fx0,fx1]x[fy0,fy1]; their slot numbers are written (ascending) to `list`.
// Returns the number kept (warp-uniform).  Lanes test 32 records per round; one ballot per round.
__device__ __forceinline__ int compact_survivors(const float4 *__restrict__ rec0, int t_begin, int t_end,
                                                 float fx0, float fx1, float fy0, float fy1,
                                                 unsigned char *__restrict__ list, int lane) {
  int n = 0;
  const unsigned lt_mask = (1u << lane) - 1u;
  for (int r = t_begin; r < t_end; r += 32) {
    const int t = r + lane;
    bool hit = false;
    if (t < t_end) {
      const float4 c = rec0[t];
      hit = !(c.x + c.z < fx0 || c.x - c.z > fx1 || c.y + c.w < fy0 || c.y - c.w > fy1);
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (hit) list[n + __popc(m & lt_mask)] = (unsigned char)t;
    n += __popc(m);
  }
  __syncwarp();
  return n;
}

}  // namespace gsr
