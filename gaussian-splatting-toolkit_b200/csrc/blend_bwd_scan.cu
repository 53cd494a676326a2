// blend_bwd_scan.cu — GAUSSIAN-PARALLEL adjoint of the per-tile alpha compositing (3 channels, FP32, 16x16 tiles).
//
// Same function as blend_backward_kernel (blend_bwd.cu; reference rasterize_backward_kernel, csrc/backward.cu:133-303)
// with the two parallel axes swapped inside a warp:
//
//   blend_backward_kernel      lane = PIXEL of the warp's 8x4 block, loop over the surviving Gaussians.  The per-pixel
//                              state (T, running colour sum) is private; the nine per-Gaussian gradient sums need a
//                              32-lane reduction PER VISIT (14 SHFL + 28 ALU + a RED, ~45 % of the loop).
//   blend_backward_scan_kernel lane = GAUSSIAN (32 consecutive survivors of the warp's compacted list), loop over the 32
//                              pixels of the block.  The nine gradient sums are private registers (no reduction, one set
//                              of REDs per 32 visits); what crosses lanes is the per-pixel state, and that is a PREFIX
//                              SCAN over the lanes: the transmittance behind Gaussian j is T_in * prod_{i<=j} 1/(1-alpha_i)
//                              (back to front), the colour sum behind it a prefix sum; both come out of ONE warp-shuffle
//                              scan of affine maps (10 SHFL + 15 ALU per pixel step, five dependent steps), with the
//                              per-pixel carry (T, s) kept in shared memory between groups of 32 survivors.
//
// The per-pixel colour bookkeeping of the reference (three running sums S_c and three v_out terms,
// backward.cu:243-262) collapses to ONE scalar because the upstream gradient is constant per pixel:
//   s = sum_c S_c v_out_c - T_final (v_out_alpha - sum_c bg_c v_out_c),   d_j = sum_c rgb_{j,c} v_out_c
//   v_alpha_j = T_j d_j - ra_j s ,   s <- s + alpha_j T_j d_j             (T_j = transmittance in front of j)
// Staging (double-buffered packed records), exact warp compaction and the early cut at max(final_idx) are those of
// blend_bwd.cu; survivors left over at the end of a staged batch stay in registers and are completed from the next
// batch, so only the last group of a tile can be partially filled.
#include <stdlib.h>

#include "blend_common.cuh"

namespace gsr {

namespace {

struct LaneGaussian {  // the Gaussian a lane owns for one group
  float x, y, A, B, C, o, r, g, b;
  int id;    // Gaussian index, -1 = empty lane
  int sidx;  // its position in the tile-sorted list (compared with the pixel's final_idx)
};

__device__ __forceinline__ void load_lane_gaussian(LaneGaussian &G, const float4 (*rec)[BLEND_THREADS], int slot,
                                                   int batch_end) {
  const float4 q0 = rec[0][slot], q1 = rec[1][slot], q2 = rec[2][slot];
  G.x = q0.x; G.y = q0.y;
  G.A = q1.x; G.B = q1.y; G.C = q1.z; G.o = q1.w;
  G.r = q2.x; G.g = q2.y; G.b = q2.z;
  G.id = __float_as_int(q2.w);
  G.sidx = batch_end - slot;
}

__device__ __forceinline__ void clear_lane_gaussian(LaneGaussian &G) {
  G.x = G.y = G.A = G.B = G.C = G.o = G.r = G.g = G.b = 0.f;
  G.id = -1;
  G.sidx = 0x7fffffff;
}

// One group: 32 lanes x 32 pixels.  pixc[p] = {v_out_r, v_out_g, v_out_b, bits(final_idx)}, pixs[p] = {px, py, T, s}.
__device__ __forceinline__ void process_group(const LaneGaussian &G, const float4 *__restrict__ pixc,
                                              float4 *__restrict__ pixs, int lane, float *__restrict__ v_xy,
                                              float *__restrict__ v_conic, float *__restrict__ v_colors,
                                              float *__restrict__ v_opacity) {
  const unsigned full = 0xffffffffu;
  float a_r = 0.f, a_g = 0.f, a_b = 0.f, a_xx = 0.f, a_xy = 0.f, a_yy = 0.f, a_x = 0.f, a_y = 0.f, a_w = 0.f;
  // Two pixels per iteration, written out so that their two shuffle chains are independent and interleave (one chain
  // alone is latency-bound: 10 dependent SHFLs; ncu of the first version: issue slots 55 % busy, short-scoreboard
  // stalls 4.1 per issue).  All shared-memory loads of the pair come first, the carry stores last.
  constexpr int U = 2;
  for (int p = 0; p < 32; p += U) {
    float4 c[U], st[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      c[u] = pixc[p + u];
      st[u] = pixs[p + u];
    }
    // Per Gaussian the pixel state (T, s) moves by an AFFINE map:  T <- ra T,  s <- s + (alpha ra d) T  (T on the right is
    // the transmittance BEFORE the Gaussian).  Maps compose associatively, (R_a, C_a) then (R_b, C_b) = (R_a R_b,
    // C_a + C_b R_a), so ONE inclusive scan of the pair (R, C) over the lanes gives both the transmittance and the colour
    // sum in front of every Gaussian — five dependent shuffle steps instead of two scans of five.
    float dx[U], dy[U], vis_e[U], alpha_e[U], ra[U], R[U], Cc[U], dj[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      dx[u] = G.x - st[u].x;
      dy[u] = G.y - st[u].y;
      const float gx = G.A * dx[u], gy = G.C * dy[u];
      const float power = dx[u] * (gx + G.B * dy[u]) + gy * dy[u];  // = -sigma log2(e)
      const float vis = exp2f(power);
      const float alpha = fminf(0.99f, G.o * vis);
      const bool valid = (G.sidx <= __float_as_int(c[u].w)) && !(power > 0.f || alpha < 1.f / 255.f);
      alpha_e[u] = valid ? alpha : 0.f;
      vis_e[u] = valid ? vis : 0.f;
      ra[u] = 1.f / (1.f - alpha_e[u]);
      dj[u] = G.r * c[u].x + G.g * c[u].y + G.b * c[u].z;
      R[u] = ra[u];
      Cc[u] = alpha_e[u] * ra[u] * dj[u];
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      float rup[U], cup[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        rup[u] = __shfl_up_sync(full, R[u], d);
        cup[u] = __shfl_up_sync(full, Cc[u], d);
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (lane >= d) {
          Cc[u] = cup[u] + Cc[u] * rup[u];
          R[u] *= rup[u];
        }
    }
    float T[U], s_after[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      T[u] = st[u].z * R[u];                    // transmittance in front of this Gaussian
      s_after[u] = st[u].w + Cc[u] * st[u].z;   // colour sum including this Gaussian
      const float fac = alpha_e[u] * T[u];
      const float s_behind = s_after[u] - fac * dj[u];
      const float v_alpha = T[u] * dj[u] - ra[u] * s_behind;
      const float w = vis_e[u] * v_alpha;
      a_r += fac * c[u].x;
      a_g += fac * c[u].y;
      a_b += fac * c[u].z;
      const float wdx = w * dx[u], wdy = w * dy[u];
      a_xx += wdx * dx[u];
      a_xy += wdx * dy[u];
      a_yy += wdy * dy[u];
      a_x += wdx;
      a_y += wdy;
      a_w += w;
    }
    if (lane == 31) {  // carry to the next group
#pragma unroll
      for (int u = 0; u < U; ++u) *reinterpret_cast<float2 *>(&pixs[p + u].z) = make_float2(T[u], s_after[u]);
    }
  }
  __syncwarp();
  if (G.id >= 0) {
    // v_sigma = -o vis v_alpha = -o w;  conic = -(2A, B, 2C) ln2
    const unsigned g = (unsigned)G.id;
    const float no = -G.o;
    const float ca = -2.f * kLn2 * G.A, cb = -kLn2 * G.B, cc = -2.f * kLn2 * G.C;
    atomicAdd(v_colors + 3u * g, a_r);
    atomicAdd(v_colors + 3u * g + 1u, a_g);
    atomicAdd(v_colors + 3u * g + 2u, a_b);
    atomicAdd(v_conic + 3u * g, 0.5f * no * a_xx);
    atomicAdd(v_conic + 3u * g + 1u, no * a_xy);
    atomicAdd(v_conic + 3u * g + 2u, 0.5f * no * a_yy);
    atomicAdd(v_xy + 2u * g, no * (ca * a_x + cb * a_y));
    atomicAdd(v_xy + 2u * g + 1u, no * (cb * a_x + cc * a_y));
    atomicAdd(v_opacity + g, a_w);
  }
}

}  // namespace

__global__ void __launch_bounds__(BLEND_THREADS, 3)
blend_backward_scan_kernel(int tiles_x, int img_w, int img_h, const int *__restrict__ gaussian_ids_sorted,
                           const int2 *__restrict__ tile_bins, const float2 *__restrict__ xys,
                           const float *__restrict__ conics, const float *__restrict__ colors,
                           const float *__restrict__ opacities, const float *__restrict__ background,
                           const float *__restrict__ final_Ts, const int *__restrict__ final_idx,
                           const float *__restrict__ v_output, const float *__restrict__ v_output_alpha,
                           float *__restrict__ v_xy, float *__restrict__ v_conic, float *__restrict__ v_colors,
                           float *__restrict__ v_opacity) {
  __shared__ float4 s_rec[2][3][BLEND_THREADS];
  __shared__ unsigned char s_list[BLEND_THREADS / 32][BLEND_THREADS];
  __shared__ float4 s_pixc[BLEND_THREADS / 32][32];
  __shared__ float4 s_pixs[BLEND_THREADS / 32][32];
  __shared__ int s_warp_max[BLEND_THREADS / 32];

  const unsigned full = 0xffffffffu;
  const int tile_x = blockIdx.x, tile_y = blockIdx.y;
  const int tile_id = tile_y * tiles_x + tile_x;
  const int tr = threadIdx.x, nthreads = BLEND_THREADS, lane = tr & 31, warp = tr >> 5;
  int lx, ly;
  map_pixel(16, lx, ly);
  const int ipx = tile_x * 16 + lx, ipy = tile_y * 16 + ly;
  const bool inside = (ipx < img_w) && (ipy < img_h);
  const int pix = inside ? (ipy * img_w + ipx) : 0;

  const float fx0 = (float)__reduce_min_sync(full, inside ? ipx : 0x7fffffff);
  const float fx1 = (float)__reduce_max_sync(full, inside ? ipx : -0x7fffffff);
  const float fy0 = (float)__reduce_min_sync(full, inside ? ipy : 0x7fffffff);
  const float fy1 = (float)__reduce_max_sync(full, inside ? ipy : -0x7fffffff);

  const int2 range = tile_bins[tile_id];
  const int bin_final = inside ? final_idx[pix] : -1;
  {
    const float T_final = inside ? final_Ts[pix] : 1.f;
    float vo_r = 0.f, vo_g = 0.f, vo_b = 0.f, vo_a = 0.f;
    if (inside) {
      vo_r = v_output[3 * (size_t)pix];
      vo_g = v_output[3 * (size_t)pix + 1];
      vo_b = v_output[3 * (size_t)pix + 2];
      vo_a = v_output_alpha[pix];
    }
    // T_final ra v_out_alpha - T_final ra sum_c bg_c v_out_c = ra c_final (backward.cu:252-256); s starts at -c_final
    const float c_final = T_final * (vo_a - (background[0] * vo_r + background[1] * vo_g + background[2] * vo_b));
    s_pixc[warp][lane] = make_float4(vo_r, vo_g, vo_b, __int_as_float(bin_final));
    s_pixs[warp][lane] = make_float4((float)ipx, (float)ipy, T_final, -c_final);
  }
  const int warp_bin_final = __reduce_max_sync(full, bin_final);
  if (lane == 0) s_warp_max[warp] = warp_bin_final;
  __syncthreads();
  int cta_bin_final = -1;
  for (int w = 0; w < (nthreads >> 5); ++w) cta_bin_final = max(cta_bin_final, s_warp_max[w]);

  const int end = min(range.y, cta_bin_final + 1);
  const int count = end - range.x;
  if (count <= 0) return;  // uniform across the CTA
  const int num_batches = (count + nthreads - 1) / nthreads;

  BlendRecord rec;
  if (end - 1 - tr >= range.x)
    rec = gather_record(gaussian_ids_sorted[end - 1 - tr], xys, conics, colors, opacities);

  LaneGaussian G;
  clear_lane_gaussian(G);
  int n_pending = 0;  // lanes [0, n_pending) hold survivors of earlier batches that wait for a full group

  for (int b = 0; b < num_batches; ++b) {
    const int buf = b & 1;
    const int batch_end = end - 1 - nthreads * b;  // sorted index held by slot 0; slot t holds batch_end - t
    if (batch_end - tr >= range.x) {
      s_rec[buf][0][tr] = rec.r0;
      s_rec[buf][1][tr] = rec.r1;
      s_rec[buf][2][tr] = rec.r2;
    }
    __syncthreads();
    {
      const int nxt = batch_end - nthreads - tr;
      if (nxt >= range.x) rec = gather_record(gaussian_ids_sorted[nxt], xys, conics, colors, opacities);
    }
    const int batch_size = min(nthreads, batch_end + 1 - range.x);
    const int t_begin = max(0, batch_end - warp_bin_final);  // slots before it are behind every pixel's last contributor
    if (t_begin >= batch_size) continue;
    const int n_list = compact_survivors(s_rec[buf][0], s_rec[buf][1], t_begin, batch_size, fx0, fx1, fy0, fy1,
                                         s_list[warp], lane);
    int li = 0;  // next unread entry of the list
    while (n_pending + (n_list - li) >= 32) {
      if (lane >= n_pending) load_lane_gaussian(G, s_rec[buf], s_list[warp][li + lane - n_pending], batch_end);
      li += 32 - n_pending;
      n_pending = 0;
      process_group(G, s_pixc[warp], s_pixs[warp], lane, v_xy, v_conic, v_colors, v_opacity);
    }
    const int rest = n_list - li;  // < 32 - n_pending
    if (lane >= n_pending && lane < n_pending + rest)
      load_lane_gaussian(G, s_rec[buf], s_list[warp][li + lane - n_pending], batch_end);
    n_pending += rest;
  }
  if (n_pending > 0) {
    if (lane >= n_pending) clear_lane_gaussian(G);
    process_group(G, s_pixc[warp], s_pixs[warp], lane, v_xy, v_conic, v_colors, v_opacity);
  }
}

// Selected with GSR_BWD_KERNEL=scan (blend_bwd_tr.cu holds the switch and the default kernel).
// Measured at cfg2 on B200 (profiles/r02): pixel-parallel 1.047 ms, this kernel 0.953 ms.  Variants tried and dropped: two
// separate scans per pixel instead of one affine-map scan (1.22 ms: latency-bound, issue slots 55 % busy); registers
// capped at 64 for 4 CTAs / SM (0.996 ms); the two pixels of an iteration packed into Blackwell's two-wide FP32
// instructions (fma.rn.f32x2 / FFMA2: 15 % fewer instructions, but 114 registers -> 2 CTAs / SM and 23 register moves per
// pixel pair: 1.106 ms).
int launch_blend_backward_scan(dim3 grid, cudaStream_t st, int img_w, int img_h, const int *gaussian_ids_sorted,
                               const int2 *tile_bins, const float2 *xys, const float *conics, const float *colors,
                               const float *opacities, const float *background, const float *final_Ts,
                               const int *final_idx, const float *v_output, const float *v_output_alpha, float *v_xy,
                               float *v_conic, float *v_colors, float *v_opacity) {
  blend_backward_scan_kernel<<<grid, BLEND_THREADS, 0, st>>>((int)grid.x, img_w, img_h, gaussian_ids_sorted, tile_bins, xys,
                                                             conics, colors, opacities, background, final_Ts, final_idx,
                                                             v_output, v_output_alpha, v_xy, v_conic, v_colors, v_opacity);
  GSR_CHECK_LAUNCH("blend_backward_scan_kernel");
  return GSR_OK;
}

}  // namespace gsr
