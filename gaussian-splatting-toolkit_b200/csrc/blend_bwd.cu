// blend_bwd.cu — adjoint of the per-tile alpha compositing (3 channels, FP32) for sm_100a.
//
// Replaces rasterize_backward_kernel (reference csrc/backward.cu:133-303, launched from
// bindings.cu:471-528).  Per pixel it replays the contributors back-to-front from (T_final, final_idx)
// exactly as the reference does (alpha clamp 0.99, clamp derivative ignored, T *= 1/(1-alpha)).
//
// What is different from the reference kernel:
//   * only the batches at or before max(final_idx) of the CTA are staged at all (the reference stages every
//     batch of the tile and skips inside);
//   * double-buffered shared-memory ring of packed records with register prefetch, one barrier per batch;
//   * warp = 8x4 pixel sub-tile; each warp compacts a batch to the Gaussians whose alpha >= 1/255 extent box
//     overlaps its pixel rectangle (ballot compaction, blend_common.cuh) and only walks the survivors;
//   * the 9 per-Gaussian partial sums are reduced across the warp with a transposing butterfly
//     (8 values in 4+2+1+1+1 = 9 shuffles, + 5 for the ninth) instead of 9 x 5 = 45 shuffles
//     (cg::reduce per value, backward.cu:275-278), and the nine totals end up in nine DIFFERENT lanes, so the
//     global accumulation is ONE predicated RED instruction per (warp, Gaussian) instead of nine serial
//     atomicAdds issued by lane 0 (backward.cu:279-300);
//   * outputs are zero-filled by this call (cudaMemsetAsync) rather than by torch::zeros in the caller.
#include "blend_common.cuh"

namespace gsr {

// Sum 8 per-lane values across the warp.  On return, lane L with (L & 3) == 0 holds in v[0] the total of
// value number ((L>>4)&1)*4 + ((L>>3)&1)*2 + ((L>>2)&1).
__device__ __forceinline__ float warp_transpose_reduce8(float v[8], int lane) {
  const unsigned full = 0xffffffffu;
  {
    const bool hi = lane & 16;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = hi ? v[i] : v[i + 4];
      const float keep = hi ? v[i + 4] : v[i];
      v[i] = keep + __shfl_xor_sync(full, send, 16);
    }
  }
  {
    const bool hi = lane & 8;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = hi ? v[i] : v[i + 2];
      const float keep = hi ? v[i + 2] : v[i];
      v[i] = keep + __shfl_xor_sync(full, send, 8);
    }
  }
  {
    const bool hi = lane & 4;
    const float send = hi ? v[0] : v[1];
    const float keep = hi ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(full, send, 4);
  }
  v[0] += __shfl_xor_sync(full, v[0], 2);
  v[0] += __shfl_xor_sync(full, v[0], 1);
  return v[0];
}

__device__ __forceinline__ float warp_sum(float x) {
  const unsigned full = 0xffffffffu;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(full, x, o);
  return x;
}

__global__ void __launch_bounds__(BLEND_THREADS)
blend_backward_kernel(int tiles_x, int img_w, int img_h, int block_width,
                      const int *__restrict__ gaussian_ids_sorted, const int2 *__restrict__ tile_bins,
                      const float2 *__restrict__ xys, const float *__restrict__ conics,
                      const float *__restrict__ colors, const float *__restrict__ opacities,
                      const float *__restrict__ background, const float *__restrict__ final_Ts,
                      const int *__restrict__ final_idx, const float *__restrict__ v_output,
                      const float *__restrict__ v_output_alpha, float *__restrict__ v_xy,
                      float *__restrict__ v_conic, float *__restrict__ v_colors,
                      float *__restrict__ v_opacity) {
  __shared__ float4 s_rec[2][3][BLEND_THREADS];
  __shared__ unsigned char s_list[BLEND_THREADS / 32][BLEND_THREADS];
  __shared__ int s_warp_max[BLEND_THREADS / 32];

  const unsigned full = 0xffffffffu;
  const int tile_x = blockIdx.x, tile_y = blockIdx.y;
  const int tile_id = tile_y * tiles_x + tile_x;
  const int tr = threadIdx.x, nthreads = blockDim.x, lane = tr & 31, warp = tr >> 5;
  int lx, ly;
  map_pixel(block_width, lx, ly);
  const int ipx = tile_x * block_width + lx, ipy = tile_y * block_width + ly;
  const bool inside = (ly < block_width) && (ipx < img_w) && (ipy < img_h);
  const float px = (float)ipx, py = (float)ipy;
  const int pix = inside ? (ipy * img_w + ipx) : 0;

  const float fx0 = (float)__reduce_min_sync(full, inside ? ipx : 0x7fffffff);
  const float fx1 = (float)__reduce_max_sync(full, inside ? ipx : -0x7fffffff);
  const float fy0 = (float)__reduce_min_sync(full, inside ? ipy : 0x7fffffff);
  const float fy1 = (float)__reduce_max_sync(full, inside ? ipy : -0x7fffffff);

  const int2 range = tile_bins[tile_id];

  const float T_final = inside ? final_Ts[pix] : 1.f;
  float T = T_final;
  // reference: bin_final = inside ? final_index : 0 (backward.cu:168); -1 for outside threads is
  // equivalent because they are never valid
  const int bin_final = inside ? final_idx[pix] : -1;
  float vo_r = 0.f, vo_g = 0.f, vo_b = 0.f, vo_a = 0.f;
  if (inside) {
    vo_r = v_output[3 * (size_t)pix];
    vo_g = v_output[3 * (size_t)pix + 1];
    vo_b = v_output[3 * (size_t)pix + 2];
    vo_a = v_output_alpha[pix];
  }
  // T_final * ra * v_out_alpha - T_final * ra * sum_c bg_c v_out_c  =  ra * c_final   (backward.cu:252-256)
  const float c_final = T_final * (vo_a - (background[0] * vo_r + background[1] * vo_g + background[2] * vo_b));
  // The upstream gradient is constant per pixel, so the reference's three running colour sums S_c (backward.cu:243-262)
  // are only ever used through  s = sum_c S_c v_out_c - c_final :  v_alpha = T d - ra s,  s += alpha T d,
  // with d = sum_c rgb_c v_out_c  (one scalar instead of three, 6 instructions instead of 13 per visit).
  float s_run = -c_final;

  const int warp_bin_final = __reduce_max_sync(full, bin_final);
  if (lane == 0) s_warp_max[warp] = warp_bin_final;
  __syncthreads();
  int cta_bin_final = -1;
  for (int w = 0; w < (nthreads >> 5); ++w) cta_bin_final = max(cta_bin_final, s_warp_max[w]);

  // only sorted indices [range.x, end) can be contributors of any pixel of this tile; walk them back to front
  const int end = min(range.y, cta_bin_final + 1);
  const int count = end - range.x;
  if (count <= 0) return;  // uniform across the CTA
  const int num_batches = (count + nthreads - 1) / nthreads;

  // per-lane destination of the reduced totals: lanes 0,4,..,28 own values 0..7, lane 1 owns value 8
  //   value 0..2 -> v_colors, 3..5 -> v_conic, 6..7 -> v_xy, 8 -> v_opacity
  float *dst_base = nullptr;
  int dst_stride = 0;
  {
    const int vi = (lane & 3) == 0 ? (((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1))
                                   : (lane == 1 ? 8 : -1);
    if (vi >= 0 && vi < 3) { dst_base = v_colors + vi; dst_stride = 3; }
    else if (vi >= 3 && vi < 6) { dst_base = v_conic + (vi - 3); dst_stride = 3; }
    else if (vi >= 6 && vi < 8) { dst_base = v_xy + (vi - 6); dst_stride = 2; }
    else if (vi == 8) { dst_base = v_opacity; dst_stride = 1; }
  }

  BlendRecord rec;
  if (end - 1 - tr >= range.x)
    rec = gather_record(gaussian_ids_sorted[end - 1 - tr], xys, conics, colors, opacities);

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
    const int t_begin = max(0, batch_end - warp_bin_final);  // slots before it are behind every lane's last contributor
    if (t_begin >= batch_size) continue;
    const int n_list = compact_survivors(s_rec[buf][0], s_rec[buf][1], t_begin, batch_size, fx0, fx1, fy0, fy1, s_list[warp], lane);
    for (int i = 0; i < n_list; ++i) {
      const int t = s_list[warp][i];
      const float4 q0 = s_rec[buf][0][t];
      const float4 q1 = s_rec[buf][1][t];
      const float dx = q0.x - px, dy = q0.y - py;
      const float gx = q1.x * dx, gy = q1.z * dy;      // A dx, C dy
      const float power = dx * (gx + q1.y * dy) + gy * dy;  // = -sigma log2(e)
      const float vis = exp2f(power);
      const float opac = q1.w;
      const float alpha = fminf(0.99f, opac * vis);
      const bool valid = inside && (batch_end - t <= bin_final) && !(power > 0.f || alpha < 1.f / 255.f);
      if (!__any_sync(full, valid)) continue;

      // Branch-free body: lanes that are not valid run the same arithmetic with alpha = vis = 0, which leaves
      // T and the running sums untouched (ra = 1, fac = 0) and makes all nine partial sums exactly zero.
      const float alpha_e = valid ? alpha : 0.f;
      const float vis_e = valid ? vis : 0.f;
      const float4 q2 = s_rec[buf][2][t];
      float v[8];
      const float ra = 1.f / (1.f - alpha_e);
      T *= ra;
      const float fac = alpha_e * T;
      v[0] = fac * vo_r;
      v[1] = fac * vo_g;
      v[2] = fac * vo_b;
      const float dcol = q2.x * vo_r + q2.y * vo_g + q2.z * vo_b;
      const float v_alpha = T * dcol - ra * s_run;
      s_run += fac * dcol;
      const float v_sigma = -opac * vis_e * v_alpha;
      // conic = -(2A, B, 2C) ln2 : v_conic = (0.5 v_sigma dx^2, v_sigma dx dy, 0.5 v_sigma dy^2)
      const float hs = 0.5f * v_sigma;
      v[3] = hs * dx * dx;
      v[4] = v_sigma * dx * dy;
      v[5] = hs * dy * dy;
      // v_xy = v_sigma * (a dx + b dy, b dx + c dy) with a = -2A ln2, b = -B ln2, c = -2C ln2
      const float ws = -kLn2 * v_sigma;
      v[6] = ws * (2.f * gx + q1.y * dy);
      v[7] = ws * (q1.y * dx + 2.f * gy);
      const float v_opac_l = vis_e * v_alpha;
      const float tot8 = warp_transpose_reduce8(v, lane);
      const float tot_op = warp_sum(v_opac_l);
      if (dst_base != nullptr) {
        const unsigned g = (unsigned)__float_as_int(q2.w);
        atomicAdd(dst_base + g * (unsigned)dst_stride, lane == 1 ? tot_op : tot8);  // 32-bit element offset
      }
    }
  }
}

// 16x16 tiles: blend_bwd_tr.cu (two-phase transposing adjoint, the default) or blend_bwd_scan.cu (warp prefix-scan adjoint,
// GSR_BWD_KERNEL=scan); this file's pixel-parallel kernel serves the other block widths and GSR_BWD_KERNEL=pixel
int blend_bwd_mode();  // 0 = pixel, 1 = scan, 8 / 16 / 32 = two-phase with that many rows per group
int launch_blend_backward_tr(int mode, dim3 grid, cudaStream_t st, int img_w, int img_h, const int *gaussian_ids_sorted,
                             const int2 *tile_bins, const float2 *xys, const float *conics, const float *colors,
                             const float *opacities, const float *background, const float *final_Ts,
                             const int *final_idx, const float *v_output, const float *v_output_alpha, float *v_xy,
                             float *v_conic, float *v_colors, float *v_opacity);
int launch_blend_backward_scan(dim3 grid, cudaStream_t st, int img_w, int img_h, const int *gaussian_ids_sorted,
                               const int2 *tile_bins, const float2 *xys, const float *conics, const float *colors,
                               const float *opacities, const float *background, const float *final_Ts,
                               const int *final_idx, const float *v_output, const float *v_output_alpha, float *v_xy,
                               float *v_conic, float *v_colors, float *v_opacity);

}  // namespace gsr

extern "C" GSR_API int gsr_rasterize_backward(unsigned img_height, unsigned img_width, unsigned block_width,
                                              int num_points, const int32_t *gaussian_ids_sorted,
                                              const int32_t *tile_bins, const float *xys, const float *conics,
                                              const float *colors, const float *opacities,
                                              const float *background, const float *final_Ts,
                                              const int32_t *final_idx, const float *v_output,
                                              const float *v_output_alpha, float *v_xy, float *v_conic,
                                              float *v_colors, float *v_opacity, void *stream) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_rasterize_backward");
  GSR_REQUIRE(block_width > 1 && block_width <= 16, GSR_ERR_INVALID_ARGUMENT,
              "block_width must be between 2 and 16 (got %u)", block_width);
  GSR_REQUIRE(img_height > 0 && img_width > 0 && num_points >= 0, GSR_ERR_INVALID_ARGUMENT, "rasterize_backward: bad sizes");
  if (num_points == 0) return GSR_OK;
  GSR_REQUIRE(gaussian_ids_sorted && tile_bins && xys && conics && colors && opacities && background && final_Ts &&
                  final_idx && v_output && v_output_alpha && v_xy && v_conic && v_colors && v_opacity,
              GSR_ERR_INVALID_ARGUMENT, "rasterize_backward: null pointer");
  GSR_REQUIRE((uintptr_t)xys % 8 == 0 && (uintptr_t)tile_bins % 8 == 0, GSR_ERR_INVALID_ARGUMENT,
              "rasterize_backward: xys / tile_bins must be 8-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  GSR_CUDA(cudaMemsetAsync(v_xy, 0, sizeof(float) * 2 * (size_t)num_points, st));
  GSR_CUDA(cudaMemsetAsync(v_conic, 0, sizeof(float) * 3 * (size_t)num_points, st));
  GSR_CUDA(cudaMemsetAsync(v_colors, 0, sizeof(float) * 3 * (size_t)num_points, st));
  GSR_CUDA(cudaMemsetAsync(v_opacity, 0, sizeof(float) * (size_t)num_points, st));
  const dim3 grid(cdiv(img_width, block_width), cdiv(img_height, block_width), 1);
  const unsigned threads = cdiv(block_width * block_width, 32) * 32;
  const int mode = block_width == 16 ? blend_bwd_mode() : 0;
  if (mode >= 8)
    return launch_blend_backward_tr(mode, grid, st, (int)img_width, (int)img_height, gaussian_ids_sorted,
                                    reinterpret_cast<const int2 *>(tile_bins), reinterpret_cast<const float2 *>(xys), conics,
                                    colors, opacities, background, final_Ts, final_idx, v_output, v_output_alpha, v_xy,
                                    v_conic, v_colors, v_opacity);
  if (mode == 1)
    return launch_blend_backward_scan(grid, st, (int)img_width, (int)img_height, gaussian_ids_sorted,
                                      reinterpret_cast<const int2 *>(tile_bins), reinterpret_cast<const float2 *>(xys),
                                      conics, colors, opacities, background, final_Ts, final_idx, v_output, v_output_alpha,
                                      v_xy, v_conic, v_colors, v_opacity);
  blend_backward_kernel<<<grid, threads, 0, st>>>(
      (int)grid.x, (int)img_width, (int)img_height, (int)block_width, gaussian_ids_sorted,
      reinterpret_cast<const int2 *>(tile_bins), reinterpret_cast<const float2 *>(xys), conics, colors, opacities,
      background, final_Ts, final_idx, v_output, v_output_alpha, v_xy, v_conic, v_colors, v_opacity);
  GSR_CHECK_LAUNCH("blend_backward_kernel");
  return GSR_OK;
}
