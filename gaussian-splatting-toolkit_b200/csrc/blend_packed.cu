// blend_packed.cu — the per-tile half of the FUSED render operator: alpha compositing of RGB and,
// in the same pass, of depth (4th channel) from the packed per-Gaussian records written by
// fused_preprocess_forward_kernel, and the adjoint accumulating one packed 12-float gradient record per Gaussian.
//
// The reference models render depth with a SECOND complete rasterize_gaussians call whose colours are the depths
// (gs_toolkit/models/vanilla_gs.py:839-855): second binning, second sort, second blend over the same lists.  Per pixel
// both passes walk the same Gaussians with the same alphas and stop at the same place, so compositing the depth as a
// fourth channel of the colour pass is the same arithmetic (forward.cu:278-395 with one more accumulator).
//
// Kernel structure = blend_fwd.cu / blend_bwd.cu (double-buffered record ring, 8x4-pixel warps, ballot compaction
// against the alpha >= 1/255 extent box, warp-uniform loops, transposing-butterfly reduction, one RED per
// (warp, Gaussian)); differences: the record is gathered with three 16-byte cp.async copies (LDGSTS) straight into the
// shared-memory ring — no register staging, no per-pair extent maths —, the depth rides in r2.w, and gradients land in
// one 48-byte record per Gaussian
// {v_x, v_y, v_opacity, v_depth | v_a, v_b, v_c, - | v_r, v_g, v_b, -}.
#include "blend_common.cuh"

namespace gsr {

// Asynchronous global -> shared copy of one 16-byte record plane entry (cp.async / LDGSTS): the gathered record goes
// straight into the shared-memory ring without passing through registers, and the copy of batch b+1 is in flight while
// batch b is composited.
__device__ __forceinline__ void cp_async16(float4 *smem_dst, const float4 *gmem_src) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// stage the record of Gaussian g into slot `slot` of ring buffer `buf`
template <int SLOTS>
__device__ __forceinline__ void stage_packed(float4 (*s_rec)[3][SLOTS], int buf, int slot, int g, int n,
                                             const float4 *__restrict__ rec) {
  cp_async16(&s_rec[buf][0][slot], rec + g);
  cp_async16(&s_rec[buf][1][slot], rec + n + g);
  cp_async16(&s_rec[buf][2][slot], rec + 2 * (size_t)n + g);
}

template <bool DEPTH>
__global__ void __launch_bounds__(BLEND_THREADS)
blend_packed_forward_kernel(int tiles_x, int img_w, int img_h, int block_width, int num_points,
                            const int *__restrict__ gaussian_ids_sorted, const int2 *__restrict__ tile_bins,
                            const float4 *__restrict__ rec, const float *__restrict__ background,
                            float *__restrict__ out_img, float *__restrict__ out_depth, float *__restrict__ final_Ts,
                            int *__restrict__ final_idx) {
  // slot kNull of every plane holds a record that never contributes (opacity 0): the survivor lists are padded with it
  constexpr int kNull = BLEND_THREADS, kUnroll = 4;
  __shared__ float4 s_rec[2][3][BLEND_THREADS + 1];
  __shared__ unsigned short s_list[BLEND_THREADS / 32][BLEND_THREADS + kUnroll + 2];

  const unsigned full = 0xffffffffu;
  const int tile_x = blockIdx.x, tile_y = blockIdx.y;
  const int tile_id = tile_y * tiles_x + tile_x;
  const int tr = threadIdx.x, nthreads = blockDim.x, lane = tr & 31, warp = tr >> 5;
  int lx, ly;
  map_pixel(block_width, lx, ly);
  const int ipx = tile_x * block_width + lx, ipy = tile_y * block_width + ly;
  const bool inside = (ly < block_width) && (ipx < img_w) && (ipy < img_h);
  const float px = (float)ipx, py = (float)ipy;
  const bool done0 = !inside;
  const float fx0 = (float)__reduce_min_sync(full, inside ? ipx : 0x7fffffff);
  const float fx1 = (float)__reduce_max_sync(full, inside ? ipx : -0x7fffffff);
  const float fy0 = (float)__reduce_min_sync(full, inside ? ipy : 0x7fffffff);
  const float fy1 = (float)__reduce_max_sync(full, inside ? ipy : -0x7fffffff);

  const int2 range = tile_bins[tile_id];
  const int num_batches = (range.y - range.x + nthreads - 1) / nthreads;
  float T = 1.f;
  int cur_idx = 0;
  float acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, acc_d = 0.f;
  // a finished pixel has slot_stop = -1; contributors need slot < slot_stop (see blend_fwd.cu)
  int slot_stop = done0 ? -1 : 0x7fffffff;
  if (tr < 6) s_rec[tr & 1][tr >> 1][kNull] = tr < 2 ? make_float4(0.f, 0.f, -1e30f, -1e30f) : make_float4(0.f, 0.f, 0.f, 0.f);

  // 2-stage cp.async ring: batch b lives in buffer b & 1; its copies were issued one iteration earlier
  if (num_batches > 0 && range.x + tr < range.y)
    stage_packed(s_rec, 0, tr, gaussian_ids_sorted[range.x + tr], num_points, rec);
  cp_async_commit();
  for (int b = 0; b < num_batches; ++b) {
    const int buf = b & 1;
    const int batch_start = range.x + nthreads * b;
    cp_async_wait_all();  // this thread's copies for batch b have landed ...
    // ... and the barrier publishes everyone's; it also tells that buffer buf ^ 1 (batch b-1) is no longer read
    if (__syncthreads_count(slot_stop < 0) >= nthreads) break;
    {
      const int nxt = batch_start + nthreads + tr;
      if (nxt < range.y) stage_packed(s_rec, buf ^ 1, tr, gaussian_ids_sorted[nxt], num_points, rec);
      cp_async_commit();
    }
    if (__all_sync(full, slot_stop < 0)) continue;
    const int batch_size = min(nthreads, range.y - batch_start);
    const unsigned short *lp = s_list[warp];
    const int n_list = compact_survivors<unsigned short, kUnroll + 2>(s_rec[buf][0], s_rec[buf][1], 0, batch_size, fx0, fx1,
                                                                      fy0, fy1, s_list[warp], lane, kNull);
    // software-pipelined walk, kUnroll survivors per trip, one all-done vote per trip (blend_fwd.cu)
    int slot = lp[0];
    float2 c0 = *reinterpret_cast<const float2 *>(&s_rec[buf][0][slot]);
    float4 q1 = s_rec[buf][1][slot];
    float4 q2 = s_rec[buf][2][slot];
    int slot_n = lp[1];
    int last_slot = -1;
    for (int i = 0; i < n_list; i += kUnroll) {
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const float2 n0 = *reinterpret_cast<const float2 *>(&s_rec[buf][0][slot_n]);
        const float4 n1 = s_rec[buf][1][slot_n];
        const float4 n2 = s_rec[buf][2][slot_n];
        const int slot_nn = lp[i + u + 2];
        const float dx = c0.x - px, dy = c0.y - py;
        const float power = dx * (q1.x * dx + q1.y * dy) + q1.z * dy * dy;
        const float alpha = fminf(0.999f, q1.w * exp2f(power));
        const bool contrib = (slot < slot_stop) && !(power > 0.f || alpha < 1.f / 255.f);
        const float next_T = T * (1.f - alpha);
        const bool stop = contrib && (next_T <= 1e-4f);
        if (stop) slot_stop = -1;
        if (contrib && !stop) {
          const float vis = alpha * T;
          acc_r += q2.x * vis;
          acc_g += q2.y * vis;
          acc_b += q2.z * vis;
          if (DEPTH) acc_d += q2.w * vis;
          T = next_T;
          last_slot = slot;
        }
        slot = slot_n; c0 = n0; q1 = n1; q2 = n2; slot_n = slot_nn;
      }
      if (__all_sync(full, slot_stop < 0)) break;
    }
    if (last_slot >= 0) cur_idx = batch_start + last_slot;
  }
  if (inside) {
    const int pix = ipy * img_w + ipx;
    final_Ts[pix] = T;
    final_idx[pix] = cur_idx;
    out_img[3 * (size_t)pix] = acc_r + T * background[0];
    out_img[3 * (size_t)pix + 1] = acc_g + T * background[1];
    out_img[3 * (size_t)pix + 2] = acc_b + T * background[2];
    if (DEPTH) out_depth[pix] = acc_d;  // depth composites against a zero background (vanilla_gs.py:850)
  }
}

__device__ __forceinline__ float pk_reduce8(float v[8], int lane) {
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

// two values: lanes with bit 4 clear end up with the total of a, lanes with bit 4 set with the total of b
__device__ __forceinline__ float pk_reduce2(float a, float b, int lane) {
  const unsigned full = 0xffffffffu;
  const bool hi = lane & 16;
  float x = (hi ? b : a) + __shfl_xor_sync(full, hi ? a : b, 16);
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) x += __shfl_xor_sync(full, x, o);
  return x;
}

template <bool DEPTH>
__global__ void __launch_bounds__(BLEND_THREADS)
blend_packed_backward_kernel(int tiles_x, int img_w, int img_h, int block_width, int num_points,
                             const int *__restrict__ gaussian_ids_sorted, const int2 *__restrict__ tile_bins,
                             const float4 *__restrict__ rec, const float *__restrict__ background,
                             const float *__restrict__ final_Ts, const int *__restrict__ final_idx,
                             const float *__restrict__ v_output, const float *__restrict__ v_output_depth,
                             const float *__restrict__ v_output_alpha, float *__restrict__ grad_rec) {
  __shared__ float4 s_rec[2][3][BLEND_THREADS];
  __shared__ int s_gid[2][BLEND_THREADS];
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
  float buf_r = 0.f, buf_g = 0.f, buf_b = 0.f, buf_d = 0.f;
  const int bin_final = inside ? final_idx[pix] : -1;
  float vo_r = 0.f, vo_g = 0.f, vo_b = 0.f, vo_d = 0.f, vo_a = 0.f;
  if (inside) {
    vo_r = v_output[3 * (size_t)pix];
    vo_g = v_output[3 * (size_t)pix + 1];
    vo_b = v_output[3 * (size_t)pix + 2];
    vo_a = v_output_alpha[pix];
    if (DEPTH) vo_d = v_output_depth[pix];
  }
  // the depth channel has a zero background, so it adds nothing to the T_final term
  const float c_final = T_final * (vo_a - (background[0] * vo_r + background[1] * vo_g + background[2] * vo_b));

  const int warp_bin_final = __reduce_max_sync(full, bin_final);
  if (lane == 0) s_warp_max[warp] = warp_bin_final;
  __syncthreads();
  int cta_bin_final = -1;
  for (int w = 0; w < (nthreads >> 5); ++w) cta_bin_final = max(cta_bin_final, s_warp_max[w]);
  const int end = min(range.y, cta_bin_final + 1);
  const int count = end - range.x;
  if (count <= 0) return;
  const int num_batches = (count + nthreads - 1) / nthreads;

  // slots of the packed gradient record owned by the lanes after the reductions:
  //   reduce8 values {v_r, v_g, v_b, v_depth, v_a, v_b(conic), v_c, v_opacity} -> lanes 0,4,..,28
  //   reduce2 values {v_x, v_y} -> lane 1 (bit 4 clear) and lane 17 (bit 4 set)
  int dst_slot = -1;
  if ((lane & 3) == 0) {
    const int vi = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    dst_slot = vi < 3 ? 8 + vi : (vi == 7 ? 2 : vi);  // {8, 9, 10, 3, 4, 5, 6, 2}
    if (!DEPTH && vi == 3) dst_slot = -1;
  } else if (lane == 1) {
    dst_slot = 0;
  } else if (lane == 17) {
    dst_slot = 1;
  }

  if (end - 1 - tr >= range.x) {
    const int gid = gaussian_ids_sorted[end - 1 - tr];
    s_gid[0][tr] = gid;
    stage_packed(s_rec, 0, tr, gid, num_points, rec);
  }
  cp_async_commit();
  for (int b = 0; b < num_batches; ++b) {
    const int buf = b & 1;
    const int batch_end = end - 1 - nthreads * b;
    cp_async_wait_all();
    __syncthreads();
    {
      const int nxt = batch_end - nthreads - tr;
      if (nxt >= range.x) {
        const int gid = gaussian_ids_sorted[nxt];
        s_gid[buf ^ 1][tr] = gid;
        stage_packed(s_rec, buf ^ 1, tr, gid, num_points, rec);
      }
      cp_async_commit();
    }
    const int batch_size = min(nthreads, batch_end + 1 - range.x);
    const int t_begin = max(0, batch_end - warp_bin_final);
    if (t_begin >= batch_size) continue;
    const int n_list = compact_survivors(s_rec[buf][0], s_rec[buf][1], t_begin, batch_size, fx0, fx1, fy0, fy1, s_list[warp], lane);
    for (int i = 0; i < n_list; ++i) {
      const int t = s_list[warp][i];
      const float4 q0 = s_rec[buf][0][t];
      const float4 q1 = s_rec[buf][1][t];
      const float dx = q0.x - px, dy = q0.y - py;
      const float gx = q1.x * dx, gy = q1.z * dy;
      const float power = dx * (gx + q1.y * dy) + gy * dy;
      const float vis = exp2f(power);
      const float opac = q1.w;
      const float alpha = fminf(0.99f, opac * vis);
      const bool valid = inside && (batch_end - t <= bin_final) && !(power > 0.f || alpha < 1.f / 255.f);
      if (!__any_sync(full, valid)) continue;

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
      v[3] = DEPTH ? fac * vo_d : 0.f;
      float v_alpha = (q2.x * T - buf_r * ra) * vo_r;
      v_alpha += (q2.y * T - buf_g * ra) * vo_g;
      v_alpha += (q2.z * T - buf_b * ra) * vo_b;
      if (DEPTH) v_alpha += (q2.w * T - buf_d * ra) * vo_d;
      v_alpha += ra * c_final;
      buf_r += q2.x * fac;
      buf_g += q2.y * fac;
      buf_b += q2.z * fac;
      if (DEPTH) buf_d += q2.w * fac;
      const float v_sigma = -opac * vis_e * v_alpha;
      const float hs = 0.5f * v_sigma;
      v[4] = hs * dx * dx;
      v[5] = v_sigma * dx * dy;
      v[6] = hs * dy * dy;
      v[7] = vis_e * v_alpha;
      const float ws = -kLn2 * v_sigma;
      const float vxl = ws * (2.f * gx + q1.y * dy);
      const float vyl = ws * (q1.y * dx + 2.f * gy);
      const float tot8 = pk_reduce8(v, lane);
      const float tot2 = pk_reduce2(vxl, vyl, lane);
      if (dst_slot >= 0) {
        const unsigned g = (unsigned)s_gid[buf][t];
        atomicAdd(grad_rec + 12u * g + (unsigned)dst_slot, (lane & 3) == 0 ? tot8 : tot2);
      }
    }
  }
}

// 16x16 tiles: blend_packed_tr.cu (two-phase transposing adjoint, default) or blend_packed_scan.cu (GSR_PACKED_BWD=scan);
// GSR_PACKED_BWD=pixel selects the kernel above
int blend_packed_bwd_mode();  // 0 = pixel, 1 = scan, 2 = tr
int launch_blend_packed_backward_tr(dim3 grid, cudaStream_t st, int img_w, int img_h, int num_points,
                                    const int *gaussian_ids_sorted, const int2 *tile_bins, const float4 *rec,
                                    const float *background, const float *final_Ts, const int *final_idx,
                                    const float *v_output, const float *v_output_depth, const float *v_output_alpha,
                                    float *grad_rec);
int launch_blend_packed_backward_scan(dim3 grid, cudaStream_t st, int img_w, int img_h, int num_points,
                                      const int *gaussian_ids_sorted, const int2 *tile_bins, const float4 *rec,
                                      const float *background, const float *final_Ts, const int *final_idx,
                                      const float *v_output, const float *v_output_depth, const float *v_output_alpha,
                                      float *grad_rec);

}  // namespace gsr

extern "C" {

GSR_API int gsr_blend_packed_forward(unsigned img_height, unsigned img_width, unsigned block_width, int num_points,
                                     const int32_t *gaussian_ids_sorted, const int32_t *tile_bins,
                                     const float *records, const float *background, float *out_img,
                                     float *out_depth /*nullable*/, float *final_Ts, int32_t *final_idx, void *stream) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_blend_packed_forward");
  GSR_REQUIRE(block_width > 1 && block_width <= 16, GSR_ERR_INVALID_ARGUMENT,
              "block_width must be between 2 and 16 (got %u)", block_width);
  GSR_REQUIRE(img_height > 0 && img_width > 0 && num_points >= 0, GSR_ERR_INVALID_ARGUMENT, "blend_packed_forward: bad sizes");
  GSR_REQUIRE(gaussian_ids_sorted && tile_bins && records && background && out_img && final_Ts && final_idx,
              GSR_ERR_INVALID_ARGUMENT, "blend_packed_forward: null pointer");
  GSR_REQUIRE((uintptr_t)records % 16 == 0 && (uintptr_t)tile_bins % 8 == 0, GSR_ERR_INVALID_ARGUMENT,
              "blend_packed_forward: records must be 16-byte, tile_bins 8-byte aligned");
  const dim3 grid(cdiv(img_width, block_width), cdiv(img_height, block_width), 1);
  const unsigned threads = cdiv(block_width * block_width, 32) * 32;
  cudaStream_t st = (cudaStream_t)stream;
  const float4 *rec = reinterpret_cast<const float4 *>(records);
  const int2 *bins = reinterpret_cast<const int2 *>(tile_bins);
  if (out_depth)
    blend_packed_forward_kernel<true><<<grid, threads, 0, st>>>((int)grid.x, (int)img_width, (int)img_height,
                                                                (int)block_width, num_points, gaussian_ids_sorted, bins,
                                                                rec, background, out_img, out_depth, final_Ts, final_idx);
  else
    blend_packed_forward_kernel<false><<<grid, threads, 0, st>>>((int)grid.x, (int)img_width, (int)img_height,
                                                                 (int)block_width, num_points, gaussian_ids_sorted, bins,
                                                                 rec, background, out_img, nullptr, final_Ts, final_idx);
  GSR_CHECK_LAUNCH("blend_packed_forward_kernel");
  return GSR_OK;
}

GSR_API int gsr_blend_packed_backward(unsigned img_height, unsigned img_width, unsigned block_width, int num_points,
                                      const int32_t *gaussian_ids_sorted, const int32_t *tile_bins,
                                      const float *records, const float *background, const float *final_Ts,
                                      const int32_t *final_idx, const float *v_output,
                                      const float *v_output_depth /*nullable*/, const float *v_output_alpha,
                                      float *grad_records, void *stream) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_blend_packed_backward");
  GSR_REQUIRE(block_width > 1 && block_width <= 16, GSR_ERR_INVALID_ARGUMENT,
              "block_width must be between 2 and 16 (got %u)", block_width);
  GSR_REQUIRE(img_height > 0 && img_width > 0 && num_points >= 0, GSR_ERR_INVALID_ARGUMENT, "blend_packed_backward: bad sizes");
  if (num_points == 0) return GSR_OK;
  GSR_REQUIRE(gaussian_ids_sorted && tile_bins && records && background && final_Ts && final_idx && v_output &&
                  v_output_alpha && grad_records,
              GSR_ERR_INVALID_ARGUMENT, "blend_packed_backward: null pointer");
  GSR_REQUIRE((uintptr_t)records % 16 == 0 && (uintptr_t)tile_bins % 8 == 0, GSR_ERR_INVALID_ARGUMENT,
              "blend_packed_backward: records must be 16-byte, tile_bins 8-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  GSR_CUDA(cudaMemsetAsync(grad_records, 0, sizeof(float) * 12 * (size_t)num_points, st));
  const dim3 grid(cdiv(img_width, block_width), cdiv(img_height, block_width), 1);
  const unsigned threads = cdiv(block_width * block_width, 32) * 32;
  const float4 *rec = reinterpret_cast<const float4 *>(records);
  const int2 *bins = reinterpret_cast<const int2 *>(tile_bins);
  const int mode = block_width == 16 ? blend_packed_bwd_mode() : 0;
  if (mode != 0)
    GSR_REQUIRE((uintptr_t)grad_records % 16 == 0, GSR_ERR_INVALID_ARGUMENT, "blend_packed_backward: grad_records must be 16-byte aligned");
  if (mode == 2)
    return launch_blend_packed_backward_tr(grid, st, (int)img_width, (int)img_height, num_points, gaussian_ids_sorted, bins,
                                           rec, background, final_Ts, final_idx, v_output, v_output_depth, v_output_alpha,
                                           grad_records);
  if (mode == 1) {
    return launch_blend_packed_backward_scan(grid, st, (int)img_width, (int)img_height, num_points, gaussian_ids_sorted, bins,
                                             rec, background, final_Ts, final_idx, v_output, v_output_depth, v_output_alpha,
                                             grad_records);
  }
  if (v_output_depth)
    blend_packed_backward_kernel<true><<<grid, threads, 0, st>>>(
        (int)grid.x, (int)img_width, (int)img_height, (int)block_width, num_points, gaussian_ids_sorted, bins, rec,
        background, final_Ts, final_idx, v_output, v_output_depth, v_output_alpha, grad_records);
  else
    blend_packed_backward_kernel<false><<<grid, threads, 0, st>>>(
        (int)grid.x, (int)img_width, (int)img_height, (int)block_width, num_points, gaussian_ids_sorted, bins, rec,
        background, final_Ts, final_idx, v_output, nullptr, v_output_alpha, grad_records);
  GSR_CHECK_LAUNCH("blend_packed_backward_kernel");
  return GSR_OK;
}
}
