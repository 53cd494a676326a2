// blend_fwd.cu — per-tile front-to-back alpha compositing (3 channels, FP32) for sm_100a.
//
// Replaces rasterize_forward (reference csrc/forward.cu:278-395, launched from bindings.cu:269-328).
//
// One CTA per tile, one thread per pixel, like the reference; per pixel the arithmetic is the reference's
// (alpha = min(0.999, o exp(-sigma)); skip if sigma < 0 or alpha < 1/255; stop BEFORE applying the Gaussian
// that would bring T to <= 1e-4; final_idx = last applied).  What is different:
//   * the tile's Gaussian list is staged in batches through a DOUBLE-buffered shared-memory ring of packed
//     48-byte records; the gather of batch b+1 is issued into registers before batch b is composited, so
//     the random 4..12-byte L2/HBM gathers overlap the FP32 work (reference: load -> barrier -> compute on a
//     single buffer, colours re-read from global memory per pixel, forward.cu:375), one barrier per batch;
//   * a warp owns an 8x4 pixel sub-tile.  Before compositing a batch each warp COMPACTS it: lanes test 32
//     records per round against the warp's pixel rectangle (conservative alpha >= 1/255 extent box, see
//     blend_common.cuh) and a ballot builds the list of survivors; the per-pixel loop then only visits
//     survivors.  On the BASELINE scene 61 % of (warp, Gaussian) pairs are rejected at 0.5 instructions each
//     instead of running the 19-instruction loop body (ncu source view, profiles/r01);
//   * the per-pixel loop is warp-uniform (votes instead of per-lane `break`), so there is no
//     BSSY/BSYNC reconvergence traffic; exp(-sigma) is one FMUL-free MUFU.EX2 because log2(e) is folded into
//     the staged conic.
#include "blend_common.cuh"

#ifndef GSR_FWD_UNROLL
#define GSR_FWD_UNROLL 4  // survivors per trip of the walk (cfg2 on B200: 4 -> 0.391 ms, 8 -> 0.400, 2 -> 0.422)
#endif
#ifndef GSR_FWD_MIN_CTAS
#define GSR_FWD_MIN_CTAS 4  // 64 registers; 5 CTAs / SM (48 registers, rematerialised pixel coordinates): 0.413 ms
#endif

#ifndef GSR_FWD_RING
#define GSR_FWD_RING 2  // buffers of the staged-record ring
#endif

namespace gsr {

// MASKS (16x16 tiles only): the staging thread evaluates its record against the eight warp blocks once (block_mask_16)
// instead of every warp testing every record (compact_survivors) — see blend_common.cuh
// SHIFT = 4: the survivor lists hold byte offsets (slot * 16) into a record plane instead of slot numbers — one address
// instruction (LEA) less per visit of an issue-bound loop (0.379 -> 0.370 ms at cfg2).
template <bool MASKS, int SHIFT>
__device__ __forceinline__ void blend_forward_body(int tiles_x, int img_w, int img_h, int block_width,
                                                   const int *__restrict__ gaussian_ids_sorted,
                                                   const int2 *__restrict__ tile_bins, const float2 *__restrict__ xys,
                                                   const float *__restrict__ conics, const float *__restrict__ colors,
                                                   const float *__restrict__ opacities,
                                                   const float *__restrict__ background, float *__restrict__ out_img,
                                                   float *__restrict__ final_Ts, int *__restrict__ final_idx) {
  // slot kNull of every plane holds a record that never contributes (opacity 0): the survivor lists are padded with it
  constexpr int kNull = BLEND_THREADS, kUnroll = GSR_FWD_UNROLL;
  __shared__ float4 s_rec[GSR_FWD_RING][3][BLEND_THREADS + 1];
  __shared__ unsigned short s_list[BLEND_THREADS / 32][BLEND_THREADS + kUnroll + 2];
  __shared__ __align__(8) unsigned char s_mask[GSR_FWD_RING][MASKS ? BLEND_THREADS : 8];

  const unsigned full = 0xffffffffu;
  const int tile_x = blockIdx.x, tile_y = blockIdx.y;
  const int tile_id = tile_y * tiles_x + tile_x;
  const int tr = threadIdx.x, nthreads = blockDim.x, lane = tr & 31, warp = tr >> 5;
  int lx, ly;
  map_pixel(block_width, lx, ly);
  const int ipx = tile_x * block_width + lx, ipy = tile_y * block_width + ly;
  const bool inside = (ly < block_width) && (ipx < img_w) && (ipy < img_h);
  const float px = (float)ipx, py = (float)ipy;
  const bool done = !inside;

  // the warp's pixel rectangle (inside pixels only); empty warps get an impossible rectangle
  const float fx0 = (float)__reduce_min_sync(full, inside ? ipx : 0x7fffffff);
  const float fx1 = (float)__reduce_max_sync(full, inside ? ipx : -0x7fffffff);
  const float fy0 = (float)__reduce_min_sync(full, inside ? ipy : 0x7fffffff);
  const float fy1 = (float)__reduce_max_sync(full, inside ? ipy : -0x7fffffff);

  const int2 range = tile_bins[tile_id];
  const int num_batches = (range.y - range.x + nthreads - 1) / nthreads;
  const float tile_x0 = (float)(tile_x * 16), tile_y0 = (float)(tile_y * 16);  // MASKS: block_width == 16

  float T = 1.f;
  int cur_idx = 0;
  float acc_r = 0.f, acc_g = 0.f, acc_b = 0.f;
  // a finished pixel has slot_stop = -1; contributors need slot < slot_stop (one ISETP instead of a flag round trip)
  int slot_stop = done ? -1 : 0x7fffffff;

  if (tr < 6 && (tr & 1) < GSR_FWD_RING)
    s_rec[tr & 1][tr >> 1][kNull] = tr < 2 ? make_float4(0.f, 0.f, -1e30f, -1e30f) : make_float4(0.f, 0.f, 0.f, 0.f);

  BlendRecord rec;
  if (num_batches > 0 && range.x + tr < range.y)
    rec = gather_record(gaussian_ids_sorted[range.x + tr], xys, conics, colors, opacities);

  for (int b = 0; b < num_batches; ++b) {
    const int buf = GSR_FWD_RING == 2 ? (b & 1) : 0;
    if (GSR_FWD_RING == 1 && b > 0) __syncthreads();  // every warp is done with the previous batch
    const int batch_start = range.x + nthreads * b;
    if (batch_start + tr < range.y) {
      s_rec[buf][0][tr] = rec.r0;
      s_rec[buf][1][tr] = rec.r1;
      s_rec[buf][2][tr] = rec.r2;
      if (MASKS) s_mask[buf][tr] = (unsigned char)block_mask_16(rec.r0, rec.r1, tile_x0, tile_y0);
    }
    // one barrier per batch: publishes buffer `buf` and counts finished pixels (forward.cu:327-329)
    if (__syncthreads_count(slot_stop < 0) >= nthreads) break;
    {
      const int nxt = batch_start + nthreads + tr;
      if (nxt < range.y) rec = gather_record(gaussian_ids_sorted[nxt], xys, conics, colors, opacities);
    }
    if (__all_sync(full, slot_stop < 0)) continue;  // this warp is finished; it still stages and syncs

    const int batch_size = min(nthreads, range.y - batch_start);
    const unsigned short *lp = s_list[warp];
    const int n_list =
        MASKS ? compact_from_masks<unsigned short, kUnroll + 2, SHIFT>(s_mask[buf], warp, 0, batch_size, s_list[warp], lane,
                                                                       kNull)
              : compact_survivors<unsigned short, kUnroll + 2, SHIFT>(s_rec[buf][0], s_rec[buf][1], 0, batch_size, fx0, fx1,
                                                                      fy0, fy1, s_list[warp], lane, kNull);
    // record planes of this buffer, addressed by list entry e = slot << SHIFT
    const char *const p0 = reinterpret_cast<const char *>(&s_rec[buf][0][0]);
    const char *const p1 = reinterpret_cast<const char *>(&s_rec[buf][1][0]);
    const char *const p2 = reinterpret_cast<const char *>(&s_rec[buf][2][0]);
    constexpr int kMul = 16 >> SHIFT;  // bytes per unit of a list entry
    // Software-pipelined walk, kUnroll survivors per trip (the list is padded with the null record, so there is no
    // remainder loop); the next record is loaded while this one is evaluated, and the all-pixels-done vote is taken
    // once per trip — visits after the last pixel has finished change nothing (no lane contributes).
    int slot = lp[0];
    float2 c0 = *reinterpret_cast<const float2 *>(p0 + slot * kMul);
    float4 q1 = *reinterpret_cast<const float4 *>(p1 + slot * kMul);
    float4 q2 = *reinterpret_cast<const float4 *>(p2 + slot * kMul);
    int slot_n = lp[1];
    int last_slot = -1;
    for (int i = 0; i < n_list; i += kUnroll) {
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const float2 n0 = *reinterpret_cast<const float2 *>(p0 + slot_n * kMul);
        const float4 n1 = *reinterpret_cast<const float4 *>(p1 + slot_n * kMul);
        const float4 n2 = *reinterpret_cast<const float4 *>(p2 + slot_n * kMul);
        const int slot_nn = lp[i + u + 2];
        const float dx = c0.x - px, dy = c0.y - py;
        const float power = dx * (q1.x * dx + q1.y * dy) + q1.z * dy * dy;  // = -sigma * log2(e)
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
          T = next_T;
          last_slot = slot;
        }
        slot = slot_n; c0 = n0; q1 = n1; q2 = n2; slot_n = slot_nn;
      }
      if (__all_sync(full, slot_stop < 0)) break;
    }
    if (last_slot >= 0) cur_idx = batch_start + (last_slot >> SHIFT);
  }

  if (inside) {
    const int pix = ipy * img_w + ipx;
    final_Ts[pix] = T;
    final_idx[pix] = cur_idx;
    out_img[3 * (size_t)pix] = acc_r + T * background[0];
    out_img[3 * (size_t)pix + 1] = acc_g + T * background[1];
    out_img[3 * (size_t)pix + 2] = acc_b + T * background[2];
  }
}

#define GSR_FWD_PARAMS                                                                                                  \
  int tiles_x, int img_w, int img_h, int block_width, const int *__restrict__ gaussian_ids_sorted,                      \
      const int2 *__restrict__ tile_bins, const float2 *__restrict__ xys, const float *__restrict__ conics,             \
      const float *__restrict__ colors, const float *__restrict__ opacities, const float *__restrict__ background,      \
      float *__restrict__ out_img, float *__restrict__ final_Ts, int *__restrict__ final_idx
#define GSR_FWD_ARGS                                                                                                    \
  tiles_x, img_w, img_h, block_width, gaussian_ids_sorted, tile_bins, xys, conics, colors, opacities, background,       \
      out_img, final_Ts, final_idx

// 16x16 tiles (the default): block masks from the staging threads
__global__ void __launch_bounds__(BLEND_THREADS, GSR_FWD_MIN_CTAS) blend_forward_kernel(GSR_FWD_PARAMS) {
  blend_forward_body<true, 4>(GSR_FWD_ARGS);
}

// any block width (and 16x16 with GSR_BLOCK_MASK=0): per-warp tests
__global__ void __launch_bounds__(BLEND_THREADS, GSR_FWD_MIN_CTAS) blend_forward_warptest_kernel(GSR_FWD_PARAMS) {
  blend_forward_body<false, 0>(GSR_FWD_ARGS);
}

}  // namespace gsr

extern "C" GSR_API int gsr_rasterize_forward(unsigned img_height, unsigned img_width, unsigned block_width,
                                             int num_points, const int32_t *gaussian_ids_sorted,
                                             const int32_t *tile_bins, const float *xys, const float *conics,
                                             const float *colors, const float *opacities,
                                             const float *background, float *out_img, float *final_Ts,
                                             int32_t *final_idx, void *stream) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_rasterize_forward");
  (void)num_points;
  GSR_REQUIRE(block_width > 1 && block_width <= 16, GSR_ERR_INVALID_ARGUMENT,
              "block_width must be between 2 and 16 (got %u)", block_width);
  GSR_REQUIRE(img_height > 0 && img_width > 0, GSR_ERR_INVALID_ARGUMENT, "rasterize_forward: empty image");
  GSR_REQUIRE(gaussian_ids_sorted && tile_bins && xys && conics && colors && opacities && background && out_img &&
                  final_Ts && final_idx,
              GSR_ERR_INVALID_ARGUMENT, "rasterize_forward: null pointer");
  GSR_REQUIRE((uintptr_t)xys % 8 == 0 && (uintptr_t)tile_bins % 8 == 0, GSR_ERR_INVALID_ARGUMENT,
              "rasterize_forward: xys / tile_bins must be 8-byte aligned");
  const dim3 grid(cdiv(img_width, block_width), cdiv(img_height, block_width), 1);
  const unsigned threads = cdiv(block_width * block_width, 32) * 32;
  static const cudaError_t carve = [] {
    const char *e = getenv("GSR_FWD_CARVEOUT");  // percent of the maximum shared-memory carveout; unset = driver default
    return (e && e[0]) ? cudaFuncSetAttribute(blend_forward_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e))
                       : cudaSuccess;
  }();
  GSR_CUDA(carve);
  if (block_width == 16 && blend_block_masks()) {
    blend_forward_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(
        (int)grid.x, (int)img_width, (int)img_height, (int)block_width, gaussian_ids_sorted,
        reinterpret_cast<const int2 *>(tile_bins), reinterpret_cast<const float2 *>(xys), conics, colors, opacities,
        background, out_img, final_Ts, final_idx);
    GSR_CHECK_LAUNCH("blend_forward_kernel");
  } else {
    blend_forward_warptest_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(
        (int)grid.x, (int)img_width, (int)img_height, (int)block_width, gaussian_ids_sorted,
        reinterpret_cast<const int2 *>(tile_bins), reinterpret_cast<const float2 *>(xys), conics, colors, opacities,
        background, out_img, final_Ts, final_idx);
    GSR_CHECK_LAUNCH("blend_forward_warptest_kernel");
  }
  return GSR_OK;
}
