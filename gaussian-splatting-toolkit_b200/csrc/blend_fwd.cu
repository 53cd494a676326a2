// blend_fwd.cu — per-tile front-to-back alpha compositing (3 channels, FP32) for sm_100a.
//
// Replaces rasterize_forward (reference csrc/forward.cu:278-395, launched from bindings.cu:269-328).
//
// One CTA per tile, one thread per pixel.  Differences from the reference kernel (results are identical
// per pixel):
//   * the per-tile Gaussian list is staged in batches through a DOUBLE-buffered shared-memory ring of
//     packed 48-byte records {x,y,opacity,ext_x | a,b,c,ext_y | r,g,b,-}; the gather of batch b+1 is issued
//     into registers before batch b is composited, so the random 4..12-byte HBM/L2 gathers overlap the FP32
//     work instead of alternating with it (reference: load -> __syncthreads -> compute, one buffer), and
//     only one __syncthreads per batch is needed;
//   * colours are staged with the record instead of being re-read from global memory per pixel
//     (forward.cu:375);
//   * a warp covers an 8x4 pixel sub-tile (16x16 tiles) and rejects, with a warp-uniform test, Gaussians
//     whose alpha >= 1/255 ellipse cannot reach the sub-tile.  The test is conservative: for a pixel at
//     horizontal distance |dx| > ext_x = sqrt(2 ln(255 o) Sigma_xx) the Mahalanobis bound
//     sigma >= dx^2 / (2 Sigma_xx) gives o * exp(-sigma) < 1/255, which the reference skips as well
//     (forward.cu:361-363), so skipped pairs contribute nothing to C, T or final_idx;
//   * shared-memory reads are 128-bit broadcasts.
#include "common.cuh"

namespace gsr {

// Conservative half-extents of the region where alpha can reach 1/255; NaN (never rejects) when the conic
// is not positive definite; negative (always rejects) when opacity < 1/255.
__device__ __forceinline__ void alpha_extents(float a, float b, float c, float opac, float &ex, float &ey) {
  const float det = a * c - b * b;             // det(conic) = 1 / det(cov2d)
  const float tau2 = 2.f * __logf(255.f * opac);  // 2 * sigma threshold
  if (!(tau2 >= 0.f)) {                        // opac < 1/255 (or NaN): can never pass the alpha test
    ex = (opac == opac) ? -1e30f : __int_as_float(0x7fc00000);
    ey = ex;
    return;
  }
  // Sigma_xx = c / det, Sigma_yy = a / det ; 0.1% + 0.01 px safety margin against rounding
  ex = sqrtf(tau2 * c / det) * 1.001f + 0.01f;
  ey = sqrtf(tau2 * a / det) * 1.001f + 0.01f;
}

struct PixelMap {
  int px, py;      // pixel coordinates
  bool inside;     // inside the image and inside the tile
};

// thread -> pixel mapping: 8x4 sub-tiles per warp for 16x16 tiles, row-major otherwise
__device__ __forceinline__ PixelMap map_pixel(int block_width, int tile_x, int tile_y, int img_w, int img_h) {
  const int tr = threadIdx.x;
  int lx, ly;
  if (block_width == 16) {
    const int w = tr >> 5, l = tr & 31;
    lx = ((w & 1) << 3) + (l & 7);
    ly = ((w >> 1) << 2) + (l >> 3);
  } else {
    lx = tr % block_width;
    ly = tr / block_width;
  }
  PixelMap m;
  m.px = tile_x * block_width + lx;
  m.py = tile_y * block_width + ly;
  m.inside = (ly < block_width) && (m.px < img_w) && (m.py < img_h);
  return m;
}

template <int MAX_THREADS>
__global__ void __launch_bounds__(MAX_THREADS)
blend_forward_kernel(int tiles_x, int img_w, int img_h, int block_width,
                     const int *__restrict__ gaussian_ids_sorted, const int2 *__restrict__ tile_bins,
                     const float2 *__restrict__ xys, const float *__restrict__ conics,
                     const float *__restrict__ colors, const float *__restrict__ opacities,
                     const float *__restrict__ background, float *__restrict__ out_img,
                     float *__restrict__ final_Ts, int *__restrict__ final_idx) {
  __shared__ float4 s_rec[2][3][MAX_THREADS];

  const int tile_x = blockIdx.x, tile_y = blockIdx.y;
  const int tile_id = tile_y * tiles_x + tile_x;
  const int tr = threadIdx.x, nthreads = blockDim.x;
  const PixelMap pm = map_pixel(block_width, tile_x, tile_y, img_w, img_h);
  const float px = (float)pm.px, py = (float)pm.py;
  bool done = !pm.inside;

  // warp's pixel rectangle (only inside pixels count); empty warps get an impossible rectangle
  const unsigned full = 0xffffffffu;
  const int wx0 = __reduce_min_sync(full, pm.inside ? pm.px : 0x7fffffff);
  const int wx1 = __reduce_max_sync(full, pm.inside ? pm.px : -0x7fffffff);
  const int wy0 = __reduce_min_sync(full, pm.inside ? pm.py : 0x7fffffff);
  const int wy1 = __reduce_max_sync(full, pm.inside ? pm.py : -0x7fffffff);
  const float fx0 = (float)wx0, fx1 = (float)wx1, fy0 = (float)wy0, fy1 = (float)wy1;

  const int2 range = tile_bins[tile_id];
  const int num_batches = (range.y - range.x + nthreads - 1) / nthreads;

  float T = 1.f;
  int cur_idx = 0;
  float3 acc = make_float3(0.f, 0.f, 0.f);

  // register-staged prefetch of one record per thread
  float4 r0, r1, r2;
  auto fetch = [&](int idx) {
    if (idx < range.y) {
      const int g = gaussian_ids_sorted[idx];
      const float2 xy = xys[g];
      const float opac = opacities[g];
      const float a = conics[3 * (size_t)g], b = conics[3 * (size_t)g + 1], c = conics[3 * (size_t)g + 2];
      float ex, ey;
      alpha_extents(a, b, c, opac, ex, ey);
      r0 = make_float4(xy.x, xy.y, opac, ex);
      r1 = make_float4(a, b, c, ey);
      r2 = make_float4(colors[3 * (size_t)g], colors[3 * (size_t)g + 1], colors[3 * (size_t)g + 2], 0.f);
    }
  };
  if (num_batches > 0) fetch(range.x + tr);

  for (int b = 0; b < num_batches; ++b) {
    const int buf = b & 1;
    const int batch_start = range.x + nthreads * b;
    if (batch_start + tr < range.y) {
      s_rec[buf][0][tr] = r0;
      s_rec[buf][1][tr] = r1;
      s_rec[buf][2][tr] = r2;
    }
    // one barrier per batch: publishes buffer `buf` and counts finished pixels (forward.cu:327-329)
    if (__syncthreads_count(done) >= nthreads) break;
    if (b + 1 < num_batches) fetch(batch_start + nthreads + tr);

    const int batch_size = min(nthreads, range.y - batch_start);
    if (!__all_sync(full, done)) {
      for (int t = 0; (t < batch_size) && !done; ++t) {
        const float4 q0 = s_rec[buf][0][t];
        const float4 q1 = s_rec[buf][1][t];
        // warp-uniform conservative reject (see header)
        if (q0.x + q0.w < fx0 || q0.x - q0.w > fx1 || q0.y + q1.w < fy0 || q0.y - q1.w > fy1) continue;
        const float dx = q0.x - px, dy = q0.y - py;
        const float sigma = 0.5f * (q1.x * dx * dx + q1.z * dy * dy) + q1.y * dx * dy;
        const float alpha = fminf(0.999f, q0.z * __expf(-sigma));
        if (sigma < 0.f || alpha < 1.f / 255.f) continue;
        const float next_T = T * (1.f - alpha);
        if (next_T <= 1e-4f) {
          done = true;
          break;
        }
        const float4 q2 = s_rec[buf][2][t];
        const float vis = alpha * T;
        acc.x += q2.x * vis;
        acc.y += q2.y * vis;
        acc.z += q2.z * vis;
        T = next_T;
        cur_idx = batch_start + t;
      }
    }
  }

  if (pm.inside) {
    const int pix = pm.py * img_w + pm.px;
    final_Ts[pix] = T;
    final_idx[pix] = cur_idx;
    out_img[3 * (size_t)pix] = acc.x + T * background[0];
    out_img[3 * (size_t)pix + 1] = acc.y + T * background[1];
    out_img[3 * (size_t)pix + 2] = acc.z + T * background[2];
  }
}

}  // namespace gsr

extern "C" GSR_API int gsr_rasterize_forward(unsigned img_height, unsigned img_width, unsigned block_width,
                                             int num_points, const int32_t *gaussian_ids_sorted,
                                             const int32_t *tile_bins, const float *xys, const float *conics,
                                             const float *colors, const float *opacities,
                                             const float *background, float *out_img, float *final_Ts,
                                             int32_t *final_idx, void *stream) {
  using namespace gsr;
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
  blend_forward_kernel<256><<<grid, threads, 0, (cudaStream_t)stream>>>(
      (int)grid.x, (int)img_width, (int)img_height, (int)block_width, gaussian_ids_sorted,
      reinterpret_cast<const int2 *>(tile_bins), reinterpret_cast<const float2 *>(xys), conics, colors, opacities,
      background, out_img, final_Ts, final_idx);
  GSR_CHECK_LAUNCH("blend_forward_kernel");
  return GSR_OK;
}
