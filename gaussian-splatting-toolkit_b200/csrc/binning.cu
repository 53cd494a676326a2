// binning.cu — tile binning for sm_100a: intersection offsets (scan), key emission, radix sort, bin edges.
//
// Replaces, in the reference's forward orchestration (rasterizer/rasterize.py:92-183 -> rasterizer/utils.py):
//   torch.cumsum(int32) + .item()          utils.py:123-124      -> gsr_cumsum_tiles_hit
//   map_gaussian_to_intersects kernel      csrc/forward.cu:94-127 -> gsr_map_gaussian_to_intersects
//   torch.sort(int64) + torch.gather       utils.py:179-180       -> gsr_sort_intersects
//   get_tile_bin_edges kernel              csrc/forward.cu:132-154 -> gsr_get_tile_bin_edges
//
// All of it is HBM-bound integer work.  The sort keeps the reference's key format ((tile << 32) | depth bits,
// int64) so sorted keys are bit-identical, but only radix-sorts the bits that can be non-zero
// (32 + ceil(log2(num_tiles))) and carries the 4-byte Gaussian id as the value instead of producing an
// 8-byte permutation and gathering through it.  Scan and sort are the hand-written kernels of radix_sort.cuh.
#include "common.cuh"
#include "radix_sort.cuh"
#include "tile_cull.cuh"

namespace gsr {

constexpr int BIN_THREADS = 256;

__global__ void __launch_bounds__(BIN_THREADS)
map_intersects_kernel(int n, const float2 *__restrict__ xys, const float *__restrict__ depths,
                      const int *__restrict__ radii, const int *__restrict__ cum_tiles_hit, int tiles_x,
                      int tiles_y, int block_width, int64_t *__restrict__ isect_ids,
                      int *__restrict__ gaussian_ids) {
  const int idx = blockIdx.x * BIN_THREADS + threadIdx.x;
  if (idx >= n) return;
  const int r = radii[idx];
  if (r <= 0) return;
  const float2 c = xys[idx];
  int x0, y0, x1, y1;
  tile_bbox(c.x, c.y, (float)r, tiles_x, tiles_y, block_width, x0, y0, x1, y1);
  int cur = (idx == 0) ? 0 : cum_tiles_hit[idx - 1];
  const int64_t depth_id = (int64_t)__float_as_int(depths[idx]);  // sign-extending, as the reference
  for (int i = y0; i < y1; ++i) {
    for (int j = x0; j < x1; ++j) {
      const int64_t tile_id = (int64_t)(i * tiles_x + j);
      isect_ids[cur] = (tile_id << 32) | depth_id;
      gaussian_ids[cur] = idx;
      ++cur;
    }
  }
}

// exact (conservative) tile culling: see tile_cull.cuh
template <bool EMIT>
__global__ void __launch_bounds__(BIN_THREADS)
tight_tiles_kernel(int n, const float2 *__restrict__ xys, const float *__restrict__ depths,
                   const int *__restrict__ radii, const float *__restrict__ conics,
                   const float *__restrict__ opacities, const int *__restrict__ cum_tiles, int tiles_x, int tiles_y,
                   int block_width, int img_w, int img_h, int *__restrict__ tiles_out,
                   int64_t *__restrict__ isect_ids, int *__restrict__ gaussian_ids) {
  const int idx = blockIdx.x * BIN_THREADS + threadIdx.x;
  if (idx >= n) return;
  const int r = radii[idx];
  int count = 0;
  if (r > 0) {
    const float2 ctr = xys[idx];
    int x0, y0, x1, y1;
    tile_bbox(ctr.x, ctr.y, (float)r, tiles_x, tiles_y, block_width, x0, y0, x1, y1);
    const CullEllipse e = make_cull_ellipse(conics[3 * (size_t)idx], conics[3 * (size_t)idx + 1],
                                            conics[3 * (size_t)idx + 2], opacities[idx], (float)(r + block_width));
    if (!e.empty) {
      int cur = 0;
      int64_t depth_id = 0;
      if (EMIT) {
        cur = (idx == 0) ? 0 : cum_tiles[idx - 1];
        depth_id = (int64_t)__float_as_int(depths[idx]);
      }
      for (int i = y0; i < y1; ++i) {
        int j0, j1;
        cull_row_range(e, ctr.x, ctr.y, i, x0, x1, block_width, j0, j1);
        if (EMIT) {
          for (int j = j0; j < j1; ++j) {
            isect_ids[cur] = ((int64_t)(i * tiles_x + j) << 32) | depth_id;
            gaussian_ids[cur] = idx;
            ++cur;
          }
        }
        count += j1 - j0;
      }
    }
  }
  if (!EMIT) tiles_out[idx] = count;
}

__global__ void __launch_bounds__(BIN_THREADS)
tile_bin_edges_kernel(int m, const int64_t *__restrict__ keys, int2 *__restrict__ tile_bins) {
  const int idx = blockIdx.x * BIN_THREADS + threadIdx.x;
  if (idx >= m) return;
  const int cur = (int)(keys[idx] >> 32);
  if (idx == 0) tile_bins[cur].x = 0;
  if (idx == m - 1) tile_bins[cur].y = m;
  if (idx == 0) return;
  const int prev = (int)(keys[idx - 1] >> 32);
  if (prev != cur) {
    tile_bins[prev].y = idx;
    tile_bins[cur].x = idx;
  }
}

static inline int key_end_bit(int num_tiles) {
  int bits = 0;
  while ((1ll << bits) < (long long)num_tiles) ++bits;
  return 32 + (bits < 1 ? 1 : bits);
}

}  // namespace gsr

extern "C" {

GSR_API size_t gsr_cumsum_workspace_bytes(int num_points) {
  return gsr::rs::scan_workspace_bytes(num_points) + 512;
}

GSR_API int gsr_cumsum_tiles_hit(int num_points, const int32_t *num_tiles_hit, int32_t *cum_tiles_hit,
                                 int32_t *total_host_pinned, void *workspace, size_t workspace_bytes,
                                 void *stream) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_cumsum_tiles_hit");
  GSR_REQUIRE(num_points >= 0, GSR_ERR_INVALID_ARGUMENT, "cumsum_tiles_hit: num_points < 0");
  if (num_points == 0) {
    if (total_host_pinned) *total_host_pinned = 0;
    return GSR_OK;
  }
  GSR_REQUIRE(num_tiles_hit && cum_tiles_hit && workspace, GSR_ERR_INVALID_ARGUMENT, "cumsum_tiles_hit: null pointer");
  const size_t need = gsr_cumsum_workspace_bytes(num_points);
  GSR_REQUIRE(workspace_bytes >= need, GSR_ERR_WORKSPACE, "cumsum_tiles_hit: workspace %zu < %zu bytes",
              workspace_bytes, need);
  GSR_REQUIRE((uintptr_t)workspace % 4 == 0, GSR_ERR_INVALID_ARGUMENT, "cumsum_tiles_hit: misaligned workspace");
  const int rc = rs::inclusive_scan(num_points, nullptr, num_tiles_hit, cum_tiles_hit, 0x7fffffff, nullptr, workspace,
                                    (cudaStream_t)stream);
  if (rc != GSR_OK) return rc;
  if (total_host_pinned)
    GSR_CUDA(cudaMemcpyAsync(total_host_pinned, cum_tiles_hit + (num_points - 1), sizeof(int32_t),
                             cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return GSR_OK;
}

GSR_API int gsr_map_gaussian_to_intersects(int num_points, int num_intersects, const float *xys,
                                           const float *depths, const int32_t *radii,
                                           const int32_t *cum_tiles_hit, unsigned tiles_x, unsigned tiles_y,
                                           unsigned block_width, int64_t *isect_ids, int32_t *gaussian_ids,
                                           void *stream) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_map_gaussian_to_intersects");
  GSR_REQUIRE(num_points >= 0 && num_intersects >= 0, GSR_ERR_INVALID_ARGUMENT, "map_gaussian_to_intersects: negative size");
  GSR_REQUIRE(block_width > 1 && block_width <= 16, GSR_ERR_INVALID_ARGUMENT,
              "block_width must be between 2 and 16 (got %u)", block_width);
  if (num_points == 0 || num_intersects == 0) return GSR_OK;
  GSR_REQUIRE(xys && depths && radii && cum_tiles_hit && isect_ids && gaussian_ids, GSR_ERR_INVALID_ARGUMENT,
              "map_gaussian_to_intersects: null pointer");
  GSR_REQUIRE((uintptr_t)xys % 8 == 0, GSR_ERR_INVALID_ARGUMENT, "map_gaussian_to_intersects: xys must be 8-byte aligned");
  map_intersects_kernel<<<cdiv(num_points, BIN_THREADS), BIN_THREADS, 0, (cudaStream_t)stream>>>(
      num_points, reinterpret_cast<const float2 *>(xys), depths, radii, cum_tiles_hit, (int)tiles_x, (int)tiles_y,
      (int)block_width, isect_ids, gaussian_ids);
  GSR_CHECK_LAUNCH("map_intersects_kernel");
  return GSR_OK;
}

GSR_API int gsr_count_tiles_tight(int num_points, const float *xys, const int32_t *radii, const float *conics,
                                  const float *opacities, unsigned img_height, unsigned img_width,
                                  unsigned block_width, int32_t *tiles_touched, void *stream) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_count_tiles_tight");
  GSR_REQUIRE(num_points >= 0, GSR_ERR_INVALID_ARGUMENT, "count_tiles_tight: num_points < 0");
  GSR_REQUIRE(block_width > 1 && block_width <= 16, GSR_ERR_INVALID_ARGUMENT,
              "block_width must be between 2 and 16 (got %u)", block_width);
  if (num_points == 0) return GSR_OK;
  GSR_REQUIRE(xys && radii && conics && opacities && tiles_touched, GSR_ERR_INVALID_ARGUMENT, "count_tiles_tight: null pointer");
  GSR_REQUIRE((uintptr_t)xys % 8 == 0, GSR_ERR_INVALID_ARGUMENT, "count_tiles_tight: xys must be 8-byte aligned");
  tight_tiles_kernel<false><<<cdiv(num_points, BIN_THREADS), BIN_THREADS, 0, (cudaStream_t)stream>>>(
      num_points, reinterpret_cast<const float2 *>(xys), nullptr, radii, conics, opacities, nullptr,
      (int)cdiv(img_width, block_width), (int)cdiv(img_height, block_width), (int)block_width, (int)img_width,
      (int)img_height, tiles_touched, nullptr, nullptr);
  GSR_CHECK_LAUNCH("tight_tiles_kernel<count>");
  return GSR_OK;
}

GSR_API int gsr_map_gaussian_to_intersects_tight(int num_points, int num_intersects, const float *xys,
                                                 const float *depths, const int32_t *radii, const float *conics,
                                                 const float *opacities, const int32_t *cum_tiles_touched,
                                                 unsigned img_height, unsigned img_width, unsigned block_width,
                                                 int64_t *isect_ids, int32_t *gaussian_ids, void *stream) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_map_gaussian_to_intersects_tight");
  GSR_REQUIRE(num_points >= 0 && num_intersects >= 0, GSR_ERR_INVALID_ARGUMENT, "map_gaussian_to_intersects_tight: negative size");
  GSR_REQUIRE(block_width > 1 && block_width <= 16, GSR_ERR_INVALID_ARGUMENT,
              "block_width must be between 2 and 16 (got %u)", block_width);
  if (num_points == 0 || num_intersects == 0) return GSR_OK;
  GSR_REQUIRE(xys && depths && radii && conics && opacities && cum_tiles_touched && isect_ids && gaussian_ids,
              GSR_ERR_INVALID_ARGUMENT, "map_gaussian_to_intersects_tight: null pointer");
  GSR_REQUIRE((uintptr_t)xys % 8 == 0, GSR_ERR_INVALID_ARGUMENT, "map_gaussian_to_intersects_tight: xys must be 8-byte aligned");
  tight_tiles_kernel<true><<<cdiv(num_points, BIN_THREADS), BIN_THREADS, 0, (cudaStream_t)stream>>>(
      num_points, reinterpret_cast<const float2 *>(xys), depths, radii, conics, opacities, cum_tiles_touched,
      (int)cdiv(img_width, block_width), (int)cdiv(img_height, block_width), (int)block_width, (int)img_width,
      (int)img_height, nullptr, isect_ids, gaussian_ids);
  GSR_CHECK_LAUNCH("tight_tiles_kernel<emit>");
  return GSR_OK;
}

GSR_API size_t gsr_sort_workspace_bytes(int num_intersects) {
  const size_t m = num_intersects > 0 ? num_intersects : 1;
  // ping-pong buffers for the keys (8 B) and ids (4 B) + histogram workspace
  return ((8 * m + 255) & ~(size_t)255) + ((4 * m + 255) & ~(size_t)255) + gsr::rs::workspace_bytes((int)m) + 512;
}

GSR_API int gsr_sort_intersects(int num_intersects, int num_tiles, const int64_t *isect_ids,
                                const int32_t *gaussian_ids, int64_t *isect_ids_sorted,
                                int32_t *gaussian_ids_sorted, void *workspace, size_t workspace_bytes,
                                void *stream) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_sort_intersects");
  GSR_REQUIRE(num_intersects >= 0 && num_tiles > 0, GSR_ERR_INVALID_ARGUMENT, "sort_intersects: bad sizes");
  if (num_intersects == 0) return GSR_OK;
  GSR_REQUIRE(isect_ids && gaussian_ids && isect_ids_sorted && gaussian_ids_sorted && workspace,
              GSR_ERR_INVALID_ARGUMENT, "sort_intersects: null pointer");
  const size_t need = gsr_sort_workspace_bytes(num_intersects);
  GSR_REQUIRE(workspace_bytes >= need, GSR_ERR_WORKSPACE, "sort_intersects: workspace %zu < %zu bytes",
              workspace_bytes, need);
  GSR_REQUIRE((uintptr_t)workspace % 8 == 0, GSR_ERR_INVALID_ARGUMENT, "sort_intersects: misaligned workspace");
  const int end_bit = key_end_bit(num_tiles);
  const size_t m = num_intersects;
  char *ws = (char *)workspace;
  unsigned long long *keys_tmp = (unsigned long long *)ws; ws += (8 * m + 255) & ~(size_t)255;
  int *vals_tmp = (int *)ws;                               ws += (4 * m + 255) & ~(size_t)255;
  void *sort_ws = ws;
  // pass 0 reads the caller's (const) arrays; the last pass must write the caller's output arrays
  const bool odd = rs::num_passes(0, end_bit) & 1;
  unsigned long long *out_k = reinterpret_cast<unsigned long long *>(isect_ids_sorted);
  unsigned long long *kx = odd ? out_k : keys_tmp, *ky = odd ? keys_tmp : out_k;
  int *vx = odd ? gaussian_ids_sorted : vals_tmp, *vy = odd ? vals_tmp : gaussian_ids_sorted;
  return rs::sort_pairs_from<unsigned long long, int>(reinterpret_cast<const unsigned long long *>(isect_ids), gaussian_ids,
                                                      kx, vx, ky, vy, num_intersects, nullptr, 0, end_bit, sort_ws,
                                                      (cudaStream_t)stream);
}

GSR_API int gsr_get_tile_bin_edges(int num_intersects, const int64_t *isect_ids_sorted, int num_tiles,
                                   int32_t *tile_bins, void *stream) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_get_tile_bin_edges");
  GSR_REQUIRE(num_intersects >= 0 && num_tiles >= 0, GSR_ERR_INVALID_ARGUMENT, "get_tile_bin_edges: negative size");
  if (num_tiles == 0) return GSR_OK;
  GSR_REQUIRE(tile_bins, GSR_ERR_INVALID_ARGUMENT, "get_tile_bin_edges: null pointer");
  GSR_REQUIRE((uintptr_t)tile_bins % 8 == 0, GSR_ERR_INVALID_ARGUMENT, "get_tile_bin_edges: tile_bins must be 8-byte aligned");
  GSR_CUDA(cudaMemsetAsync(tile_bins, 0, sizeof(int32_t) * 2 * (size_t)num_tiles, (cudaStream_t)stream));
  if (num_intersects == 0) return GSR_OK;
  GSR_REQUIRE(isect_ids_sorted, GSR_ERR_INVALID_ARGUMENT, "get_tile_bin_edges: null pointer");
  tile_bin_edges_kernel<<<cdiv(num_intersects, BIN_THREADS), BIN_THREADS, 0, (cudaStream_t)stream>>>(
      num_intersects, isect_ids_sorted, reinterpret_cast<int2 *>(tile_bins));
  GSR_CHECK_LAUNCH("tile_bin_edges_kernel");
  return GSR_OK;
}
}
