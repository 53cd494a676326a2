// binning_fast.cu — the internal tile-binning path of rasterize_gaussians (no counterpart binding in the
// reference; it replaces the SEQUENCE cumsum -> map_gaussian_to_intersects -> torch.sort(int64) ->
// torch.gather -> get_tile_bin_edges of rasterizer/rasterize.py:106-138 + utils.py:106-182 as a whole).
//
// The reference sorts M 64-bit (tile | depth) keys with an int64 permutation payload: 8 radix passes over
// 16-byte pairs.  The order it defines is: tile ascending, then depth ascending, ties in Gaussian-index order
// (CUB's radix sort is stable).  The same order is produced here with far less traffic by a two-level sort:
//   1. sort the N Gaussians once by depth (32-bit keys, stable => ties stay in index order);
//   2. count, per Gaussian, the tiles of its bounding box that can be reached with
//      alpha >= 1/255 (exact tile culling, see binning.cu) and remember them in a 64-bit mask;
//   3. inclusive scan of the counts read through the depth permutation -> offsets in depth order; the total M
//      goes to a pinned host word;
//   4. emit (tile id, Gaussian id) pairs in depth order;
//   5. STABLE radix sort of the pairs by tile id only: ceil(log2(T)) bits = 2 passes over 8-byte pairs;
//   6. bin edges.
// All of it is HBM-bound integer work; see DESIGN.md for the byte counts.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <thrust/iterator/permutation_iterator.h>

#include "common.cuh"
#include "tile_cull.cuh"

namespace gsr {

constexpr int FB_THREADS = 256;

// depth keys: visible Gaussians sort by the IEEE bits of their (positive) depth, exactly the low 32 bits of the
// reference key (forward.cu:116); culled ones go to the end
__global__ void __launch_bounds__(FB_THREADS)
depth_keys_kernel(int n, const float *__restrict__ depths, const int *__restrict__ radii,
                  unsigned *__restrict__ keys, int *__restrict__ ids) {
  const int i = blockIdx.x * FB_THREADS + threadIdx.x;
  if (i >= n) return;
  keys[i] = radii[i] > 0 ? (unsigned)__float_as_int(depths[i]) : 0xffffffffu;
  ids[i] = i;
}

// per Gaussian g (original order => coalesced loads): number of reachable tiles and their mask (bit k = k-th tile
// of the bounding box in row-major order; only meaningful for boxes of at most 64 tiles — larger boxes are
// re-tested tile by tile when emitting).
__global__ void __launch_bounds__(FB_THREADS)
count_tiles_kernel(int n, const float2 *__restrict__ xys, const int *__restrict__ radii,
                   const float *__restrict__ conics, const float *__restrict__ opacities, int tiles_x, int tiles_y,
                   int block_width, int img_w, int img_h, int *__restrict__ counts,
                   unsigned long long *__restrict__ masks) {
  const int g = blockIdx.x * FB_THREADS + threadIdx.x;
  if (g >= n) return;
  const int r = radii[g];
  int count = 0;
  unsigned long long mask = 0ull;
  if (r > 0) {
    const float2 ctr = xys[g];
    int x0, y0, x1, y1;
    tile_bbox(ctr.x, ctr.y, (float)r, tiles_x, tiles_y, block_width, x0, y0, x1, y1);
    const int bw_tiles = x1 - x0, area = bw_tiles * (y1 - y0);
    const CullEllipse e = make_cull_ellipse(conics[3 * (size_t)g], conics[3 * (size_t)g + 1], conics[3 * (size_t)g + 2],
                                            opacities[g]);
    if (e.never_cull) {
      count = area;
      mask = ~0ull;
    } else if (!e.empty) {
      for (int i = y0; i < y1; ++i) {
        int j0, j1;
        cull_row_range(e, ctr.x, ctr.y, i, x0, x1, block_width, j0, j1);
        const int cnt = j1 - j0;
        if (cnt > 0 && area <= 64) {  // boxes with more than 64 tiles are re-derived row by row by the emit kernel
          const int k0 = (i - y0) * bw_tiles + (j0 - x0);
          mask |= ((cnt >= 64) ? ~0ull : ((1ull << cnt) - 1ull)) << k0;
        }
        count += cnt;
      }
    }
  }
  counts[g] = count;
  masks[g] = mask;
}

__global__ void __launch_bounds__(FB_THREADS)
emit_sorted_kernel(int n, const int *__restrict__ perm, const float2 *__restrict__ xys,
                   const int *__restrict__ radii, const float *__restrict__ conics,
                   const float *__restrict__ opacities, const int *__restrict__ cum,
                   const unsigned long long *__restrict__ masks, int tiles_x, int tiles_y, int block_width,
                   int img_w, int img_h, unsigned *__restrict__ tile_keys, int *__restrict__ gaussian_ids) {
  const int j = blockIdx.x * FB_THREADS + threadIdx.x;
  if (j >= n) return;
  const int end = cum[j];
  int cur = (j == 0) ? 0 : cum[j - 1];
  if (cur == end) return;
  const int g = perm[j];
  const float2 ctr = xys[g];
  int x0, y0, x1, y1;
  tile_bbox(ctr.x, ctr.y, (float)radii[g], tiles_x, tiles_y, block_width, x0, y0, x1, y1);
  const int bw_tiles = x1 - x0, area = bw_tiles * (y1 - y0);
  if (end - cur == area) {  // every tile of the box is kept
    for (int i = y0; i < y1; ++i)
      for (int jx = x0; jx < x1; ++jx) {
        tile_keys[cur] = (unsigned)(i * tiles_x + jx);
        gaussian_ids[cur] = g;
        ++cur;
      }
  } else if (area > 64) {  // big box, partially culled: repeat the count kernel's row ranges
    const CullEllipse e = make_cull_ellipse(conics[3 * (size_t)g], conics[3 * (size_t)g + 1], conics[3 * (size_t)g + 2],
                                            opacities[g]);
    for (int i = y0; i < y1; ++i) {
      int j0, j1;
      cull_row_range(e, ctr.x, ctr.y, i, x0, x1, block_width, j0, j1);
      for (int jx = j0; jx < j1 && cur < end; ++jx) {
        tile_keys[cur] = (unsigned)(i * tiles_x + jx);
        gaussian_ids[cur] = g;
        ++cur;
      }
    }
  } else {
    unsigned long long m = masks[g];
    while (m) {
      const int k = __ffsll((long long)m) - 1;
      m &= m - 1;
      const int i = y0 + k / bw_tiles, jx = x0 + k % bw_tiles;
      tile_keys[cur] = (unsigned)(i * tiles_x + jx);
      gaussian_ids[cur] = g;
      ++cur;
    }
  }
}

__global__ void __launch_bounds__(FB_THREADS)
bin_edges_u32_kernel(int m, const unsigned *__restrict__ keys, int2 *__restrict__ tile_bins) {
  const int idx = blockIdx.x * FB_THREADS + threadIdx.x;
  if (idx >= m) return;
  const int cur = (int)keys[idx];
  if (idx == 0) tile_bins[cur].x = 0;
  if (idx == m - 1) tile_bins[cur].y = m;
  if (idx == 0) return;
  const int prev = (int)keys[idx - 1];
  if (prev != cur) {
    tile_bins[prev].y = idx;
    tile_bins[cur].x = idx;
  }
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
static inline int bits_for(int n) {
  int bits = 1;
  while ((1ll << bits) < (long long)n) ++bits;
  return bits;
}

}  // namespace gsr

extern "C" {

GSR_API size_t gsr_bin_prepare_workspace_bytes(int num_points) {
  using namespace gsr;
  const size_t n = num_points > 0 ? num_points : 1;
  size_t sort_b = 0, scan_b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_b, (const unsigned *)nullptr, (unsigned *)nullptr,
                                  (const int *)nullptr, (int *)nullptr, (int)n, 0, 32);
  cub::DeviceScan::InclusiveSum(nullptr, scan_b,
                                thrust::make_permutation_iterator((const int *)nullptr, (const int *)nullptr),
                                (int *)nullptr, (int)n);
  // keys_in, keys_out, ids_in, counts + cub temp
  return 4 * align256(4 * n) + align256(sort_b > scan_b ? sort_b : scan_b) + 256;
}

GSR_API int gsr_bin_prepare(int num_points, const float *xys, const float *depths, const int32_t *radii,
                            const float *conics, const float *opacities, unsigned img_height,
                            unsigned img_width, unsigned block_width, int32_t *perm, int32_t *cum_tiles,
                            uint64_t *masks, int32_t *total_host_pinned, void *workspace, size_t workspace_bytes,
                            void *stream) {
  using namespace gsr;
  GSR_REQUIRE(num_points >= 0, GSR_ERR_INVALID_ARGUMENT, "bin_prepare: num_points < 0");
  GSR_REQUIRE(block_width > 1 && block_width <= 16, GSR_ERR_INVALID_ARGUMENT,
              "block_width must be between 2 and 16 (got %u)", block_width);
  if (num_points == 0) {
    if (total_host_pinned) *total_host_pinned = 0;
    return GSR_OK;
  }
  GSR_REQUIRE(xys && depths && radii && conics && opacities && perm && cum_tiles && masks && workspace,
              GSR_ERR_INVALID_ARGUMENT, "bin_prepare: null pointer");
  GSR_REQUIRE((uintptr_t)xys % 8 == 0 && (uintptr_t)masks % 8 == 0 && (uintptr_t)workspace % 16 == 0,
              GSR_ERR_INVALID_ARGUMENT, "bin_prepare: misaligned pointer");
  GSR_REQUIRE(workspace_bytes >= gsr_bin_prepare_workspace_bytes(num_points), GSR_ERR_WORKSPACE,
              "bin_prepare: workspace %zu < %zu bytes", workspace_bytes, gsr_bin_prepare_workspace_bytes(num_points));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = num_points;
  char *ws = (char *)workspace;
  unsigned *keys_in = (unsigned *)ws;  ws += align256(4 * n);
  unsigned *keys_out = (unsigned *)ws; ws += align256(4 * n);
  int *ids_in = (int *)ws;             ws += align256(4 * n);
  int *counts = (int *)ws;             ws += align256(4 * n);
  void *cub_ws = ws;
  size_t cub_bytes = workspace_bytes - (size_t)(ws - (char *)workspace);
  const unsigned grid = cdiv(num_points, FB_THREADS);
  depth_keys_kernel<<<grid, FB_THREADS, 0, st>>>(num_points, depths, radii, keys_in, ids_in);
  GSR_CHECK_LAUNCH("depth_keys_kernel");
  GSR_CUDA(cub::DeviceRadixSort::SortPairs(cub_ws, cub_bytes, keys_in, keys_out, ids_in, perm, num_points, 0, 32, st));
  const int tiles_x = cdiv(img_width, block_width), tiles_y = cdiv(img_height, block_width);
  count_tiles_kernel<<<grid, FB_THREADS, 0, st>>>(num_points, reinterpret_cast<const float2 *>(xys), radii, conics,
                                                  opacities, tiles_x, tiles_y, (int)block_width, (int)img_width,
                                                  (int)img_height, counts,
                                                  reinterpret_cast<unsigned long long *>(masks));
  GSR_CHECK_LAUNCH("count_tiles_kernel");
  // inclusive scan of the counts read THROUGH the depth permutation: cum[j] = sum_{i<=j} counts[perm[i]]
  cub_bytes = workspace_bytes - (size_t)(ws - (char *)workspace);
  GSR_CUDA(cub::DeviceScan::InclusiveSum(cub_ws, cub_bytes,
                                         thrust::make_permutation_iterator((const int *)counts, (const int *)perm),
                                         cum_tiles, num_points, st));
  if (total_host_pinned)
    GSR_CUDA(cudaMemcpyAsync(total_host_pinned, cum_tiles + (num_points - 1), sizeof(int32_t),
                             cudaMemcpyDeviceToHost, st));
  return GSR_OK;
}

GSR_API size_t gsr_bin_emit_workspace_bytes(int num_intersects) {
  using namespace gsr;
  const size_t m = num_intersects > 0 ? num_intersects : 1;
  size_t sort_b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_b, (const unsigned *)nullptr, (unsigned *)nullptr,
                                  (const int *)nullptr, (int *)nullptr, (int)m, 0, 32);
  return 3 * align256(4 * m) + align256(sort_b) + 256;  // keys_in, keys_out, ids_in + cub temp
}

GSR_API int gsr_bin_emit_sort(int num_points, int num_intersects, const float *xys, const int32_t *radii,
                              const float *conics, const float *opacities, const int32_t *perm,
                              const int32_t *cum_tiles, const uint64_t *masks,
                              unsigned img_height, unsigned img_width, unsigned block_width,
                              int32_t *gaussian_ids_sorted, int32_t *tile_bins, void *workspace,
                              size_t workspace_bytes, void *stream) {
  using namespace gsr;
  GSR_REQUIRE(num_points >= 0 && num_intersects >= 0, GSR_ERR_INVALID_ARGUMENT, "bin_emit_sort: negative size");
  GSR_REQUIRE(block_width > 1 && block_width <= 16, GSR_ERR_INVALID_ARGUMENT,
              "block_width must be between 2 and 16 (got %u)", block_width);
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles_x = cdiv(img_width, block_width), tiles_y = cdiv(img_height, block_width);
  const int num_tiles = tiles_x * tiles_y;
  GSR_REQUIRE(tile_bins, GSR_ERR_INVALID_ARGUMENT, "bin_emit_sort: null pointer");
  GSR_CUDA(cudaMemsetAsync(tile_bins, 0, sizeof(int32_t) * 2 * (size_t)num_tiles, st));
  if (num_points == 0 || num_intersects == 0) return GSR_OK;
  GSR_REQUIRE(xys && radii && conics && opacities && perm && cum_tiles && masks && gaussian_ids_sorted && workspace,
              GSR_ERR_INVALID_ARGUMENT, "bin_emit_sort: null pointer");
  GSR_REQUIRE(workspace_bytes >= gsr_bin_emit_workspace_bytes(num_intersects), GSR_ERR_WORKSPACE,
              "bin_emit_sort: workspace %zu < %zu bytes", workspace_bytes, gsr_bin_emit_workspace_bytes(num_intersects));
  const size_t m = num_intersects;
  char *ws = (char *)workspace;
  unsigned *keys_in = (unsigned *)ws;  ws += align256(4 * m);
  unsigned *keys_out = (unsigned *)ws; ws += align256(4 * m);
  int *ids_in = (int *)ws;             ws += align256(4 * m);
  void *cub_ws = ws;
  size_t cub_bytes = workspace_bytes - (size_t)(ws - (char *)workspace);
  emit_sorted_kernel<<<cdiv(num_points, FB_THREADS), FB_THREADS, 0, st>>>(
      num_points, perm, reinterpret_cast<const float2 *>(xys), radii, conics, opacities, cum_tiles,
      reinterpret_cast<const unsigned long long *>(masks), tiles_x, tiles_y, (int)block_width, (int)img_width,
      (int)img_height, keys_in, ids_in);
  GSR_CHECK_LAUNCH("emit_sorted_kernel");
  GSR_CUDA(cub::DeviceRadixSort::SortPairs(cub_ws, cub_bytes, keys_in, keys_out, ids_in, gaussian_ids_sorted,
                                           num_intersects, 0, bits_for(num_tiles), st));
  bin_edges_u32_kernel<<<cdiv(num_intersects, FB_THREADS), FB_THREADS, 0, st>>>(num_intersects, keys_out,
                                                                                reinterpret_cast<int2 *>(tile_bins));
  GSR_CHECK_LAUNCH("bin_edges_u32_kernel");
  return GSR_OK;
}
}
