// binning_device.cu — the tile binning of rasterize_gaussians as ONE asynchronous call with a DEVICE-side pair count.
//
// Same result as gsr_bin_prepare + gsr_bin_emit_sort (binning_fast.cu) — and therefore the same per-tile order as the
// reference's cumsum -> .item() -> map_gaussian_to_intersects -> torch.sort(int64) -> torch.gather ->
// get_tile_bin_edges (rasterizer/rasterize.py:106-138, utils.py:106-182), with exact tile culling — but
//   * the number of (Gaussian, tile) pairs M never travels to the host: every kernel after the scan reads it from device
//     memory and runs on a grid sized for the caller's CAPACITY; the call neither synchronises nor allocates, so a view
//     (and a whole training iteration) can be captured in a CUDA graph.  If M exceeds the capacity the pair list is
//     truncated (the farthest Gaussians are dropped), meta[1] is set and the caller learns it asynchronously;
//   * both sorts are the hand-written radix sort of radix_sort.cuh (no library sort), the scan is hand-written too;
//   * the depth keys are produced by the kernel that counts the tiles (one pass over the projected Gaussians).
//
//   prep_kernel        per Gaussian: depth key (IEEE bits of the positive depth; culled -> 0xffffffff), id,
//                      number of reachable tiles + 64-bit mask of them (tile_cull.cuh)
//   rs::sort_pairs     Gaussians by depth, 32 bits = 4 passes (stable: ties stay in index order, like the reference's
//                      stable 64-bit sort)
//   perm scan          cum[j] = sum_{i<=j} counts[perm[i]]: tile sums -> one-block scan of the sums (+ M, overflow flag)
//                      -> in-tile scans
//   emit_kernel        (tile id, Gaussian id) pairs in depth order, clipped to the capacity
//   rs::sort_pairs     pairs by tile id, ceil(log2 T) bits = 2 passes for T <= 65536, count read from the device
//   bin_edges_kernel   tile_bins from the sorted tile ids, count read from the device
#include "common.cuh"
#include "radix_sort.cuh"
#include "tile_cull.cuh"

namespace gsr {
namespace {

constexpr int BD_THREADS = 256;

__global__ void __launch_bounds__(BD_THREADS)
prep_kernel(int n, const float2 *__restrict__ xys, const float *__restrict__ depths, const int *__restrict__ radii,
            const float *__restrict__ conics, const float *__restrict__ opacities, int tiles_x, int tiles_y,
            int block_width, unsigned *__restrict__ keys, int *__restrict__ ids, int *__restrict__ counts,
            unsigned long long *__restrict__ masks) {
  const int g = blockIdx.x * BD_THREADS + threadIdx.x;
  if (g >= n) return;
  const int r = radii[g];
  int count = 0;
  unsigned long long mask = 0ull;
  unsigned key = 0xffffffffu;
  if (r > 0) {
    key = (unsigned)__float_as_int(depths[g]);  // the low 32 bits of the reference key (forward.cu:116)
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
        if (cnt > 0 && area <= 64) {  // larger boxes are re-derived row by row when emitting
          const int k0 = (i - y0) * bw_tiles + (j0 - x0);
          mask |= ((cnt >= 64) ? ~0ull : ((1ull << cnt) - 1ull)) << k0;
        }
        count += cnt;
      }
    }
  }
  keys[g] = key;
  ids[g] = g;
  counts[g] = count;
  masks[g] = mask;
}

__global__ void __launch_bounds__(BD_THREADS)
emit_capped_kernel(int n, const int *__restrict__ perm, const float2 *__restrict__ xys, const int *__restrict__ radii,
                   const float *__restrict__ conics, const float *__restrict__ opacities, const int *__restrict__ cum,
                   const unsigned long long *__restrict__ masks, int tiles_x, int tiles_y, int block_width, int capacity,
                   unsigned *__restrict__ tile_keys, int *__restrict__ gaussian_ids) {
  const int j = blockIdx.x * BD_THREADS + threadIdx.x;
  if (j >= n) return;
  const int end = min(cum[j], capacity);
  int cur = (j == 0) ? 0 : cum[j - 1];
  if (cur >= end) return;
  const int g = perm[j];
  const float2 ctr = xys[g];
  int x0, y0, x1, y1;
  tile_bbox(ctr.x, ctr.y, (float)radii[g], tiles_x, tiles_y, block_width, x0, y0, x1, y1);
  const int bw_tiles = x1 - x0, area = bw_tiles * (y1 - y0);
  if (cum[j] - cur == area) {  // every tile of the box is kept
    for (int i = y0; i < y1; ++i)
      for (int jx = x0; jx < x1 && cur < end; ++jx) {
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
    while (m && cur < end) {
      const int k = __ffsll((long long)m) - 1;
      m &= m - 1;
      const int i = y0 + k / bw_tiles, jx = x0 + k % bw_tiles;
      tile_keys[cur] = (unsigned)(i * tiles_x + jx);
      gaussian_ids[cur] = g;
      ++cur;
    }
  }
}

__global__ void __launch_bounds__(BD_THREADS)
bin_edges_dev_kernel(int capacity, const int *__restrict__ m_dev, const unsigned *__restrict__ keys,
                     int2 *__restrict__ tile_bins) {
  const int m = m_dev ? min(capacity, *m_dev) : capacity;
  const int idx = blockIdx.x * BD_THREADS + threadIdx.x;
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

inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }
inline int bits_for_tiles(int n) {
  int bits = 1;
  while ((1ll << bits) < (long long)n) ++bits;
  return bits;
}

}  // namespace
}  // namespace gsr

extern "C" {

// ---- asynchronous form: device-side M, capacity-bounded pair buffers ------------------------------------------------
GSR_API size_t gsr_bin_device_workspace_bytes(int num_points, int capacity) {
  using namespace gsr;
  const size_t n = num_points > 0 ? num_points : 1, c = capacity > 0 ? capacity : 1;
  // per Gaussian: depth keys a/b, ids a/b, counts, cum (4 B each) + masks (8 B); per pair: tile keys a/b, ids tmp (4 B each)
  return 6 * al256(4 * n) + al256(8 * n) + al256(rs::scan_workspace_bytes((int)n)) + 3 * al256(4 * c) +
         al256(rs::workspace_bytes((int)(n > c ? n : c))) + 256;
}

GSR_API int gsr_bin_gaussians_device(int num_points, const float *xys, const float *depths, const int32_t *radii,
                                     const float *conics, const float *opacities, unsigned img_height,
                                     unsigned img_width, unsigned block_width, int capacity,
                                     int32_t *gaussian_ids_sorted, int32_t *tile_bins, int32_t *meta,
                                     int32_t *meta_host_pinned, void *workspace, size_t workspace_bytes, void *stream) {
  using namespace gsr;
  GSR_REQUIRE(num_points >= 0 && capacity >= 0, GSR_ERR_INVALID_ARGUMENT, "bin_gaussians_device: negative size");
  GSR_REQUIRE(block_width > 1 && block_width <= 16, GSR_ERR_INVALID_ARGUMENT,
              "block_width must be between 2 and 16 (got %u)", block_width);
  GSR_REQUIRE(tile_bins && meta, GSR_ERR_INVALID_ARGUMENT, "bin_gaussians_device: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles_x = cdiv(img_width, block_width), tiles_y = cdiv(img_height, block_width);
  const int num_tiles = tiles_x * tiles_y;
  GSR_CUDA(cudaMemsetAsync(tile_bins, 0, sizeof(int32_t) * 2 * (size_t)num_tiles, st));
  if (num_points == 0 || capacity == 0) {
    GSR_CUDA(cudaMemsetAsync(meta, 0, 4 * sizeof(int32_t), st));
    if (meta_host_pinned) GSR_CUDA(cudaMemcpyAsync(meta_host_pinned, meta, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    return GSR_OK;
  }
  GSR_REQUIRE(xys && depths && radii && conics && opacities && gaussian_ids_sorted && workspace,
              GSR_ERR_INVALID_ARGUMENT, "bin_gaussians_device: null pointer");
  GSR_REQUIRE((uintptr_t)xys % 8 == 0 && (uintptr_t)workspace % 16 == 0 && (uintptr_t)tile_bins % 8 == 0,
              GSR_ERR_INVALID_ARGUMENT, "bin_gaussians_device: misaligned pointer");
  GSR_REQUIRE(workspace_bytes >= gsr_bin_device_workspace_bytes(num_points, capacity), GSR_ERR_WORKSPACE,
              "bin_gaussians_device: workspace %zu < %zu bytes", workspace_bytes,
              gsr_bin_device_workspace_bytes(num_points, capacity));
  const size_t n = num_points, c = capacity;
  char *ws = (char *)workspace;
  unsigned *dkeys_a = (unsigned *)ws; ws += al256(4 * n);
  unsigned *dkeys_b = (unsigned *)ws; ws += al256(4 * n);
  int *ids_a = (int *)ws;             ws += al256(4 * n);
  int *ids_b = (int *)ws;             ws += al256(4 * n);
  int *counts = (int *)ws;            ws += al256(4 * n);
  int *cum = (int *)ws;               ws += al256(4 * n);
  unsigned long long *masks = (unsigned long long *)ws; ws += al256(8 * n);
  void *scan_ws = ws;                 ws += al256(rs::scan_workspace_bytes(num_points));
  unsigned *tkeys_a = (unsigned *)ws; ws += al256(4 * c);
  unsigned *tkeys_b = (unsigned *)ws; ws += al256(4 * c);
  int *pids_tmp = (int *)ws;          ws += al256(4 * c);
  void *sort_ws = ws;

  const unsigned grid_n = cdiv(num_points, BD_THREADS);
  prep_kernel<<<grid_n, BD_THREADS, 0, st>>>(num_points, reinterpret_cast<const float2 *>(xys), depths, radii, conics,
                                             opacities, tiles_x, tiles_y, (int)block_width, dkeys_a, ids_a, counts, masks);
  GSR_CHECK_LAUNCH("prep_kernel");
  // Gaussians by depth (32 bits = 4 passes: back in the A buffers)
  int rc = rs::sort_pairs<unsigned, int>(dkeys_a, ids_a, dkeys_b, ids_b, num_points, nullptr, 0, 32, sort_ws, st);
  if (rc != GSR_OK) return rc;
  const int *perm = (rs::num_passes(0, 32) & 1) ? ids_b : ids_a;
  rc = rs::inclusive_scan(num_points, perm, counts, cum, capacity, meta, scan_ws, st);
  if (rc != GSR_OK) return rc;
  if (meta_host_pinned)
    GSR_CUDA(cudaMemcpyAsync(meta_host_pinned, meta, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  // pairs by tile id: the result must END in gaussian_ids_sorted
  const int tile_bits = bits_for_tiles(num_tiles);
  const bool odd = rs::num_passes(0, tile_bits) & 1;
  int *pids_first = odd ? pids_tmp : gaussian_ids_sorted;  // the buffer the emit kernel writes
  int *pids_second = odd ? gaussian_ids_sorted : pids_tmp;
  emit_capped_kernel<<<grid_n, BD_THREADS, 0, st>>>(num_points, perm, reinterpret_cast<const float2 *>(xys), radii, conics,
                                                    opacities, cum, masks, tiles_x, tiles_y, (int)block_width, capacity,
                                                    tkeys_a, pids_first);
  GSR_CHECK_LAUNCH("emit_capped_kernel");
  rc = rs::sort_pairs<unsigned, int>(tkeys_a, pids_first, tkeys_b, pids_second, capacity, meta + 2, 0, tile_bits, sort_ws, st);
  if (rc != GSR_OK) return rc;
  const unsigned *sorted_keys = odd ? tkeys_b : tkeys_a;
  bin_edges_dev_kernel<<<cdiv(capacity, BD_THREADS), BD_THREADS, 0, st>>>(capacity, meta + 2, sorted_keys,
                                                                          reinterpret_cast<int2 *>(tile_bins));
  GSR_CHECK_LAUNCH("bin_edges_dev_kernel");
  return GSR_OK;
}

// ---- two-call form with a host-visible M (the caller synchronises between the calls and sizes the outputs exactly) ----
GSR_API size_t gsr_bin_prepare_workspace_bytes(int num_points) {
  using namespace gsr;
  const size_t n = num_points > 0 ? num_points : 1;
  // depth keys a/b, ids b, counts + scan + sort workspaces + meta
  return 4 * al256(4 * n) + al256(rs::scan_workspace_bytes((int)n)) + al256(rs::workspace_bytes((int)n)) + 512;
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
  unsigned *dkeys_a = (unsigned *)ws; ws += al256(4 * n);
  unsigned *dkeys_b = (unsigned *)ws; ws += al256(4 * n);
  int *ids_b = (int *)ws;             ws += al256(4 * n);
  int *counts = (int *)ws;            ws += al256(4 * n);
  void *scan_ws = ws;                 ws += al256(rs::scan_workspace_bytes(num_points));
  int *meta = (int *)ws;              ws += 256;
  void *sort_ws = ws;
  const int tiles_x = cdiv(img_width, block_width), tiles_y = cdiv(img_height, block_width);
  static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "mask type");
  prep_kernel<<<cdiv(num_points, BD_THREADS), BD_THREADS, 0, st>>>(
      num_points, reinterpret_cast<const float2 *>(xys), depths, radii, conics, opacities, tiles_x, tiles_y,
      (int)block_width, dkeys_a, perm, counts, reinterpret_cast<unsigned long long *>(masks));
  GSR_CHECK_LAUNCH("prep_kernel");
  // 4 passes: the sorted ids end in the A buffer = the caller's `perm`
  int rc = rs::sort_pairs<unsigned, int>(dkeys_a, perm, dkeys_b, ids_b, num_points, nullptr, 0, 32, sort_ws, st);
  if (rc != GSR_OK) return rc;
  static_assert(((32 + 7) / 8) % 2 == 0, "depth sort must end in the A buffers");
  rc = rs::inclusive_scan(num_points, perm, counts, cum_tiles, 0x7fffffff, meta, scan_ws, st);
  if (rc != GSR_OK) return rc;
  if (total_host_pinned)
    GSR_CUDA(cudaMemcpyAsync(total_host_pinned, meta, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  return GSR_OK;
}

GSR_API size_t gsr_bin_emit_workspace_bytes(int num_intersects) {
  using namespace gsr;
  const size_t m = num_intersects > 0 ? num_intersects : 1;
  return 3 * al256(4 * m) + al256(rs::workspace_bytes((int)m)) + 256;  // tile keys a/b, ids tmp + sort workspace
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
  unsigned *tkeys_a = (unsigned *)ws; ws += al256(4 * m);
  unsigned *tkeys_b = (unsigned *)ws; ws += al256(4 * m);
  int *pids_tmp = (int *)ws;          ws += al256(4 * m);
  void *sort_ws = ws;
  const int tile_bits = bits_for_tiles(num_tiles);
  const bool odd = rs::num_passes(0, tile_bits) & 1;
  int *pids_first = odd ? pids_tmp : gaussian_ids_sorted;
  int *pids_second = odd ? gaussian_ids_sorted : pids_tmp;
  emit_capped_kernel<<<cdiv(num_points, BD_THREADS), BD_THREADS, 0, st>>>(
      num_points, perm, reinterpret_cast<const float2 *>(xys), radii, conics, opacities, cum_tiles,
      reinterpret_cast<const unsigned long long *>(masks), tiles_x, tiles_y, (int)block_width, num_intersects, tkeys_a,
      pids_first);
  GSR_CHECK_LAUNCH("emit_capped_kernel");
  int rc = rs::sort_pairs<unsigned, int>(tkeys_a, pids_first, tkeys_b, pids_second, num_intersects, nullptr, 0, tile_bits,
                                         sort_ws, st);
  if (rc != GSR_OK) return rc;
  bin_edges_dev_kernel<<<cdiv(num_intersects, BD_THREADS), BD_THREADS, 0, st>>>(
      num_intersects, nullptr, odd ? tkeys_b : tkeys_a, reinterpret_cast<int2 *>(tile_bins));
  GSR_CHECK_LAUNCH("bin_edges_dev_kernel");
  return GSR_OK;
}
}
