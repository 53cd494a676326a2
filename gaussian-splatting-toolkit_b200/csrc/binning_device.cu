// binning_device.cu — the internal tile binning of rasterize_gaussians: per-tile depth-ordered Gaussian lists with
// exact tile culling, WITHOUT a global sort and without the host ever needing the number of pairs.
//
// The reference builds the lists by sorting M 64-bit keys (tile << 32 | depth bits) globally: cumsum -> .item() ->
// map_gaussian_to_intersects -> torch.sort(int64) -> torch.gather -> get_tile_bin_edges (rasterizer/rasterize.py:106-138,
// utils.py:106-182).  The order that defines is: per tile, depth ascending (IEEE bits of the positive depth), ties in
// Gaussian-index order.  The tile part of the key only PARTITIONS the pairs; only the order INSIDE a tile needs a
// comparison sort, and a tile holds a few hundred pairs.  So (a counting sort by tile + small independent sorts):
//
//   bin_count_blocks_kernel  block b owns a contiguous range of Gaussians; per Gaussian the tiles of its bounding box
//                      that can be reached with alpha >= 1/255 (exact tile culling, tile_cull.cuh) are counted into a
//                      SHARED-MEMORY histogram over the T tiles (packed 16-bit counters, shared atomics — no global
//                      atomics), the 64-bit tile mask is kept; the block's histogram row goes to base[b][0..T)
//   bin_colscan_kernel per tile: exclusive prefix of its column over the blocks (in place) -> offset of every block's
//                      pairs inside the tile's segment; column total -> tile_count[t]
//   tile_scan_kernel   one block: exclusive scan over the T tiles -> tile_bins [T,2] (clipped to the pair capacity),
//                      segment starts, meta = {M, overflow, min(M, capacity)}
//   bin_fill_blocks_kernel   same blocks again: slot = tile_start[t] + base[b][t] + (shared-memory cursor of the block
//                      for t)++; writes the 64-bit key (depth bits << 32 | Gaussian id) — unordered inside the
//                      (block, tile) group, every group in its own range of the tile's segment
//   tile_sort_warp_kernel    one WARP per tile with <= 1024 pairs: bitonic sort in registers (E = 4..32 keys per lane;
//                      in-register stages + shuffle stages, no shared memory, no barrier); the low words (Gaussian
//                      ids) go to gaussian_ids_sorted.  <false>: all tiles up to 512 pairs; <true>: the 513..1024-pair
//                      tiles, from a list tile_scan_kernel builds
//   tile_sort_long_kernel    tiles with more pairs (list-driven persistent grid, 1024-thread CTAs): all-ascending bitonic
//                      network with virtual padding — in shared memory up to 8192 pairs, chunked shared memory + a few
//                      in-place global steps above (any length; object-centric scenes put thousands of pairs in a tile)
//   (images with more than 48 K tiles do not fit the shared-memory histogram: bin_count_kernel / bin_fill_kernel do the
//   same with global atomics)
//
// Keys are unique (the id is part of the key), so the result is deterministic and identical to the reference's order
// although the fill order is not.  No kernel's grid depends on M: the call neither synchronises nor allocates and can
// be captured in a CUDA graph; if M exceeds the caller's capacity, meta[1] is set (the lists are truncated).
// Compared with round 1 (depth sort of N + emit + stable sort of M pairs by tile with cub::DeviceRadixSort + bin edges)
// this moves 8 M bytes of keys once instead of ~50 M and drops both library sorts.
#include <stdlib.h>

#include "common.cuh"
#include "radix_sort.cuh"
#include "tile_cull.cuh"

namespace gsr {
namespace {

inline int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (cached[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cached[dev] = n > 0 ? n : 148;
  }
  return cached[dev];
}

constexpr int BD_THREADS = 256;

// count pass of one Gaussian
template <typename F>
__device__ __forceinline__ u64 count_one(int g, const float2 *__restrict__ xys, const int *__restrict__ radii,
                                         const float *__restrict__ conics, const float *__restrict__ opacities, int tiles_x,
                                         int tiles_y, int block_width, F f) {
  const int r = radii[g];
  const float2 ctr = xys[g];
  const float ca = conics[3 * (size_t)g], cb = conics[3 * (size_t)g + 1], cc = conics[3 * (size_t)g + 2];
  const float opac = opacities[g];
  u64 mask = 0ull;
  if (r > 0) {
    int x0, y0, x1, y1;
    tile_bbox(ctr.x, ctr.y, (float)r, tiles_x, tiles_y, block_width, x0, y0, x1, y1);
    if ((x1 - x0) * (y1 - y0) > 0) cull_tiles(ctr, r, ca, cb, cc, opac, x0, y0, x1, y1, tiles_x, block_width, mask, f);
  }
  return mask;
}

// fill pass of one Gaussian: the cached mask is walked; the conic and the opacity are read only for the rare boxes of
// more than 64 tiles, whose decision is evaluated again.  (Measured and dropped: forcing radius / mask / centre / depth
// to be requested together with ld.relaxed.gpu loads — the compiler sinks plain loads behind the branches, three
// dependent round trips — made both passes SLOWER in the steady state, +9 us count / +8 us fill at cfg2, +0.18 ms at
// cfg4: the L1-bypassing loads cost more than the shorter chain saves, gpurun_out/r2_run49_*.)
template <typename F>
__device__ __forceinline__ void fill_one(int g, const float2 *__restrict__ xys, const float *__restrict__ depths,
                                         const int *__restrict__ radii, const float *__restrict__ conics,
                                         const float *__restrict__ opacities, const u64 *__restrict__ masks, int tiles_x,
                                         int tiles_y, int block_width, F f /* f(tile, key) */) {
  const int r = radii[g];
  const u64 mask = masks[g];
  const float2 ctr = xys[g];
  // the low 32 bits of the reference key are the IEEE bits of the depth (forward.cu:116); the id breaks ties in index order
  const u64 key = ((u64)(unsigned)__float_as_int(depths[g]) << 32) | (u64)(unsigned)g;
  if (r <= 0) return;
  int x0, y0, x1, y1;
  tile_bbox(ctr.x, ctr.y, (float)r, tiles_x, tiles_y, block_width, x0, y0, x1, y1);
  const int area = (x1 - x0) * (y1 - y0);
  if (area <= 0) return;
  if (area <= 64) {
    walk_tile_mask(mask, x0, y0, x1 - x0, tiles_x, [&](int tile) { f(tile, key); });
  } else {
    u64 unused;
    cull_tiles(ctr, r, conics[3 * (size_t)g], conics[3 * (size_t)g + 1], conics[3 * (size_t)g + 2], opacities[g], x0, y0, x1,
               y1, tiles_x, block_width, unused, [&](int tile) { f(tile, key); });
  }
}

__global__ void __launch_bounds__(BD_THREADS)
bin_count_kernel(int n, const float2 *__restrict__ xys, const int *__restrict__ radii, const float *__restrict__ conics,
                 const float *__restrict__ opacities, int tiles_x, int tiles_y, int block_width,
                 u64 *__restrict__ masks, unsigned *__restrict__ tile_count) {
  const int g = blockIdx.x * BD_THREADS + threadIdx.x;
  if (g >= n) return;
  masks[g] = count_one(g, xys, radii, conics, opacities, tiles_x, tiles_y, block_width,
                              [&](int tile) { atomicAdd(tile_count + tile, 1u); });
}

constexpr int SORT_WARP_SMALL = 512;  // tiles up to this size: tile_sort_warp_kernel<false> over all tiles
constexpr int SORT_WARP_MAX = 1024;   // (SORT_WARP_SMALL, SORT_WARP_MAX]: tile_sort_warp_kernel<true> over the `mid` list
                                      // above: tile_sort_long_kernel over the `long` list
// lists = {mid_count, long_count, long_cursor, 0, mid[T], long[T]}
__host__ __device__ inline int *list_mid(int *lists) { return lists + 4; }
__host__ __device__ inline int *list_long(int *lists, int num_tiles) { return lists + 4 + num_tiles; }

// one block of 1024 threads: exclusive scan of tile_count -> tile_bins (clipped to capacity), cursors = segment starts,
// meta, and the lists of the tiles whose sort needs more than the short-tile kernel.  The counts are staged through
// shared memory in coalesced chunks (a thread's serial walk over its tiles used to wait on one global load per tile).
constexpr int SCAN_CHUNK = 8192;
__global__ void __launch_bounds__(1024)
tile_scan_kernel(int num_tiles, const unsigned *__restrict__ tile_count, int capacity, int2 *__restrict__ tile_bins,
                 unsigned *__restrict__ cursors, int *__restrict__ meta, int *__restrict__ lists) {
  __shared__ unsigned s_cnt[SCAN_CHUNK];
  __shared__ unsigned long long s_warp[32];
  __shared__ int s_nmid, s_nlong;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int PER = SCAN_CHUNK / 1024;
  if (tid == 0) s_nmid = s_nlong = 0;
  unsigned long long carry = 0;  // pairs in the chunks before this one
  const unsigned long long cap = (unsigned long long)capacity;
  int *mid = list_mid(lists), *lng = list_long(lists, num_tiles);
  for (int c0 = 0; c0 < num_tiles; c0 += SCAN_CHUNK) {
    __syncthreads();  // the previous chunk's s_cnt / s_warp are no longer read
    for (int i = tid; i < SCAN_CHUNK; i += 1024) s_cnt[i] = (c0 + i < num_tiles) ? tile_count[c0 + i] : 0u;
    __syncthreads();
    unsigned v[PER];
    unsigned long long sum = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      v[i] = s_cnt[tid * PER + i];
      sum += v[i];
    }
    unsigned long long inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long up = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += up;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    unsigned long long base = 0, total = 0;
    for (int w = 0; w < 32; ++w) {
      const unsigned long long t = s_warp[w];
      if (w < warp) base += t;
      total += t;
    }
    unsigned long long run = carry + base + inc - sum;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int tile = c0 + tid * PER + i;
      const unsigned long long nxt = run + v[i];
      int len = 0;
      if (tile < num_tiles) {
        const int b = (int)min(run, cap), e = (int)min(nxt, cap);
        tile_bins[tile] = (e > b) ? make_int2(b, e) : make_int2(0, 0);  // empty tiles are (0, 0), like the reference's zeros
        cursors[tile] = (unsigned)min(run, 0xffffffffull);
        len = e - b;
      }
      // list appends: one shared-memory atomic per warp and list instead of one per tile on the same counter (at cfg4
      // nearly all 32 400 tiles are `mid` tiles)
      const bool is_long = len > SORT_WARP_MAX, is_mid = !is_long && len > SORT_WARP_SMALL;
      const unsigned m_mid = __ballot_sync(0xffffffffu, is_mid), m_long = __ballot_sync(0xffffffffu, is_long);
      const unsigned below = (1u << lane) - 1u;
      if (m_mid) {
        int at = 0;
        if (lane == 0) at = atomicAdd(&s_nmid, __popc(m_mid));
        at = __shfl_sync(0xffffffffu, at, 0);
        if (is_mid) mid[at + __popc(m_mid & below)] = tile;
      }
      if (m_long) {
        int at = 0;
        if (lane == 0) at = atomicAdd(&s_nlong, __popc(m_long));
        at = __shfl_sync(0xffffffffu, at, 0);
        if (is_long) lng[at + __popc(m_long & below)] = tile;
      }
      run = nxt;
    }
    carry += total;
  }
  __syncthreads();
  if (tid == 0) {
    // the reference keeps M in an int32 too (torch.cumsum(dtype=int32), rasterizer/utils.py:123)
    meta[0] = (int)min(carry, 0x7fffffffull);
    meta[1] = carry > cap ? 1 : 0;
    meta[2] = (int)min(carry, cap);
    meta[3] = 0;
    lists[0] = s_nmid;
    lists[1] = s_nlong;
    lists[2] = 0;
    lists[3] = 0;
  }
}

__global__ void __launch_bounds__(BD_THREADS)
bin_fill_kernel(int n, const float2 *__restrict__ xys, const float *__restrict__ depths, const int *__restrict__ radii,
                const float *__restrict__ conics, const float *__restrict__ opacities, const u64 *__restrict__ masks,
                int tiles_x, int tiles_y, int block_width, int capacity, unsigned *__restrict__ cursors,
                u64 *__restrict__ keys) {
  const int g = blockIdx.x * BD_THREADS + threadIdx.x;
  if (g >= n) return;
  fill_one(g, xys, depths, radii, conics, opacities, masks, tiles_x, tiles_y, block_width, [&](int tile, u64 key) {
    const unsigned pos = atomicAdd(cursors + tile, 1u);
    if (pos < (unsigned)capacity) keys[pos] = key;
  });
}

// ---- block-privatised counting / filling (no global atomics) --------------------------------------------------------
constexpr int SMEM_HIST_MAX_TILES = 49152;  // count: packed u16 counters (96 KB); fill: 32-bit cursors (192 KB) of dynamic smem

__device__ __forceinline__ unsigned smem_count_inc(unsigned *s_hist, int tile) {
  // packed 16-bit counters: returns the counter's value BEFORE the increment
  const unsigned old = atomicAdd(s_hist + (tile >> 1), (tile & 1) ? 0x10000u : 1u);
  return (tile & 1) ? (old >> 16) : (old & 0xffffu);
}

#ifndef GSR_BLK_THREADS
#define GSR_BLK_THREADS 1024
#endif
constexpr int BLK_THREADS = GSR_BLK_THREADS;  // few, fat blocks: one shared-memory histogram per block, full occupancy per SM

// WIDE: one 32-bit counter per tile (an increment is LEA + ATOMS instead of the ~8 instructions of the packed form) —
// used when two blocks per SM still fit (4 T bytes each); the packed 16-bit counters otherwise
template <bool WIDE>
__global__ void __launch_bounds__(BLK_THREADS)
bin_count_blocks_kernel(int n, int per_block, const float2 *__restrict__ xys, const int *__restrict__ radii,
                        const float *__restrict__ conics, const float *__restrict__ opacities, int tiles_x, int tiles_y,
                        int block_width, u64 *__restrict__ masks, unsigned *__restrict__ base /*[B][T]*/) {
  extern __shared__ unsigned s_hist[];
  const int num_tiles = tiles_x * tiles_y, words = WIDE ? num_tiles : (num_tiles + 1) >> 1;
  for (int i = threadIdx.x; i < words; i += BLK_THREADS) s_hist[i] = 0u;
  __syncthreads();
  const int g0 = blockIdx.x * per_block, g1 = min(n, g0 + per_block);
  for (int g = g0 + threadIdx.x; g < g1; g += BLK_THREADS)
    masks[g] = count_one(g, xys, radii, conics, opacities, tiles_x, tiles_y, block_width, [&](int tile) {
      if (WIDE) atomicAdd(s_hist + tile, 1u);
      else smem_count_inc(s_hist, tile);
    });
  __syncthreads();
  unsigned *row = base + (size_t)blockIdx.x * num_tiles;
  for (int t = threadIdx.x; t < num_tiles; t += BLK_THREADS)
    row[t] = WIDE ? s_hist[t] : (s_hist[t >> 1] >> ((t & 1) * 16)) & 0xffffu;
}

// exclusive prefix of every column of base [B][T] over the blocks, in place; column totals -> tile_count.
// CTA = 32 columns x 32 row chunks: coalesced 128-byte rows, the chunk sums are scanned through shared memory.
__global__ void __launch_bounds__(1024)
bin_colscan_kernel(int num_blocks, int num_tiles, unsigned *__restrict__ base, unsigned *__restrict__ tile_count) {
  __shared__ unsigned s_sum[32][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int t = blockIdx.x * 32 + cx;
  const int per = (num_blocks + 31) / 32;
  const int b0 = min(num_blocks, ry * per), b1 = min(num_blocks, b0 + per);
  unsigned sum = 0;
  if (t < num_tiles)
    for (int b = b0; b < b1; ++b) sum += base[(size_t)b * num_tiles + t];
  s_sum[ry][cx] = sum;
  __syncthreads();
  if (ry == 0) {  // one warp: thread cx scans its column's 32 chunk sums
    unsigned run = 0;
#pragma unroll
    for (int y = 0; y < 32; ++y) {
      const unsigned v = s_sum[y][cx];
      s_sum[y][cx] = run;
      run += v;
    }
    if (t < num_tiles) tile_count[t] = run;
  }
  __syncthreads();
  if (t < num_tiles) {
    unsigned run = s_sum[ry][cx];
    for (int b = b0; b < b1; ++b) {
      const size_t i = (size_t)b * num_tiles + t;
      const unsigned c = base[i];
      base[i] = run;
      run += c;
    }
  }
}

__global__ void __launch_bounds__(BLK_THREADS)
bin_fill_blocks_kernel(int n, int per_block, const float2 *__restrict__ xys, const float *__restrict__ depths,
                       const int *__restrict__ radii, const float *__restrict__ conics, const float *__restrict__ opacities,
                       const u64 *__restrict__ masks, int tiles_x, int tiles_y, int block_width, int capacity,
                       const unsigned *__restrict__ tile_start, const unsigned *__restrict__ base, u64 *__restrict__ keys) {
  // one 32-bit write cursor per tile in shared memory, initialised (coalesced) to the start of this block's range inside
  // the tile's segment: a pair then costs one shared-memory atomic — the per-pair gathers of tile_start[] and of this
  // block's row of base[] (37 MB at cfg4, 0.8 ms) are gone
  extern __shared__ unsigned s_cur[];
  const int num_tiles = tiles_x * tiles_y;
  const unsigned *row = base + (size_t)blockIdx.x * num_tiles;
  for (int i = threadIdx.x; i < num_tiles; i += BLK_THREADS) s_cur[i] = tile_start[i] + row[i];
  __syncthreads();
  const int g0 = blockIdx.x * per_block, g1 = min(n, g0 + per_block);
  for (int g = g0 + threadIdx.x; g < g1; g += BLK_THREADS)
    fill_one(g, xys, depths, radii, conics, opacities, masks, tiles_x, tiles_y, block_width, [&](int tile, u64 key) {
      const unsigned pos = atomicAdd(&s_cur[tile], 1u);
      if (pos < (unsigned)capacity) keys[pos] = key;
    });
}

// ---- per-tile sort, one warp per tile, keys in registers --------------------------------------------------------------
// E keys per lane, element index e = lane * E + r.  Bitonic network over 32 E elements: the stages whose partner
// distance j is below E exchange registers of one lane, the others exchange with lane ^ (j / E) through shuffles.
template <int E>
__device__ __forceinline__ void warp_bitonic_sort(u64 (&k)[E], int lane) {
  const unsigned full = 0xffffffffu;
  for (int k2 = 2; k2 <= 32 * E; k2 <<= 1) {
    for (int j = k2 >> 1; j >= E; j >>= 1) {
      // partner lane = lane ^ (j / E); k2 >= 2 E here, so the direction bit of element lane * E + r does not depend on r
      const int lm = j / E;
      const bool take_min = ((lane & lm) == 0) == (((lane * E) & k2) == 0);
#pragma unroll
      for (int r = 0; r < E; ++r) {
        const u64 other = __shfl_xor_sync(full, k[r], lm);
        if ((k[r] > other) == take_min) k[r] = other;
      }
    }
#pragma unroll
    for (int jj = E >> 1; jj > 0; jj >>= 1) {
      if (jj < k2) {
#pragma unroll
        for (int r = 0; r < E; ++r) {
          if ((r & jj) == 0) {
            const bool up = ((lane * E + r) & k2) == 0;
            const u64 a = k[r], b = k[r | jj];
            if ((a > b) == up) {
              k[r] = b;
              k[r | jj] = a;
            }
          }
        }
      }
    }
  }
}

template <int E>
__device__ __forceinline__ void warp_sort_tile(const u64 *__restrict__ seg, int n, int lane, int *__restrict__ ids_out) {
  u64 k[E];
#pragma unroll
  for (int r = 0; r < E; ++r) {
    const int e = lane * E + r;
    k[r] = e < n ? seg[e] : ~0ull;
  }
  warp_bitonic_sort<E>(k, lane);
#pragma unroll
  for (int r = 0; r < E; ++r) {
    const int e = lane * E + r;
    if (e < n) ids_out[e] = (int)(unsigned)k[r];
  }
}

#ifndef GSR_SORT_WARPS
#define GSR_SORT_WARPS 4
#endif
constexpr int SORT_WARPS = GSR_SORT_WARPS;  // warps (= tiles) per CTA of tile_sort_warp_kernel

// BIG = false: every tile with 1..512 pairs (E <= 16, ~90 registers), one warp per tile; BIG = true: the tiles of the
// `mid` list (513..1024 pairs, E = 32, ~170 registers), a fixed grid striding over the list — two kernels so that the
// common short tiles are not held to the occupancy of the 64-register key array, and the second costs a few
// microseconds when the list is short
template <bool BIG>
__global__ void __launch_bounds__(32 * SORT_WARPS)
tile_sort_warp_kernel(int num_tiles, const int2 *__restrict__ tile_bins, const u64 *__restrict__ keys,
                      const int *__restrict__ lists, int *__restrict__ ids_out) {
  const int w = blockIdx.x * SORT_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (BIG) {
    // one warp per list entry (the grid covers a list of all tiles: at cfg4 nearly every tile is a mid tile; warps beyond
    // the list leave at once)
    if (w >= lists[0]) return;
    const int2 range = tile_bins[list_mid(const_cast<int *>(lists))[w]];
    warp_sort_tile<32>(keys + range.x, range.y - range.x, lane, ids_out + range.x);
  } else {
    if (w >= num_tiles) return;
    const int2 range = tile_bins[w];
    const int n = range.y - range.x;
    const u64 *seg = keys + range.x;
    int *out = ids_out + range.x;
    if (n <= 0 || n > SORT_WARP_SMALL) return;
    if (n <= 128) warp_sort_tile<4>(seg, n, lane, out);
    else if (n <= 256) warp_sort_tile<8>(seg, n, lane, out);
    else warp_sort_tile<16>(seg, n, lane, out);
  }
}

// ---- long tiles (more than SORT_WARP_MAX pairs) ----------------------------------------------------------------------
// A persistent grid of 1024-thread CTAs pulls tiles from the `long` list.  Bitonic network in its all-ascending form
// (stage k = one "flip" step, partner i ^ (k - 1), then "disperse" steps j = k/4 .. 1, partner i + j): every
// compare-exchange puts the smaller key at the lower index, so the padding up to a power of two is VIRTUAL — a pair whose
// upper index is >= n is skipped — and a tile of any length is sorted in place.  Up to LONG_CHUNK keys the whole network
// runs in shared memory; longer tiles sort LONG_CHUNK-key chunks in shared memory, then for every stage k > LONG_CHUNK
// run the steps whose partner distance spans chunks directly on the (L2-resident) segment and the remaining steps of
// the stage chunk by chunk in shared memory again.
constexpr int LONG_THREADS = 1024;
constexpr int LONG_CHUNK = 8192;  // 64 KB of 64-bit keys: two CTAs per SM

__device__ __forceinline__ void cmp_swap(u64 &a, u64 &b) {
  if (a > b) {
    const u64 t = a;
    a = b;
    b = t;
  }
}

// Shared-memory index of key i: one key of padding after every 8, so that the 8-key groups of the register passes start
// in different banks (thread g reads keys 8g .. 8g+7: unpadded, 32 lanes would hit two bank pairs, 16-way conflicts)
__device__ __forceinline__ int sp(int i) { return i + (i >> 3); }
constexpr int LONG_SMEM_BYTES = (LONG_CHUNK + LONG_CHUNK / 8) * 8;

// The three innermost disperse steps (j = 4, 2, 1) of a stage touch aligned groups of 8 keys only: one thread takes a
// group through all three in registers (one pass and one barrier instead of three).  FIRST: the stages k = 2, 4, 8
// instead (flip partners inside the group).  Keys at or beyond n_valid read as +inf and are not written back; every
// exchange is ascending, so they never move below a real key.
template <bool FIRST>
__device__ __forceinline__ void smem_groups_of_8(u64 *s, int m, int n_valid) {
  for (int g = threadIdx.x; g < (m >> 3); g += LONG_THREADS) {
    const int base = g << 3;
    if (base >= n_valid) continue;
    u64 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (base + i < n_valid) ? s[sp(base + i)] : ~0ull;
    if (FIRST) {
      cmp_swap(v[0], v[1]); cmp_swap(v[2], v[3]); cmp_swap(v[4], v[5]); cmp_swap(v[6], v[7]);   // k = 2
      cmp_swap(v[0], v[3]); cmp_swap(v[1], v[2]); cmp_swap(v[4], v[7]); cmp_swap(v[5], v[6]);   // k = 4: flip
      cmp_swap(v[0], v[1]); cmp_swap(v[2], v[3]); cmp_swap(v[4], v[5]); cmp_swap(v[6], v[7]);   //        j = 1
      cmp_swap(v[0], v[7]); cmp_swap(v[1], v[6]); cmp_swap(v[2], v[5]); cmp_swap(v[3], v[4]);   // k = 8: flip
    } else {
      cmp_swap(v[0], v[4]); cmp_swap(v[1], v[5]); cmp_swap(v[2], v[6]); cmp_swap(v[3], v[7]);   // j = 4
    }
    cmp_swap(v[0], v[2]); cmp_swap(v[1], v[3]); cmp_swap(v[4], v[6]); cmp_swap(v[5], v[7]);     // j = 2
    cmp_swap(v[0], v[1]); cmp_swap(v[2], v[3]); cmp_swap(v[4], v[5]); cmp_swap(v[6], v[7]);     // j = 1
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (base + i < n_valid) s[sp(base + i)] = v[i];
  }
  __syncthreads();
}

// steps of stage k >= 16 (flip first when `flip`) with partner distance < m on s[0, m); n_valid = real keys in s
__device__ __forceinline__ void smem_stage(u64 *s, int m, int n_valid, int k, bool flip) {
  const int tid = threadIdx.x;
  if (flip) {
    const int h = k >> 1;
    for (int t = tid; t < (m >> 1); t += LONG_THREADS) {
      const int lo = ((t & ~(h - 1)) << 1) | (t & (h - 1)), hi = lo ^ (k - 1);
      if (hi < n_valid) {
        u64 a = s[sp(lo)], b = s[sp(hi)];
        if (a > b) { s[sp(lo)] = b; s[sp(hi)] = a; }
      }
    }
    __syncthreads();
  }
  for (int j = flip ? (k >> 2) : min(k >> 2, m >> 1); j >= 8; j >>= 1) {
    for (int t = tid; t < (m >> 1); t += LONG_THREADS) {
      const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1)), hi = lo + j;
      if (hi < n_valid) {
        u64 a = s[sp(lo)], b = s[sp(hi)];
        if (a > b) { s[sp(lo)] = b; s[sp(hi)] = a; }
      }
    }
    __syncthreads();
  }
  smem_groups_of_8<false>(s, m, n_valid);
}

// the whole network up to stage k_max on s[0, m)
__device__ __forceinline__ void smem_sort(u64 *s, int m, int n_valid, int k_max) {
  smem_groups_of_8<true>(s, m, n_valid);
  for (int k = 16; k <= k_max; k <<= 1) smem_stage(s, m, n_valid, k, true);
}

__global__ void __launch_bounds__(LONG_THREADS)
tile_sort_long_kernel(int num_tiles, const int2 *__restrict__ tile_bins, u64 *__restrict__ keys, int *__restrict__ lists,
                      int *__restrict__ ids_out) {
  extern __shared__ __align__(16) unsigned char long_smem[];
  u64 *s = reinterpret_cast<u64 *>(long_smem);
  __shared__ int s_next;
  const int tid = threadIdx.x;
  const int n_long = lists[1];
  const int *lng = list_long(lists, num_tiles);
  for (;;) {
    __syncthreads();
    if (tid == 0) s_next = atomicAdd(&lists[2], 1);
    __syncthreads();
    const int item = s_next;
    if (item >= n_long) return;
    const int2 range = tile_bins[lng[item]];
    const int n = range.y - range.x;
    u64 *seg = keys + range.x;
    int *out = ids_out + range.x;
    int P = 2048;
    while (P < n) P <<= 1;
    if (P <= LONG_CHUNK) {
      for (int i = tid; i < n; i += LONG_THREADS) s[sp(i)] = seg[i];
      __syncthreads();
      smem_sort(s, P, n, P);
      for (int i = tid; i < n; i += LONG_THREADS) out[i] = (int)(unsigned)s[sp(i)];
      continue;
    }
    const int nchunks = (n + LONG_CHUNK - 1) / LONG_CHUNK;
    for (int c = 0; c < nchunks; ++c) {  // sort every chunk
      const int base = c * LONG_CHUNK, nv = min(LONG_CHUNK, n - base);
      for (int i = tid; i < nv; i += LONG_THREADS) s[sp(i)] = seg[base + i];
      __syncthreads();
      smem_sort(s, LONG_CHUNK, nv, LONG_CHUNK);
      for (int i = tid; i < nv; i += LONG_THREADS) seg[base + i] = s[sp(i)];
      __syncthreads();
    }
    for (int k = 2 * LONG_CHUNK; k <= P; k <<= 1) {
      {  // flip step of stage k on the segment
        const int h = k >> 1;
        for (int t = tid; t < (P >> 1); t += LONG_THREADS) {
          const int lo = ((t & ~(h - 1)) << 1) | (t & (h - 1)), hi = lo ^ (k - 1);
          if (hi < n) {
            u64 a = seg[lo], b = seg[hi];
            if (a > b) { seg[lo] = b; seg[hi] = a; }
          }
        }
        __syncthreads();
      }
      for (int j = k >> 2; j >= LONG_CHUNK; j >>= 1) {  // disperse steps that span chunks
        for (int t = tid; t < (P >> 1); t += LONG_THREADS) {
          const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1)), hi = lo + j;
          if (hi < n) {
            u64 a = seg[lo], b = seg[hi];
            if (a > b) { seg[lo] = b; seg[hi] = a; }
          }
        }
        __syncthreads();
      }
      const bool last = (k == P);
      for (int c = 0; c < nchunks; ++c) {  // disperse steps j = LONG_CHUNK / 2 .. 1, chunk by chunk
        const int base = c * LONG_CHUNK, nv = min(LONG_CHUNK, n - base);
        for (int i = tid; i < nv; i += LONG_THREADS) s[sp(i)] = seg[base + i];
        __syncthreads();
        smem_stage(s, LONG_CHUNK, nv, k, false);
        if (last)
          for (int i = tid; i < nv; i += LONG_THREADS) out[base + i] = (int)(unsigned)s[sp(i)];
        else
          for (int i = tid; i < nv; i += LONG_THREADS) seg[base + i] = s[sp(i)];
        __syncthreads();
      }
    }
  }
}

inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }
inline int bits_for(long long n) {
  int bits = 1;
  while ((1ll << bits) < n) ++bits;
  return bits;
}

struct BinLayout {  // carved from the caller's workspace
  u64 *masks, *keys;
  unsigned *tile_count, *cursors, *base;
  int *lists;  // {mid_count, long_count, long_cursor, 0, mid[T], long[T]} (tile_scan_kernel)
  int per_block, num_blocks;  // Gaussians per block / blocks of the shared-memory-histogram kernels (0 = global atomics)
};

// Gaussians per block: two blocks per SM (2 x 148), a multiple of the warp size, at most 65535 (16-bit counters).
// (Rounding up to a multiple of the block size — 3379 -> 4096 at cfg2 — left 245 blocks for 296 slots: a third of the SMs
// ran one block while the others ran two, ncu "0.8 full waves", SMs active 74 % of the kernel.)
inline void block_partition(int num_points, int num_tiles, int &per_block, int &num_blocks) {
  per_block = num_blocks = 0;
  if (num_tiles > SMEM_HIST_MAX_TILES || num_points <= 0) return;
  constexpr long long kBlocks = 148ll * (2048 / BLK_THREADS);
  long long per = ((long long)num_points + kBlocks - 1) / kBlocks;
  per = ((per + 31) / 32) * 32;
  if (per < 2 * BLK_THREADS) per = 2 * BLK_THREADS;
  if (per > 65280) per = 65280;
  per_block = (int)per;
  num_blocks = (int)(((long long)num_points + per - 1) / per);
}

inline size_t layout_bytes(int num_points, int capacity, int num_tiles) {
  const size_t n = num_points > 0 ? num_points : 1, c = capacity > 0 ? capacity : 1, t = num_tiles > 0 ? num_tiles : 1;
  int per_block, num_blocks;
  block_partition(num_points, num_tiles, per_block, num_blocks);
  return al256(8 * n) + al256(8 * c) + 2 * al256(4 * t) + al256(4 * t * (size_t)(num_blocks > 0 ? num_blocks : 1)) +
         al256(4 * (2 * t + 4)) + 256;
}

inline BinLayout carve(void *workspace, int num_points, int capacity, int num_tiles) {
  const size_t n = num_points > 0 ? num_points : 1, c = capacity > 0 ? capacity : 1, t = num_tiles > 0 ? num_tiles : 1;
  char *ws = (char *)workspace;
  BinLayout L;
  block_partition(num_points, num_tiles, L.per_block, L.num_blocks);
  // the capacity-independent part first, so that gsr_bin_count (capacity 1) and gsr_bin_fill_sort agree on it
  L.masks = (u64 *)ws;           ws += al256(8 * n);
  L.tile_count = (unsigned *)ws; ws += al256(4 * t);
  L.cursors = (unsigned *)ws;    ws += al256(4 * t);
  L.base = (unsigned *)ws;       ws += al256(4 * t * (size_t)(L.num_blocks > 0 ? L.num_blocks : 1));
  L.lists = (int *)ws;           ws += al256(4 * (2 * t + 4));
  L.keys = (u64 *)ws;
  return L;
}

// count + scan: tile_bins / segment starts / meta for `capacity`
int run_count(int num_points, const float *xys, const int32_t *radii, const float *conics, const float *opacities,
              int tiles_x, int tiles_y, unsigned block_width, int capacity, const BinLayout &L, int32_t *tile_bins,
              int32_t *meta, cudaStream_t st) {
  const int num_tiles = tiles_x * tiles_y;
  if (L.num_blocks > 0) {
    static const bool allow_wide = [] { const char *e = getenv("GSR_COUNT_WIDE"); return !(e && e[0] == '0'); }();
    const bool wide = allow_wide && sizeof(unsigned) * (size_t)num_tiles <= 100 * 1024;
    const size_t smem = wide ? sizeof(unsigned) * (size_t)num_tiles : sizeof(unsigned) * (size_t)((num_tiles + 1) >> 1);
    if (wide) {
      if (smem > 48 * 1024)
        GSR_CUDA(cudaFuncSetAttribute(bin_count_blocks_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      bin_count_blocks_kernel<true><<<L.num_blocks, BLK_THREADS, smem, st>>>(
          num_points, L.per_block, reinterpret_cast<const float2 *>(xys), radii, conics, opacities, tiles_x, tiles_y,
          (int)block_width, L.masks, L.base);
    } else {
      if (smem > 48 * 1024)
        GSR_CUDA(cudaFuncSetAttribute(bin_count_blocks_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      bin_count_blocks_kernel<false><<<L.num_blocks, BLK_THREADS, smem, st>>>(
          num_points, L.per_block, reinterpret_cast<const float2 *>(xys), radii, conics, opacities, tiles_x, tiles_y,
          (int)block_width, L.masks, L.base);
    }
    GSR_CHECK_LAUNCH("bin_count_blocks_kernel");
    bin_colscan_kernel<<<cdiv(num_tiles, 32), 1024, 0, st>>>(L.num_blocks, num_tiles, L.base, L.tile_count);
    GSR_CHECK_LAUNCH("bin_colscan_kernel");
  } else {
    GSR_CUDA(cudaMemsetAsync(L.tile_count, 0, sizeof(unsigned) * (size_t)num_tiles, st));
    if (num_points > 0) {
      bin_count_kernel<<<cdiv(num_points, BD_THREADS), BD_THREADS, 0, st>>>(
          num_points, reinterpret_cast<const float2 *>(xys), radii, conics, opacities, tiles_x, tiles_y, (int)block_width,
          L.masks, L.tile_count);
      GSR_CHECK_LAUNCH("bin_count_kernel");
    }
  }
  tile_scan_kernel<<<1, 1024, 0, st>>>(num_tiles, L.tile_count, capacity, reinterpret_cast<int2 *>(tile_bins), L.cursors, meta,
                                       L.lists);
  GSR_CHECK_LAUNCH("tile_scan_kernel");
  return GSR_OK;
}

int run_fill_sort(int num_points, const float *xys, const float *depths, const int32_t *radii, const float *conics,
                  const float *opacities, int tiles_x, int tiles_y, unsigned block_width, int capacity, const BinLayout &L,
                  const int32_t *tile_bins, int32_t *gaussian_ids_sorted, cudaStream_t st) {
  if (num_points <= 0 || capacity <= 0) return GSR_OK;
  const int num_tiles = tiles_x * tiles_y;
  if (L.num_blocks > 0) {
    const size_t smem = sizeof(unsigned) * (size_t)num_tiles;  // one 32-bit cursor per tile (<= 192 KB)
    if (smem > 48 * 1024)
      GSR_CUDA(cudaFuncSetAttribute(bin_fill_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bin_fill_blocks_kernel<<<L.num_blocks, BLK_THREADS, smem, st>>>(
        num_points, L.per_block, reinterpret_cast<const float2 *>(xys), depths, radii, conics, opacities, L.masks, tiles_x,
        tiles_y, (int)block_width, capacity, L.cursors, L.base, L.keys);
    GSR_CHECK_LAUNCH("bin_fill_blocks_kernel");
  } else {
    bin_fill_kernel<<<cdiv(num_points, BD_THREADS), BD_THREADS, 0, st>>>(
        num_points, reinterpret_cast<const float2 *>(xys), depths, radii, conics, opacities, L.masks, tiles_x, tiles_y,
        (int)block_width, capacity, L.cursors, L.keys);
    GSR_CHECK_LAUNCH("bin_fill_kernel");
  }
  const int2 *bins = reinterpret_cast<const int2 *>(tile_bins);
  tile_sort_warp_kernel<false><<<cdiv(num_tiles, SORT_WARPS), 32 * SORT_WARPS, 0, st>>>(num_tiles, bins, L.keys, L.lists,
                                                                                        gaussian_ids_sorted);
  GSR_CHECK_LAUNCH("tile_sort_warp_kernel<small>");
  // the two list-driven kernels: fixed grids, so nothing here depends on a device-side count
  tile_sort_warp_kernel<true><<<cdiv(num_tiles, SORT_WARPS), 32 * SORT_WARPS, 0, st>>>(num_tiles, bins, L.keys, L.lists,
                                                                                       gaussian_ids_sorted);
  GSR_CHECK_LAUNCH("tile_sort_warp_kernel<mid>");
  static const cudaError_t attr = cudaFuncSetAttribute(tile_sort_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                       LONG_SMEM_BYTES);
  GSR_CUDA(attr);
  tile_sort_long_kernel<<<2 * num_sms(), LONG_THREADS, LONG_SMEM_BYTES, st>>>(num_tiles, bins, L.keys, L.lists,
                                                                                   gaussian_ids_sorted);
  GSR_CHECK_LAUNCH("tile_sort_long_kernel");
  return GSR_OK;
}

}  // namespace
}  // namespace gsr

extern "C" {

GSR_API size_t gsr_bin_device_workspace_bytes(int num_points, int capacity, unsigned img_height, unsigned img_width,
                                              unsigned block_width) {
  using namespace gsr;
  if (block_width < 2) block_width = 2;
  return layout_bytes(num_points, capacity, (int)(cdiv(img_width, block_width) * cdiv(img_height, block_width)));
}

// Whole binning, asynchronous: M stays on the device (meta), buffers hold `capacity` pairs.
GSR_API int gsr_bin_gaussians_device(int num_points, const float *xys, const float *depths, const int32_t *radii,
                                     const float *conics, const float *opacities, unsigned img_height,
                                     unsigned img_width, unsigned block_width, int capacity,
                                     int32_t *gaussian_ids_sorted, int32_t *tile_bins, int32_t *meta,
                                     int32_t *meta_host_pinned, void *workspace, size_t workspace_bytes, void *stream) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_bin_gaussians_device");
  GSR_REQUIRE(num_points >= 0 && capacity >= 0, GSR_ERR_INVALID_ARGUMENT, "bin_gaussians_device: negative size");
  GSR_REQUIRE(block_width > 1 && block_width <= 16, GSR_ERR_INVALID_ARGUMENT,
              "block_width must be between 2 and 16 (got %u)", block_width);
  GSR_REQUIRE(img_height > 0 && img_width > 0, GSR_ERR_INVALID_ARGUMENT, "bin_gaussians_device: empty image");
  GSR_REQUIRE(tile_bins && meta && workspace, GSR_ERR_INVALID_ARGUMENT, "bin_gaussians_device: null pointer");
  GSR_REQUIRE(num_points == 0 || (xys && depths && radii && conics && opacities), GSR_ERR_INVALID_ARGUMENT,
              "bin_gaussians_device: null pointer");
  GSR_REQUIRE(capacity == 0 || gaussian_ids_sorted, GSR_ERR_INVALID_ARGUMENT, "bin_gaussians_device: null pointer");
  GSR_REQUIRE((uintptr_t)xys % 8 == 0 && (uintptr_t)workspace % 16 == 0 && (uintptr_t)tile_bins % 8 == 0,
              GSR_ERR_INVALID_ARGUMENT, "bin_gaussians_device: misaligned pointer");
  const int tiles_x = cdiv(img_width, block_width), tiles_y = cdiv(img_height, block_width);
  GSR_REQUIRE(workspace_bytes >= layout_bytes(num_points, capacity, tiles_x * tiles_y), GSR_ERR_WORKSPACE,
              "bin_gaussians_device: workspace %zu < %zu bytes", workspace_bytes,
              layout_bytes(num_points, capacity, tiles_x * tiles_y));
  cudaStream_t st = (cudaStream_t)stream;
  const BinLayout L = carve(workspace, num_points, capacity, tiles_x * tiles_y);
  int rc = run_count(num_points, xys, radii, conics, opacities, tiles_x, tiles_y, block_width, capacity, L, tile_bins, meta, st);
  if (rc != GSR_OK) return rc;
  if (meta_host_pinned)
    GSR_CUDA(cudaMemcpyAsync(meta_host_pinned, meta, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  return run_fill_sort(num_points, xys, depths, radii, conics, opacities, tiles_x, tiles_y, block_width, capacity, L, tile_bins,
                       gaussian_ids_sorted, st);
}

// Two-call form with a host-visible M: gsr_bin_count (capacity "unbounded": tile_bins are exact), the caller reads
// meta_host_pinned[0] = M after synchronising, allocates gaussian_ids_sorted [M] and a workspace for capacity M, and
// calls gsr_bin_fill_sort with the SAME masks / cursors (kept in the first workspace: pass it as `count_workspace`).
GSR_API size_t gsr_bin_count_workspace_bytes(int num_points, unsigned img_height, unsigned img_width, unsigned block_width) {
  return gsr_bin_device_workspace_bytes(num_points, 1, img_height, img_width, block_width);
}

GSR_API int gsr_bin_count(int num_points, const float *xys, const int32_t *radii, const float *conics,
                          const float *opacities, unsigned img_height, unsigned img_width, unsigned block_width,
                          int32_t *tile_bins, int32_t *meta, int32_t *meta_host_pinned, void *count_workspace,
                          size_t workspace_bytes, void *stream) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_bin_count");
  GSR_REQUIRE(num_points >= 0, GSR_ERR_INVALID_ARGUMENT, "bin_count: num_points < 0");
  GSR_REQUIRE(block_width > 1 && block_width <= 16, GSR_ERR_INVALID_ARGUMENT,
              "block_width must be between 2 and 16 (got %u)", block_width);
  GSR_REQUIRE(img_height > 0 && img_width > 0, GSR_ERR_INVALID_ARGUMENT, "bin_count: empty image");
  GSR_REQUIRE(tile_bins && meta && count_workspace, GSR_ERR_INVALID_ARGUMENT, "bin_count: null pointer");
  GSR_REQUIRE(num_points == 0 || (xys && radii && conics && opacities), GSR_ERR_INVALID_ARGUMENT, "bin_count: null pointer");
  GSR_REQUIRE((uintptr_t)xys % 8 == 0 && (uintptr_t)count_workspace % 16 == 0 && (uintptr_t)tile_bins % 8 == 0,
              GSR_ERR_INVALID_ARGUMENT, "bin_count: misaligned pointer");
  const int tiles_x = cdiv(img_width, block_width), tiles_y = cdiv(img_height, block_width);
  GSR_REQUIRE(workspace_bytes >= layout_bytes(num_points, 1, tiles_x * tiles_y), GSR_ERR_WORKSPACE,
              "bin_count: workspace %zu < %zu bytes", workspace_bytes, layout_bytes(num_points, 1, tiles_x * tiles_y));
  cudaStream_t st = (cudaStream_t)stream;
  const BinLayout L = carve(count_workspace, num_points, 1, tiles_x * tiles_y);
  const int rc = run_count(num_points, xys, radii, conics, opacities, tiles_x, tiles_y, block_width, 0x7fffffff, L, tile_bins,
                           meta, st);
  if (rc != GSR_OK) return rc;
  if (meta_host_pinned)
    GSR_CUDA(cudaMemcpyAsync(meta_host_pinned, meta, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  return GSR_OK;
}

GSR_API size_t gsr_bin_fill_workspace_bytes(int num_intersects) {
  const size_t c = num_intersects > 0 ? num_intersects : 1;
  return ((8 * c + 255) & ~(size_t)255) + 256;  // keys
}

GSR_API int gsr_bin_fill_sort(int num_points, int num_intersects, const float *xys, const float *depths,
                              const int32_t *radii, const float *conics, const float *opacities,
                              unsigned img_height, unsigned img_width, unsigned block_width, const int32_t *tile_bins,
                              void *count_workspace, int32_t *gaussian_ids_sorted, void *fill_workspace,
                              size_t fill_workspace_bytes, void *stream) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_bin_fill_sort");
  GSR_REQUIRE(num_points >= 0 && num_intersects >= 0, GSR_ERR_INVALID_ARGUMENT, "bin_fill_sort: negative size");
  GSR_REQUIRE(block_width > 1 && block_width <= 16, GSR_ERR_INVALID_ARGUMENT,
              "block_width must be between 2 and 16 (got %u)", block_width);
  if (num_points == 0 || num_intersects == 0) return GSR_OK;
  GSR_REQUIRE(xys && depths && radii && conics && opacities && tile_bins && count_workspace && gaussian_ids_sorted &&
                  fill_workspace,
              GSR_ERR_INVALID_ARGUMENT, "bin_fill_sort: null pointer");
  GSR_REQUIRE(fill_workspace_bytes >= gsr_bin_fill_workspace_bytes(num_intersects), GSR_ERR_WORKSPACE,
              "bin_fill_sort: workspace %zu < %zu bytes", fill_workspace_bytes, gsr_bin_fill_workspace_bytes(num_intersects));
  GSR_REQUIRE((uintptr_t)fill_workspace % 16 == 0, GSR_ERR_INVALID_ARGUMENT, "bin_fill_sort: misaligned workspace");
  const int tiles_x = cdiv(img_width, block_width), tiles_y = cdiv(img_height, block_width);
  BinLayout L = carve(count_workspace, num_points, 1, tiles_x * tiles_y);
  L.keys = (u64 *)fill_workspace;
  return run_fill_sort(num_points, xys, depths, radii, conics, opacities, tiles_x, tiles_y, block_width, num_intersects, L,
                       tile_bins, gaussian_ids_sorted, (cudaStream_t)stream);
}
}
