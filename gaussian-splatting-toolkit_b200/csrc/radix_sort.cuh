// radix_sort.cuh — hand-written stable LSD radix sort of (key, value) pairs for sm_100a, with the element count read
// from DEVICE memory (so that a sort can follow a scan whose total the host never sees: no stream synchronisation, the
// whole binning is capturable in a CUDA graph).
//
// It replaces the two library sorts of the tile binning (round 1: cub::DeviceRadixSort): the depth sort of the N
// Gaussians (32-bit keys) and the stable sort of the M (tile, Gaussian) pairs by tile id (ceil(log2 T) bits), which
// together replace the reference's torch.sort(int64) + torch.gather (rasterizer/utils.py:179-180).
//
// One pass over `bits` <= 8 key bits = three launches on a fixed grid of `tiles_cap` = ceil(n_cap / 4096) blocks:
//   hist_kernel     per-tile digit histogram -> tile_hist[digit][tile]               (reads the keys: 4 or 8 B / element)
//   row_scan_kernel one block per digit: exclusive scan of its row over the tiles, row total -> digit_total[digit]
//   scatter_kernel  per tile: digit start = exclusive scan of digit_total (in shared memory) + the tile's row entry;
//                   STABLE rank of every element inside the tile: a warp owns 512 CONSECUTIVE elements and walks them
//                   in 16 rounds of 32; per round `__match_any_sync` on the digit gives the peers, the rank among
//                   them is a popcount, the warp's running per-digit counters live in shared memory; a cross-warp
//                   exclusive prefix per digit finishes the in-tile offsets; elements go straight to their final slot
//                   (runs of equal digits are contiguous, L2 combines the sectors).
// All of it is HBM / L2-bound integer work: 12 (u32 key, i32 value) or 20 (u64 key) bytes read + 8 / 12 written per
// element and pass.
#pragma once
#include "common.cuh"

namespace gsr {
namespace rs {

constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
constexpr int ITEMS = 16;
constexpr int TILE = THREADS * ITEMS;  // 4096 elements per block
constexpr int MAX_RADIX = 256;

__device__ __forceinline__ int count_of(int n_cap, const int *__restrict__ n_dev) {
  return n_dev ? min(n_cap, max(0, *n_dev)) : n_cap;
}

// exclusive scan of one value per thread over the 256 threads of the block; returns the exclusive prefix, *total = sum
__device__ __forceinline__ unsigned block_exclusive_scan(unsigned v, unsigned *s_warp /*[WARPS]*/, unsigned *total) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned up = __shfl_up_sync(full, inc, d);
    if (lane >= d) inc += up;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  unsigned base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < WARPS; ++w) {
    const unsigned t = s_warp[w];
    if (w < warp) base += t;
    tot += t;
  }
  __syncthreads();
  *total = tot;
  return base + inc - v;
}

template <typename K>
__global__ void __launch_bounds__(THREADS)
hist_kernel(const K *__restrict__ keys, int n_cap, const int *__restrict__ n_dev, int shift, int bits,
            unsigned *__restrict__ tile_hist, int tiles_cap) {
  __shared__ unsigned s_hist[WARPS][MAX_RADIX];
  const int n = count_of(n_cap, n_dev);
  const int tile = blockIdx.x, tid = threadIdx.x, warp = tid >> 5;
  const int radix = 1 << bits;
  const unsigned mask = (unsigned)radix - 1u;
  for (int i = tid; i < WARPS * MAX_RADIX; i += THREADS) (&s_hist[0][0])[i] = 0u;
  __syncthreads();
  const long long base = (long long)tile * TILE;
  if (base < n) {
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      const long long idx = base + k * THREADS + tid;
      if (idx < n) atomicAdd(&s_hist[warp][(unsigned)(keys[idx] >> shift) & mask], 1u);
    }
  }
  __syncthreads();
  for (int d = tid; d < radix; d += THREADS) {
    unsigned sum = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) sum += s_hist[w][d];
    tile_hist[(size_t)d * tiles_cap + tile] = sum;  // every block writes its column: no memset between passes
  }
}

static __global__ void __launch_bounds__(THREADS)
row_scan_kernel(unsigned *__restrict__ tile_hist, int tiles_cap, unsigned *__restrict__ digit_total) {
  __shared__ unsigned s_warp[WARPS];
  unsigned *row = tile_hist + (size_t)blockIdx.x * tiles_cap;
  const int per = (tiles_cap + THREADS - 1) / THREADS;
  const int lo = min(tiles_cap, (int)threadIdx.x * per), hi = min(tiles_cap, lo + per);
  unsigned sum = 0;
  for (int i = lo; i < hi; ++i) sum += row[i];
  unsigned total;
  unsigned run = block_exclusive_scan(sum, s_warp, &total);
  for (int i = lo; i < hi; ++i) {
    const unsigned v = row[i];
    row[i] = run;
    run += v;
  }
  if (threadIdx.x == 0) digit_total[blockIdx.x] = total;
}

template <typename K, typename V>
__global__ void __launch_bounds__(THREADS)
scatter_kernel(const K *__restrict__ keys_in, const V *__restrict__ vals_in, K *__restrict__ keys_out,
               V *__restrict__ vals_out, int n_cap, const int *__restrict__ n_dev, int shift, int bits,
               const unsigned *__restrict__ tile_hist, int tiles_cap, const unsigned *__restrict__ digit_total) {
  __shared__ unsigned s_warp_hist[WARPS][MAX_RADIX];
  __shared__ unsigned s_base[MAX_RADIX];
  __shared__ unsigned s_warp[WARPS];
  const unsigned full = 0xffffffffu;
  const int n = count_of(n_cap, n_dev);
  const int tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long base = (long long)tile * TILE;
  if (base >= n) return;  // uniform
  const int radix = 1 << bits;
  const unsigned mask = (unsigned)radix - 1u;
  for (int i = tid; i < WARPS * MAX_RADIX; i += THREADS) (&s_warp_hist[0][0])[i] = 0u;
  {
    unsigned total;
    const unsigned mine = tid < radix ? digit_total[tid] : 0u;
    const unsigned start = block_exclusive_scan(mine, s_warp, &total);  // contains the barriers
    if (tid < radix) s_base[tid] = start + tile_hist[(size_t)tid * tiles_cap + tile];
  }
  __syncthreads();

  K key[ITEMS];
  V val[ITEMS];
  unsigned short rank[ITEMS];
  const long long wbase = base + (long long)warp * (32 * ITEMS);
  const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) {
    const long long idx = wbase + r * 32 + lane;
    const bool valid = idx < n;
    unsigned d = 0xffffffffu;
    if (valid) {
      key[r] = keys_in[idx];
      val[r] = vals_in[idx];
      d = (unsigned)(key[r] >> shift) & mask;
    }
    const unsigned peers = __match_any_sync(full, d);
    const unsigned rnk = __popc(peers & lt_mask);
    unsigned prev = 0;
    if (valid) prev = s_warp_hist[warp][d];
    __syncwarp();
    if (valid && rnk == 0) s_warp_hist[warp][d] = prev + __popc(peers);
    __syncwarp();
    rank[r] = (unsigned short)(prev + rnk);
  }
  __syncthreads();
  if (tid < radix) {
    unsigned run = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
      const unsigned t = s_warp_hist[w][tid];
      s_warp_hist[w][tid] = run;
      run += t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) {
    const long long idx = wbase + r * 32 + lane;
    if (idx < n) {
      const unsigned d = (unsigned)(key[r] >> shift) & mask;
      const unsigned pos = s_base[d] + s_warp_hist[warp][d] + rank[r];
      keys_out[pos] = key[r];
      vals_out[pos] = val[r];
    }
  }
}

static inline int tiles_for(int n_cap) { return n_cap > 0 ? (n_cap + TILE - 1) / TILE : 1; }
// workspace of one sort: tile_hist [MAX_RADIX][tiles_cap] + digit_total [MAX_RADIX]
static inline size_t workspace_bytes(int n_cap) {
  return ((size_t)MAX_RADIX * tiles_for(n_cap) + MAX_RADIX) * sizeof(unsigned) + 256;
}
static inline int num_passes(int begin_bit, int end_bit) { return (end_bit - begin_bit + 7) / 8; }

// Sorts bits [begin_bit, end_bit) of the keys, stable.  Pass 0 reads (keys_in, vals_in) — which may be const caller
// data — and writes the X buffers; later passes alternate Y, X, Y, ...: the result is in the X buffers when
// num_passes() is odd and in the Y buffers when it is even.  n = min(n_cap, *n_dev) when n_dev != NULL, else n_cap.
template <typename K, typename V>
static int sort_pairs_from(const K *keys_in, const V *vals_in, K *keys_x, V *vals_x, K *keys_y, V *vals_y, int n_cap,
                           const int *n_dev, int begin_bit, int end_bit, void *workspace, cudaStream_t st) {
  if (n_cap <= 0 || end_bit <= begin_bit) return GSR_OK;
  const int tiles_cap = tiles_for(n_cap);
  unsigned *tile_hist = (unsigned *)workspace;
  unsigned *digit_total = tile_hist + (size_t)MAX_RADIX * tiles_cap;
  const int passes = num_passes(begin_bit, end_bit);
  int bit = begin_bit;
  for (int p = 0; p < passes; ++p) {
    const int left = end_bit - bit;
    const int bits = (left + (passes - p) - 1) / (passes - p);  // spread the bits evenly over the remaining passes
    K *ko = (p & 1) ? keys_y : keys_x;
    V *vo = (p & 1) ? vals_y : vals_x;
    hist_kernel<K><<<tiles_cap, THREADS, 0, st>>>(keys_in, n_cap, n_dev, bit, bits, tile_hist, tiles_cap);
    GSR_CHECK_LAUNCH("rs::hist_kernel");
    row_scan_kernel<<<1 << bits, THREADS, 0, st>>>(tile_hist, tiles_cap, digit_total);
    GSR_CHECK_LAUNCH("rs::row_scan_kernel");
    scatter_kernel<K, V><<<tiles_cap, THREADS, 0, st>>>(keys_in, vals_in, ko, vo, n_cap, n_dev, bit, bits, tile_hist,
                                                        tiles_cap, digit_total);
    GSR_CHECK_LAUNCH("rs::scatter_kernel");
    keys_in = ko;
    vals_in = vo;
    bit += bits;
  }
  return GSR_OK;
}

// In-place flavour: the data starts in the A buffers and ends there when num_passes() is even, in B when it is odd.
template <typename K, typename V>
static int sort_pairs(K *keys_a, V *vals_a, K *keys_b, V *vals_b, int n_cap, const int *n_dev, int begin_bit,
                      int end_bit, void *workspace, cudaStream_t st) {
  return sort_pairs_from<K, V>(keys_a, vals_a, keys_b, vals_b, keys_a, vals_a, n_cap, n_dev, begin_bit, end_bit, workspace, st);
}

// ---- inclusive scan of int32 counts, optionally read through a permutation: out[j] = sum_{i<=j} in[perm ? perm[i] : i] ----
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = THREADS * SCAN_ITEMS;  // 4096

static __global__ void __launch_bounds__(THREADS)
scan_tile_sums_kernel(int n, const int *__restrict__ perm, const int *__restrict__ in, unsigned *__restrict__ tile_sums) {
  __shared__ unsigned s_warp[WARPS];
  const int base = blockIdx.x * SCAN_TILE;
  unsigned sum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const int j = base + k * THREADS + threadIdx.x;
    if (j < n) sum += (unsigned)in[perm ? perm[j] : j];
  }
  unsigned total;
  block_exclusive_scan(sum, s_warp, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// one block: exclusive scan of the tile sums in place; meta (nullable) = {total, total > capacity, min(total, capacity), 0}
static __global__ void __launch_bounds__(THREADS)
scan_of_sums_kernel(int num_tiles, unsigned *__restrict__ tile_sums, int capacity, int *__restrict__ meta) {
  __shared__ unsigned s_warp[WARPS];
  unsigned carry = 0;
  for (int base = 0; base < num_tiles; base += THREADS) {
    const int i = base + threadIdx.x;
    const unsigned v = i < num_tiles ? tile_sums[i] : 0u;
    unsigned total;
    const unsigned ex = block_exclusive_scan(v, s_warp, &total);
    if (i < num_tiles) tile_sums[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0 && meta) {
    // the reference keeps the total in an int32 too (torch.cumsum(dtype=int32), rasterizer/utils.py:123)
    const long long m = (long long)carry;
    meta[0] = (int)min(m, 0x7fffffffLL);
    meta[1] = m > (long long)capacity ? 1 : 0;
    meta[2] = (int)min(m, (long long)capacity);
    meta[3] = 0;
  }
}

static __global__ void __launch_bounds__(THREADS)
scan_tiles_kernel(int n, const int *__restrict__ perm, const int *__restrict__ in,
                  const unsigned *__restrict__ tile_offsets, int *__restrict__ out) {
  __shared__ unsigned s_warp[WARPS];
  // thread t owns SCAN_ITEMS CONSECUTIVE elements, so one block scan of the thread sums finishes the tile
  const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  unsigned v[SCAN_ITEMS];
  unsigned sum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const int j = base + k;
    v[k] = j < n ? (unsigned)in[perm ? perm[j] : j] : 0u;
    sum += v[k];
  }
  unsigned total;
  unsigned run = tile_offsets[blockIdx.x] + block_exclusive_scan(sum, s_warp, &total);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const int j = base + k;
    run += v[k];
    if (j < n) out[j] = (int)run;
  }
}

static inline int scan_tiles_for(int n) { return n > 0 ? (n + SCAN_TILE - 1) / SCAN_TILE : 1; }
static inline size_t scan_workspace_bytes(int n) { return (size_t)scan_tiles_for(n) * sizeof(unsigned) + 256; }

// out[j] = inclusive sum; meta (nullable, DEVICE int[4]) as in scan_of_sums_kernel.  `workspace`: scan_workspace_bytes(n).
static int inclusive_scan(int n, const int *perm, const int *in, int *out, int capacity, int *meta, void *workspace,
                          cudaStream_t st) {
  if (n <= 0) return GSR_OK;
  unsigned *tile_sums = (unsigned *)workspace;
  const int tiles = scan_tiles_for(n);
  scan_tile_sums_kernel<<<tiles, THREADS, 0, st>>>(n, perm, in, tile_sums);
  GSR_CHECK_LAUNCH("rs::scan_tile_sums_kernel");
  scan_of_sums_kernel<<<1, THREADS, 0, st>>>(tiles, tile_sums, capacity, meta);
  GSR_CHECK_LAUNCH("rs::scan_of_sums_kernel");
  scan_tiles_kernel<<<tiles, THREADS, 0, st>>>(n, perm, in, tile_sums, out);
  GSR_CHECK_LAUNCH("rs::scan_tiles_kernel");
  return GSR_OK;
}

}  // namespace rs
}  // namespace gsr
