// peer_reduce.cu — all-reduce (sum) of a replicated FP32 buffer over NVLink peer memory, in two kernels:
//
//   gsr_peer_reduce_scatter   rank r sums slice r of every rank's buffer (coalesced 16-byte loads from the peers' mapped
//                             buffers) and stores the total into slice r of its OWN buffer, in place;
//   gsr_peer_all_gather       rank r copies the reduced slice w of every peer w into slice w of its own buffer.
//
// The caller separates the two with a device-side barrier over all ranks (every slice reduced before anyone gathers) and
// brackets them with "all published" / "all consumed" barriers (rasterizer/view_parallel.py: symmetric-memory signal pads,
// stream-ordered, no host synchronisation).  In place is safe: during the first kernel rank r writes only slice r of its own
// buffer while the peers read only their own slices of it; during the second it writes only the other slices while the
// peers read only slice r.  Used for the 11 N non-SH gradient floats of the view-parallel exchange: at 8
// ranks a GPU receives 2 x 7/8 x 44 MB instead of running a 44 MB NCCL all-reduce beside the peer-load SH adjoint.
//
// This is plumbing of the multi-GPU extension; the reference's counterpart is torch DDP's bucketed
// all-reduce (pipelines/base_pipeline.py:202-207).
#include "common.cuh"

namespace gsr {

constexpr int PEER_MAX_RANKS = 16;
struct PeerBufs {
  float *buf[PEER_MAX_RANKS];
};

// floats per slice: a multiple of 4 (16-byte loads), slices [r * L, min((r + 1) * L, n))
__host__ __device__ inline long long peer_slice_len(long long n, int world) {
  const long long per = (n + world - 1) / world;
  return (per + 3) & ~3ll;
}

template <int WMAX>  // smallest of 2, 4, 8, 16 that holds the world: bounds the unrolled per-rank registers
__global__ void __launch_bounds__(256)
peer_reduce_scatter_kernel(PeerBufs bufs, int world, int rank, long long n) {
  const long long L = peer_slice_len(n, world);
  const long long lo = (long long)rank * L, hi = min(n, lo + L);
  if (lo >= hi) return;
  const long long nvec = (hi - lo) >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  // two float4 per thread and trip, every rank's load issued before the first add: NVLink round trips (~3 us) need many
  // bytes in flight per SM
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += 2 * stride) {
    const long long i1 = i + stride;
    const bool two = i1 < nvec;
    float4 v0[WMAX], v1[WMAX];
#pragma unroll
    for (int w = 0; w < WMAX; ++w)  // unrolled with a guard: the pointer table stays in the parameter space
      if (w < world) {
        v0[w] = *reinterpret_cast<const float4 *>(bufs.buf[w] + lo + 4 * i);
        if (two) v1[w] = *reinterpret_cast<const float4 *>(bufs.buf[w] + lo + 4 * i1);
      }
    // fixed summation order (rank 0, 1, ...): every replica of the result is identical bit for bit
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
#pragma unroll
    for (int w = 0; w < WMAX; ++w)
      if (w < world) {
        a0.x += v0[w].x; a0.y += v0[w].y; a0.z += v0[w].z; a0.w += v0[w].w;
        if (two) { a1.x += v1[w].x; a1.y += v1[w].y; a1.z += v1[w].z; a1.w += v1[w].w; }
      }
    *reinterpret_cast<float4 *>(bufs.buf[rank] + lo + 4 * i) = a0;
    if (two) *reinterpret_cast<float4 *>(bufs.buf[rank] + lo + 4 * i1) = a1;
  }
  // tail of the last slice (n not a multiple of 4)
  const long long t0 = lo + (nvec << 2);
  if (blockIdx.x == 0 && threadIdx.x < hi - t0) {
    float acc = 0.f;
    for (int w = 0; w < world; ++w) acc += bufs.buf[w][t0 + threadIdx.x];
    bufs.buf[rank][t0 + threadIdx.x] = acc;
  }
}

__global__ void __launch_bounds__(256)
peer_all_gather_kernel(PeerBufs bufs, int world, int rank, long long n) {
  const long long L = peer_slice_len(n, world);
  // blockIdx.y = which peer (skipping this rank)
  const int w = blockIdx.y + (blockIdx.y >= rank ? 1 : 0);
  const long long lo = (long long)w * L, hi = min(n, lo + L);
  if (lo >= hi) return;
  const long long nvec = (hi - lo) >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const float *src = bufs.buf[w];
  float *dst = bufs.buf[rank];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += 4 * stride) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i + u * stride < nvec) v[u] = *reinterpret_cast<const float4 *>(src + lo + 4 * (i + u * stride));
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i + u * stride < nvec) *reinterpret_cast<float4 *>(dst + lo + 4 * (i + u * stride)) = v[u];
  }
  const long long t0 = lo + (nvec << 2);
  if (blockIdx.x == 0 && threadIdx.x < hi - t0) dst[t0 + threadIdx.x] = src[t0 + threadIdx.x];
}

// ---- push-based variants: remote STORES (posted: no NVLink round trip per access), reductions from local memory ----------
//   peer_push_kernel              dst[w][0 .. n_w) <- src[w * src_stride .. )   for every rank w (blockIdx.y); src_stride = 0
//                                 broadcasts one source, src_stride = L scatters slice w of the source to rank w
//   peer_reduce_broadcast_kernel  acc = sum over slots (fixed order) of this rank's LOCAL staging slots; the total is
//                                 stored into every rank's destination (the reduced slice lands in all replicas)
__global__ void __launch_bounds__(256)
peer_push_kernel(PeerBufs dsts, const float *__restrict__ src, long long src_stride, long long n_per_dst, long long n_total) {
  const int w = blockIdx.y;
  const long long off = (long long)w * src_stride;
  // scatter mode: the last slices may be shorter / empty
  const long long n = src_stride > 0 ? max(0ll, min(n_per_dst, n_total - off)) : n_per_dst;
  if (n <= 0) return;
  const float *s = src + off;
  float *d = dsts.buf[w];
  const long long nvec = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += 4 * stride) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i + u * stride < nvec) v[u] = *reinterpret_cast<const float4 *>(s + 4 * (i + u * stride));
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i + u * stride < nvec) *reinterpret_cast<float4 *>(d + 4 * (i + u * stride)) = v[u];
  }
  const long long t0 = nvec << 2;
  if (blockIdx.x == 0 && threadIdx.x < n - t0) d[t0 + threadIdx.x] = s[t0 + threadIdx.x];
}

template <int WMAX>
__global__ void __launch_bounds__(256)
peer_reduce_broadcast_kernel(PeerBufs dsts, int world, const float *__restrict__ slots, long long slot_stride, long long n) {
  const long long nvec = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    float4 v[WMAX];
#pragma unroll
    for (int w = 0; w < WMAX; ++w)
      if (w < world) v[w] = *reinterpret_cast<const float4 *>(slots + (long long)w * slot_stride + 4 * i);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int w = 0; w < WMAX; ++w)  // fixed order: bit-identical replicas
      if (w < world) { a.x += v[w].x; a.y += v[w].y; a.z += v[w].z; a.w += v[w].w; }
#pragma unroll
    for (int w = 0; w < WMAX; ++w)
      if (w < world) *reinterpret_cast<float4 *>(dsts.buf[w] + 4 * i) = a;
  }
  const long long t0 = nvec << 2;
  if (blockIdx.x == 0 && threadIdx.x < n - t0) {
    float a = 0.f;
    for (int w = 0; w < world; ++w) a += slots[(long long)w * slot_stride + t0 + threadIdx.x];
    for (int w = 0; w < world; ++w) dsts.buf[w][t0 + threadIdx.x] = a;
  }
}

static int fill_bufs(PeerBufs &b, int world, int rank, float *const *bufs_host, const char *what) {
  GSR_REQUIRE(world >= 1 && world <= PEER_MAX_RANKS, GSR_ERR_UNSUPPORTED, "%s: world %d not in [1,%d]", what, world, PEER_MAX_RANKS);
  GSR_REQUIRE(rank >= 0 && rank < world, GSR_ERR_INVALID_ARGUMENT, "%s: rank %d not in [0,%d)", what, rank, world);
  GSR_REQUIRE(bufs_host != nullptr, GSR_ERR_INVALID_ARGUMENT, "%s: null pointer table", what);
  for (int w = 0; w < world; ++w) {
    GSR_REQUIRE(bufs_host[w] != nullptr && (uintptr_t)bufs_host[w] % 16 == 0, GSR_ERR_INVALID_ARGUMENT,
                "%s: buffer of rank %d is null or not 16-byte aligned", what, w);
    b.buf[w] = bufs_host[w];
  }
  return GSR_OK;
}

}  // namespace gsr

extern "C" {

GSR_API int gsr_peer_reduce_scatter(int world, int rank, float *const *bufs_host, long long num_floats, void *stream) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_peer_reduce_scatter");
  PeerBufs b;
  const int rc = fill_bufs(b, world, rank, bufs_host, "peer_reduce_scatter");
  if (rc != GSR_OK) return rc;
  GSR_REQUIRE(num_floats >= 0, GSR_ERR_INVALID_ARGUMENT, "peer_reduce_scatter: negative size");
  if (num_floats == 0) return GSR_OK;
  const long long L = peer_slice_len(num_floats, world);
  const int blocks = (int)std::min<long long>(8 * 148, std::max<long long>(1, (L / 8 + 255) / 256));
  cudaStream_t st = (cudaStream_t)stream;
  if (world <= 2) peer_reduce_scatter_kernel<2><<<blocks, 256, 0, st>>>(b, world, rank, num_floats);
  else if (world <= 4) peer_reduce_scatter_kernel<4><<<blocks, 256, 0, st>>>(b, world, rank, num_floats);
  else if (world <= 8) peer_reduce_scatter_kernel<8><<<blocks, 256, 0, st>>>(b, world, rank, num_floats);
  else peer_reduce_scatter_kernel<16><<<blocks, 256, 0, st>>>(b, world, rank, num_floats);
  GSR_CHECK_LAUNCH("peer_reduce_scatter_kernel");
  return GSR_OK;
}

GSR_API int gsr_peer_all_gather(int world, int rank, float *const *bufs_host, long long num_floats, void *stream) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_peer_all_gather");
  PeerBufs b;
  const int rc = fill_bufs(b, world, rank, bufs_host, "peer_all_gather");
  if (rc != GSR_OK) return rc;
  GSR_REQUIRE(num_floats >= 0, GSR_ERR_INVALID_ARGUMENT, "peer_all_gather: negative size");
  if (num_floats == 0 || world == 1) return GSR_OK;
  const long long L = peer_slice_len(num_floats, world);
  const int bx = (int)std::min<long long>(std::max(148, 8 * 148 / (world - 1)), std::max<long long>(1, (L / 16 + 255) / 256));
  peer_all_gather_kernel<<<dim3(bx, world - 1, 1), 256, 0, (cudaStream_t)stream>>>(b, world, rank, num_floats);
  GSR_CHECK_LAUNCH("peer_all_gather_kernel");
  return GSR_OK;
}

GSR_API int gsr_peer_push(int world, float *const *dsts_host, const float *src, long long src_stride, long long n_per_dst,
                          long long n_total, void *stream) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_peer_push");
  PeerBufs b;
  const int rc = fill_bufs(b, world, 0, dsts_host, "peer_push");
  if (rc != GSR_OK) return rc;
  GSR_REQUIRE(src != nullptr && (uintptr_t)src % 16 == 0 && src_stride >= 0 && src_stride % 4 == 0 && n_per_dst >= 0,
              GSR_ERR_INVALID_ARGUMENT, "peer_push: bad source / stride / size");
  if (n_per_dst == 0) return GSR_OK;
  const int bx = (int)std::min<long long>(std::max(64, 8 * 148 / world), std::max<long long>(1, (n_per_dst / 16 + 255) / 256));
  peer_push_kernel<<<dim3(bx, world, 1), 256, 0, (cudaStream_t)stream>>>(b, src, src_stride, n_per_dst, n_total);
  GSR_CHECK_LAUNCH("peer_push_kernel");
  return GSR_OK;
}

GSR_API int gsr_peer_reduce_broadcast(int world, float *const *dsts_host, const float *slots, long long slot_stride,
                                      long long num_floats, void *stream) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_peer_reduce_broadcast");
  PeerBufs b;
  const int rc = fill_bufs(b, world, 0, dsts_host, "peer_reduce_broadcast");
  if (rc != GSR_OK) return rc;
  GSR_REQUIRE(slots != nullptr && (uintptr_t)slots % 16 == 0 && slot_stride % 4 == 0 && num_floats >= 0,
              GSR_ERR_INVALID_ARGUMENT, "peer_reduce_broadcast: bad slots / stride / size");
  if (num_floats == 0) return GSR_OK;
  const int blocks = (int)std::min<long long>(8 * 148, std::max<long long>(1, (num_floats / 4 + 255) / 256));
  cudaStream_t st = (cudaStream_t)stream;
  if (world <= 2) peer_reduce_broadcast_kernel<2><<<blocks, 256, 0, st>>>(b, world, slots, slot_stride, num_floats);
  else if (world <= 4) peer_reduce_broadcast_kernel<4><<<blocks, 256, 0, st>>>(b, world, slots, slot_stride, num_floats);
  else if (world <= 8) peer_reduce_broadcast_kernel<8><<<blocks, 256, 0, st>>>(b, world, slots, slot_stride, num_floats);
  else peer_reduce_broadcast_kernel<16><<<blocks, 256, 0, st>>>(b, world, slots, slot_stride, num_floats);
  GSR_CHECK_LAUNCH("peer_reduce_broadcast_kernel");
  return GSR_OK;
}
}
