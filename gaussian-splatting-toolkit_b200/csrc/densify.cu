// densify.cu — adaptive density control of the Gaussian set as device-side compaction, sm_100a.
//
// Replaces the torch formulation of gs_toolkit/models/vanilla_gs.py:
//   after_train       :344-372   running statistics (sum of |d loss / d xy|, visibility count, max screen radius)
//   refinement_after  :381-497   split / duplicate / cull decisions, torch.cat of the new Gaussians, boolean-mask
//                                indexing of every parameter and of every Adam moment (remove_from_optim :282-300,
//                                dup_in_optim :308-337), opacity reset
//   split_gaussians   :537-581, dup_gaussians :583-592, cull_gaussians :499-535
// The reference walks the 59 parameter floats + 118 Adam-moment floats of every Gaussian through ~100 torch kernels
// (exp / max / compare / cat / index / nonzero ...), with four host synchronisations (`.sum().item()`, torch.where,
// boolean indexing).  Here:
//   gsr_densify_stats_update : one elementwise kernel per training view;
//   gsr_densify_plan         : one classification kernel + ONE exclusive scan of four counters (CUB) -> per-Gaussian
//                              flags, ranks and the four totals (one 16-byte host read sizes the new set);
//   gsr_densify_apply        : one map kernel (destination row -> source row) + ONE multi-tensor gather kernel that
//                              writes the compacted parameters AND the compacted Adam moments (zeros for new
//                              Gaussians), applying the split transform (mean + R(q) (exp(s) * z), s - log 1.6) on the fly.
// Row order of the result is exactly the reference's: [surviving originals | split samples (sample-major: all first
// samples, then all second samples ...) | duplicates], each in increasing source index.
// HBM-bound: plan 40·N B, apply ≈ (2·4·W + 4)·N' B for W floats per row (W = 177 with both Adam moments).
// Compiled WITHOUT --use_fast_math (threshold decisions must see IEEE exp / log / division; csrc/Makefile).
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "common.cuh"

namespace gsr {

constexpr int DF_SPLIT = 1, DF_DUP = 2, DF_KEEP_ORIG = 4, DF_KEEP_SPLIT = 8, DF_KEEP_DUP = 16;
constexpr unsigned MAP_KIND_SHIFT = 30, MAP_ORIG = 0u, MAP_SPLIT = 1u, MAP_DUP = 2u;

// ------------------------------------------------------------------ after_train (vanilla_gs.py:344-372)
// first != 0  <=>  the three statistics are None in the reference: xys_grad_norm = |grad| for EVERY Gaussian,
// vis_counts = 1 for every Gaussian (:355-357), max_2Dsize = 0 then the visible update (:367-372).
__global__ void __launch_bounds__(256)
densify_stats_kernel(int n, const float *__restrict__ xys_grad, int grad_stride, const int *__restrict__ radii,
                     float max_dim, int first, float *__restrict__ grad_norm, float *__restrict__ vis_counts,
                     float *__restrict__ max_2dsize) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gx = xys_grad[(size_t)i * grad_stride], gy = xys_grad[(size_t)i * grad_stride + 1];
  const float g = sqrtf(gx * gx + gy * gy);
  const int r = radii[i];
  const bool vis = r > 0;
  if (first) {
    grad_norm[i] = g;
    vis_counts[i] = 1.f;
    max_2dsize[i] = vis ? fmaxf(0.f, (float)r / max_dim) : 0.f;
  } else if (vis) {
    vis_counts[i] = vis_counts[i] + 1.f;
    grad_norm[i] = g + grad_norm[i];
    max_2dsize[i] = fmaxf(max_2dsize[i], (float)r / max_dim);
  }
}

// ------------------------------------------------------------------ plan
struct PlanParams {
  int do_densify;
  float max_dim;  // max(last_size): avg_grad_norm = (sum / count) * 0.5 * max_dim
  float densify_grad_thresh, densify_size_thresh;
  int use_split_screen;
  float split_screen_size;
  float cull_alpha_thresh;
  int cull_big;  // step > refine_every * reset_alpha_every
  float cull_scale_thresh;
  int use_cull_screen;
  float cull_screen_size;
};

__device__ __forceinline__ float sigmoidf_exact(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(256)
densify_classify_kernel(int n, const float *__restrict__ scales_raw, const float *__restrict__ opac_raw,
                        const float *__restrict__ grad_norm, const float *__restrict__ vis_counts,
                        const float *__restrict__ max_2dsize, const PlanParams P, uint8_t *__restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float s0 = scales_raw[3 * (size_t)i], s1 = scales_raw[3 * (size_t)i + 1], s2 = scales_raw[3 * (size_t)i + 2];
  const float e0 = expf(s0), e1 = expf(s1), e2 = expf(s2);
  float emax = fmaxf(e0, fmaxf(e1, e2));
  const float m2 = max_2dsize ? max_2dsize[i] : 0.f;
  bool split = false, dup = false;
  if (P.do_densify) {
    const float avg = ((grad_norm[i] / vis_counts[i]) * 0.5f) * P.max_dim;  // :401-405
    const bool high = avg > P.densify_grad_thresh;                           // :406
    split = emax > P.densify_size_thresh;                                    // :407-410
    if (P.use_split_screen) split |= (m2 > P.split_screen_size);             // :411-414
    split &= high;
    if (split) {
      // split_gaussians overwrites the scales of the split Gaussians BEFORE the duplication mask is evaluated
      // (:567-569 then :419-423): a split Gaussian whose reduced scale falls under the threshold is duplicated too.
      const float t0 = logf(e0 / 1.6f), t1 = logf(e1 / 1.6f), t2 = logf(e2 / 1.6f);
      emax = fmaxf(expf(t0), fmaxf(expf(t1), expf(t2)));
    }
    dup = (emax <= P.densify_size_thresh) && high;
  }
  // cull_gaussians (:499-535) on the concatenated set: originals carry their max_2Dsize, new rows carry 0 (:433-441)
  bool cull_new = sigmoidf_exact(opac_raw[i]) < P.cull_alpha_thresh;
  bool cull_orig = cull_new;
  if (P.cull_big) {
    const bool toobig = emax > P.cull_scale_thresh;
    cull_new |= toobig;
    cull_orig |= toobig || (P.use_cull_screen && (m2 > P.cull_screen_size));
  }
  cull_orig |= split;  // splits_mask (:449-461)
  int f = 0;
  if (split) f |= DF_SPLIT;
  if (dup) f |= DF_DUP;
  if (!cull_orig) f |= DF_KEEP_ORIG;
  if (split && !cull_new) f |= DF_KEEP_SPLIT;
  if (dup && !cull_new) f |= DF_KEEP_DUP;
  flags[i] = (uint8_t)f;
}

struct Int4Sum {
  __host__ __device__ __forceinline__ int4 operator()(const int4 &a, const int4 &b) const {
    return make_int4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
};
struct FlagsToCounts {
  __host__ __device__ __forceinline__ int4 operator()(const uint8_t &f) const {
    return make_int4((f & DF_SPLIT) ? 1 : 0, (f & DF_KEEP_ORIG) ? 1 : 0, (f & DF_KEEP_SPLIT) ? 1 : 0,
                     (f & DF_KEEP_DUP) ? 1 : 0);
  }
};
using CountsIter = cub::TransformInputIterator<int4, FlagsToCounts, const uint8_t *>;

__global__ void densify_totals_kernel(int n, const uint8_t *__restrict__ flags, const int4 *__restrict__ ranks,
                                      int *__restrict__ counts) {
  const int4 r = ranks[n - 1];
  const int4 c = FlagsToCounts()(flags[n - 1]);
  counts[0] = r.x + c.x;
  counts[1] = r.y + c.y;
  counts[2] = r.z + c.z;
  counts[3] = r.w + c.w;
}

// ------------------------------------------------------------------ apply
__global__ void __launch_bounds__(256)
densify_map_kernel(int n, int n_samples, int n_keep_orig, int n_keep_split, const uint8_t *__restrict__ flags,
                   const int4 *__restrict__ ranks, unsigned *__restrict__ map) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int f = flags[i];
  if (!(f & (DF_KEEP_ORIG | DF_KEEP_SPLIT | DF_KEEP_DUP))) return;
  const int4 r = ranks[i];
  if (f & DF_KEEP_ORIG) map[r.y] = (unsigned)i | (MAP_ORIG << MAP_KIND_SHIFT);
  if (f & DF_KEEP_SPLIT)
    for (int s = 0; s < n_samples; ++s)
      map[(size_t)n_keep_orig + (size_t)s * n_keep_split + r.z] = (unsigned)i | (MAP_SPLIT << MAP_KIND_SHIFT);
  if (f & DF_KEEP_DUP)
    map[(size_t)n_keep_orig + (size_t)n_samples * n_keep_split + r.w] = (unsigned)i | (MAP_DUP << MAP_KIND_SHIFT);
}

struct GatherSegment {
  const float *src;
  float *dst;
  int width;          // floats per row
  int kind;           // gsr_densify_kind
  int rows_per_cta;
  unsigned magic;     // ceil(2^32 / width)
  int vec_ok;         // dst 16-byte aligned => 128-bit stores
};
struct GatherLaunch {
  GatherSegment seg[GSR_DENSIFY_MAX_TENSORS];
  int num_segments;
  int new_n, n_samples, n_split, n_keep_orig, n_keep_split;
  const unsigned *map;
  const uint8_t *flags;
  const int4 *ranks;
  const float *samples;     // [n_samples * n_split, 3]
  const float *means, *scales_raw, *quats_raw;
};

constexpr int GATHER_MAX_ROWS = 4096;   // rows per CTA (width-1 tensors); wider rows use 4096 / width, at least 32

// One CTA moves `rows_per_cta` destination rows of ONE tensor: the row map is staged in shared memory with one
// coalesced read, then the elements are streamed with destination-contiguous (fully coalesced) writes; the loop is
// specialised per tensor kind and unrolled so that several independent gathers are in flight per thread.
template <int KIND>
__device__ __forceinline__ float gather_value(const GatherLaunch &L, const GatherSegment &S, const unsigned *s_map,
                                              long long row0, int lr, int c) {
  const int w = S.width;
  const unsigned code = s_map[lr];
  const unsigned from = code & ((1u << MAP_KIND_SHIFT) - 1u), kind = code >> MAP_KIND_SHIFT;
  if (KIND == GSR_DENSIFY_ZERO_NEW) return (kind == MAP_ORIG) ? __ldg(S.src + (size_t)from * w + c) : 0.f;
  float val = __ldg(S.src + (size_t)from * w + c);
  if (KIND == GSR_DENSIFY_SCALES) {
    if (L.flags[from] & DF_SPLIT) val = logf(expf(val) / 1.6f);  // :563-569 (children, and the duplicate of a split)
  } else if (KIND == GSR_DENSIFY_MEANS) {
    if (kind == MAP_SPLIT) {
      // new_means = R(q/|q|) (exp(scales) * z) + mean  (:543-553); z = samples[s * n_split + split rank]
      const int s = (int)((row0 + lr - L.n_keep_orig) / L.n_keep_split);
      const float *z = L.samples + ((size_t)s * L.n_split + L.ranks[from].x) * 3;
      const float *sc = L.scales_raw + (size_t)from * 3, *q = L.quats_raw + (size_t)from * 4;
      float R[9];
      quat_to_rotmat(q[0], q[1], q[2], q[3], R);
      const float v0 = expf(sc[0]) * z[0], v1 = expf(sc[1]) * z[1], v2 = expf(sc[2]) * z[2];
      val = (R[3 * c] * v0 + R[3 * c + 1] * v1 + R[3 * c + 2] * v2) + val;
    }
  }
  return val;
}

// The CTA's destination range is contiguous (rows_per_cta rows of width w starting at a 16-byte boundary when the
// tensor is 16-byte aligned: rows_per_cta is a multiple of 4): every thread gathers 4 consecutive destination
// elements (possibly from two source rows) and writes them with one 128-bit store.
template <int KIND>
__device__ __forceinline__ void gather_rows(const GatherLaunch &L, const GatherSegment &S, const unsigned *s_map,
                                            long long row0, int total) {
  const int w = S.width;
  const unsigned magic = S.magic;   // ceil(2^32 / w): j / w == umulhi(j, magic) for j < 2^17, w <= 4096
  float *__restrict__ dst = S.dst + row0 * w;
  const int nvec = S.vec_ok ? (total >> 2) : 0;
#pragma unroll 2
  for (int q = threadIdx.x; q < nvec; q += blockDim.x) {
    const int j = q << 2;
    int lr = (w == 1) ? j : (int)__umulhi((unsigned)j, magic), c = j - lr * w;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[e] = gather_value<KIND>(L, S, s_map, row0, lr, c);
      if (++c == w) { c = 0; ++lr; }
    }
    reinterpret_cast<float4 *>(dst)[q] = make_float4(v[0], v[1], v[2], v[3]);
  }
  for (int j = (nvec << 2) + threadIdx.x; j < total; j += blockDim.x) {
    const int lr = (w == 1) ? j : (int)__umulhi((unsigned)j, magic), c = j - lr * w;
    dst[j] = gather_value<KIND>(L, S, s_map, row0, lr, c);
  }
}

__global__ void __launch_bounds__(256)
densify_gather_kernel(const __grid_constant__ GatherLaunch L) {
  __shared__ unsigned s_map[GATHER_MAX_ROWS];
  const GatherSegment &S = L.seg[blockIdx.y];
  const long long row0 = (long long)blockIdx.x * S.rows_per_cta;
  if (row0 >= L.new_n) return;
  const int rows = (int)min((long long)S.rows_per_cta, (long long)L.new_n - row0);
  for (int i = threadIdx.x; i < rows; i += blockDim.x) s_map[i] = L.map[row0 + i];
  __syncthreads();
  const int total = rows * S.width;
  switch (S.kind) {
    case GSR_DENSIFY_ZERO_NEW: gather_rows<GSR_DENSIFY_ZERO_NEW>(L, S, s_map, row0, total); break;
    case GSR_DENSIFY_SCALES: gather_rows<GSR_DENSIFY_SCALES>(L, S, s_map, row0, total); break;
    case GSR_DENSIFY_MEANS: gather_rows<GSR_DENSIFY_MEANS>(L, S, s_map, row0, total); break;
    default: gather_rows<GSR_DENSIFY_COPY>(L, S, s_map, row0, total); break;
  }
}

}  // namespace gsr

extern "C" {

GSR_API int gsr_densify_stats_update(int num_points, const float *xys_grad, int xys_grad_stride, const int32_t *radii,
                                     float max_dim, int first, float *xys_grad_norm, float *vis_counts,
                                     float *max_2Dsize, void *stream) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_densify_stats_update");
  GSR_REQUIRE(num_points >= 0 && xys_grad_stride >= 2 && max_dim > 0.f, GSR_ERR_INVALID_ARGUMENT,
              "densify_stats_update: bad sizes (N=%d, stride=%d, max_dim=%g)", num_points, xys_grad_stride, max_dim);
  if (num_points == 0) return GSR_OK;
  GSR_REQUIRE(xys_grad && radii && xys_grad_norm && vis_counts && max_2Dsize, GSR_ERR_INVALID_ARGUMENT,
              "densify_stats_update: null pointer");
  densify_stats_kernel<<<cdiv(num_points, 256), 256, 0, (cudaStream_t)stream>>>(
      num_points, xys_grad, xys_grad_stride, radii, max_dim, first, xys_grad_norm, vis_counts, max_2Dsize);
  GSR_CHECK_LAUNCH("densify_stats_kernel");
  return GSR_OK;
}

GSR_API size_t gsr_densify_plan_workspace_bytes(int num_points) {
  using namespace gsr;
  size_t tmp = 0;
  CountsIter it((const uint8_t *)nullptr, FlagsToCounts());
  cub::DeviceScan::ExclusiveScan((void *)nullptr, tmp, it, (int4 *)nullptr, Int4Sum(), make_int4(0, 0, 0, 0),
                                 num_points > 0 ? num_points : 1);
  return tmp + 256;
}

GSR_API int gsr_densify_plan(int num_points, const float *scales_raw, const float *opacities_raw,
                             const float *xys_grad_norm, const float *vis_counts, const float *max_2Dsize,
                             int do_densify, float max_dim, float densify_grad_thresh, float densify_size_thresh,
                             int use_split_screen, float split_screen_size, float cull_alpha_thresh, int cull_big,
                             float cull_scale_thresh, int use_cull_screen, float cull_screen_size, uint8_t *flags,
                             int32_t *ranks, int32_t *counts, void *workspace, size_t workspace_bytes, void *stream) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_densify_plan");
  GSR_REQUIRE(num_points >= 1, GSR_ERR_INVALID_ARGUMENT, "densify_plan: num_points must be >= 1 (got %d)", num_points);
  GSR_REQUIRE(scales_raw && opacities_raw && flags && ranks && counts && workspace, GSR_ERR_INVALID_ARGUMENT,
              "densify_plan: null pointer");
  GSR_REQUIRE(!do_densify || (xys_grad_norm && vis_counts && max_2Dsize), GSR_ERR_INVALID_ARGUMENT,
              "densify_plan: densification needs the three running statistics (vanilla_gs.py:396-400)");
  GSR_REQUIRE(!(use_cull_screen || use_split_screen) || max_2Dsize, GSR_ERR_INVALID_ARGUMENT,
              "densify_plan: screen-size tests need max_2Dsize (vanilla_gs.py:521)");
  GSR_REQUIRE(((uintptr_t)ranks & 15u) == 0, GSR_ERR_INVALID_ARGUMENT, "densify_plan: ranks must be 16-byte aligned");
  PlanParams P;
  P.do_densify = do_densify;
  P.max_dim = max_dim;
  P.densify_grad_thresh = densify_grad_thresh;
  P.densify_size_thresh = densify_size_thresh;
  P.use_split_screen = do_densify && use_split_screen;
  P.split_screen_size = split_screen_size;
  P.cull_alpha_thresh = cull_alpha_thresh;
  P.cull_big = cull_big;
  P.cull_scale_thresh = cull_scale_thresh;
  P.use_cull_screen = use_cull_screen;
  P.cull_screen_size = cull_screen_size;
  cudaStream_t st = (cudaStream_t)stream;
  densify_classify_kernel<<<cdiv(num_points, 256), 256, 0, st>>>(num_points, scales_raw, opacities_raw, xys_grad_norm,
                                                                  vis_counts, max_2Dsize, P, flags);
  GSR_CHECK_LAUNCH("densify_classify_kernel");
  const size_t need = gsr_densify_plan_workspace_bytes(num_points);
  GSR_REQUIRE(workspace_bytes >= need, GSR_ERR_WORKSPACE, "densify_plan: workspace too small (%zu < %zu)",
              workspace_bytes, need);
  void *tmp = (void *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  size_t tmp_bytes = workspace_bytes - ((uintptr_t)tmp - (uintptr_t)workspace);
  CountsIter it(flags, FlagsToCounts());
  GSR_CUDA(cub::DeviceScan::ExclusiveScan(tmp, tmp_bytes, it, reinterpret_cast<int4 *>(ranks), Int4Sum(),
                                          make_int4(0, 0, 0, 0), num_points, st));
  densify_totals_kernel<<<1, 1, 0, st>>>(num_points, flags, reinterpret_cast<const int4 *>(ranks), counts);
  GSR_CHECK_LAUNCH("densify_totals_kernel");
  return GSR_OK;
}

GSR_API int gsr_densify_apply(int num_points, int n_split_samples, const int32_t *counts_host, const uint8_t *flags,
                              const int32_t *ranks, const float *samples, const float *means, const float *scales_raw,
                              const float *quats_raw, int num_tensors, const float *const *src_host,
                              float *const *dst_host, const int32_t *widths_host, const int32_t *kinds_host,
                              uint32_t *map, void *stream) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_densify_apply");
  GSR_REQUIRE(num_points >= 1 && n_split_samples >= 1, GSR_ERR_INVALID_ARGUMENT, "densify_apply: bad sizes");
  GSR_REQUIRE(counts_host && flags && ranks && src_host && dst_host && widths_host && kinds_host,
              GSR_ERR_INVALID_ARGUMENT, "densify_apply: null pointer");
  GSR_REQUIRE(num_tensors >= 0 && num_tensors <= GSR_DENSIFY_MAX_TENSORS, GSR_ERR_INVALID_ARGUMENT,
              "densify_apply: num_tensors must be in [0,%d] (got %d)", GSR_DENSIFY_MAX_TENSORS, num_tensors);
  const int n_split = counts_host[0], n_keep_orig = counts_host[1], n_keep_split = counts_host[2],
            n_keep_dup = counts_host[3];
  GSR_REQUIRE(n_split >= 0 && n_keep_orig >= 0 && n_keep_split >= 0 && n_keep_dup >= 0 && n_split <= num_points &&
                  n_keep_orig <= num_points && n_keep_split <= n_split && n_keep_dup <= num_points,
              GSR_ERR_INVALID_ARGUMENT, "densify_apply: inconsistent counts");
  const long long new_n = (long long)n_keep_orig + (long long)n_split_samples * n_keep_split + n_keep_dup;
  GSR_REQUIRE(new_n < (1ll << MAP_KIND_SHIFT), GSR_ERR_UNSUPPORTED, "densify_apply: more than 2^30 Gaussians");
  if (new_n == 0) return GSR_OK;
  GSR_REQUIRE(map, GSR_ERR_INVALID_ARGUMENT, "densify_apply: null map");
  cudaStream_t st = (cudaStream_t)stream;
  densify_map_kernel<<<cdiv(num_points, 256), 256, 0, st>>>(num_points, n_split_samples, n_keep_orig, n_keep_split,
                                                             flags, reinterpret_cast<const int4 *>(ranks), map);
  GSR_CHECK_LAUNCH("densify_map_kernel");
  if (num_tensors == 0) return GSR_OK;
  GatherLaunch L;
  L.num_segments = num_tensors;
  L.new_n = (int)new_n;
  L.n_samples = n_split_samples;
  L.n_split = n_split;
  L.n_keep_orig = n_keep_orig;
  L.n_keep_split = n_keep_split > 0 ? n_keep_split : 1;
  L.map = map;
  L.flags = flags;
  L.ranks = reinterpret_cast<const int4 *>(ranks);
  L.samples = samples;
  L.means = means;
  L.scales_raw = scales_raw;
  L.quats_raw = quats_raw;
  unsigned grid_x = 1;
  for (int k = 0; k < num_tensors; ++k) {
    GSR_REQUIRE(src_host[k] && dst_host[k] && widths_host[k] >= 1 && widths_host[k] <= 4096, GSR_ERR_INVALID_ARGUMENT,
                "densify_apply: bad tensor %d (width %d)", k, widths_host[k]);
    GSR_REQUIRE(kinds_host[k] >= GSR_DENSIFY_COPY && kinds_host[k] <= GSR_DENSIFY_ZERO_NEW, GSR_ERR_INVALID_ARGUMENT,
                "densify_apply: bad kind %d for tensor %d", kinds_host[k], k);
    if (kinds_host[k] == GSR_DENSIFY_MEANS) {
      GSR_REQUIRE(widths_host[k] == 3, GSR_ERR_INVALID_ARGUMENT, "densify_apply: the means tensor must have width 3");
      GSR_REQUIRE(n_keep_split == 0 || (samples && means && scales_raw && quats_raw), GSR_ERR_INVALID_ARGUMENT,
                  "densify_apply: the split transform needs samples, means, scales_raw and quats_raw");
    }
    GatherSegment &S = L.seg[k];
    S.src = src_host[k];
    S.dst = dst_host[k];
    S.width = widths_host[k];
    S.kind = kinds_host[k];
    const int rpc = (GATHER_MAX_ROWS / S.width) & ~3;   // multiple of 4 rows => 16-byte aligned CTA ranges
    S.rows_per_cta = rpc < 32 ? 32 : rpc;
    S.vec_ok = ((uintptr_t)S.dst & 15u) == 0;
    S.magic = S.width == 1 ? 0u : (unsigned)((0x100000000ull + (unsigned)S.width - 1) / (unsigned)S.width);
    const unsigned gx = (unsigned)((new_n + S.rows_per_cta - 1) / S.rows_per_cta);
    grid_x = gx > grid_x ? gx : grid_x;
  }
  densify_gather_kernel<<<dim3(grid_x, (unsigned)num_tensors, 1), 256, 0, st>>>(L);
  GSR_CHECK_LAUNCH("densify_gather_kernel");
  return GSR_OK;
}
}
