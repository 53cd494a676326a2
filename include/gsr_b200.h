/*
 * gsr_b200.h — C ABI of the B200-native Gaussian-splatting rasterizer (libgsr_b200.so).
 *
 * This is the drop-in boundary for the hot path of Gaussian-Splatting-Toolkit's `rasterizer` package.
 * Every entry point below replaces one pybind binding of the reference's native module
 * (gs_toolkit/gs_components/rasterizer/cuda/csrc/ext.cpp:6-17, C++ signatures in bindings.h:19-115),
 * or one ATen library call the reference's Python wrapper makes between those bindings
 * (rasterizer/utils.py:123 torch.cumsum, :179-180 torch.sort + torch.gather).
 *
 * Conventions
 *  - Plain C: pointers + sizes only. No torch / pybind types cross this boundary.
 *  - Unless a name ends in `_host`, every pointer is a DEVICE pointer (cudaMalloc'd or carved from the
 *    caller's allocator, e.g. torch's caching allocator) valid on the current CUDA device, and `stream`
 *    is a `cudaStream_t` passed as `void*` (NULL = legacy default stream).  Calls are asynchronous with
 *    respect to the host: they enqueue work on `stream` and return.
 *  - All floating point is IEEE binary32, all arrays dense row-major ("contiguous"):
 *    means3d/scales [N,3], quats [N,4] (w,x,y,z), viewmat row-major 3x4 or 4x4 (first 12 floats used),
 *    projmat row-major 4x4 (= P * V), xys [N,2], conics [N,3] (a,b,c upper-triangular inverse cov2d),
 *    cov3d [N,6] upper-triangular, colors [N,C], opacities [N], images [H,W,C], final_Ts / final_idx [H,W].
 *    radii, num_tiles_hit, cum_tiles_hit, gaussian_ids, tile_bins [T,2], final_idx are int32;
 *    isect_ids are int64 = (tile_id << 32) | bits(depth).
 *  - Outputs are fully written by the call (no pre-zeroing by the caller is required; where the
 *    reference relies on torch::zeros the call zero-fills on `stream` itself).
 *  - Return value: 0 (GSR_OK) on success, a negative gsr_status otherwise; gsr_last_error() returns a
 *    thread-local human-readable message for the last failure.  Launch errors are checked
 *    (cudaGetLastError) after every launch — the reference never checks.
 *  - Thread-safety: no global mutable state besides the thread-local error string; re-entrant per device.
 */
#ifndef GSR_B200_H
#define GSR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define GSR_API __declspec(dllexport)
#else
#define GSR_API __attribute__((visibility("default")))
#endif

typedef enum gsr_status {
  GSR_OK = 0,
  GSR_ERR_INVALID_ARGUMENT = -1, /* bad sizes / block_width outside [2,16] / null pointer */
  GSR_ERR_CUDA = -2,             /* a CUDA runtime call or kernel launch failed */
  GSR_ERR_WORKSPACE = -3,        /* workspace too small */
  GSR_ERR_UNSUPPORTED = -4       /* e.g. SH degree > 4, channels > GSR_MAX_CHANNELS */
} gsr_status;

#define GSR_MAX_CHANNELS 32

/* library identification */
GSR_API const char *gsr_version(void);      /* "0.1.2+b200.<n>": tracks rasterizer/version.py:1 */
GSR_API const char *gsr_last_error(void);   /* thread-local message of the last non-zero return */
GSR_API int gsr_built_for_sm(void);         /* 100 : the only arch in the fatbin is sm_100a */
/* number of kernels of THIS library launched by the process so far (every launch site is counted after its
 * cudaGetLastError check; library kernels such as CUB's and memsets are not) — bench.py's `gpu_launches` */
GSR_API unsigned long long gsr_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Spherical harmonics  — replaces compute_sh_forward / compute_sh_backward
 * (bindings.h:23-33, bindings.cu:58-103, kernels sh.cuh:33-224)
 * degree = degree of the stored coefficients (K = (degree+1)^2 bases, degree <= 4),
 * degrees_to_use <= degree.  coeffs [N,K,3], viewdirs [N,3] (normalised inside), colors [N,3].
 * Backward writes ALL K*3 entries of v_coeffs (zeros for bases above degrees_to_use).
 * ---------------------------------------------------------------------------------------------- */
GSR_API int gsr_compute_sh_forward(int num_points, int degree, int degrees_to_use,
                                   const float *viewdirs, const float *coeffs, float *colors,
                                   void *stream);
GSR_API int gsr_compute_sh_backward(int num_points, int degree, int degrees_to_use,
                                    const float *viewdirs, const float *v_colors, float *v_coeffs,
                                    void *stream);

/* Multi-view SH adjoint — the compute half of the view-parallel gradient exchange (no counterpart in the
 * reference, whose multi-GPU path is torch DDP's all-reduce of every parameter gradient,
 * gs_toolkit/pipelines/base_pipeline.py:202-207):  v_coeffs[n,k,c] = sum_v Y_k(means3d[n] - cam_v) v_colors_v[n,c].
 * Ranks exchange the 3-float colour gradients (12 B / Gaussian) instead of all-reducing the 3K-float SH gradients
 * (192 B / Gaussian at degree 3); every rank then evaluates all views' outer products while writing v_coeffs once.
 * cam_positions [V,3] is a DEVICE array (e.g. the all-gathered camera centres); v_colors_views_host [V] is a HOST
 * array (read before return) of V DEVICE pointers, each to [N,3] floats — local memory, slices of an all-gathered
 * buffer, or peer-GPU memory mapped over NVLink.  1 <= num_views <= 16. */
GSR_API int gsr_compute_sh_backward_multiview(int num_points, int degree, int degrees_to_use, int num_views,
                                              const float *means3d, const float *cam_positions,
                                              const float *const *v_colors_views_host, float *v_coeffs,
                                              void *stream);
/* Same, with the camera centres also addressed through a HOST table of V DEVICE pointers (3 floats each): this is
 * the form used when every rank publishes {v_colors, camera centre} in NVLink-mapped symmetric memory and the kernel
 * loads the peers' data directly (fused all-gather + adjoint, no staging copy). */
GSR_API int gsr_compute_sh_backward_multiview_ptrs(int num_points, int degree, int degrees_to_use, int num_views,
                                                   const float *means3d, const float *const *cam_views_host,
                                                   const float *const *v_colors_views_host, float *v_coeffs,
                                                   void *stream);

/* All-reduce (sum) of a replicated FP32 buffer over NVLink peer memory, in two kernels (multi-GPU plumbing of SURVEY §8(e);
 * the reference's counterpart is torch DDP's bucketed all-reduce, pipelines/base_pipeline.py:202-207).  bufs_host[w] is
 * rank w's buffer as mapped into THIS process (symmetric memory), all of num_floats floats and 16-byte aligned.
 * reduce_scatter: slice `rank` of this rank's buffer <- sum over ranks of that slice (fixed order: identical replicas);
 * all_gather: the other slices of this rank's buffer <- the owners' reduced slices.  The caller puts a barrier over all
 * ranks before, between and after the two calls (rasterizer/view_parallel.py). */
GSR_API int gsr_peer_reduce_scatter(int world, int rank, float *const *bufs_host, long long num_floats, void *stream);
GSR_API int gsr_peer_all_gather(int world, int rank, float *const *bufs_host, long long num_floats, void *stream);
/* Push-based forms (remote stores are posted; reductions read local memory).  peer_push: for every rank w,
 * dsts_host[w][0 .. n_w) <- src[w * src_stride ..): src_stride = 0 broadcasts n_per_dst floats, src_stride > 0 scatters
 * slice w (n_w = min(n_per_dst, n_total - w * src_stride)).  peer_reduce_broadcast: sums the `world` local staging slots
 * (slot w at slots + w * slot_stride, fixed order) and stores the total into dsts_host[w] of every rank.  Pointers 16-byte
 * aligned, strides multiples of 4 floats; barriers between the steps are the caller's (rasterizer/view_parallel.py). */
GSR_API int gsr_peer_push(int world, float *const *dsts_host, const float *src, long long src_stride, long long n_per_dst,
                          long long n_total, void *stream);
GSR_API int gsr_peer_reduce_broadcast(int world, float *const *dsts_host, const float *slots, long long slot_stride,
                                      long long num_floats, void *stream);

/* ------------------------------------------------------------------------------------------------
 * EWA projection — replaces project_gaussians_forward / project_gaussians_backward
 * (bindings.h:35-55, bindings.cu:105-216, kernels forward.cu:13-90, backward.cu:305-453)
 * Forward outputs: cov3d [N,6], xys [N,2], depths [N], radii [N] i32, conics [N,3], compensation [N],
 * num_tiles_hit [N] i32.  Culled Gaussians get radii = num_tiles_hit = 0 and zeros elsewhere, except
 * that cov3d / conics are written as far as the reference kernel writes them before its early returns
 * (forward.cu:46-47,68).
 * Backward outputs: v_cov2d [N,3], v_cov3d [N,6] (returned by the reference binding, discarded by its
 * Python wrapper; may be NULL here), v_mean3d [N,3], v_scale [N,3], v_quat [N,4]; zero where radii<=0.
 * ---------------------------------------------------------------------------------------------- */
GSR_API int gsr_project_gaussians_forward(int num_points, const float *means3d, const float *scales,
                                          float glob_scale, const float *quats, const float *viewmat,
                                          const float *projmat, float fx, float fy, float cx, float cy,
                                          unsigned img_height, unsigned img_width, unsigned block_width,
                                          float clip_thresh, float *cov3d, float *xys, float *depths,
                                          int32_t *radii, float *conics, float *compensation,
                                          int32_t *num_tiles_hit, void *stream);
GSR_API int gsr_project_gaussians_backward(int num_points, const float *means3d, const float *scales,
                                           float glob_scale, const float *quats, const float *viewmat,
                                           const float *projmat, float fx, float fy, float cx, float cy,
                                           unsigned img_height, unsigned img_width, const float *cov3d,
                                           const int32_t *radii, const float *conics,
                                           const float *compensation, const float *v_xy,
                                           const float *v_depth, const float *v_conic,
                                           const float *v_compensation, float *v_cov2d /*nullable*/,
                                           float *v_cov3d /*nullable*/, float *v_mean3d, float *v_scale,
                                           float *v_quat, void *stream);

/* compute_cov2d_bounds (bindings.h:19-21, bindings.cu:19-56): cov2d [N,3] -> conics [N,3], radii [N] f32 */
GSR_API int gsr_compute_cov2d_bounds(int num_pts, const float *covs2d, float *conics, float *radii,
                                     void *stream);

/* ------------------------------------------------------------------------------------------------
 * Tile binning
 * gsr_cumsum_tiles_hit   replaces torch.cumsum(num_tiles_hit, int32) (rasterizer/utils.py:123).
 *                        If total_host_pinned != NULL the total M is also written there (a pinned
 *                        HOST int32 the caller reads after synchronising `stream`; replaces `.item()`).
 * gsr_map_gaussian_to_intersects  replaces the binding of the same name (bindings.h:57-61,
 *                        bindings.cu:218-251, kernel forward.cu:94-127).
 * gsr_sort_intersects    replaces torch.sort(int64) + torch.gather (rasterizer/utils.py:179-180):
 *                        stable ascending radix sort of the 64-bit keys carrying the int32 Gaussian ids;
 *                        only the bits that can be set (32 depth bits + ceil(log2(num_tiles)) tile bits)
 *                        are sorted.
 * gsr_get_tile_bin_edges replaces the binding of the same name (bindings.h:63-66, bindings.cu:253-267,
 *                        kernel forward.cu:132-154); tile_bins [num_tiles,2] is zero-filled first.
 * Workspace: call the *_workspace_bytes query, hand in a device buffer of at least that size.
 * ---------------------------------------------------------------------------------------------- */
GSR_API size_t gsr_cumsum_workspace_bytes(int num_points);
GSR_API int gsr_cumsum_tiles_hit(int num_points, const int32_t *num_tiles_hit, int32_t *cum_tiles_hit,
                                 int32_t *total_host_pinned /*nullable*/, void *workspace,
                                 size_t workspace_bytes, void *stream);
GSR_API int gsr_map_gaussian_to_intersects(int num_points, int num_intersects, const float *xys,
                                           const float *depths, const int32_t *radii,
                                           const int32_t *cum_tiles_hit, unsigned tiles_x,
                                           unsigned tiles_y, unsigned block_width, int64_t *isect_ids,
                                           int32_t *gaussian_ids, void *stream);
/* Exact tile culling (no counterpart in the reference; used inside rasterize_gaussians in place of
 * num_tiles_hit + gsr_map_gaussian_to_intersects).  Of the tiles in the reference's bounding box
 * (helpers.cuh:11-34) only those are kept in which at least one pixel can reach alpha >= 1/255
 * (forward.cu:360-363) — the reference `continue`s on every pixel of the dropped (Gaussian, tile) pairs, so
 * images, final_Ts and all gradients are unchanged while the sort and both blend kernels see fewer pairs.
 * gsr_count_tiles_tight writes the per-Gaussian kept-tile count; after gsr_cumsum_tiles_hit on those counts
 * gsr_map_gaussian_to_intersects_tight emits keys / ids for exactly the kept pairs, in the reference's order. */
GSR_API int gsr_count_tiles_tight(int num_points, const float *xys, const int32_t *radii, const float *conics,
                                  const float *opacities, unsigned img_height, unsigned img_width,
                                  unsigned block_width, int32_t *tiles_touched, void *stream);
GSR_API int gsr_map_gaussian_to_intersects_tight(int num_points, int num_intersects, const float *xys,
                                                 const float *depths, const int32_t *radii, const float *conics,
                                                 const float *opacities, const int32_t *cum_tiles_touched,
                                                 unsigned img_height, unsigned img_width, unsigned block_width,
                                                 int64_t *isect_ids, int32_t *gaussian_ids, void *stream);
/* Internal binning of rasterize_gaussians (replaces, as a whole, the reference's forward orchestration
 * cumsum -> .item() -> map_gaussian_to_intersects -> torch.sort(int64) -> torch.gather -> get_tile_bin_edges,
 * rasterizer/rasterize.py:106-138 + utils.py:106-182).  Same per-tile order as the reference (depth ascending, ties in
 * Gaussian-index order), with exact tile culling as above, obtained WITHOUT a global sort: pairs are counted per tile
 * (atomics), the tile counts are scanned into tile_bins, the pairs are written unordered into their tile's segment as
 * 64-bit keys (depth bits << 32 | Gaussian id) and every segment is sorted on its own (one CTA per tile, shared
 * memory; segments above 4096 pairs fall back to a radix sort in global memory).  csrc/binning_device.cu.
 *
 * gsr_bin_gaussians_device — the whole binning in ONE asynchronous call; the pair count M stays on the DEVICE: no kernel
 *   grid depends on it, the call neither synchronises nor allocates, so a whole view can be captured in a CUDA graph
 *   (the reference's `.item()`, utils.py:124, is the host read this removes).
 *     gaussian_ids_sorted [capacity] i32 (entries >= min(M, capacity) unspecified), tile_bins [T,2] i32,
 *     meta  DEVICE int32[4] = {M, overflow (M > capacity), min(M, capacity), 0}; if meta_host_pinned != NULL the four
 *     words are also copied there on `stream` (inspect them after an event / at the next call: `overflow` means the lists
 *     were truncated and the caller must repeat the view with a larger capacity).
 * gsr_bin_count + gsr_bin_fill_sort — the same in two calls for callers that want exactly-sized outputs: after
 *   gsr_bin_count the caller synchronises, reads M = meta_host_pinned[0], allocates gaussian_ids_sorted [M] and a fill
 *   workspace, and calls gsr_bin_fill_sort with the SAME count workspace (it holds the culling masks and fill cursors). */
GSR_API size_t gsr_bin_device_workspace_bytes(int num_points, int capacity, unsigned img_height, unsigned img_width,
                                              unsigned block_width);
GSR_API int gsr_bin_gaussians_device(int num_points, const float *xys, const float *depths, const int32_t *radii,
                                     const float *conics, const float *opacities, unsigned img_height,
                                     unsigned img_width, unsigned block_width, int capacity,
                                     int32_t *gaussian_ids_sorted, int32_t *tile_bins, int32_t *meta,
                                     int32_t *meta_host_pinned /*nullable*/, void *workspace, size_t workspace_bytes,
                                     void *stream);
GSR_API size_t gsr_bin_count_workspace_bytes(int num_points, unsigned img_height, unsigned img_width,
                                             unsigned block_width);
GSR_API int gsr_bin_count(int num_points, const float *xys, const int32_t *radii, const float *conics,
                          const float *opacities, unsigned img_height, unsigned img_width, unsigned block_width,
                          int32_t *tile_bins, int32_t *meta, int32_t *meta_host_pinned /*nullable*/,
                          void *count_workspace, size_t workspace_bytes, void *stream);
GSR_API size_t gsr_bin_fill_workspace_bytes(int num_intersects);
GSR_API int gsr_bin_fill_sort(int num_points, int num_intersects, const float *xys, const float *depths,
                              const int32_t *radii, const float *conics, const float *opacities,
                              unsigned img_height, unsigned img_width, unsigned block_width, const int32_t *tile_bins,
                              void *count_workspace, int32_t *gaussian_ids_sorted, void *fill_workspace,
                              size_t fill_workspace_bytes, void *stream);
GSR_API size_t gsr_sort_workspace_bytes(int num_intersects);
GSR_API int gsr_sort_intersects(int num_intersects, int num_tiles, const int64_t *isect_ids,
                                const int32_t *gaussian_ids, int64_t *isect_ids_sorted,
                                int32_t *gaussian_ids_sorted, void *workspace, size_t workspace_bytes,
                                void *stream);
GSR_API int gsr_get_tile_bin_edges(int num_intersects, const int64_t *isect_ids_sorted, int num_tiles,
                                   int32_t *tile_bins, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Per-tile alpha compositing and its adjoint — replaces rasterize_forward / rasterize_backward
 * (bindings.h:68-115, bindings.cu:269-328,471-528, kernels forward.cu:278-395, backward.cu:133-303)
 * and the N-D variants nd_rasterize_forward / nd_rasterize_backward (bindings.cu:330-469,
 * kernels forward.cu:159-276, backward.cu:23-131).
 * Forward outputs: out_img [H,W,C], final_Ts [H,W], final_idx [H,W] i32 (absolute index into the sorted
 * list of the last contributing Gaussian).
 * Backward outputs: v_xy [N,2], v_conic [N,3], v_colors [N,C], v_opacity [N] — zero-filled by the call
 * and accumulated with block-reduced atomics.
 * The 3-channel calls compute in FP32.  The N-D calls (C != 3, 1 <= C <= GSR_MAX_CHANNELS) reproduce the
 * reference's numerics: binary16 colour accumulators and a backward that excludes the last contributor
 * (backward.cu:64-65) — see DESIGN.md "N-D variants".
 * ---------------------------------------------------------------------------------------------- */
GSR_API int gsr_rasterize_forward(unsigned img_height, unsigned img_width, unsigned block_width,
                                  int num_points, const int32_t *gaussian_ids_sorted,
                                  const int32_t *tile_bins, const float *xys, const float *conics,
                                  const float *colors, const float *opacities, const float *background,
                                  float *out_img, float *final_Ts, int32_t *final_idx, void *stream);
GSR_API int gsr_rasterize_backward(unsigned img_height, unsigned img_width, unsigned block_width,
                                   int num_points, const int32_t *gaussian_ids_sorted,
                                   const int32_t *tile_bins, const float *xys, const float *conics,
                                   const float *colors, const float *opacities, const float *background,
                                   const float *final_Ts, const int32_t *final_idx,
                                   const float *v_output, const float *v_output_alpha, float *v_xy,
                                   float *v_conic, float *v_colors, float *v_opacity, void *stream);
GSR_API int gsr_nd_rasterize_forward(unsigned img_height, unsigned img_width, unsigned block_width,
                                     unsigned channels, int num_points,
                                     const int32_t *gaussian_ids_sorted, const int32_t *tile_bins,
                                     const float *xys, const float *conics, const float *colors,
                                     const float *opacities, const float *background, float *out_img,
                                     float *final_Ts, int32_t *final_idx, void *stream);
GSR_API int gsr_nd_rasterize_backward(unsigned img_height, unsigned img_width, unsigned block_width,
                                      unsigned channels, int num_points,
                                      const int32_t *gaussian_ids_sorted, const int32_t *tile_bins,
                                      const float *xys, const float *conics, const float *colors,
                                      const float *opacities, const float *background,
                                      const float *final_Ts, const int32_t *final_idx,
                                      const float *v_output, const float *v_output_alpha, float *v_xy,
                                      float *v_conic, float *v_colors, float *v_opacity, void *stream);

/* ------------------------------------------------------------------------------------------------
 * FUSED render operator (SURVEY 8(f1) — the model-side glue folded into the operator; no single counterpart in the
 * reference: it replaces the sequence gs_toolkit/models/vanilla_gs.py:759-855 issues per view).
 *
 * gsr_fused_preprocess_forward: per Gaussian, from the RAW model parameters — exp(scales_raw), normalised
 *   quats_raw, sigmoid(opacities_raw), SH colour from {features_dc [N,3], features_rest [N,K-1,3]} in direction
 *   means3d - camera centre, clamp(rgb + 0.5, min=0), EWA projection — to packed blend records
 *   `records` [3][N] float4 planes {x,y,ext_x,ext_y | A,B,C,opacity | r,g,b,depth} plus xys [N,2], depths [N],
 *   radii [N] i32, conics [N,3], opacities [N] (activated), clamp_mask [N] i32 (bit c set iff rgb_c + 0.5 > 0).
 * gsr_blend_packed_forward / _backward: compositing of RGB and (if out_depth / v_output_depth != NULL) depth as a
 *   fourth channel in the same pass (the reference models run a second rasterize_gaussians call for depth,
 *   vanilla_gs.py:839-855); the adjoint accumulates grad_records [N,12] =
 *   {v_x, v_y, v_opacity, v_depth | v_a, v_b, v_c, - | v_r, v_g, v_b, -} (zero-filled by the call).
 * gsr_fused_preprocess_backward: grad_records (+ optional v_xys_extra [N,2]) -> gradients of the six raw parameter
 *   tensors (chain rules of exp / normalise / sigmoid / clamp included).
 * compensation [N] (nullable): non-NULL selects rasterize_mode = "antialiased" (vanilla_gs.py:813-816): the forward
 *   stores the EWA compensation factor there and uses opacity = sigmoid(raw) * compensation; the backward (same
 *   array) routes d/d compensation through the projection adjoint.  NULL = "classic".
 * ---------------------------------------------------------------------------------------------- */
GSR_API int gsr_fused_preprocess_forward(int num_points, int sh_degree, int degrees_to_use, const float *means3d,
                                         const float *scales_raw, const float *quats_raw, const float *opacities_raw,
                                         const float *features_dc, const float *features_rest, const float *viewmat,
                                         const float *projmat, float glob_scale, float fx, float fy, float cx, float cy,
                                         unsigned img_height, unsigned img_width, unsigned block_width,
                                         float clip_thresh, float *records, float *xys, float *depths, int32_t *radii,
                                         float *conics, float *opacities, int32_t *clamp_mask,
                                         float *compensation /*nullable*/, void *stream);
GSR_API int gsr_fused_preprocess_backward(int num_points, int sh_degree, int degrees_to_use, const float *means3d,
                                          const float *scales_raw, const float *quats_raw, const float *opacities_raw,
                                          const float *viewmat, const float *projmat, float glob_scale, float fx,
                                          float fy, unsigned img_height, unsigned img_width, const int32_t *radii,
                                          const float *conics, const int32_t *clamp_mask,
                                          const float *compensation /*nullable*/, const float *grad_records,
                                          const float *v_xys_extra /*nullable*/, float *v_means3d, float *v_scales_raw,
                                          float *v_quats_raw, float *v_opacities_raw, float *v_features_dc,
                                          float *v_features_rest, void *stream);
GSR_API int gsr_blend_packed_forward(unsigned img_height, unsigned img_width, unsigned block_width, int num_points,
                                     const int32_t *gaussian_ids_sorted, const int32_t *tile_bins,
                                     const float *records, const float *background, float *out_img,
                                     float *out_depth /*nullable*/, float *final_Ts, int32_t *final_idx, void *stream);
GSR_API int gsr_blend_packed_backward(unsigned img_height, unsigned img_width, unsigned block_width, int num_points,
                                      const int32_t *gaussian_ids_sorted, const int32_t *tile_bins,
                                      const float *records, const float *background, const float *final_Ts,
                                      const int32_t *final_idx, const float *v_output,
                                      const float *v_output_depth /*nullable*/, const float *v_output_alpha,
                                      float *grad_records, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Fused photometric loss — replaces gs_toolkit/models/vanilla_gs.py:926-934:
 *   loss = (1 - ssim_lambda) * mean|gt - pred| + ssim_lambda * (1 - SSIM(gt, pred))
 * with SSIM = pytorch_msssim.SSIM(data_range=1.0, size_average=True, channel=3) (third-party, not vendored in the
 * reference; algorithm restated in csrc/loss.cu).  pred, gt: [H,W,3] (the rasterizer's layout), H, W >= 11.
 *   gsr_l1_ssim_forward : maps [3][H-10][W-10][3] (derivative maps kept for the adjoint), partials [n][2] scratch with
 *                         n = gsr_l1_ssim_num_partials(H, W) (per-CTA {sum SSIM, sum |pred-gt|}, reduced in FP64 by a
 *                         one-CTA finalisation kernel => deterministic), loss_l1_ssim [3] = {loss, L1, SSIM}.
 *   gsr_l1_ssim_backward: v_pred [H,W,3] = v_loss * d loss / d pred  (v_loss: device scalar, NULL = 1).
 * ---------------------------------------------------------------------------------------------- */
GSR_API int gsr_l1_ssim_num_partials(unsigned img_height, unsigned img_width);
GSR_API int gsr_l1_ssim_forward(unsigned img_height, unsigned img_width, float ssim_lambda, const float *pred,
                                const float *gt, float *maps, float *partials, float *loss_l1_ssim, void *stream);
GSR_API int gsr_l1_ssim_backward(unsigned img_height, unsigned img_width, float ssim_lambda, const float *pred,
                                 const float *gt, const float *maps, const float *v_loss /*nullable: 1.0*/,
                                 float *v_pred, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Optimizer step — replaces the six torch.optim.Adam(lr, eps=1e-15).step() calls of
 * gs_toolkit/engine/optimizers.py:173-180 (groups / learning rates: configs/method_configs.py:98-125) by ONE launch
 * over up to GSR_ADAM_MAX_SEGMENTS tensors.  The `_host` arrays are HOST arrays of length num_segments holding device
 * pointers / element counts / the group's learning rate / the group's step number t >= 1 (the value of
 * state["step"] AFTER the increment, torch/optim/adam.py).  Arithmetic of torch's Adam (amsgrad = False,
 * weight_decay = 0): m += (g - m)(1 - beta1); v = v beta2 + (1 - beta2) g g;
 * p -= lr / (1 - beta1^t) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps), with g multiplied by `grad_scale` first
 * (1 / world_size turns the summed view-parallel gradient into DDP's mean, pipelines/base_pipeline.py:202-207).
 * gsr_opacity_reset: opacities = min(opacities, max_logit) and the group's Adam moments = 0
 * (gs_toolkit/models/vanilla_gs.py:472-489).
 * ---------------------------------------------------------------------------------------------- */
#define GSR_ADAM_MAX_SEGMENTS 8
GSR_API int gsr_adam_step_multi(int num_segments, float *const *params_host, const float *const *grads_host,
                                float *const *exp_avg_host, float *const *exp_avg_sq_host,
                                const int64_t *numels_host, const double *lrs_host, const int64_t *steps_host,
                                double beta1, double beta2, double eps, float grad_scale, void *stream);
GSR_API int gsr_opacity_reset(int num_points, float max_logit, float *opacities_raw, float *exp_avg /*nullable*/,
                              float *exp_avg_sq /*nullable*/, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Adaptive density control — replaces gs_toolkit/models/vanilla_gs.py after_train :344-372,
 * refinement_after :381-497, cull_gaussians :499-535, split_gaussians :537-581, dup_gaussians :583-592 and the Adam
 * state surgery remove_from_optim :282-300 / dup_in_optim :308-337.
 *  gsr_densify_stats_update : per view. xys_grad is read with a row stride of `xys_grad_stride` floats (2 for a dense
 *      [N,2] tensor, 12 for the grad_records of the fused operator); max_dim = max(image width, height);
 *      first != 0 when the statistics are unset (None in the reference).
 *  gsr_densify_plan : flags [N] u8 (bit0 split, bit1 duplicate, bit2 original survives, bit3 its split samples
 *      survive, bit4 its duplicate survives), ranks [N,4] i32 (exclusive ranks among {split, surviving originals,
 *      surviving splits, surviving duplicates}; 16-byte aligned), counts [4] i32 on the DEVICE = the four totals.
 *      do_densify = 0 is the cull-only branch (:462-466): the three statistics may then be NULL unless
 *      use_cull_screen.  cull_big = (step > refine_every * reset_alpha_every) (:513).
 *  gsr_densify_apply : counts_host = the four totals read back by the caller; N' = counts[1] + n_split_samples *
 *      counts[2] + counts[3] rows are written to every dst tensor (row order: surviving originals, split samples
 *      sample-major, duplicates).  samples [n_split_samples * counts[0], 3] are the standard-normal draws of
 *      split_gaussians (:543-545), row s * counts[0] + (rank of the source among ALL split Gaussians).
 *      kinds: COPY (new rows copy their source), MEANS (split samples get mean + R(q)(exp(s) z)), SCALES (rows whose
 *      source was split get log(exp(s) / 1.6)), ZERO_NEW (Adam moments: zeros for new rows).  src and dst must not
 *      alias.  map [N'] u32 is scratch (destination row -> source row).
 * ---------------------------------------------------------------------------------------------- */
#define GSR_DENSIFY_MAX_TENSORS 24
typedef enum gsr_densify_kind {
  GSR_DENSIFY_COPY = 0,
  GSR_DENSIFY_MEANS = 1,
  GSR_DENSIFY_SCALES = 2,
  GSR_DENSIFY_ZERO_NEW = 3
} gsr_densify_kind;
GSR_API int gsr_densify_stats_update(int num_points, const float *xys_grad, int xys_grad_stride, const int32_t *radii,
                                     float max_dim, int first, float *xys_grad_norm, float *vis_counts,
                                     float *max_2Dsize, void *stream);
GSR_API size_t gsr_densify_plan_workspace_bytes(int num_points);
GSR_API int gsr_densify_plan(int num_points, const float *scales_raw, const float *opacities_raw,
                             const float *xys_grad_norm, const float *vis_counts, const float *max_2Dsize,
                             int do_densify, float max_dim, float densify_grad_thresh, float densify_size_thresh,
                             int use_split_screen, float split_screen_size, float cull_alpha_thresh, int cull_big,
                             float cull_scale_thresh, int use_cull_screen, float cull_screen_size, uint8_t *flags,
                             int32_t *ranks, int32_t *counts, void *workspace, size_t workspace_bytes, void *stream);
GSR_API int gsr_densify_apply(int num_points, int n_split_samples, const int32_t *counts_host, const uint8_t *flags,
                              const int32_t *ranks, const float *samples, const float *means, const float *scales_raw,
                              const float *quats_raw, int num_tensors, const float *const *src_host,
                              float *const *dst_host, const int32_t *widths_host, const int32_t *kinds_host,
                              uint32_t *map, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GSR_B200_H */
