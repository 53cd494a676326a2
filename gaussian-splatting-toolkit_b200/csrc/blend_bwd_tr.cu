// blend_bwd_tr.cu — TWO-PHASE adjoint of the per-tile alpha compositing (3 channels, FP32, 16x16 tiles): the per-(pixel,
// Gaussian) weights are TRANSPOSED through shared memory instead of being reduced (blend_bwd.cu) or scanned
// (blend_bwd_scan.cu) with warp shuffles.
//
// Same function as blend_backward_kernel (reference rasterize_backward_kernel, csrc/backward.cu:133-303).  A warp owns an
// 8x4 pixel block and the compacted list of the Gaussians that can reach it; the list is consumed in groups of G rows:
//
//   phase 1  lane = PIXEL.  Walk the group's Gaussians back to front exactly as the reference does (the per-pixel state
//            T, s lives in registers, records are broadcast from the staged ring) and write for every (Gaussian j, pixel p)
//            the two numbers the parameter gradients are linear in,
//                w_jp   = vis_jp * v_alpha_jp      (gradient of the loss w.r.t. log-opacity weight; v_sigma = -o w)
//                fac_jp = alpha_jp * T_jp          (weight of the pixel's upstream colour gradient)
//            as one float2 into a [G][33] matrix (row = Gaussian; the odd row stride keeps both phases conflict-free).
//   phase 2  lane = GAUSSIAN (32 / G lanes share a row and split its 32 pixels).  Each lane sums its row:
//                v_rgb += fac * v_out,  (sum w dx^2, sum w dx dy, sum w dy^2, sum w dx, sum w dy, sum w)
//            in private registers (the geometric sums through pixel-grid moments of w: 3 instructions per pixel), the 32 / G partial sums are combined with log2(32 / G) shuffles, and ONE set of nine
//            RED.ADD per (warp, Gaussian) goes to global memory.
//
// Per (warp, Gaussian) visit this costs ~36 + ~10 instructions and no dependent shuffle chain, against 116 for the
// shuffle butterfly of blend_bwd.cu and 72 (10 of them SHFL on the critical path) for the affine-map scan of
// blend_bwd_scan.cu.  Staging (double-buffered packed records, register prefetch), exact warp compaction and the early cut
// at max(final_idx) are those of blend_bwd.cu (tried and dropped: staging two batches ahead so that tiles of <= 512 pairs
// need no barrier after the prologue — same time, the warps of a CTA finish unevenly either way); rows left over at the end of a staged batch stay in the matrix and the
// group is completed from the next batch, so only the last group of a (warp, tile) is partially filled.
// Measured and dropped (round 2, gpurun_out/r2_run50 / r2_run51): 16-bit lists holding byte offsets as in the forward
// (two address instructions less per visit, 72 instead of 76 per trip of two) — 0.694 -> 0.763 ms, with AND without the
// shift: the wider lists take the CTA from 63.6 to 65.7 KB of shared memory and the kernel loses more than the
// instructions save; 8 / 32 rows per group 0.793 / 0.729 ms; phase-1 unroll 1 / 4: 0.730 / 0.693 ms (2 = 0.694).
#include <stdlib.h>

#include "blend_common.cuh"

#ifndef GSR_TR_UNROLL
#define GSR_TR_UNROLL 2  // visits per trip of the phase-1 walk
#endif
#ifndef GSR_TR_RING
#define GSR_TR_RING 1  // buffers of the staged-record ring (see TrSmem)
#endif
#define GSR_PRAGMA_(x) _Pragma(#x)
#define GSR_PRAGMA_UNROLL(n) GSR_PRAGMA_(unroll n)

namespace gsr {

namespace {

constexpr int kRowStride = 33;  // float2 elements per matrix row (32 pixels + 1 pad)

// Shared memory per CTA decides the L1 / shared-memory split of the SM, and the L1 size decides how many of the
// scattered record gathers can be in flight: forcing the largest carveout (28 KB of L1) takes this kernel from 0.692 to
// 0.759 ms and the forward from 0.369 to 0.428 ms (gpurun_out/r2_run52_*).  A SINGLE-buffered record ring (one more
// barrier per batch; the next batch still waits in registers) brings the CTA from 63.6 to 51.3 KB, the carveout from
// 196 to 164 KB, and the kernel to 0.677 ms.
template <int G>
struct TrSmem {
  float4 rec[GSR_TR_RING][3][BLEND_THREADS];
  float2 wf[BLEND_THREADS / 32][G * kRowStride];
  float4 vout[BLEND_THREADS / 32][32];
  unsigned char list[BLEND_THREADS / 32][BLEND_THREADS + 8];  // + 8: the pipelined phase 1 reads up to two entries ahead
  alignas(8) unsigned char mask[GSR_TR_RING][BLEND_THREADS];  // MASKS: per staged record, the warp blocks it can reach (block_mask_16)
  int warp_max[BLEND_THREADS / 32];
};

struct RowGaussian {  // the Gaussian whose matrix row a lane sums in phase 2
  float x, y, A, B, C, o;
  int id;
};

// phase 2: lane (row = lane % G, sub = lane / G) sums pixels [sub * G, sub * G + G) of its row
template <int G>
__device__ __forceinline__ void sum_rows(const RowGaussian &R, int rows, const float2 *__restrict__ wf,
                                         const float4 *__restrict__ vout, int lane, float x0, float y0,
                                         float *__restrict__ v_xy, float *__restrict__ v_conic,
                                         float *__restrict__ v_colors, float *__restrict__ v_opacity) {
  const unsigned full = 0xffffffffu;
  const int row = lane & (G - 1), sub = lane / G;
  const float2 *wrow = wf + row * kRowStride + sub * G;
  const float4 *vo = vout + sub * G;
  // Pixel p = sub * G + 8 r + c sits at (x0 + c, y0 + sub * NR + r).  With the block-centred constants cx = c - 3.5 and
  // cy = r - (NR - 1) / 2 the geometric sums are polynomials in the Gaussian's offset (gx, gy) from the centre of the
  // lane's pixels, dx = gx - cx, dy = gy - cy, whose coefficients are MOMENTS of w over the pixels:
  //   sum w dx^2 = gx^2 W - 2 gx Mx + Mxx,  sum w dx dy = gx gy W - gx My - gy Mx + Mxy,  sum w dx = gx W - Mx, ...
  // cx, cx^2 are compile-time constants and cy is constant per pixel row, so a pixel costs 1 FADD + 2 FFMA (row sums W_r,
  // X_r = sum w cx, XX_r = sum w cx^2) instead of the 10 instructions of the direct form.
  constexpr int NR = G / 8;
  float a_r = 0.f, a_g = 0.f, a_b = 0.f;
  float Wr[NR], Xr[NR], XXr[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    Wr[r] = Xr[r] = XXr[r] = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float2 e = wrow[8 * r + c];
      const float4 v = vo[8 * r + c];
      a_r += e.y * v.x;
      a_g += e.y * v.y;
      a_b += e.y * v.z;
      const float cx = (float)c - 3.5f;
      Wr[r] += e.x;
      Xr[r] += e.x * cx;
      XXr[r] += e.x * (cx * cx);
    }
  }
  float W = 0.f, Mx = 0.f, Mxx = 0.f, My = 0.f, Myy = 0.f, Mxy = 0.f;
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const float cy = (float)r - 0.5f * (float)(NR - 1);
    W += Wr[r];
    Mx += Xr[r];
    Mxx += XXr[r];
    My += cy * Wr[r];
    Myy += (cy * cy) * Wr[r];
    Mxy += cy * Xr[r];
  }
  const float gx = R.x - (x0 + 3.5f), gy = R.y - (y0 + (float)(sub * NR) + 0.5f * (float)(NR - 1));
  float a_xx = gx * (gx * W - 2.f * Mx) + Mxx;
  float a_xy = gx * (gy * W - My) - gy * Mx + Mxy;
  float a_yy = gy * (gy * W - 2.f * My) + Myy;
  float a_x = gx * W - Mx;
  float a_y = gy * W - My;
  float a_w = W;
#pragma unroll
  for (int o = G; o < 32; o <<= 1) {
    a_r += __shfl_xor_sync(full, a_r, o);
    a_g += __shfl_xor_sync(full, a_g, o);
    a_b += __shfl_xor_sync(full, a_b, o);
    a_xx += __shfl_xor_sync(full, a_xx, o);
    a_xy += __shfl_xor_sync(full, a_xy, o);
    a_yy += __shfl_xor_sync(full, a_yy, o);
    a_x += __shfl_xor_sync(full, a_x, o);
    a_y += __shfl_xor_sync(full, a_y, o);
    a_w += __shfl_xor_sync(full, a_w, o);
  }
  if (sub == 0 && row < rows) {
    // v_sigma = -o vis v_alpha = -o w;  conic = -(2A, B, 2C) ln2
    const unsigned g = (unsigned)R.id;
    const float no = -R.o;
    const float ca = -2.f * kLn2 * R.A, cb = -kLn2 * R.B, cc = -2.f * kLn2 * R.C;
    atomicAdd(v_colors + 3u * g, a_r);
    atomicAdd(v_colors + 3u * g + 1u, a_g);
    atomicAdd(v_colors + 3u * g + 2u, a_b);
    atomicAdd(v_conic + 3u * g, 0.5f * no * a_xx);
    atomicAdd(v_conic + 3u * g + 1u, no * a_xy);
    atomicAdd(v_conic + 3u * g + 2u, 0.5f * no * a_yy);
    atomicAdd(v_xy + 2u * g, no * (ca * a_x + cb * a_y));
    atomicAdd(v_xy + 2u * g + 1u, no * (cb * a_x + cc * a_y));
    atomicAdd(v_opacity + g, a_w);
  }
}

#define GSR_TR_PARAMS                                                                                                   \
  int tiles_x, int img_w, int img_h, const int *__restrict__ gaussian_ids_sorted, const int2 *__restrict__ tile_bins,   \
      const float2 *__restrict__ xys, const float *__restrict__ conics, const float *__restrict__ colors,               \
      const float *__restrict__ opacities, const float *__restrict__ background, const float *__restrict__ final_Ts,    \
      const int *__restrict__ final_idx, const float *__restrict__ v_output, const float *__restrict__ v_output_alpha,  \
      float *__restrict__ v_xy, float *__restrict__ v_conic, float *__restrict__ v_colors, float *__restrict__ v_opacity
#define GSR_TR_ARGS                                                                                                     \
  tiles_x, img_w, img_h, gaussian_ids_sorted, tile_bins, xys, conics, colors, opacities, background, final_Ts,          \
      final_idx, v_output, v_output_alpha, v_xy, v_conic, v_colors, v_opacity

// MASKS: the staging thread evaluates its record against the eight warp blocks once (block_mask_16, blend_common.cuh)
// instead of every warp testing every record (compact_survivors)
template <int G, bool MASKS>
__device__ __forceinline__ void blend_backward_tr_body(GSR_TR_PARAMS) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TrSmem<G> &S = *reinterpret_cast<TrSmem<G> *>(smem_raw);

  const unsigned full = 0xffffffffu;
  const int tile_x = blockIdx.x, tile_y = blockIdx.y;
  const int tile_id = tile_y * tiles_x + tile_x;
  const int tr = threadIdx.x, nthreads = BLEND_THREADS, lane = tr & 31, warp = tr >> 5;
  int lx, ly;
  map_pixel(16, lx, ly);
  const int ipx = tile_x * 16 + lx, ipy = tile_y * 16 + ly;
  const bool inside = (ipx < img_w) && (ipy < img_h);
  const float px = (float)ipx, py = (float)ipy;
  const int pix = inside ? (ipy * img_w + ipx) : 0;
  const float x0 = (float)(tile_x * 16 + ((warp & 1) << 3)), y0 = (float)(tile_y * 16 + ((warp >> 1) << 2));
  const float tile_x0 = (float)(tile_x * 16), tile_y0 = (float)(tile_y * 16);

  const float fx0 = (float)__reduce_min_sync(full, inside ? ipx : 0x7fffffff);
  const float fx1 = (float)__reduce_max_sync(full, inside ? ipx : -0x7fffffff);
  const float fy0 = (float)__reduce_min_sync(full, inside ? ipy : 0x7fffffff);
  const float fy1 = (float)__reduce_max_sync(full, inside ? ipy : -0x7fffffff);

  const int2 range = tile_bins[tile_id];
  // reference: bin_final = inside ? final_index : 0 (backward.cu:168); -1 for outside threads is equivalent (never valid)
  const int bin_final = inside ? final_idx[pix] : -1;
  float T = inside ? final_Ts[pix] : 1.f;
  float vo_r = 0.f, vo_g = 0.f, vo_b = 0.f, vo_a = 0.f;
  if (inside) {
    vo_r = v_output[3 * (size_t)pix];
    vo_g = v_output[3 * (size_t)pix + 1];
    vo_b = v_output[3 * (size_t)pix + 2];
    vo_a = v_output_alpha[pix];
  }
  // the reference's three running colour sums collapse to one scalar (see blend_bwd.cu): s starts at -c_final
  float s_run = -T * (vo_a - (background[0] * vo_r + background[1] * vo_g + background[2] * vo_b));
  S.vout[warp][lane] = make_float4(vo_r, vo_g, vo_b, 0.f);

  const int warp_bin_final = __reduce_max_sync(full, bin_final);
  if (lane == 0) S.warp_max[warp] = warp_bin_final;
  __syncthreads();
  int cta_bin_final = -1;
  for (int w = 0; w < (nthreads >> 5); ++w) cta_bin_final = max(cta_bin_final, S.warp_max[w]);

  const int end = min(range.y, cta_bin_final + 1);
  const int count = end - range.x;
  if (count <= 0) return;  // uniform across the CTA
  const int num_batches = (count + nthreads - 1) / nthreads;

  BlendRecord rec;
  if (end - 1 - tr >= range.x)
    rec = gather_record(gaussian_ids_sorted[end - 1 - tr], xys, conics, colors, opacities);

  RowGaussian R;
  R.x = R.y = R.A = R.B = R.C = R.o = 0.f;
  R.id = 0;
  int rows = 0;  // rows of the matrix already written (phase 1 done) and waiting for a full group
  float2 *const wf = S.wf[warp];
  const int my_row = lane & (G - 1);

  for (int b = 0; b < num_batches; ++b) {
    const int buf = GSR_TR_RING == 2 ? (b & 1) : 0;
    if (GSR_TR_RING == 1 && b > 0) __syncthreads();  // every warp is done with the previous batch
    const int batch_end = end - 1 - nthreads * b;  // sorted index held by slot 0; slot t holds batch_end - t
    if (batch_end - tr >= range.x) {
      S.rec[buf][0][tr] = rec.r0;
      S.rec[buf][1][tr] = rec.r1;
      S.rec[buf][2][tr] = rec.r2;
      if (MASKS) S.mask[buf][tr] = (unsigned char)block_mask_16(rec.r0, rec.r1, tile_x0, tile_y0);
    }
    __syncthreads();
    {
      const int nxt = batch_end - nthreads - tr;
      if (nxt >= range.x) rec = gather_record(gaussian_ids_sorted[nxt], xys, conics, colors, opacities);
    }
    const int batch_size = min(nthreads, batch_end + 1 - range.x);
    const int t_begin = max(0, batch_end - warp_bin_final);  // slots before it are behind every pixel's last contributor
    if (t_begin >= batch_size) continue;
    const int n_list = MASKS ? compact_from_masks(S.mask[buf], warp, t_begin, batch_size, S.list[warp], lane)
                             : compact_survivors(S.rec[buf][0], S.rec[buf][1], t_begin, batch_size, fx0, fx1, fy0, fy1,
                                                 S.list[warp], lane);
    const unsigned char *list = S.list[warp];
    const int slot_min = batch_end - bin_final;  // slot t holds sorted index batch_end - t <= bin_final  <=>  t >= slot_min
    int li = 0;
    while (li < n_list) {
      const int take = min(G - rows, n_list - li);
      if (my_row >= rows && my_row < rows + take) {  // the lanes that will sum the new rows keep their Gaussian
        const int slot = list[li + my_row - rows];
        const float4 q0 = S.rec[buf][0][slot], q1 = S.rec[buf][1][slot];
        R.x = q0.x; R.y = q0.y;
        R.A = q1.x; R.B = q1.y; R.C = q1.z; R.o = q1.w;
        R.id = __float_as_int(S.rec[buf][2][slot].w);
      }
      // ---- phase 1: lane = pixel ----  (software-pipelined: the next record is loaded while this one is evaluated)
      float2 *dst = wf + rows * kRowStride + lane;
      const unsigned char *lp = list + li;
      int slot = lp[0];
      float2 c0 = *reinterpret_cast<const float2 *>(&S.rec[buf][0][slot]);
      float4 q1 = S.rec[buf][1][slot];
      float4 q2 = S.rec[buf][2][slot];
      int slot_n = lp[1];  // the list is padded: reading one or two entries past its end is harmless
GSR_PRAGMA_UNROLL(GSR_TR_UNROLL)
      for (int k = 0; k < take; ++k) {
        const float2 n0 = *reinterpret_cast<const float2 *>(&S.rec[buf][0][slot_n]);
        const float4 n1 = S.rec[buf][1][slot_n];
        const float4 n2 = S.rec[buf][2][slot_n];
        const int slot_nn = lp[k + 2];
        const float dx = c0.x - px, dy = c0.y - py;
        const float gx = q1.x * dx, gy = q1.z * dy;           // A dx, C dy
        const float power = dx * (gx + q1.y * dy) + gy * dy;  // = -sigma log2(e)
        const float vis = exp2f(power);
        const float alpha = fminf(0.99f, q1.w * vis);
        const bool valid = (slot >= slot_min) && !(power > 0.f || alpha < 1.f / 255.f);
        const float alpha_e = valid ? alpha : 0.f;
        const float vis_e = valid ? vis : 0.f;
        const float ra = 1.f / (1.f - alpha_e);
        T *= ra;  // transmittance in front of this Gaussian
        const float fac = alpha_e * T;
        const float dcol = q2.x * vo_r + q2.y * vo_g + q2.z * vo_b;
        const float v_alpha = T * dcol - ra * s_run;
        s_run += fac * dcol;
        *dst = make_float2(vis_e * v_alpha, fac);
        dst += kRowStride;
        slot = slot_n; c0 = n0; q1 = n1; q2 = n2; slot_n = slot_nn;
      }
      rows += take;
      li += take;
      if (rows == G) {
        __syncwarp();
        sum_rows<G>(R, G, wf, S.vout[warp], lane, x0, y0, v_xy, v_conic, v_colors, v_opacity);
        __syncwarp();
        rows = 0;
      }
    }
  }
  if (rows > 0) {
    __syncwarp();
    sum_rows<G>(R, rows, wf, S.vout[warp], lane, x0, y0, v_xy, v_conic, v_colors, v_opacity);
  }
}

// the default: block masks from the staging threads
template <int G, int MIN_CTAS>
__global__ void __launch_bounds__(BLEND_THREADS, MIN_CTAS) blend_backward_tr_kernel(GSR_TR_PARAMS) {
  blend_backward_tr_body<G, true>(GSR_TR_ARGS);
}

// GSR_BLOCK_MASK=0: per-warp tests
template <int G, int MIN_CTAS>
__global__ void __launch_bounds__(BLEND_THREADS, MIN_CTAS) blend_backward_tr_warptest_kernel(GSR_TR_PARAMS) {
  blend_backward_tr_body<G, false>(GSR_TR_ARGS);
}

template <int G, int MIN_CTAS>
int launch_tr(dim3 grid, cudaStream_t st, int img_w, int img_h, const int *gaussian_ids_sorted, const int2 *tile_bins,
              const float2 *xys, const float *conics, const float *colors, const float *opacities,
              const float *background, const float *final_Ts, const int *final_idx, const float *v_output,
              const float *v_output_alpha, float *v_xy, float *v_conic, float *v_colors, float *v_opacity) {
  if (!blend_block_masks()) {
    static const cudaError_t attr0 =
        cudaFuncSetAttribute(blend_backward_tr_warptest_kernel<G, MIN_CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(TrSmem<G>));
    GSR_CUDA(attr0);
    blend_backward_tr_warptest_kernel<G, MIN_CTAS><<<grid, BLEND_THREADS, sizeof(TrSmem<G>), st>>>(
        (int)grid.x, img_w, img_h, gaussian_ids_sorted, tile_bins, xys, conics, colors, opacities, background, final_Ts,
        final_idx, v_output, v_output_alpha, v_xy, v_conic, v_colors, v_opacity);
    GSR_CHECK_LAUNCH("blend_backward_tr_warptest_kernel");
    return GSR_OK;
  }
  static const cudaError_t attr = cudaFuncSetAttribute(blend_backward_tr_kernel<G, MIN_CTAS>,
                                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TrSmem<G>));
  GSR_CUDA(attr);
  static const cudaError_t carve = [] {
    const char *e = getenv("GSR_TR_CARVEOUT");  // percent of the maximum shared-memory carveout; unset = driver default
    return (e && e[0]) ? cudaFuncSetAttribute(blend_backward_tr_kernel<G, MIN_CTAS>,
                                              cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e))
                       : cudaSuccess;
  }();
  GSR_CUDA(carve);
  blend_backward_tr_kernel<G, MIN_CTAS><<<grid, BLEND_THREADS, sizeof(TrSmem<G>), st>>>(
      (int)grid.x, img_w, img_h, gaussian_ids_sorted, tile_bins, xys, conics, colors, opacities, background, final_Ts,
      final_idx, v_output, v_output_alpha, v_xy, v_conic, v_colors, v_opacity);
  GSR_CHECK_LAUNCH("blend_backward_tr_kernel");
  return GSR_OK;
}

}  // namespace

// GSR_BWD_KERNEL = tr (default: this file, G = 16) | tr8 | tr32 | scan (blend_bwd_scan.cu) | pixel (blend_bwd.cu) — read once
int blend_bwd_mode() {
  static const int v = [] {
    const char *e = getenv("GSR_BWD_KERNEL");
    if (!e || !e[0]) return 16;
    if (e[0] == 'p') return 0;
    if (e[0] == 's') return 1;
    if (e[0] == 't' && e[1] == 'r') {
      if (e[2] == '8') return 8;
      if (e[2] == '3') return 32;
      return 16;
    }
    return 16;
  }();
  return v;
}

int launch_blend_backward_tr(int mode, dim3 grid, cudaStream_t st, int img_w, int img_h, const int *gaussian_ids_sorted,
                             const int2 *tile_bins, const float2 *xys, const float *conics, const float *colors,
                             const float *opacities, const float *background, const float *final_Ts,
                             const int *final_idx, const float *v_output, const float *v_output_alpha, float *v_xy,
                             float *v_conic, float *v_colors, float *v_opacity) {
  if (mode == 8)
    return launch_tr<8, 4>(grid, st, img_w, img_h, gaussian_ids_sorted, tile_bins, xys, conics, colors, opacities, background,
                           final_Ts, final_idx, v_output, v_output_alpha, v_xy, v_conic, v_colors, v_opacity);
  if (mode == 32)
    return launch_tr<32, 2>(grid, st, img_w, img_h, gaussian_ids_sorted, tile_bins, xys, conics, colors, opacities,
                            background, final_Ts, final_idx, v_output, v_output_alpha, v_xy, v_conic, v_colors, v_opacity);
  return launch_tr<16, 3>(grid, st, img_w, img_h, gaussian_ids_sorted, tile_bins, xys, conics, colors, opacities, background,
                          final_Ts, final_idx, v_output, v_output_alpha, v_xy, v_conic, v_colors, v_opacity);
}

}  // namespace gsr
