// blend_packed_tr.cu — two-phase transposing adjoint (see blend_bwd_tr.cu) of the FUSED operator's compositing: RGB plus
// depth as a fourth channel, packed 48-byte records {x, y, ext_x, ext_y | A, B, C, o | r, g, b, z} staged with cp.async
// (LDGSTS, no register staging), ten per-Gaussian sums leaving as THREE 16-byte vector reductions
// (red.global.add.v4.f32) into the packed gradient record {v_x, v_y, v_opacity, v_depth | v_a, v_b, v_c, - | v_r, v_g, v_b, -}.
// Default adjoint of gsr_blend_packed_backward for 16x16 tiles (GSR_PACKED_BWD = tr | scan | pixel).
#include <stdlib.h>

#include "blend_common.cuh"

namespace gsr {

namespace {

constexpr int kRowStride = 33;  // float2 elements per matrix row (32 pixels + 1 pad)
constexpr int G = 16;           // rows (Gaussians) per group

template <int BATCH>
struct PkTrSmem {
  float4 rec[2][3][BATCH];
  int gid[2][BATCH];
  float2 wf[BLEND_THREADS / 32][G * kRowStride];
  float4 vout[BLEND_THREADS / 32][32];  // {v_out_r, v_out_g, v_out_b, v_out_depth}
  unsigned char list[BLEND_THREADS / 32][BLEND_THREADS + 8];
  int warp_max[BLEND_THREADS / 32];
};

struct RowGaussian {
  float x, y, A, B, C, o;
  int id;
};

__device__ __forceinline__ void pkt_cp_async16(float4 *smem_dst, const float4 *gmem_src) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void pkt_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void pkt_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void pkt_red_add_v4(float *addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// phase 2 (see blend_bwd_tr.cu): lane (row = lane % G, sub = lane / G) sums pixels [sub * G, sub * G + G) of its row
template <bool DEPTH>
__device__ __forceinline__ void pkt_sum_rows(const RowGaussian &R, int rows, const float2 *__restrict__ wf,
                                             const float4 *__restrict__ vout, int lane, float x0, float y0,
                                             float *__restrict__ grad_rec) {
  const unsigned full = 0xffffffffu;
  const int row = lane & (G - 1), sub = lane / G;
  const float2 *wrow = wf + row * kRowStride + sub * G;
  const float4 *vo = vout + sub * G;
  constexpr int NR = G / 8;
  float a_r = 0.f, a_g = 0.f, a_b = 0.f, a_d = 0.f;
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
      if (DEPTH) a_d += e.y * v.w;
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
    if (DEPTH) a_d += __shfl_xor_sync(full, a_d, o);
    a_xx += __shfl_xor_sync(full, a_xx, o);
    a_xy += __shfl_xor_sync(full, a_xy, o);
    a_yy += __shfl_xor_sync(full, a_yy, o);
    a_x += __shfl_xor_sync(full, a_x, o);
    a_y += __shfl_xor_sync(full, a_y, o);
    a_w += __shfl_xor_sync(full, a_w, o);
  }
  if (sub == 0 && row < rows) {
    float *dst = grad_rec + 12u * (unsigned)R.id;
    const float no = -R.o;
    const float ca = -2.f * kLn2 * R.A, cb = -kLn2 * R.B, cc = -2.f * kLn2 * R.C;
    pkt_red_add_v4(dst, no * (ca * a_x + cb * a_y), no * (cb * a_x + cc * a_y), a_w, a_d);
    pkt_red_add_v4(dst + 4, 0.5f * no * a_xx, no * a_xy, 0.5f * no * a_yy, 0.f);
    pkt_red_add_v4(dst + 8, a_r, a_g, a_b, 0.f);
  }
}

// BATCH = records per ring stage.  <256, 3>: 66.6 KB of shared memory, three CTAs / SM; <128, 4>: 53.3 KB and 64 registers,
// four CTAs / SM (32 warps) at twice the number of batch barriers.
template <bool DEPTH, int BATCH, int MIN_CTAS>
__global__ void __launch_bounds__(BLEND_THREADS, MIN_CTAS)
blend_packed_backward_tr_kernel(int tiles_x, int img_w, int img_h, int num_points,
                                const int *__restrict__ gaussian_ids_sorted, const int2 *__restrict__ tile_bins,
                                const float4 *__restrict__ rec, const float *__restrict__ background,
                                const float *__restrict__ final_Ts, const int *__restrict__ final_idx,
                                const float *__restrict__ v_output, const float *__restrict__ v_output_depth,
                                const float *__restrict__ v_output_alpha, float *__restrict__ grad_rec) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PkTrSmem<BATCH> &S = *reinterpret_cast<PkTrSmem<BATCH> *>(smem_raw);

  const unsigned full = 0xffffffffu;
  const int tile_x = blockIdx.x, tile_y = blockIdx.y;
  const int tile_id = tile_y * tiles_x + tile_x;
  const int tr = threadIdx.x, nthreads = BATCH, lane = tr & 31, warp = tr >> 5;  // nthreads: records per batch
  int lx, ly;
  map_pixel(16, lx, ly);
  const int ipx = tile_x * 16 + lx, ipy = tile_y * 16 + ly;
  const bool inside = (ipx < img_w) && (ipy < img_h);
  const float px = (float)ipx, py = (float)ipy;
  const int pix = inside ? (ipy * img_w + ipx) : 0;
  const float x0 = (float)(tile_x * 16 + ((warp & 1) << 3)), y0 = (float)(tile_y * 16 + ((warp >> 1) << 2));

  const float fx0 = (float)__reduce_min_sync(full, inside ? ipx : 0x7fffffff);
  const float fx1 = (float)__reduce_max_sync(full, inside ? ipx : -0x7fffffff);
  const float fy0 = (float)__reduce_min_sync(full, inside ? ipy : 0x7fffffff);
  const float fy1 = (float)__reduce_max_sync(full, inside ? ipy : -0x7fffffff);

  const int2 range = tile_bins[tile_id];
  const int bin_final = inside ? final_idx[pix] : -1;
  float T = inside ? final_Ts[pix] : 1.f;
  float vo_r = 0.f, vo_g = 0.f, vo_b = 0.f, vo_d = 0.f, vo_a = 0.f;
  if (inside) {
    vo_r = v_output[3 * (size_t)pix];
    vo_g = v_output[3 * (size_t)pix + 1];
    vo_b = v_output[3 * (size_t)pix + 2];
    vo_a = v_output_alpha[pix];
    if (DEPTH) vo_d = v_output_depth[pix];
  }
  // the depth channel has a zero background, so it adds nothing to the T_final term
  float s_run = -T * (vo_a - (background[0] * vo_r + background[1] * vo_g + background[2] * vo_b));
  S.vout[warp][lane] = make_float4(vo_r, vo_g, vo_b, vo_d);

  const int warp_bin_final = __reduce_max_sync(full, bin_final);
  if (lane == 0) S.warp_max[warp] = warp_bin_final;
  __syncthreads();
  int cta_bin_final = -1;
  for (int w = 0; w < BLEND_THREADS / 32; ++w) cta_bin_final = max(cta_bin_final, S.warp_max[w]);

  const int end = min(range.y, cta_bin_final + 1);
  const int count = end - range.x;
  if (count <= 0) return;  // uniform across the CTA
  const int num_batches = (count + nthreads - 1) / nthreads;

  auto stage = [&](int buf, int sorted_index) {
    const int gid = gaussian_ids_sorted[sorted_index];
    S.gid[buf][tr] = gid;
    pkt_cp_async16(&S.rec[buf][0][tr], rec + gid);
    pkt_cp_async16(&S.rec[buf][1][tr], rec + num_points + gid);
    pkt_cp_async16(&S.rec[buf][2][tr], rec + 2 * (size_t)num_points + gid);
  };
  if (tr < BATCH && end - 1 - tr >= range.x) stage(0, end - 1 - tr);
  pkt_commit();

  RowGaussian R;
  R.x = R.y = R.A = R.B = R.C = R.o = 0.f;
  R.id = 0;
  int rows = 0;
  float2 *const wf = S.wf[warp];
  const int my_row = lane & (G - 1);

  for (int b = 0; b < num_batches; ++b) {
    const int buf = b & 1;
    const int batch_end = end - 1 - nthreads * b;  // sorted index held by slot 0; slot t holds batch_end - t
    pkt_wait_all();
    __syncthreads();
    {
      const int nxt = batch_end - nthreads - tr;
      if (tr < BATCH && nxt >= range.x) stage(buf ^ 1, nxt);
      pkt_commit();
    }
    const int batch_size = min(nthreads, batch_end + 1 - range.x);
    const int t_begin = max(0, batch_end - warp_bin_final);  // slots before it are behind every pixel's last contributor
    if (t_begin >= batch_size) continue;
    const int n_list = compact_survivors(S.rec[buf][0], S.rec[buf][1], t_begin, batch_size, fx0, fx1, fy0, fy1,
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
        R.id = S.gid[buf][slot];
      }
      // ---- phase 1: lane = pixel ----  (software-pipelined: the next record is loaded while this one is evaluated)
      float2 *dst = wf + rows * kRowStride + lane;
      const unsigned char *lp = list + li;
      int slot = lp[0];
      float2 c0 = *reinterpret_cast<const float2 *>(&S.rec[buf][0][slot]);
      float4 q1 = S.rec[buf][1][slot];
      float4 q2 = S.rec[buf][2][slot];
      int slot_n = lp[1];  // the list is padded: reading one or two entries past its end is harmless
#pragma unroll 2
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
        float dcol = q2.x * vo_r + q2.y * vo_g + q2.z * vo_b;
        if (DEPTH) dcol += q2.w * vo_d;
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
        pkt_sum_rows<DEPTH>(R, G, wf, S.vout[warp], lane, x0, y0, grad_rec);
        __syncwarp();
        rows = 0;
      }
    }
  }
  if (rows > 0) {
    __syncwarp();
    pkt_sum_rows<DEPTH>(R, rows, wf, S.vout[warp], lane, x0, y0, grad_rec);
  }
}

}  // namespace

// GSR_PACKED_BWD = tr (default: this file) | scan (blend_packed_scan.cu) | pixel (blend_packed.cu) — read once
int blend_packed_bwd_mode() {
  static const int v = [] {
    const char *e = getenv("GSR_PACKED_BWD");
    if (e && e[0] == 'p') return 0;
    if (e && e[0] == 's') return 1;
    return 2;
  }();
  return v;
}

template <bool DEPTH, int BATCH, int MIN_CTAS>
static int launch_pkt(dim3 grid, cudaStream_t st, int img_w, int img_h, int num_points, const int *gaussian_ids_sorted,
                      const int2 *tile_bins, const float4 *rec, const float *background, const float *final_Ts,
                      const int *final_idx, const float *v_output, const float *v_output_depth, const float *v_output_alpha,
                      float *grad_rec) {
  static const cudaError_t attr = cudaFuncSetAttribute(blend_packed_backward_tr_kernel<DEPTH, BATCH, MIN_CTAS>,
                                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PkTrSmem<BATCH>));
  GSR_CUDA(attr);
  blend_packed_backward_tr_kernel<DEPTH, BATCH, MIN_CTAS><<<grid, BLEND_THREADS, sizeof(PkTrSmem<BATCH>), st>>>(
      (int)grid.x, img_w, img_h, num_points, gaussian_ids_sorted, tile_bins, rec, background, final_Ts, final_idx, v_output,
      v_output_depth, v_output_alpha, grad_rec);
  GSR_CHECK_LAUNCH("blend_packed_backward_tr_kernel");
  return GSR_OK;
}

int launch_blend_packed_backward_tr(dim3 grid, cudaStream_t st, int img_w, int img_h, int num_points,
                                    const int *gaussian_ids_sorted, const int2 *tile_bins, const float4 *rec,
                                    const float *background, const float *final_Ts, const int *final_idx,
                                    const float *v_output, const float *v_output_depth, const float *v_output_alpha,
                                    float *grad_rec) {
  // GSR_PACKED_TR_BATCH = 256 (default, three CTAs / SM) | 128 (four CTAs / SM: measured 20 us SLOWER at cfg2) — read once
  static const int batch = [] {
    const char *e = getenv("GSR_PACKED_TR_BATCH");
    return (e && e[0] == '1') ? 128 : 256;
  }();
#define GSR_PKT_ARGS grid, st, img_w, img_h, num_points, gaussian_ids_sorted, tile_bins, rec, background, final_Ts, final_idx, \
                     v_output, v_output_depth, v_output_alpha, grad_rec
  if (batch == 256) return v_output_depth ? launch_pkt<true, 256, 3>(GSR_PKT_ARGS) : launch_pkt<false, 256, 3>(GSR_PKT_ARGS);
  return v_output_depth ? launch_pkt<true, 128, 4>(GSR_PKT_ARGS) : launch_pkt<false, 128, 4>(GSR_PKT_ARGS);
#undef GSR_PKT_ARGS
}

}  // namespace gsr
