// blend_packed_scan.cu — Gaussian-parallel adjoint of the FUSED operator's compositing (RGB + depth as 4th channel),
// the counterpart of blend_bwd_scan.cu for the packed records of blend_packed.cu (16x16 tiles).
//
// lane = one of 32 consecutive survivors of the warp's compacted list, loop over the 32 pixels of the warp's 8x4 block;
// the per-pixel state (T, s) travels across the lanes as ONE inclusive warp-shuffle scan of affine maps (see
// blend_bwd_scan.cu); the ten per-Gaussian gradient sums stay in registers and leave as THREE 16-byte vector
// reductions (red.global.add.v4.f32) into the packed 48-byte gradient record
//   {v_x, v_y, v_opacity, v_depth | v_a, v_b, v_c, - | v_r, v_g, v_b, -}
// — one set per 32 visits instead of one 10-lane RED per visit after a 32-lane butterfly.
// Staging (cp.async ring of packed records) and compaction are those of blend_packed.cu.
#include <stdlib.h>

#include "blend_common.cuh"

namespace gsr {

namespace {

__device__ __forceinline__ void pks_cp_async16(float4 *smem_dst, const float4 *gmem_src) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void pks_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void pks_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void pks_stage(float4 (*s_rec)[3][BLEND_THREADS], int buf, int slot, int g, int n,
                                          const float4 *__restrict__ rec) {
  pks_cp_async16(&s_rec[buf][0][slot], rec + g);
  pks_cp_async16(&s_rec[buf][1][slot], rec + n + g);
  pks_cp_async16(&s_rec[buf][2][slot], rec + 2 * (size_t)n + g);
}
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct LaneRec {  // the Gaussian a lane owns for one group
  float x, y, A, B, C, o, r, g, b, z;
  int id, sidx;
};

template <bool DEPTH>
__device__ __forceinline__ void pks_process_group(const LaneRec &G, const float4 *__restrict__ pixc,
                                                  float4 *__restrict__ pixs, const float *__restrict__ pixd, int lane,
                                                  float *__restrict__ grad_rec) {
  const unsigned full = 0xffffffffu;
  float a_r = 0.f, a_g = 0.f, a_b = 0.f, a_d = 0.f, a_xx = 0.f, a_xy = 0.f, a_yy = 0.f, a_x = 0.f, a_y = 0.f, a_w = 0.f;
  constexpr int U = 2;  // two pixels per iteration: two independent shuffle chains
  for (int p = 0; p < 32; p += U) {
    float4 c[U], st[U];
    float vd[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      c[u] = pixc[p + u];
      st[u] = pixs[p + u];
      vd[u] = DEPTH ? pixd[p + u] : 0.f;
    }
    float dx[U], dy[U], vis_e[U], alpha_e[U], ra[U], R[U], Cc[U], dj[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      dx[u] = G.x - st[u].x;
      dy[u] = G.y - st[u].y;
      const float gx = G.A * dx[u], gy = G.C * dy[u];
      const float power = dx[u] * (gx + G.B * dy[u]) + gy * dy[u];
      const float vis = exp2f(power);
      const float alpha = fminf(0.99f, G.o * vis);
      const bool valid = (G.sidx <= __float_as_int(c[u].w)) && !(power > 0.f || alpha < 1.f / 255.f);
      alpha_e[u] = valid ? alpha : 0.f;
      vis_e[u] = valid ? vis : 0.f;
      ra[u] = 1.f / (1.f - alpha_e[u]);
      dj[u] = G.r * c[u].x + G.g * c[u].y + G.b * c[u].z;
      if (DEPTH) dj[u] += G.z * vd[u];
      R[u] = ra[u];
      Cc[u] = alpha_e[u] * ra[u] * dj[u];
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      float rup[U], cup[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        rup[u] = __shfl_up_sync(full, R[u], d);
        cup[u] = __shfl_up_sync(full, Cc[u], d);
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (lane >= d) {
          Cc[u] = cup[u] + Cc[u] * rup[u];
          R[u] *= rup[u];
        }
    }
    float T[U], s_after[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      T[u] = st[u].z * R[u];
      s_after[u] = st[u].w + Cc[u] * st[u].z;
      const float fac = alpha_e[u] * T[u];
      const float s_behind = s_after[u] - fac * dj[u];
      const float v_alpha = T[u] * dj[u] - ra[u] * s_behind;
      const float w = vis_e[u] * v_alpha;
      a_r += fac * c[u].x;
      a_g += fac * c[u].y;
      a_b += fac * c[u].z;
      if (DEPTH) a_d += fac * vd[u];
      const float wdx = w * dx[u], wdy = w * dy[u];
      a_xx += wdx * dx[u];
      a_xy += wdx * dy[u];
      a_yy += wdy * dy[u];
      a_x += wdx;
      a_y += wdy;
      a_w += w;
    }
    if (lane == 31) {
#pragma unroll
      for (int u = 0; u < U; ++u) *reinterpret_cast<float2 *>(&pixs[p + u].z) = make_float2(T[u], s_after[u]);
    }
  }
  __syncwarp();
  if (G.id >= 0) {
    float *dst = grad_rec + 12u * (unsigned)G.id;
    const float no = -G.o;
    const float ca = -2.f * kLn2 * G.A, cb = -kLn2 * G.B, cc = -2.f * kLn2 * G.C;
    red_add_v4(dst, no * (ca * a_x + cb * a_y), no * (cb * a_x + cc * a_y), a_w, a_d);
    red_add_v4(dst + 4, 0.5f * no * a_xx, no * a_xy, 0.5f * no * a_yy, 0.f);
    red_add_v4(dst + 8, a_r, a_g, a_b, 0.f);
  }
}

}  // namespace

template <bool DEPTH>
__global__ void __launch_bounds__(BLEND_THREADS, 3)
blend_packed_backward_scan_kernel(int tiles_x, int img_w, int img_h, int num_points,
                                  const int *__restrict__ gaussian_ids_sorted, const int2 *__restrict__ tile_bins,
                                  const float4 *__restrict__ rec, const float *__restrict__ background,
                                  const float *__restrict__ final_Ts, const int *__restrict__ final_idx,
                                  const float *__restrict__ v_output, const float *__restrict__ v_output_depth,
                                  const float *__restrict__ v_output_alpha, float *__restrict__ grad_rec) {
  __shared__ float4 s_rec[2][3][BLEND_THREADS];
  __shared__ int s_gid[2][BLEND_THREADS];
  __shared__ unsigned char s_list[BLEND_THREADS / 32][BLEND_THREADS];
  __shared__ float4 s_pixc[BLEND_THREADS / 32][32];
  __shared__ float4 s_pixs[BLEND_THREADS / 32][32];
  __shared__ float s_pixd[BLEND_THREADS / 32][32];
  __shared__ int s_warp_max[BLEND_THREADS / 32];

  const unsigned full = 0xffffffffu;
  const int tile_x = blockIdx.x, tile_y = blockIdx.y;
  const int tile_id = tile_y * tiles_x + tile_x;
  const int tr = threadIdx.x, nthreads = BLEND_THREADS, lane = tr & 31, warp = tr >> 5;
  int lx, ly;
  map_pixel(16, lx, ly);
  const int ipx = tile_x * 16 + lx, ipy = tile_y * 16 + ly;
  const bool inside = (ipx < img_w) && (ipy < img_h);
  const int pix = inside ? (ipy * img_w + ipx) : 0;
  const float fx0 = (float)__reduce_min_sync(full, inside ? ipx : 0x7fffffff);
  const float fx1 = (float)__reduce_max_sync(full, inside ? ipx : -0x7fffffff);
  const float fy0 = (float)__reduce_min_sync(full, inside ? ipy : 0x7fffffff);
  const float fy1 = (float)__reduce_max_sync(full, inside ? ipy : -0x7fffffff);

  const int2 range = tile_bins[tile_id];
  const int bin_final = inside ? final_idx[pix] : -1;
  {
    const float T_final = inside ? final_Ts[pix] : 1.f;
    float vo_r = 0.f, vo_g = 0.f, vo_b = 0.f, vo_d = 0.f, vo_a = 0.f;
    if (inside) {
      vo_r = v_output[3 * (size_t)pix];
      vo_g = v_output[3 * (size_t)pix + 1];
      vo_b = v_output[3 * (size_t)pix + 2];
      vo_a = v_output_alpha[pix];
      if (DEPTH) vo_d = v_output_depth[pix];
    }
    // the depth channel has a zero background, so it adds nothing to the T_final term
    const float c_final = T_final * (vo_a - (background[0] * vo_r + background[1] * vo_g + background[2] * vo_b));
    s_pixc[warp][lane] = make_float4(vo_r, vo_g, vo_b, __int_as_float(bin_final));
    s_pixs[warp][lane] = make_float4((float)ipx, (float)ipy, T_final, -c_final);
    s_pixd[warp][lane] = vo_d;
  }
  const int warp_bin_final = __reduce_max_sync(full, bin_final);
  if (lane == 0) s_warp_max[warp] = warp_bin_final;
  __syncthreads();
  int cta_bin_final = -1;
  for (int w = 0; w < (nthreads >> 5); ++w) cta_bin_final = max(cta_bin_final, s_warp_max[w]);
  const int end = min(range.y, cta_bin_final + 1);
  const int count = end - range.x;
  if (count <= 0) return;
  const int num_batches = (count + nthreads - 1) / nthreads;

  LaneRec G;
  auto clear = [&]() {
    G.x = G.y = G.A = G.B = G.C = G.o = G.r = G.g = G.b = G.z = 0.f;
    G.id = -1;
    G.sidx = 0x7fffffff;
  };
  auto load = [&](int buf, int slot, int batch_end) {
    const float4 q0 = s_rec[buf][0][slot], q1 = s_rec[buf][1][slot], q2 = s_rec[buf][2][slot];
    G.x = q0.x; G.y = q0.y;
    G.A = q1.x; G.B = q1.y; G.C = q1.z; G.o = q1.w;
    G.r = q2.x; G.g = q2.y; G.b = q2.z; G.z = q2.w;
    G.id = s_gid[buf][slot];
    G.sidx = batch_end - slot;
  };
  clear();
  int n_pending = 0;

  if (end - 1 - tr >= range.x) {
    const int gid = gaussian_ids_sorted[end - 1 - tr];
    s_gid[0][tr] = gid;
    pks_stage(s_rec, 0, tr, gid, num_points, rec);
  }
  pks_commit();
  for (int b = 0; b < num_batches; ++b) {
    const int buf = b & 1;
    const int batch_end = end - 1 - nthreads * b;
    pks_wait_all();
    __syncthreads();
    {
      const int nxt = batch_end - nthreads - tr;
      if (nxt >= range.x) {
        const int gid = gaussian_ids_sorted[nxt];
        s_gid[buf ^ 1][tr] = gid;
        pks_stage(s_rec, buf ^ 1, tr, gid, num_points, rec);
      }
      pks_commit();
    }
    const int batch_size = min(nthreads, batch_end + 1 - range.x);
    const int t_begin = max(0, batch_end - warp_bin_final);
    if (t_begin >= batch_size) continue;
    const int n_list = compact_survivors(s_rec[buf][0], s_rec[buf][1], t_begin, batch_size, fx0, fx1, fy0, fy1,
                                         s_list[warp], lane);
    int li = 0;
    while (n_pending + (n_list - li) >= 32) {
      if (lane >= n_pending) load(buf, s_list[warp][li + lane - n_pending], batch_end);
      li += 32 - n_pending;
      n_pending = 0;
      pks_process_group<DEPTH>(G, s_pixc[warp], s_pixs[warp], s_pixd[warp], lane, grad_rec);
    }
    const int rest = n_list - li;
    if (lane >= n_pending && lane < n_pending + rest) load(buf, s_list[warp][li + lane - n_pending], batch_end);
    n_pending += rest;
  }
  if (n_pending > 0) {
    if (lane >= n_pending) clear();
    pks_process_group<DEPTH>(G, s_pixc[warp], s_pixs[warp], s_pixd[warp], lane, grad_rec);
  }
}

// selected with GSR_PACKED_BWD=scan (blend_packed_tr.cu holds the switch and the default kernel)
int launch_blend_packed_backward_scan(dim3 grid, cudaStream_t st, int img_w, int img_h, int num_points,
                                      const int *gaussian_ids_sorted, const int2 *tile_bins, const float4 *rec,
                                      const float *background, const float *final_Ts, const int *final_idx,
                                      const float *v_output, const float *v_output_depth, const float *v_output_alpha,
                                      float *grad_rec) {
  if (v_output_depth)
    blend_packed_backward_scan_kernel<true><<<grid, BLEND_THREADS, 0, st>>>((int)grid.x, img_w, img_h, num_points,
                                                                            gaussian_ids_sorted, tile_bins, rec, background,
                                                                            final_Ts, final_idx, v_output, v_output_depth,
                                                                            v_output_alpha, grad_rec);
  else
    blend_packed_backward_scan_kernel<false><<<grid, BLEND_THREADS, 0, st>>>((int)grid.x, img_w, img_h, num_points,
                                                                             gaussian_ids_sorted, tile_bins, rec, background,
                                                                             final_Ts, final_idx, v_output, nullptr,
                                                                             v_output_alpha, grad_rec);
  GSR_CHECK_LAUNCH("blend_packed_backward_scan_kernel");
  return GSR_OK;
}

}  // namespace gsr
