// fused.cu — the per-Gaussian half of the FUSED render operator (the model-side glue fused into
// the operator).  One kernel replaces, per view, everything gs_toolkit/models/vanilla_gs.py:759-820 does between
// the parameters and the blend:
//     torch.exp(scales) · quats / |quats| · torch.cat(features_dc, features_rest) · viewdirs = means - cam ·
//     spherical_harmonics(...) · clamp(rgb + 0.5, min=0) · torch.sigmoid(opacities) · project_gaussians(...)
// and writes, per Gaussian, one packed 48-byte blend record {x,y,ext_x,ext_y | A,B,C,opacity | r,g,b,depth} (three
// float4 planes) plus the arrays the binning and the caller need (xys, depths, radii, conics, activated opacity,
// clamp mask).  The adjoint kernel does the reverse in one pass: packed blend gradients -> gradients of the six RAW
// parameter tensors (log-scales, unnormalised quaternions, logit opacities, features_dc, features_rest, means).
//
// HBM-bound: forward reads 56 + 12 K bytes and writes 84 per Gaussian (320 B at K = 16) where the unfused chain
// moves about 830 B (it materialises exp / normalise / sigmoid / cat / viewdirs / clamp tensors, and the SH and
// projection kernels each re-read their inputs).
#include "blend_common.cuh"
#include "project_math.cuh"
#include "sh_math.cuh"

namespace gsr {

constexpr int FU_THREADS = 128;

struct FusedCam {
  float V[12];
  float PM[16];
  float cam[3];
};

__device__ __forceinline__ void load_cam(FusedCam &c, const float *__restrict__ viewmat,
                                         const float *__restrict__ projmat) {
  if (threadIdx.x < 12) c.V[threadIdx.x] = viewmat[threadIdx.x];
  if (threadIdx.x >= 32 && threadIdx.x < 48) c.PM[threadIdx.x - 32] = projmat[threadIdx.x - 32];
  __syncthreads();
  if (threadIdx.x < 3) {
    // camera centre = -R^T t
    const int k = threadIdx.x;
    c.cam[k] = -(c.V[k] * c.V[3] + c.V[4 + k] * c.V[7] + c.V[8 + k] * c.V[11]);
  }
  __syncthreads();
}

// coalesced global <-> shared movement of [rows][row_len] float blocks (row stride `stride` in shared memory)
__device__ __forceinline__ void fu_stage_in(const float *__restrict__ g, float *s, int count, int row_len, int stride,
                                            bool vec_ok) {
  const int tid = threadIdx.x;
  const int nvec = vec_ok ? (count >> 2) : 0;
  const float4 *g4 = reinterpret_cast<const float4 *>(g);
  for (int i = tid; i < nvec; i += FU_THREADS) {
    const float4 v = __ldg(g4 + i);
    const int f = i << 2;
    int r = f / row_len, c = f - r * row_len;
    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s[r * stride + c] = vv[j];
      if (++c == row_len) { c = 0; ++r; }
    }
  }
  for (int f = (nvec << 2) + tid; f < count; f += FU_THREADS) {
    const int r = f / row_len, c = f - r * row_len;
    s[r * stride + c] = __ldg(g + f);
  }
}

__device__ __forceinline__ void fu_stage_out(float *__restrict__ g, const float *s, int count, int row_len, int stride,
                                             bool vec_ok) {
  const int tid = threadIdx.x;
  const int nvec = vec_ok ? (count >> 2) : 0;
  float4 *g4 = reinterpret_cast<float4 *>(g);
  for (int i = tid; i < nvec; i += FU_THREADS) {
    const int f = i << 2;
    int r = f / row_len, c = f - r * row_len;
    float vv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      vv[j] = s[r * stride + c];
      if (++c == row_len) { c = 0; ++r; }
    }
    g4[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
  }
  for (int f = (nvec << 2) + tid; f < count; f += FU_THREADS) {
    const int r = f / row_len, c = f - r * row_len;
    g[f] = s[r * stride + c];
  }
}

// K = number of SH bases stored (features_dc is basis 0, features_rest the K-1 others); rest_len = 3 (K - 1)
__global__ void __launch_bounds__(FU_THREADS)
fused_preprocess_forward_kernel(int n, int K, int deg_use, const float *__restrict__ means3d,
                                const float *__restrict__ scales_raw, const float *__restrict__ quats_raw,
                                const float *__restrict__ opacities_raw, const float *__restrict__ features_dc,
                                const float *__restrict__ features_rest, const float *__restrict__ viewmat,
                                const float *__restrict__ projmat, float glob_scale, float fx, float fy, float cx,
                                float cy, float tan_fovx, float tan_fovy, int img_w, int img_h, int tiles_x,
                                int tiles_y, int block_width,
                                float clip_thresh, float4 *__restrict__ rec0, float4 *__restrict__ rec1,
                                float4 *__restrict__ rec2, float *__restrict__ xys, float *__restrict__ depths,
                                int *__restrict__ radii, float *__restrict__ conics, float *__restrict__ opac_act,
                                int *__restrict__ clamp_mask, float *__restrict__ compensation, int vec_ok) {
  extern __shared__ float smem[];
  __shared__ FusedCam cam;
  load_cam(cam, viewmat, projmat);
  const int rest_len = 3 * (K - 1), stride = rest_len | 1;
  const int g0 = blockIdx.x * FU_THREADS;
  const int rows = min(FU_THREADS, n - g0);
  if (rest_len > 0) fu_stage_in(features_rest + (size_t)g0 * rest_len, smem, rows * rest_len, rest_len, stride, vec_ok != 0);
  __syncthreads();
  const int tid = threadIdx.x;
  if (tid >= rows) return;
  const int g = g0 + tid;
  const size_t G = (size_t)g;

  const float px = means3d[3 * G], py = means3d[3 * G + 1], pz = means3d[3 * G + 2];
  // activations (models/vanilla_gs.py:759-769,813): exp, sigmoid; the quaternion is normalised inside quat_to_rotmat
  const float s0 = glob_scale * __expf(scales_raw[3 * G]), s1 = glob_scale * __expf(scales_raw[3 * G + 1]),
              s2 = glob_scale * __expf(scales_raw[3 * G + 2]);
  const float4 q = reinterpret_cast<const float4 *>(quats_raw)[g];
  float opac = 1.f / (1.f + __expf(-opacities_raw[g]));

  const ProjFwd p = project_one(px, py, pz, s0, s1, s2, q.x, q.y, q.z, q.w, cam.V, cam.PM, fx, fy, cx, cy, tan_fovx,
                                tan_fovy, img_w, img_h, tiles_x, tiles_y, block_width, clip_thresh);
  // rasterize_mode == "antialiased": opacities = sigmoid(raw) * compensation (vanilla_gs.py:813-816)
  if (compensation != nullptr) {
    compensation[g] = p.comp;
    opac *= p.comp;
  }

  // SH colour (sh.cuh:33-98) + clamp(rgb + 0.5, min=0) (vanilla_gs.py:806-807)
  float Y[25];
  sh_basis_all(deg_use, px - cam.cam[0], py - cam.cam[1], pz - cam.cam[2], Y);
  const int Ku = deg_use == 0 ? 1 : deg_use == 1 ? 4 : deg_use == 2 ? 9 : deg_use == 3 ? 16 : 25;
  const float *rest = smem + tid * stride;
  float rgb[3];
  int mask = 0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float acc = Y[0] * features_dc[3 * G + c];
#pragma unroll
    for (int k = 1; k < 25; ++k)
      if (k < Ku) acc += Y[k] * rest[3 * (k - 1) + c];
    const float v = acc + 0.5f;
    if (v > 0.f) mask |= 1 << c;
    rgb[c] = fmaxf(v, 0.f);
  }

  const bool visible = p.radius > 0;
  float ex = -1e30f, ey = -1e30f;
  if (visible) alpha_extents(p.conic[0], p.conic[1], p.conic[2], opac, ex, ey);
  rec0[g] = make_float4(p.x, p.y, ex, ey);
  rec1[g] = make_float4(-0.5f * kLog2e * p.conic[0], -kLog2e * p.conic[1], -0.5f * kLog2e * p.conic[2], opac);
  rec2[g] = make_float4(rgb[0], rgb[1], rgb[2], p.depth);
  reinterpret_cast<float2 *>(xys)[g] = make_float2(p.x, p.y);
  depths[g] = p.depth;
  radii[g] = p.radius;
  conics[3 * G] = p.conic[0];
  conics[3 * G + 1] = p.conic[1];
  conics[3 * G + 2] = p.conic[2];
  opac_act[g] = opac;
  clamp_mask[g] = mask;
}

// grad_rec [N,12]: {v_x, v_y, v_opacity, v_depth | v_a, v_b, v_c, - | v_r, v_g, v_b, -} accumulated by the packed blend
// adjoint; v_xys_extra (nullable) = a gradient that reached the returned xys tensor from outside the operator.
__global__ void __launch_bounds__(FU_THREADS)
fused_preprocess_backward_kernel(int n, int K, int deg_use, const float *__restrict__ means3d,
                                 const float *__restrict__ scales_raw, const float *__restrict__ quats_raw,
                                 const float *__restrict__ opacities_raw, const float *__restrict__ viewmat,
                                 const float *__restrict__ projmat, float glob_scale, float fx, float fy, int img_w,
                                 int img_h, const int *__restrict__ radii, const float *__restrict__ conics,
                                 const int *__restrict__ clamp_mask, const float *__restrict__ compensation,
                                 const float4 *__restrict__ grad_rec,
                                 const float *__restrict__ v_xys_extra, float *__restrict__ v_means3d,
                                 float *__restrict__ v_scales_raw, float *__restrict__ v_quats_raw,
                                 float *__restrict__ v_opacities_raw, float *__restrict__ v_features_dc,
                                 float *__restrict__ v_features_rest, int vec_ok) {
  extern __shared__ float smem[];
  __shared__ FusedCam cam;
  load_cam(cam, viewmat, projmat);
  const int rest_len = 3 * (K - 1), stride = rest_len | 1;
  const int g0 = blockIdx.x * FU_THREADS;
  const int rows = min(FU_THREADS, n - g0);
  const int tid = threadIdx.x;
  if (tid < rows) {
    const int g = g0 + tid;
    const size_t G = (size_t)g;
    const float4 ga = grad_rec[3 * G], gb = grad_rec[3 * G + 1], gc = grad_rec[3 * G + 2];
    const bool visible = radii[g] > 0;
    const float px = means3d[3 * G], py = means3d[3 * G + 1], pz = means3d[3 * G + 2];
    const float e0 = __expf(scales_raw[3 * G]), e1 = __expf(scales_raw[3 * G + 1]), e2 = __expf(scales_raw[3 * G + 2]);
    const float4 q = reinterpret_cast<const float4 *>(quats_raw)[g];

    // cov3d is recomputed (forward.cu:445-464) instead of being stored and re-read
    float c3[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (visible) {
      float R[9];
      quat_to_rotmat(q.x, q.y, q.z, q.w, R);
      const float s0 = glob_scale * e0, s1 = glob_scale * e1, s2 = glob_scale * e2;
      const float M[9] = {R[0] * s0, R[1] * s1, R[2] * s2, R[3] * s0, R[4] * s1, R[5] * s2, R[6] * s0, R[7] * s1, R[8] * s2};
      c3[0] = M[0] * M[0] + M[1] * M[1] + M[2] * M[2];
      c3[1] = M[0] * M[3] + M[1] * M[4] + M[2] * M[5];
      c3[2] = M[0] * M[6] + M[1] * M[7] + M[2] * M[8];
      c3[3] = M[3] * M[3] + M[4] * M[4] + M[5] * M[5];
      c3[4] = M[3] * M[6] + M[4] * M[7] + M[5] * M[8];
      c3[5] = M[6] * M[6] + M[7] * M[7] + M[8] * M[8];
    }
    float vx = ga.x, vy = ga.y;
    if (v_xys_extra != nullptr) {
      vx += v_xys_extra[2 * G];
      vy += v_xys_extra[2 * G + 1];
    }
    // antialiased: opacity = sigmoid(raw) * comp  =>  v_sigmoid = v_opacity * comp, v_comp = v_opacity * sigmoid(raw)
    const float sig = 1.f / (1.f + __expf(-opacities_raw[g]));
    const float comp = compensation != nullptr ? compensation[g] : 1.f;
    const float v_comp = compensation != nullptr ? ga.z * sig : 0.f;
    const ProjBwd pg = project_one_vjp(visible, px, py, pz, e0, e1, e2, glob_scale, q.x, q.y, q.z, q.w, cam.V, cam.PM, fx,
                                       fy, img_w, img_h, c3, conics[3 * G], conics[3 * G + 1], conics[3 * G + 2], comp,
                                       vx, vy, ga.w, gb.x, gb.y, gb.z, v_comp);
    v_means3d[3 * G] = pg.mean[0];
    v_means3d[3 * G + 1] = pg.mean[1];
    v_means3d[3 * G + 2] = pg.mean[2];
    // scale = exp(raw)  =>  v_raw = v_scale * scale
    v_scales_raw[3 * G] = pg.scale[0] * e0;
    v_scales_raw[3 * G + 1] = pg.scale[1] * e1;
    v_scales_raw[3 * G + 2] = pg.scale[2] * e2;
    // q_hat = q / |q|  =>  v_q = (v_qhat - q_hat <q_hat, v_qhat>) / |q|   (autograd of quats / quats.norm, vanilla_gs.py:769)
    {
      const float inv = rsqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
      const float h0 = q.x * inv, h1 = q.y * inv, h2 = q.z * inv, h3 = q.w * inv;
      const float d = h0 * pg.quat[0] + h1 * pg.quat[1] + h2 * pg.quat[2] + h3 * pg.quat[3];
      reinterpret_cast<float4 *>(v_quats_raw)[g] = make_float4((pg.quat[0] - h0 * d) * inv, (pg.quat[1] - h1 * d) * inv,
                                                               (pg.quat[2] - h2 * d) * inv, (pg.quat[3] - h3 * d) * inv);
    }
    // opacity = sigmoid(raw)  =>  v_raw = v_opacity * o (1 - o)
    v_opacities_raw[g] = (ga.z * comp) * sig * (1.f - sig);
    // SH adjoint (sh.cuh:100-186) with the clamp mask; basis 0 -> features_dc, the others -> features_rest
    float Y[25];
    sh_basis_all(deg_use, px - cam.cam[0], py - cam.cam[1], pz - cam.cam[2], Y);
    const int Ku = deg_use == 0 ? 1 : deg_use == 1 ? 4 : deg_use == 2 ? 9 : deg_use == 3 ? 16 : 25;
    const int mask = clamp_mask[g];
    const float v0 = (mask & 1) ? gc.x : 0.f, v1 = (mask & 2) ? gc.y : 0.f, v2 = (mask & 4) ? gc.z : 0.f;
    v_features_dc[3 * G] = Y[0] * v0;
    v_features_dc[3 * G + 1] = Y[0] * v1;
    v_features_dc[3 * G + 2] = Y[0] * v2;
    float *o = smem + tid * stride;
#pragma unroll
    for (int k = 1; k < 25; ++k) {
      if (k < K) {
        const float y = (k < Ku) ? Y[k] : 0.f;
        o[3 * (k - 1)] = y * v0;
        o[3 * (k - 1) + 1] = y * v1;
        o[3 * (k - 1) + 2] = y * v2;
      }
    }
  }
  __syncthreads();
  if (rest_len > 0) fu_stage_out(v_features_rest + (size_t)g0 * rest_len, smem, rows * rest_len, rest_len, stride, vec_ok != 0);
}

}  // namespace gsr

extern "C" {

GSR_API int gsr_fused_preprocess_forward(int num_points, int sh_degree, int degrees_to_use, const float *means3d,
                                         const float *scales_raw, const float *quats_raw, const float *opacities_raw,
                                         const float *features_dc, const float *features_rest, const float *viewmat,
                                         const float *projmat, float glob_scale, float fx, float fy, float cx, float cy,
                                         unsigned img_height, unsigned img_width, unsigned block_width,
                                         float clip_thresh, float *records, float *xys, float *depths, int32_t *radii,
                                         float *conics, float *opacities, int32_t *clamp_mask,
                                         float *compensation /*nullable*/, void *stream) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_fused_preprocess_forward");
  GSR_REQUIRE(num_points >= 0, GSR_ERR_INVALID_ARGUMENT, "fused_preprocess_forward: num_points < 0");
  GSR_REQUIRE(sh_degree >= 0 && sh_degree <= 4, GSR_ERR_UNSUPPORTED, "fused_preprocess_forward: sh_degree %d not in [0,4]", sh_degree);
  GSR_REQUIRE(degrees_to_use >= 0 && degrees_to_use <= sh_degree, GSR_ERR_INVALID_ARGUMENT,
              "fused_preprocess_forward: degrees_to_use %d not in [0,%d]", degrees_to_use, sh_degree);
  GSR_REQUIRE(block_width > 1 && block_width <= 16, GSR_ERR_INVALID_ARGUMENT,
              "block_width must be between 2 and 16 (got %u)", block_width);
  if (num_points == 0) return GSR_OK;
  GSR_REQUIRE(means3d && scales_raw && quats_raw && opacities_raw && features_dc && viewmat && projmat && records &&
                  xys && depths && radii && conics && opacities && clamp_mask && (sh_degree == 0 || features_rest),
              GSR_ERR_INVALID_ARGUMENT, "fused_preprocess_forward: null pointer");
  GSR_REQUIRE((uintptr_t)quats_raw % 16 == 0 && (uintptr_t)records % 16 == 0 && (uintptr_t)xys % 8 == 0,
              GSR_ERR_INVALID_ARGUMENT, "fused_preprocess_forward: quats / records must be 16-byte, xys 8-byte aligned");
  const int K = (sh_degree + 1) * (sh_degree + 1);
  const int rest_len = 3 * (K - 1);
  const size_t smem = (size_t)FU_THREADS * (rest_len | 1) * sizeof(float);
  const int vec_ok = ((uintptr_t)features_rest % 16 == 0) ? 1 : 0;
  float4 *rec = reinterpret_cast<float4 *>(records);
  fused_preprocess_forward_kernel<<<cdiv(num_points, FU_THREADS), FU_THREADS, smem, (cudaStream_t)stream>>>(
      num_points, K, degrees_to_use, means3d, scales_raw, quats_raw, opacities_raw, features_dc, features_rest, viewmat,
      projmat, glob_scale, fx, fy, cx, cy, tan_half_fov(img_width, fx), tan_half_fov(img_height, fy), (int)img_width,
      (int)img_height, (int)cdiv(img_width, block_width),
      (int)cdiv(img_height, block_width), (int)block_width, clip_thresh, rec, rec + num_points, rec + 2 * (size_t)num_points,
      xys, depths, radii, conics, opacities, clamp_mask, compensation, vec_ok);
  GSR_CHECK_LAUNCH("fused_preprocess_forward_kernel");
  return GSR_OK;
}

GSR_API int gsr_fused_preprocess_backward(int num_points, int sh_degree, int degrees_to_use, const float *means3d,
                                          const float *scales_raw, const float *quats_raw, const float *opacities_raw,
                                          const float *viewmat, const float *projmat, float glob_scale, float fx,
                                          float fy, unsigned img_height, unsigned img_width, const int32_t *radii,
                                          const float *conics, const int32_t *clamp_mask,
                                          const float *compensation /*nullable*/, const float *grad_records,
                                          const float *v_xys_extra, float *v_means3d, float *v_scales_raw,
                                          float *v_quats_raw, float *v_opacities_raw, float *v_features_dc,
                                          float *v_features_rest, void *stream) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_fused_preprocess_backward");
  GSR_REQUIRE(num_points >= 0, GSR_ERR_INVALID_ARGUMENT, "fused_preprocess_backward: num_points < 0");
  GSR_REQUIRE(sh_degree >= 0 && sh_degree <= 4, GSR_ERR_UNSUPPORTED, "fused_preprocess_backward: sh_degree %d not in [0,4]", sh_degree);
  GSR_REQUIRE(degrees_to_use >= 0 && degrees_to_use <= sh_degree, GSR_ERR_INVALID_ARGUMENT,
              "fused_preprocess_backward: degrees_to_use %d not in [0,%d]", degrees_to_use, sh_degree);
  if (num_points == 0) return GSR_OK;
  GSR_REQUIRE(means3d && scales_raw && quats_raw && opacities_raw && viewmat && projmat && radii && conics && clamp_mask &&
                  grad_records && v_means3d && v_scales_raw && v_quats_raw && v_opacities_raw && v_features_dc &&
                  (sh_degree == 0 || v_features_rest),
              GSR_ERR_INVALID_ARGUMENT, "fused_preprocess_backward: null pointer");
  GSR_REQUIRE((uintptr_t)quats_raw % 16 == 0 && (uintptr_t)v_quats_raw % 16 == 0 && (uintptr_t)grad_records % 16 == 0,
              GSR_ERR_INVALID_ARGUMENT, "fused_preprocess_backward: quats / grad_records must be 16-byte aligned");
  const int K = (sh_degree + 1) * (sh_degree + 1);
  const int rest_len = 3 * (K - 1);
  const size_t smem = (size_t)FU_THREADS * (rest_len | 1) * sizeof(float);
  const int vec_ok = ((uintptr_t)v_features_rest % 16 == 0) ? 1 : 0;
  fused_preprocess_backward_kernel<<<cdiv(num_points, FU_THREADS), FU_THREADS, smem, (cudaStream_t)stream>>>(
      num_points, K, degrees_to_use, means3d, scales_raw, quats_raw, opacities_raw, viewmat, projmat, glob_scale, fx, fy,
      (int)img_width, (int)img_height, radii, conics, clamp_mask, compensation,
      reinterpret_cast<const float4 *>(grad_records),
      v_xys_extra, v_means3d, v_scales_raw, v_quats_raw, v_opacities_raw, v_features_dc, v_features_rest, vec_ok);
  GSR_CHECK_LAUNCH("fused_preprocess_backward_kernel");
  return GSR_OK;
}
}
