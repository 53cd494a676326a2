// project.cu — per-Gaussian 3D->2D EWA projection, its adjoint, and cov2d bounds, for sm_100a.
//
// Replaces project_gaussians_forward_kernel (reference csrc/forward.cu:13-90 with device functions
// forward.cu:398-464 and helpers.cuh:7-59,93-122,144-159,202-219), project_gaussians_backward_kernel
// (backward.cu:305-347 with VJPs backward.cu:350-453, helpers.cuh:62-90,125-142,161-200) and
// compute_cov2d_bounds_kernel (bindings.cu:19-37).
//
// HBM-bound streaming kernels, one thread per Gaussian.  The 3x3 algebra is written out on scalars
// (row-major) instead of the reference's vendored glm; the expression order of every quantity that feeds
// a discrete decision (radius ceil, tile-bbox truncation, near clip) follows the reference so that the
// integer outputs agree.  The forward writes every output for every Gaussian (zeros for culled ones), so
// the 7 torch::zeros memsets of the reference binding (bindings.cu:126-139) are not needed.
#include "common.cuh"

namespace gsr {

constexpr int PROJ_THREADS = 256;

struct CamParams {
  float V[12];   // viewmat, row-major 3x4
  float PM[16];  // projmat, row-major 4x4
};

__global__ void __launch_bounds__(PROJ_THREADS)
project_forward_kernel(int n, const float *__restrict__ means3d, const float *__restrict__ scales,
                       float glob_scale, const float *__restrict__ quats,
                       const float *__restrict__ viewmat, const float *__restrict__ projmat, float fx,
                       float fy, float cx, float cy, float tan_fovx, float tan_fovy, int img_w, int img_h,
                       int tiles_x, int tiles_y, int block_width, float clip_thresh, float *__restrict__ cov3d,
                       float *__restrict__ xys, float *__restrict__ depths, int *__restrict__ radii,
                       float *__restrict__ conics, float *__restrict__ compensation,
                       int *__restrict__ num_tiles_hit, int quats_vec) {
  __shared__ CamParams cam;
  if (threadIdx.x < 12) cam.V[threadIdx.x] = viewmat[threadIdx.x];
  if (threadIdx.x >= 32 && threadIdx.x < 48) cam.PM[threadIdx.x - 32] = projmat[threadIdx.x - 32];
  __syncthreads();
  const int idx = blockIdx.x * PROJ_THREADS + threadIdx.x;
  if (idx >= n) return;
  const float *V = cam.V, *PM = cam.PM;

  float o_cov3d[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float o_conic[3] = {0.f, 0.f, 0.f};
  float o_x = 0.f, o_y = 0.f, o_depth = 0.f, o_comp = 0.f;
  int o_radius = 0, o_tiles = 0;

  // every input of the Gaussian is requested before the first use (one exposed round trip to HBM instead of two: the
  // quaternion and the scales used to be loaded behind the near-plane branch)
  const float px = means3d[3 * (size_t)idx], py = means3d[3 * (size_t)idx + 1], pz = means3d[3 * (size_t)idx + 2];
  float4 q;
  if (quats_vec) {
    q = reinterpret_cast<const float4 *>(quats)[idx];
  } else {
    q = make_float4(quats[4 * (size_t)idx], quats[4 * (size_t)idx + 1], quats[4 * (size_t)idx + 2],
                    quats[4 * (size_t)idx + 3]);
  }
  const float sc0 = scales[3 * (size_t)idx], sc1 = scales[3 * (size_t)idx + 1], sc2 = scales[3 * (size_t)idx + 2];
  // clip_near_plane (helpers.cuh:210-219)
  const float vx = V[0] * px + V[1] * py + V[2] * pz + V[3];
  const float vy = V[4] * px + V[5] * py + V[6] * pz + V[7];
  const float vz = V[8] * px + V[9] * py + V[10] * pz + V[11];
  do {
    if (vz <= clip_thresh) break;

    // scale_rot_to_cov3d (forward.cu:445-464): M = R * S, Sigma = M * M^T
    float R[9];
    quat_to_rotmat(q.x, q.y, q.z, q.w, R);
    const float s0 = glob_scale * sc0, s1 = glob_scale * sc1, s2 = glob_scale * sc2;
    float M[9] = {R[0] * s0, R[1] * s1, R[2] * s2, R[3] * s0, R[4] * s1, R[5] * s2, R[6] * s0, R[7] * s1, R[8] * s2};
    o_cov3d[0] = M[0] * M[0] + M[1] * M[1] + M[2] * M[2];
    o_cov3d[1] = M[0] * M[3] + M[1] * M[4] + M[2] * M[5];
    o_cov3d[2] = M[0] * M[6] + M[1] * M[7] + M[2] * M[8];
    o_cov3d[3] = M[3] * M[3] + M[4] * M[4] + M[5] * M[5];
    o_cov3d[4] = M[3] * M[6] + M[4] * M[7] + M[5] * M[8];
    o_cov3d[5] = M[6] * M[6] + M[7] * M[7] + M[8] * M[8];

    // project_cov3d_ewa (forward.cu:398-442)
    // tan_fov: `0.5 * img_size.x / fx` in DOUBLE, rounded once, like the reference (host side: tan_half_fov)
    const float lim_x = 1.3f * tan_fovx, lim_y = 1.3f * tan_fovy;
    const float tz = vz;
    const float tx = tz * fminf(lim_x, fmaxf(-lim_x, vx / tz));
    const float ty = tz * fminf(lim_y, fmaxf(-lim_y, vy / tz));
    const float rz = 1.f / tz, rz2 = rz * rz;
    // T = J * W (only the two non-zero rows of J)
    const float j00 = fx * rz, j02 = -fx * tx * rz2, j11 = fy * rz, j12 = -fy * ty * rz2;
    const float T0[3] = {j00 * V[0] + j02 * V[8], j00 * V[1] + j02 * V[9], j00 * V[2] + j02 * V[10]};
    const float T1[3] = {j11 * V[4] + j12 * V[8], j11 * V[5] + j12 * V[9], j11 * V[6] + j12 * V[10]};
    // TV = T * Sigma
    const float *c3 = o_cov3d;
    const float a0 = T0[0] * c3[0] + T0[1] * c3[1] + T0[2] * c3[2];
    const float a1 = T0[0] * c3[1] + T0[1] * c3[3] + T0[2] * c3[4];
    const float a2 = T0[0] * c3[2] + T0[1] * c3[4] + T0[2] * c3[5];
    const float b0 = T1[0] * c3[0] + T1[1] * c3[1] + T1[2] * c3[2];
    const float b1 = T1[0] * c3[1] + T1[1] * c3[3] + T1[2] * c3[4];
    const float b2 = T1[0] * c3[2] + T1[1] * c3[4] + T1[2] * c3[5];
    const float c00 = a0 * T0[0] + a1 * T0[1] + a2 * T0[2];
    const float c01 = b0 * T0[0] + b1 * T0[1] + b2 * T0[2];  // the reference reads cov[0][1] = column 0, row 1 of glm's (T V) T^T
    const float c11 = b0 * T1[0] + b1 * T1[1] + b2 * T1[2];
    const float det_orig = c00 * c11 - c01 * c01;
    const float cxx = c00 + 0.3f, cxy = c01, cyy = c11 + 0.3f;
    const float det_blur = cxx * cyy - cxy * cxy;
    const float comp = sqrtf(fmaxf(0.f, det_orig / det_blur));

    float ca, cb, cc, radius;
    if (!cov2d_to_conic_radius(cxx, cxy, cyy, ca, cb, cc, radius)) break;
    o_conic[0] = ca;
    o_conic[1] = cb;
    o_conic[2] = cc;

    // project_pix (helpers.cuh:114-122)
    const float hx = PM[0] * px + PM[1] * py + PM[2] * pz + PM[3];
    const float hy = PM[4] * px + PM[5] * py + PM[6] * pz + PM[7];
    const float hw = PM[12] * px + PM[13] * py + PM[14] * pz + PM[15];
    const float rw = 1.f / (hw + 1e-6f);
    const float ux = 0.5f * (float)img_w * (hx * rw) + cx - 0.5f;
    const float uy = 0.5f * (float)img_h * (hy * rw) + cy - 0.5f;
    int x0, y0, x1, y1;
    tile_bbox(ux, uy, radius, tiles_x, tiles_y, block_width, x0, y0, x1, y1);
    const int area = (x1 - x0) * (y1 - y0);
    if (area <= 0) break;
    o_tiles = area;
    o_depth = vz;
    o_radius = (int)radius;
    o_x = ux;
    o_y = uy;
    o_comp = comp;
  } while (0);

#pragma unroll
  for (int k = 0; k < 6; ++k) cov3d[6 * (size_t)idx + k] = o_cov3d[k];
#pragma unroll
  for (int k = 0; k < 3; ++k) conics[3 * (size_t)idx + k] = o_conic[k];
  reinterpret_cast<float2 *>(xys)[idx] = make_float2(o_x, o_y);
  depths[idx] = o_depth;
  radii[idx] = o_radius;
  compensation[idx] = o_comp;
  num_tiles_hit[idx] = o_tiles;
}

__global__ void __launch_bounds__(PROJ_THREADS)
project_backward_kernel(int n, const float *__restrict__ means3d, const float *__restrict__ scales,
                        float glob_scale, const float *__restrict__ quats,
                        const float *__restrict__ viewmat, const float *__restrict__ projmat, float fx,
                        float fy, int img_w, int img_h, const float *__restrict__ cov3d,
                        const int *__restrict__ radii, const float *__restrict__ conics,
                        const float *__restrict__ compensation, const float *__restrict__ v_xy,
                        const float *__restrict__ v_depth, const float *__restrict__ v_conic,
                        const float *__restrict__ v_compensation, float *__restrict__ v_cov2d,
                        float *__restrict__ v_cov3d, float *__restrict__ v_mean3d,
                        float *__restrict__ v_scale, float *__restrict__ v_quat) {
  __shared__ CamParams cam;
  if (threadIdx.x < 12) cam.V[threadIdx.x] = viewmat[threadIdx.x];
  if (threadIdx.x >= 32 && threadIdx.x < 48) cam.PM[threadIdx.x - 32] = projmat[threadIdx.x - 32];
  __syncthreads();
  const int idx = blockIdx.x * PROJ_THREADS + threadIdx.x;
  if (idx >= n) return;
  const float *V = cam.V, *PM = cam.PM;
  const size_t I = (size_t)idx;

  float o_mean[3] = {0.f, 0.f, 0.f}, o_scale[3] = {0.f, 0.f, 0.f}, o_quat[4] = {0.f, 0.f, 0.f, 0.f};
  float o_cov2d[3] = {0.f, 0.f, 0.f}, o_cov3d[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

  if (radii[idx] > 0) {
    const float px = means3d[3 * I], py = means3d[3 * I + 1], pz = means3d[3 * I + 2];
    // project_pix_vjp (helpers.cuh:125-142)
    const float hx = PM[0] * px + PM[1] * py + PM[2] * pz + PM[3];
    const float hy = PM[4] * px + PM[5] * py + PM[6] * pz + PM[7];
    const float hw = PM[12] * px + PM[13] * py + PM[14] * pz + PM[15];
    const float rw = 1.f / (hw + 1e-6f);
    const float2 vxy = reinterpret_cast<const float2 *>(v_xy)[idx];
    const float gx = 0.5f * (float)img_w * vxy.x, gy = 0.5f * (float)img_h * vxy.y;
    const float vt0 = gx * rw, vt1 = gy * rw, vt3 = -(gx * hx + gy * hy) * rw * rw;
#pragma unroll
    for (int k = 0; k < 3; ++k) o_mean[k] = PM[k] * vt0 + PM[4 + k] * vt1 + PM[12 + k] * vt3;
    // depth (backward.cu:325-331)
    const float vz = v_depth[idx];
    o_mean[0] += V[8] * vz;
    o_mean[1] += V[9] * vz;
    o_mean[2] += V[10] * vz;

    // cov2d_to_conic_vjp (helpers.cuh:62-74): Y = -X G X
    const float ca = conics[3 * I], cb = conics[3 * I + 1], cc = conics[3 * I + 2];
    const float g00 = v_conic[3 * I], g01 = v_conic[3 * I + 1] / 2.f, g11 = v_conic[3 * I + 2];
    const float xg00 = ca * g00 + cb * g01, xg01 = ca * g01 + cb * g11;
    const float xg10 = cb * g00 + cc * g01, xg11 = cb * g01 + cc * g11;
    const float y00 = -(xg00 * ca + xg01 * cb), y01 = -(xg00 * cb + xg01 * cc);
    const float y10 = -(xg10 * ca + xg11 * cb), y11 = -(xg10 * cb + xg11 * cc);
    o_cov2d[0] = y00;
    o_cov2d[1] = y01 + y10;
    o_cov2d[2] = y11;
    // cov2d_to_compensation_vjp (helpers.cuh:76-90)
    {
      const float comp = compensation[idx], vcomp = v_compensation[idx];
      const float inv_det = ca * cc - cb * cb;
      const float om = 1.f - comp * comp;
      const float u = vcomp * 0.5f / (comp + 1e-6f);
      o_cov2d[0] += u * (om * ca - 0.3f * inv_det);
      o_cov2d[1] += 2.f * u * (om * cb);
      o_cov2d[2] += u * (om * cc - 0.3f * inv_det);
    }

    // project_cov3d_ewa_vjp (backward.cu:350-423) — t is NOT clamped here (reference quirk)
    const float tx = V[0] * px + V[1] * py + V[2] * pz + V[3];
    const float ty = V[4] * px + V[5] * py + V[6] * pz + V[7];
    const float tz = V[8] * px + V[9] * py + V[10] * pz + V[11];
    const float rz = 1.f / tz, rz2 = rz * rz, rz3 = rz2 * rz;
    const float W[9] = {V[0], V[1], V[2], V[4], V[5], V[6], V[8], V[9], V[10]};
    const float J[9] = {fx * rz, 0.f, -fx * tx * rz2, 0.f, fy * rz, -fy * ty * rz2, 0.f, 0.f, 0.f};
    const float *c3 = cov3d + 6 * I;
    const float Vm[9] = {c3[0], c3[1], c3[2], c3[1], c3[3], c3[4], c3[2], c3[4], c3[5]};
    const float Gc[9] = {o_cov2d[0], 0.5f * o_cov2d[1], 0.f, 0.5f * o_cov2d[1], o_cov2d[2], 0.f, 0.f, 0.f, 0.f};
    float T[9], Tt[9], tmp[9], vV[9], vT[9];
    mat3_mul(J, W, T);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) Tt[3 * j + i] = T[3 * i + j];
    mat3_mul(Tt, Gc, tmp);
    mat3_mul(tmp, T, vV);
    o_cov3d[0] = vV[0];
    o_cov3d[1] = vV[1] + vV[3];
    o_cov3d[2] = vV[2] + vV[6];
    o_cov3d[3] = vV[4];
    o_cov3d[4] = vV[5] + vV[7];
    o_cov3d[5] = vV[8];
    // v_T = G T V^T + G^T T V = 2 G T V (G, V symmetric)
    mat3_mul(Gc, T, tmp);
    mat3_mul(tmp, Vm, vT);
    // v_J = v_T W^T ; only entries (0,0),(0,2),(1,1),(1,2) are used
    const float vJ00 = 2.f * (vT[0] * W[0] + vT[1] * W[1] + vT[2] * W[2]);
    const float vJ02 = 2.f * (vT[0] * W[6] + vT[1] * W[7] + vT[2] * W[8]);
    const float vJ11 = 2.f * (vT[3] * W[3] + vT[4] * W[4] + vT[5] * W[5]);
    const float vJ12 = 2.f * (vT[3] * W[6] + vT[4] * W[7] + vT[5] * W[8]);
    const float v_t0 = -fx * rz2 * vJ02;
    const float v_t1 = -fy * rz2 * vJ12;
    const float v_t2 = -fx * rz2 * vJ00 + 2.f * fx * tx * rz3 * vJ02 - fy * rz2 * vJ11 + 2.f * fy * ty * rz3 * vJ12;
    o_mean[0] += W[0] * v_t0 + W[3] * v_t1 + W[6] * v_t2;
    o_mean[1] += W[1] * v_t0 + W[4] * v_t1 + W[7] * v_t2;
    o_mean[2] += W[2] * v_t0 + W[5] * v_t1 + W[8] * v_t2;

    // scale_rot_to_cov3d_vjp (backward.cu:425-453)
    const float vVs[9] = {o_cov3d[0], 0.5f * o_cov3d[1], 0.5f * o_cov3d[2], 0.5f * o_cov3d[1], o_cov3d[3],
                          0.5f * o_cov3d[4], 0.5f * o_cov3d[2], 0.5f * o_cov3d[4], o_cov3d[5]};
    const float qw = quats[4 * I], qx = quats[4 * I + 1], qy = quats[4 * I + 2], qz = quats[4 * I + 3];
    float R[9];
    quat_to_rotmat(qw, qx, qy, qz, R);
    const float S[3] = {glob_scale * scales[3 * I], glob_scale * scales[3 * I + 1], glob_scale * scales[3 * I + 2]};
    float M[9], vM[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) M[3 * r + c] = R[3 * r + c] * S[c];
    mat3_mul(vVs, M, vM);
#pragma unroll
    for (int k = 0; k < 9; ++k) vM[k] *= 2.f;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      o_scale[c] = (R[c] * vM[c] + R[3 + c] * vM[3 + c] + R[6 + c] * vM[6 + c]) * glob_scale;
    float D[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) D[3 * r + c] = vM[3 * r + c] * S[c];
    // quat_to_rotmat_vjp (helpers.cuh:161-200), D[3*i+j] = v_R(row i, col j)
    const float s = rsqrtf(qw * qw + qx * qx + qy * qy + qz * qz);
    const float w = qw * s, x = qx * s, y = qy * s, z = qz * s;
    o_quat[0] = 2.f * (x * (D[7] - D[5]) + y * (D[2] - D[6]) + z * (D[3] - D[1]));
    o_quat[1] = 2.f * (-2.f * x * (D[4] + D[8]) + y * (D[3] + D[1]) + z * (D[6] + D[2]) + w * (D[7] - D[5]));
    o_quat[2] = 2.f * (x * (D[3] + D[1]) - 2.f * y * (D[0] + D[8]) + z * (D[7] + D[5]) + w * (D[2] - D[6]));
    o_quat[3] = 2.f * (x * (D[6] + D[2]) + y * (D[7] + D[5]) - 2.f * z * (D[0] + D[4]) + w * (D[3] - D[1]));
  }

#pragma unroll
  for (int k = 0; k < 3; ++k) {
    v_mean3d[3 * I + k] = o_mean[k];
    v_scale[3 * I + k] = o_scale[k];
  }
  reinterpret_cast<float4 *>(v_quat)[idx] = make_float4(o_quat[0], o_quat[1], o_quat[2], o_quat[3]);
  if (v_cov2d) {
#pragma unroll
    for (int k = 0; k < 3; ++k) v_cov2d[3 * I + k] = o_cov2d[k];
  }
  if (v_cov3d) {
#pragma unroll
    for (int k = 0; k < 6; ++k) v_cov3d[6 * I + k] = o_cov3d[k];
  }
}

__global__ void __launch_bounds__(PROJ_THREADS)
cov2d_bounds_kernel(int n, const float *__restrict__ covs2d, float *__restrict__ conics,
                    float *__restrict__ radii) {
  const int idx = blockIdx.x * PROJ_THREADS + threadIdx.x;
  if (idx >= n) return;
  float ca = 0.f, cb = 0.f, cc = 0.f, r = 0.f;
  cov2d_to_conic_radius(covs2d[3 * (size_t)idx], covs2d[3 * (size_t)idx + 1], covs2d[3 * (size_t)idx + 2], ca, cb, cc, r);
  conics[3 * (size_t)idx] = ca;
  conics[3 * (size_t)idx + 1] = cb;
  conics[3 * (size_t)idx + 2] = cc;
  radii[idx] = r;
}

}  // namespace gsr

extern "C" {

GSR_API int gsr_project_gaussians_forward(int num_points, const float *means3d, const float *scales,
                                          float glob_scale, const float *quats, const float *viewmat,
                                          const float *projmat, float fx, float fy, float cx, float cy,
                                          unsigned img_height, unsigned img_width, unsigned block_width,
                                          float clip_thresh, float *cov3d, float *xys, float *depths,
                                          int32_t *radii, float *conics, float *compensation,
                                          int32_t *num_tiles_hit, void *stream) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_project_gaussians_forward");
  GSR_REQUIRE(num_points >= 0, GSR_ERR_INVALID_ARGUMENT, "project_gaussians_forward: num_points < 0");
  GSR_REQUIRE(block_width > 1 && block_width <= 16, GSR_ERR_INVALID_ARGUMENT,
              "block_width must be between 2 and 16 (got %u)", block_width);
  GSR_REQUIRE(img_height > 0 && img_width > 0, GSR_ERR_INVALID_ARGUMENT, "project_gaussians_forward: empty image");
  if (num_points == 0) return GSR_OK;
  GSR_REQUIRE(means3d && scales && quats && viewmat && projmat && cov3d && xys && depths && radii && conics &&
                  compensation && num_tiles_hit,
              GSR_ERR_INVALID_ARGUMENT, "project_gaussians_forward: null pointer");
  GSR_REQUIRE((uintptr_t)xys % 8 == 0, GSR_ERR_INVALID_ARGUMENT, "project_gaussians_forward: xys must be 8-byte aligned");
  const int tiles_x = cdiv(img_width, block_width), tiles_y = cdiv(img_height, block_width);
  project_forward_kernel<<<cdiv(num_points, PROJ_THREADS), PROJ_THREADS, 0, (cudaStream_t)stream>>>(
      num_points, means3d, scales, glob_scale, quats, viewmat, projmat, fx, fy, cx, cy, tan_half_fov(img_width, fx),
      tan_half_fov(img_height, fy), (int)img_width, (int)img_height, tiles_x, tiles_y, (int)block_width, clip_thresh, cov3d,
      xys, depths, radii, conics, compensation, num_tiles_hit, (uintptr_t)quats % 16 == 0 ? 1 : 0);
  GSR_CHECK_LAUNCH("project_forward_kernel");
  return GSR_OK;
}

GSR_API int gsr_project_gaussians_backward(int num_points, const float *means3d, const float *scales,
                                           float glob_scale, const float *quats, const float *viewmat,
                                           const float *projmat, float fx, float fy, float cx, float cy,
                                           unsigned img_height, unsigned img_width, const float *cov3d,
                                           const int32_t *radii, const float *conics,
                                           const float *compensation, const float *v_xy,
                                           const float *v_depth, const float *v_conic,
                                           const float *v_compensation, float *v_cov2d, float *v_cov3d,
                                           float *v_mean3d, float *v_scale, float *v_quat, void *stream) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_project_gaussians_backward");
  (void)cx;
  (void)cy;
  GSR_REQUIRE(num_points >= 0, GSR_ERR_INVALID_ARGUMENT, "project_gaussians_backward: num_points < 0");
  if (num_points == 0) return GSR_OK;
  GSR_REQUIRE(means3d && scales && quats && viewmat && projmat && cov3d && radii && conics && compensation &&
                  v_xy && v_depth && v_conic && v_compensation && v_mean3d && v_scale && v_quat,
              GSR_ERR_INVALID_ARGUMENT, "project_gaussians_backward: null pointer");
  GSR_REQUIRE((uintptr_t)v_xy % 8 == 0 && (uintptr_t)v_quat % 16 == 0, GSR_ERR_INVALID_ARGUMENT,
              "project_gaussians_backward: v_xy must be 8-byte and v_quat 16-byte aligned");
  project_backward_kernel<<<cdiv(num_points, PROJ_THREADS), PROJ_THREADS, 0, (cudaStream_t)stream>>>(
      num_points, means3d, scales, glob_scale, quats, viewmat, projmat, fx, fy, (int)img_width, (int)img_height,
      cov3d, radii, conics, compensation, v_xy, v_depth, v_conic, v_compensation, v_cov2d, v_cov3d, v_mean3d,
      v_scale, v_quat);
  GSR_CHECK_LAUNCH("project_backward_kernel");
  return GSR_OK;
}

GSR_API int gsr_compute_cov2d_bounds(int num_pts, const float *covs2d, float *conics, float *radii, void *stream) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_compute_cov2d_bounds");
  GSR_REQUIRE(num_pts >= 0, GSR_ERR_INVALID_ARGUMENT, "compute_cov2d_bounds: num_pts < 0");
  if (num_pts == 0) return GSR_OK;
  GSR_REQUIRE(covs2d && conics && radii, GSR_ERR_INVALID_ARGUMENT, "compute_cov2d_bounds: null pointer");
  cov2d_bounds_kernel<<<cdiv(num_pts, PROJ_THREADS), PROJ_THREADS, 0, (cudaStream_t)stream>>>(num_pts, covs2d,
                                                                                               conics, radii);
  GSR_CHECK_LAUNCH("cov2d_bounds_kernel");
  return GSR_OK;
}
}
