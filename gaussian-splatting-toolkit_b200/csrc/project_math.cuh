// project_math.cuh — the per-Gaussian EWA projection and its adjoint as device functions on scalars, for kernels
// that fuse the projection with other per-Gaussian work (fused.cu).  Same arithmetic, in the same order, as
// project_forward_kernel / project_backward_kernel of project.cu (which restate reference csrc/forward.cu:13-90,
// :398-464, backward.cu:305-453, helpers.cuh).
#pragma once
#include "common.cuh"

namespace gsr {

struct ProjFwd {
  float cov3d[6];
  float conic[3];
  float x, y, depth, comp;
  int radius, tiles;
};

// s0,s1,s2 = glob_scale * (activated) scales; V = viewmat (12 floats), PM = projmat (16 floats)
GSR_HD ProjFwd project_one(float px, float py, float pz, float s0, float s1, float s2, float qw,
                                               float qx, float qy, float qz, const float *V, const float *PM, float fx,
                                               float fy, float cx, float cy, float tan_fovx, float tan_fovy, int img_w,
                                               int img_h, int tiles_x, int tiles_y, int block_width,
                                               float clip_thresh) {
  float o_cov3d[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float o_conic[3] = {0.f, 0.f, 0.f};
  float o_x = 0.f, o_y = 0.f, o_depth = 0.f, o_comp = 0.f;
  int o_radius = 0, o_tiles = 0;

  // clip_near_plane (helpers.cuh:210-219)
  const float vx = V[0] * px + V[1] * py + V[2] * pz + V[3];
  const float vy = V[4] * px + V[5] * py + V[6] * pz + V[7];
  const float vz = V[8] * px + V[9] * py + V[10] * pz + V[11];
  do {
    if (vz <= clip_thresh) break;

    // scale_rot_to_cov3d (forward.cu:445-464): M = R * S, Sigma = M * M^T
    float R[9];
    quat_to_rotmat(qw, qx, qy, qz, R);
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

  ProjFwd o;
#pragma unroll
  for (int k = 0; k < 6; ++k) o.cov3d[k] = o_cov3d[k];
#pragma unroll
  for (int k = 0; k < 3; ++k) o.conic[k] = o_conic[k];
  o.x = o_x; o.y = o_y; o.depth = o_depth; o.comp = o_comp; o.radius = o_radius; o.tiles = o_tiles;
  return o;
}

struct ProjBwd {
  float mean[3], scale[3], quat[4];
};

// sc0..2 = (activated) scales before glob_scale; c3 = cov3d[6]; (ca,cb,cc) = conic; outputs are zero when !visible
GSR_HD ProjBwd project_one_vjp(bool visible, float px, float py, float pz, float sc0, float sc1,
                                                   float sc2, float glob_scale, float qw, float qx, float qy, float qz,
                                                   const float *V, const float *PM, float fx, float fy, int img_w,
                                                   int img_h, const float *c3, float ca, float cb, float cc, float comp,
                                                   float v_x, float v_y, float v_depth, float v_ca, float v_cb,
                                                   float v_cc, float v_comp) {
  float o_mean[3] = {0.f, 0.f, 0.f}, o_scale[3] = {0.f, 0.f, 0.f}, o_quat[4] = {0.f, 0.f, 0.f, 0.f};
  float o_cov2d[3] = {0.f, 0.f, 0.f}, o_cov3d[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

  if (visible) {
    // project_pix_vjp (helpers.cuh:125-142)
    const float hx = PM[0] * px + PM[1] * py + PM[2] * pz + PM[3];
    const float hy = PM[4] * px + PM[5] * py + PM[6] * pz + PM[7];
    const float hw = PM[12] * px + PM[13] * py + PM[14] * pz + PM[15];
    const float rw = 1.f / (hw + 1e-6f);
    const float2 vxy = make_float2(v_x, v_y);
    const float gx = 0.5f * (float)img_w * vxy.x, gy = 0.5f * (float)img_h * vxy.y;
    const float vt0 = gx * rw, vt1 = gy * rw, vt3 = -(gx * hx + gy * hy) * rw * rw;
#pragma unroll
    for (int k = 0; k < 3; ++k) o_mean[k] = PM[k] * vt0 + PM[4 + k] * vt1 + PM[12 + k] * vt3;
    // depth (backward.cu:325-331)
    const float vz = v_depth;
    o_mean[0] += V[8] * vz;
    o_mean[1] += V[9] * vz;
    o_mean[2] += V[10] * vz;

    // cov2d_to_conic_vjp (helpers.cuh:62-74): Y = -X G X
    const float g00 = v_ca, g01 = v_cb / 2.f, g11 = v_cc;
    const float xg00 = ca * g00 + cb * g01, xg01 = ca * g01 + cb * g11;
    const float xg10 = cb * g00 + cc * g01, xg11 = cb * g01 + cc * g11;
    const float y00 = -(xg00 * ca + xg01 * cb), y01 = -(xg00 * cb + xg01 * cc);
    const float y10 = -(xg10 * ca + xg11 * cb), y11 = -(xg10 * cb + xg11 * cc);
    o_cov2d[0] = y00;
    o_cov2d[1] = y01 + y10;
    o_cov2d[2] = y11;
    // cov2d_to_compensation_vjp (helpers.cuh:76-90)
    {
      const float vcomp = v_comp;
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
    float R[9];
    quat_to_rotmat(qw, qx, qy, qz, R);
    const float S[3] = {glob_scale * sc0, glob_scale * sc1, glob_scale * sc2};
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

  ProjBwd g;
#pragma unroll
  for (int k = 0; k < 3; ++k) { g.mean[k] = o_mean[k]; g.scale[k] = o_scale[k]; }
#pragma unroll
  for (int k = 0; k < 4; ++k) g.quat[k] = o_quat[k];
  return g;
}

}  // namespace gsr
