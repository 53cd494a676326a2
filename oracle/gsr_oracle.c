/*
 * gsr_oracle.c — TEST INFRASTRUCTURE.  CPU restatement (plain C, binary32 arithmetic, OpenMP over
 * Gaussians / tiles) of the reference rasterizer hot path.  It is the checker for libgsr_b200.so and
 * the "port" CPU baseline of bench.py; it is never linked into, imported by, or called from the product
 * path (gaussian-splatting-toolkit_b200/).
 *
 * Every function cites the reference source it restates.  Paths are relative to
 * /root/reference/gs_toolkit/gs_components/rasterizer/cuda/csrc/ (CSRC) or .../rasterizer/ (RAST).
 * The arithmetic follows the CUDA kernels (NOT RAST/_torch_impl.py, which has known defects),
 * including the analytic-backward quirks of the CUDA code.
 *
 * Parity status: pinned against (i) RAST/_torch_impl.py forward outputs generated in the build
 * container (tests/golden/torch_impl_*.npz, script tests/golden/gen_golden_torch_impl.py) and
 * (ii) outputs of the compiled, unmodified reference CUDA extension run on a B200
 * (tests/golden/refcuda_*.npz, script tests/golden/gen_golden_ref_cuda.py).
 *
 * Deliberate differences from the CUDA reference (documented, do not affect the 1e-4 parity bar):
 *   - expf() instead of __expf(); IEEE division / sqrt instead of --use_fast_math approximations;
 *   - per-Gaussian gradient sums of the blend adjoint are accumulated in binary64 (the reference uses
 *     order-nondeterministic binary32 atomics), so the oracle is the low-noise value both
 *     implementations are compared with;
 *   - culled Gaussians have every projection output zero-filled except cov3d / conics, which are
 *     written exactly where the reference kernel writes them before its early returns.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------------
 * Spherical harmonics — CSRC/sh.cuh:6-31 (constants, num_sh_bases), :33-98 (forward), :100-186 (vjp)
 * ---------------------------------------------------------------------------------------------- */
static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f,  -0.4570457994644658f,
                               0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};
static const float SH_C4[9] = {2.5033429417967046f,  -1.7701307697799304f, 0.9461746957575601f,
                               -0.6690465435572892f, 0.10578554691520431f, -0.6690465435572892f,
                               0.47308734787878004f, -1.7701307697799304f, 0.6258357354491761f};

ORC_API int orc_num_sh_bases(int degree) { /* sh.cuh:21-31 */
  if (degree == 0) return 1;
  if (degree == 1) return 4;
  if (degree == 2) return 9;
  if (degree == 3) return 16;
  return 25;
}

/* basis values Y_k(dir) for k < (deg+1)^2; dir normalised here as sh.cuh:44-48 does */
static void sh_basis(int deg, const float *d, float *Y) {
  Y[0] = SH_C0;
  if (deg < 1) return;
  float norm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  float x = d[0] / norm, y = d[1] / norm, z = d[2] / norm;
  Y[1] = -SH_C1 * y;
  Y[2] = SH_C1 * z;
  Y[3] = -SH_C1 * x;
  if (deg < 2) return;
  float xx = x * x, xy = x * y, xz = x * z, yy = y * y, yz = y * z, zz = z * z;
  Y[4] = SH_C2[0] * xy;
  Y[5] = SH_C2[1] * yz;
  Y[6] = SH_C2[2] * (2.f * zz - xx - yy);
  Y[7] = SH_C2[3] * xz;
  Y[8] = SH_C2[4] * (xx - yy);
  if (deg < 3) return;
  Y[9] = SH_C3[0] * y * (3.f * xx - yy);
  Y[10] = SH_C3[1] * xy * z;
  Y[11] = SH_C3[2] * y * (4.f * zz - xx - yy);
  Y[12] = SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
  Y[13] = SH_C3[4] * x * (4.f * zz - xx - yy);
  Y[14] = SH_C3[5] * z * (xx - yy);
  Y[15] = SH_C3[6] * x * (xx - 3.f * yy);
  if (deg < 4) return;
  Y[16] = SH_C4[0] * xy * (xx - yy);
  Y[17] = SH_C4[1] * yz * (3.f * xx - yy);
  Y[18] = SH_C4[2] * xy * (7.f * zz - 1.f);
  Y[19] = SH_C4[3] * yz * (7.f * zz - 3.f);
  Y[20] = SH_C4[4] * (zz * (35.f * zz - 30.f) + 3.f);
  Y[21] = SH_C4[5] * xz * (7.f * zz - 3.f);
  Y[22] = SH_C4[6] * (xx - yy) * (7.f * zz - 1.f);
  Y[23] = SH_C4[7] * xz * (xx - 3.f * yy);
  Y[24] = SH_C4[8] * (xx * (xx - 3.f * yy) - yy * (3.f * xx - yy));
}

/* sh.cuh:188-205 + :33-98 ; coeffs [N,K,3] with K=(degree+1)^2, colors [N,3] */
ORC_API void orc_sh_forward(int n, int degree, int degrees_to_use, const float *viewdirs,
                            const float *coeffs, float *colors) {
  int K = orc_num_sh_bases(degree), Ku = orc_num_sh_bases(degrees_to_use);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) {
    float Y[25];
    sh_basis(degrees_to_use, viewdirs + 3 * (size_t)i, Y);
    const float *c = coeffs + (size_t)i * K * 3;
    for (int ch = 0; ch < 3; ++ch) {
      /* same association as sh.cuh: band sums are added band by band */
      float acc = SH_C0 * c[ch];
      for (int band = 1; band <= degrees_to_use; ++band) {
        float s = 0.f;
        for (int k = band * band; k < (band + 1) * (band + 1); ++k) s += Y[k] * c[3 * k + ch];
        acc += s;
      }
      (void)Ku;
      colors[3 * (size_t)i + ch] = acc;
    }
  }
}

/* sh.cuh:207-224 + :100-186 ; v_coeffs [N,K,3] fully written (zeros above degrees_to_use, as the
 * reference's torch::zeros does, bindings.cu:95-96) */
ORC_API void orc_sh_backward(int n, int degree, int degrees_to_use, const float *viewdirs,
                             const float *v_colors, float *v_coeffs) {
  int K = orc_num_sh_bases(degree), Ku = orc_num_sh_bases(degrees_to_use);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) {
    float Y[25];
    sh_basis(degrees_to_use, viewdirs + 3 * (size_t)i, Y);
    float *v = v_coeffs + (size_t)i * K * 3;
    for (int k = 0; k < K; ++k)
      for (int ch = 0; ch < 3; ++ch)
        v[3 * k + ch] = (k < Ku) ? Y[k] * v_colors[3 * (size_t)i + ch] : 0.f;
  }
}

/* ------------------------------------------------------------------------------------------------
 * small 3x3 helpers (row-major), replacing the vendored glm of the reference
 * ---------------------------------------------------------------------------------------------- */
static void quat_to_rotmat(const float *q /*wxyz*/, float R[9]) { /* helpers.cuh:144-159 */
  float s = 1.f / sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  float w = q[0] * s, x = q[1] * s, y = q[2] * s, z = q[3] * s;
  R[0] = 1.f - 2.f * (y * y + z * z);
  R[1] = 2.f * (x * y - w * z);
  R[2] = 2.f * (x * z + w * y);
  R[3] = 2.f * (x * y + w * z);
  R[4] = 1.f - 2.f * (x * x + z * z);
  R[5] = 2.f * (y * z - w * x);
  R[6] = 2.f * (x * z - w * y);
  R[7] = 2.f * (y * z + w * x);
  R[8] = 1.f - 2.f * (x * x + y * y);
}

static void mat3_mul(const float *A, const float *B, float *C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
static void mat3_T(const float *A, float *B) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) B[3 * j + i] = A[3 * i + j];
}

/* helpers.cuh:11-34 get_bbox / get_tile_bbox : (int) truncates toward zero, clamp to [0, tiles] */
static void tile_bbox(float cx, float cy, float radius, int tiles_x, int tiles_y, int bw, int *x0,
                      int *y0, int *x1, int *y1) {
  float tcx = cx / (float)bw, tcy = cy / (float)bw, tr = radius / (float)bw;
  int a;
  a = (int)(tcx - tr);      *x0 = a < 0 ? 0 : (a > tiles_x ? tiles_x : a);
  a = (int)(tcx + tr + 1);  *x1 = a < 0 ? 0 : (a > tiles_x ? tiles_x : a);
  a = (int)(tcy - tr);      *y0 = a < 0 ? 0 : (a > tiles_y ? tiles_y : a);
  a = (int)(tcy + tr + 1);  *y1 = a < 0 ? 0 : (a > tiles_y ? tiles_y : a);
}

/* helpers.cuh:36-59 */
static int cov2d_bounds(const float cov2d[3], float conic[3], float *radius) {
  float det = cov2d[0] * cov2d[2] - cov2d[1] * cov2d[1];
  if (det == 0.f) return 0;
  float inv_det = 1.f / det;
  conic[0] = cov2d[2] * inv_det;
  conic[1] = -cov2d[1] * inv_det;
  conic[2] = cov2d[0] * inv_det;
  float b = 0.5f * (cov2d[0] + cov2d[2]);
  float v1 = b + sqrtf(fmaxf(0.1f, b * b - det));
  float v2 = b - sqrtf(fmaxf(0.1f, b * b - det));
  *radius = ceilf(3.f * sqrtf(fmaxf(v1, v2)));
  return 1;
}

/* bindings.cu:19-56 compute_cov2d_bounds_kernel: conics/radii outputs are torch::zeros, written even
 * when det==0 with whatever compute_cov2d_bounds left (uninitialised in the reference); here 0. */
ORC_API void orc_compute_cov2d_bounds(int n, const float *covs2d, float *conics, float *radii) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) {
    float c[3] = {0.f, 0.f, 0.f}, r = 0.f;
    cov2d_bounds(covs2d + 3 * (size_t)i, c, &r);
    conics[3 * (size_t)i] = c[0];
    conics[3 * (size_t)i + 1] = c[1];
    conics[3 * (size_t)i + 2] = c[2];
    radii[i] = r;
  }
}

/* ------------------------------------------------------------------------------------------------
 * project_gaussians_forward_kernel — CSRC/forward.cu:13-90 with device fns :398-464 and
 * helpers.cuh:7-9,93-122,202-219.  viewmat: first 12 floats (row-major 3x4), projmat 4x4 row-major.
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_project_forward(int n, const float *means3d, const float *scales, float glob_scale,
                                 const float *quats, const float *viewmat, const float *projmat,
                                 float fx, float fy, float cx, float cy, int img_height, int img_width,
                                 int block_width, float clip_thresh, float *cov3d, float *xys,
                                 float *depths, int32_t *radii, float *conics, float *compensation,
                                 int32_t *num_tiles_hit) {
  const int tiles_x = (img_width + block_width - 1) / block_width;
  const int tiles_y = (img_height + block_width - 1) / block_width;
  const float *V = viewmat;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) {
    size_t I = (size_t)i;
    radii[i] = 0;
    num_tiles_hit[i] = 0;
    depths[i] = 0.f;
    compensation[i] = 0.f;
    xys[2 * I] = xys[2 * I + 1] = 0.f;
    for (int k = 0; k < 6; ++k) cov3d[6 * I + k] = 0.f;
    for (int k = 0; k < 3; ++k) conics[3 * I + k] = 0.f;

    const float *p = means3d + 3 * I;
    /* clip_near_plane, helpers.cuh:210-219 (<=) */
    float pv[3] = {V[0] * p[0] + V[1] * p[1] + V[2] * p[2] + V[3],
                   V[4] * p[0] + V[5] * p[1] + V[6] * p[2] + V[7],
                   V[8] * p[0] + V[9] * p[1] + V[10] * p[2] + V[11]};
    if (pv[2] <= clip_thresh) continue;

    /* scale_rot_to_cov3d, forward.cu:445-464 : M = R*S, Sigma = M*M^T */
    float R[9], M[9], Mt[9], S3[9];
    quat_to_rotmat(quats + 4 * I, R);
    const float *sc = scales + 3 * I;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) M[3 * r + c] = R[3 * r + c] * (glob_scale * sc[c]);
    mat3_T(M, Mt);
    mat3_mul(M, Mt, S3);
    float *cv = cov3d + 6 * I;
    cv[0] = S3[0]; cv[1] = S3[1]; cv[2] = S3[2]; cv[3] = S3[4]; cv[4] = S3[5]; cv[5] = S3[8];

    /* project_cov3d_ewa, forward.cu:398-442 */
    float tan_fovx = 0.5f * (float)img_width / fx, tan_fovy = 0.5f * (float)img_height / fy;
    float t[3] = {pv[0], pv[1], pv[2]};
    float lim_x = 1.3f * tan_fovx, lim_y = 1.3f * tan_fovy;
    t[0] = t[2] * fminf(lim_x, fmaxf(-lim_x, t[0] / t[2]));
    t[1] = t[2] * fminf(lim_y, fmaxf(-lim_y, t[1] / t[2]));
    float rz = 1.f / t[2], rz2 = rz * rz;
    float J[9] = {fx * rz, 0.f, -fx * t[0] * rz2, 0.f, fy * rz, -fy * t[1] * rz2, 0.f, 0.f, 0.f};
    float W[9] = {V[0], V[1], V[2], V[4], V[5], V[6], V[8], V[9], V[10]};
    float T[9], Tt[9], Vm[9] = {cv[0], cv[1], cv[2], cv[1], cv[3], cv[4], cv[2], cv[4], cv[5]};
    float TV[9], C[9];
    mat3_mul(J, W, T);
    mat3_T(T, Tt);
    mat3_mul(T, Vm, TV);
    mat3_mul(TV, Tt, C);
    float c00 = C[0], c11 = C[4], c01 = C[1];
    float det_orig = c00 * c11 - c01 * c01;
    float cov2d[3] = {c00 + 0.3f, c01, c11 + 0.3f};
    float det_blur = cov2d[0] * cov2d[2] - cov2d[1] * cov2d[1];
    float comp = sqrtf(fmaxf(0.f, det_orig / det_blur));

    float conic[3], radius;
    if (!cov2d_bounds(cov2d, conic, &radius)) continue;
    conics[3 * I] = conic[0]; conics[3 * I + 1] = conic[1]; conics[3 * I + 2] = conic[2];

    /* project_pix, helpers.cuh:114-122 */
    const float *PM = projmat;
    float hx = PM[0] * p[0] + PM[1] * p[1] + PM[2] * p[2] + PM[3];
    float hy = PM[4] * p[0] + PM[5] * p[1] + PM[6] * p[2] + PM[7];
    float hw = PM[12] * p[0] + PM[13] * p[1] + PM[14] * p[2] + PM[15];
    float rw = 1.f / (hw + 1e-6f);
    float px = 0.5f * (float)img_width * (hx * rw) + cx - 0.5f;
    float py = 0.5f * (float)img_height * (hy * rw) + cy - 0.5f;

    int x0, y0, x1, y1;
    tile_bbox(px, py, radius, tiles_x, tiles_y, block_width, &x0, &y0, &x1, &y1);
    int area = (x1 - x0) * (y1 - y0);
    if (area <= 0) continue;
    num_tiles_hit[i] = area;
    depths[i] = pv[2];
    radii[i] = (int)radius;
    xys[2 * I] = px; xys[2 * I + 1] = py;
    compensation[i] = comp;
  }
}

/* ------------------------------------------------------------------------------------------------
 * project_gaussians_backward_kernel — CSRC/backward.cu:305-347, VJPs backward.cu:350-453,
 * helpers.cuh:62-90 (conic / compensation vjp), :125-142 (project_pix_vjp), :161-200 (quat vjp).
 * Outputs zero where radii <= 0 (reference: torch::zeros + early return).
 * Quirks kept: unclamped t in the EWA vjp (backward.cu:367-369); compensation vjp divides by
 * (comp + 1e-6) (helpers.cuh:86); rw = 1/(w + 1e-6).
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_project_backward(int n, const float *means3d, const float *scales, float glob_scale,
                                  const float *quats, const float *viewmat, const float *projmat,
                                  float fx, float fy, float cx, float cy, int img_height,
                                  int img_width, const float *cov3d, const int32_t *radii,
                                  const float *conics, const float *compensation, const float *v_xy,
                                  const float *v_depth, const float *v_conic,
                                  const float *v_compensation, float *v_cov2d, float *v_cov3d,
                                  float *v_mean3d, float *v_scale, float *v_quat) {
  (void)cx; (void)cy;
  const float *V = viewmat, *PM = projmat;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) {
    size_t I = (size_t)i;
    for (int k = 0; k < 3; ++k) v_cov2d[3 * I + k] = v_mean3d[3 * I + k] = v_scale[3 * I + k] = 0.f;
    for (int k = 0; k < 6; ++k) v_cov3d[6 * I + k] = 0.f;
    for (int k = 0; k < 4; ++k) v_quat[4 * I + k] = 0.f;
    if (radii[i] <= 0) continue;
    const float *p = means3d + 3 * I;

    /* project_pix_vjp */
    float hx = PM[0] * p[0] + PM[1] * p[1] + PM[2] * p[2] + PM[3];
    float hy = PM[4] * p[0] + PM[5] * p[1] + PM[6] * p[2] + PM[7];
    float hw = PM[12] * p[0] + PM[13] * p[1] + PM[14] * p[2] + PM[15];
    float rw = 1.f / (hw + 1e-6f);
    float gx = 0.5f * (float)img_width * v_xy[2 * I], gy = 0.5f * (float)img_height * v_xy[2 * I + 1];
    float vt[4] = {gx * rw, gy * rw, 0.f, -(gx * hx + gy * hy) * rw * rw};
    float vm[3];
    for (int k = 0; k < 3; ++k)
      vm[k] = PM[k] * vt[0] + PM[4 + k] * vt[1] + PM[8 + k] * vt[2] + PM[12 + k] * vt[3];
    /* depth */
    float vz = v_depth[i];
    vm[0] += V[8] * vz; vm[1] += V[9] * vz; vm[2] += V[10] * vz;

    /* cov2d_to_conic_vjp: Y = -X G X */
    const float *co = conics + 3 * I, *vc = v_conic + 3 * I;
    float X[4] = {co[0], co[1], co[1], co[2]};
    float G[4] = {vc[0], vc[1] / 2.f, vc[1] / 2.f, vc[2]};
    float XG[4] = {X[0] * G[0] + X[1] * G[2], X[0] * G[1] + X[1] * G[3], X[2] * G[0] + X[3] * G[2],
                   X[2] * G[1] + X[3] * G[3]};
    float Y[4] = {-(XG[0] * X[0] + XG[1] * X[2]), -(XG[0] * X[1] + XG[1] * X[3]),
                  -(XG[2] * X[0] + XG[3] * X[2]), -(XG[2] * X[1] + XG[3] * X[3])};
    float vcov2d[3] = {Y[0], Y[1] + Y[2], Y[3]};
    /* cov2d_to_compensation_vjp */
    {
      float comp = compensation[i], vcomp = v_compensation[i];
      float inv_det = co[0] * co[2] - co[1] * co[1];
      float om = 1.f - comp * comp;
      float u = vcomp * 0.5f / (comp + 1e-6f);
      vcov2d[0] += u * (om * co[0] - 0.3f * inv_det);
      vcov2d[1] += 2.f * u * (om * co[1]);
      vcov2d[2] += u * (om * co[2] - 0.3f * inv_det);
    }
    v_cov2d[3 * I] = vcov2d[0]; v_cov2d[3 * I + 1] = vcov2d[1]; v_cov2d[3 * I + 2] = vcov2d[2];

    /* project_cov3d_ewa_vjp (unclamped t) */
    float W[9] = {V[0], V[1], V[2], V[4], V[5], V[6], V[8], V[9], V[10]};
    float t[3] = {V[0] * p[0] + V[1] * p[1] + V[2] * p[2] + V[3],
                  V[4] * p[0] + V[5] * p[1] + V[6] * p[2] + V[7],
                  V[8] * p[0] + V[9] * p[1] + V[10] * p[2] + V[11]};
    float rz = 1.f / t[2], rz2 = rz * rz, rz3 = rz2 * rz;
    float J[9] = {fx * rz, 0.f, -fx * t[0] * rz2, 0.f, fy * rz, -fy * t[1] * rz2, 0.f, 0.f, 0.f};
    const float *cv = cov3d + 6 * I;
    float Vm[9] = {cv[0], cv[1], cv[2], cv[1], cv[3], cv[4], cv[2], cv[4], cv[5]};
    float Gc[9] = {vcov2d[0], 0.5f * vcov2d[1], 0.f, 0.5f * vcov2d[1], vcov2d[2], 0.f, 0.f, 0.f, 0.f};
    float T[9], Tt[9], tmp[9], vV[9], vT[9], tmp2[9];
    mat3_mul(J, W, T);
    mat3_T(T, Tt);
    mat3_mul(Tt, Gc, tmp);
    mat3_mul(tmp, T, vV); /* v_V = T^T G T */
    float *v3 = v_cov3d + 6 * I;
    v3[0] = vV[0]; v3[1] = vV[1] + vV[3]; v3[2] = vV[2] + vV[6];
    v3[3] = vV[4]; v3[4] = vV[5] + vV[7]; v3[5] = vV[8];
    /* v_T = G T V^T + G^T T V  (G, V symmetric) */
    mat3_mul(Gc, T, tmp);
    mat3_mul(tmp, Vm, vT);
    for (int k = 0; k < 9; ++k) vT[k] *= 2.f;
    /* v_J = v_T W^T */
    float Wt[9], vJ[9];
    mat3_T(W, Wt);
    mat3_mul(vT, Wt, vJ);
    (void)tmp2;
    float v_t[3] = {-fx * rz2 * vJ[2], -fy * rz2 * vJ[5],
                    -fx * rz2 * vJ[0] + 2.f * fx * t[0] * rz3 * vJ[2] - fy * rz2 * vJ[4] +
                        2.f * fy * t[1] * rz3 * vJ[5]};
    /* v_mean += W^T v_t */
    vm[0] += W[0] * v_t[0] + W[3] * v_t[1] + W[6] * v_t[2];
    vm[1] += W[1] * v_t[0] + W[4] * v_t[1] + W[7] * v_t[2];
    vm[2] += W[2] * v_t[0] + W[5] * v_t[1] + W[8] * v_t[2];
    v_mean3d[3 * I] = vm[0]; v_mean3d[3 * I + 1] = vm[1]; v_mean3d[3 * I + 2] = vm[2];

    /* scale_rot_to_cov3d_vjp, backward.cu:425-453 */
    float vVs[9] = {v3[0], 0.5f * v3[1], 0.5f * v3[2], 0.5f * v3[1], v3[3], 0.5f * v3[4],
                    0.5f * v3[2], 0.5f * v3[4], v3[5]};
    float R[9], M[9], vM[9];
    quat_to_rotmat(quats + 4 * I, R);
    const float *sc = scales + 3 * I;
    float S[3] = {glob_scale * sc[0], glob_scale * sc[1], glob_scale * sc[2]};
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) M[3 * r + c] = R[3 * r + c] * S[c];
    mat3_mul(vVs, M, vM);
    for (int k = 0; k < 9; ++k) vM[k] *= 2.f;
    for (int c = 0; c < 3; ++c)
      v_scale[3 * I + c] = (R[c] * vM[c] + R[3 + c] * vM[3 + c] + R[6 + c] * vM[6 + c]) * glob_scale;
    float D[9]; /* v_R = v_M * S ; D[3*i+j] = v_R(row i, col j) */
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) D[3 * r + c] = vM[3 * r + c] * S[c];
    const float *q = quats + 4 * I;
    float s = 1.f / sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    float w = q[0] * s, x = q[1] * s, y = q[2] * s, z = q[3] * s;
#define Dm(i, j) D[3 * (i) + (j)]
    v_quat[4 * I + 0] =
        2.f * (x * (Dm(2, 1) - Dm(1, 2)) + y * (Dm(0, 2) - Dm(2, 0)) + z * (Dm(1, 0) - Dm(0, 1)));
    v_quat[4 * I + 1] = 2.f * (-2.f * x * (Dm(1, 1) + Dm(2, 2)) + y * (Dm(1, 0) + Dm(0, 1)) +
                               z * (Dm(2, 0) + Dm(0, 2)) + w * (Dm(2, 1) - Dm(1, 2)));
    v_quat[4 * I + 2] = 2.f * (x * (Dm(1, 0) + Dm(0, 1)) - 2.f * y * (Dm(0, 0) + Dm(2, 2)) +
                               z * (Dm(2, 1) + Dm(1, 2)) + w * (Dm(0, 2) - Dm(2, 0)));
    v_quat[4 * I + 3] = 2.f * (x * (Dm(2, 0) + Dm(0, 2)) + y * (Dm(2, 1) + Dm(1, 2)) -
                               2.f * z * (Dm(0, 0) + Dm(1, 1)) + w * (Dm(1, 0) - Dm(0, 1)));
#undef Dm
  }
}

/* ------------------------------------------------------------------------------------------------
 * Binning — RAST/utils.py:106-125 (cumsum), CSRC/forward.cu:94-127 (key emission),
 * RAST/utils.py:179-180 (sort + gather; CUB radix sort is stable), CSRC/forward.cu:132-154 (bin edges)
 * ---------------------------------------------------------------------------------------------- */
ORC_API int64_t orc_cumsum(int n, const int32_t *num_tiles_hit, int32_t *cum) {
  int32_t acc = 0; /* int32 like torch.cumsum(dtype=int32) */
  for (int i = 0; i < n; ++i) {
    acc += num_tiles_hit[i];
    cum[i] = acc;
  }
  return n > 0 ? (int64_t)cum[n - 1] : 0;
}

ORC_API void orc_map_intersects(int n, const float *xys, const float *depths, const int32_t *radii,
                                const int32_t *cum_tiles_hit, int tiles_x, int tiles_y,
                                int block_width, int64_t *isect_ids, int32_t *gaussian_ids) {
#pragma omp parallel for schedule(dynamic, 1024)
  for (int i = 0; i < n; ++i) {
    if (radii[i] <= 0) continue;
    int x0, y0, x1, y1;
    tile_bbox(xys[2 * (size_t)i], xys[2 * (size_t)i + 1], (float)radii[i], tiles_x, tiles_y,
              block_width, &x0, &y0, &x1, &y1);
    int32_t cur = (i == 0) ? 0 : cum_tiles_hit[i - 1];
    int32_t bits;
    memcpy(&bits, depths + i, 4);
    int64_t depth_id = (int64_t)bits; /* sign-extending, as (int64_t)*(int32_t*)&depth */
    for (int r = y0; r < y1; ++r)
      for (int c = x0; c < x1; ++c) {
        int64_t tile_id = (int64_t)(r * tiles_x + c);
        isect_ids[cur] = (tile_id << 32) | depth_id;
        gaussian_ids[cur] = i;
        ++cur;
      }
  }
}

/* stable LSD radix sort on the 64-bit keys viewed as signed (torch.sort on int64): flip the sign bit */
ORC_API void orc_sort_intersects(int64_t m, const int64_t *keys_in, const int32_t *vals_in,
                                 int64_t *keys_out, int32_t *vals_out) {
  if (m <= 0) return;
  uint64_t *ka = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)m);
  uint64_t *kb = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)m);
  int32_t *va = (int32_t *)malloc(sizeof(int32_t) * (size_t)m);
  int32_t *vb = (int32_t *)malloc(sizeof(int32_t) * (size_t)m);
  for (int64_t i = 0; i < m; ++i) {
    ka[i] = (uint64_t)keys_in[i] ^ 0x8000000000000000ull;
    va[i] = vals_in[i];
  }
  size_t *hist = (size_t *)malloc(sizeof(size_t) * 65537);
  for (int pass = 0; pass < 4; ++pass) {
    int shift = 16 * pass;
    memset(hist, 0, sizeof(size_t) * 65537);
    for (int64_t i = 0; i < m; ++i) hist[((ka[i] >> shift) & 0xFFFF) + 1]++;
    for (int d = 0; d < 65536; ++d) hist[d + 1] += hist[d];
    for (int64_t i = 0; i < m; ++i) {
      size_t pos = hist[(ka[i] >> shift) & 0xFFFF]++;
      kb[pos] = ka[i];
      vb[pos] = va[i];
    }
    uint64_t *tk = ka; ka = kb; kb = tk;
    int32_t *tv = va; va = vb; vb = tv;
  }
  for (int64_t i = 0; i < m; ++i) {
    keys_out[i] = (int64_t)(ka[i] ^ 0x8000000000000000ull);
    vals_out[i] = va[i];
  }
  free(ka); free(kb); free(va); free(vb); free(hist);
}

ORC_API void orc_tile_bin_edges(int64_t m, const int64_t *sorted_keys, int num_tiles,
                                int32_t *tile_bins) {
  memset(tile_bins, 0, sizeof(int32_t) * 2 * (size_t)num_tiles);
  for (int64_t idx = 0; idx < m; ++idx) {
    int32_t cur = (int32_t)(sorted_keys[idx] >> 32);
    if (idx == 0) tile_bins[2 * cur] = 0;
    if (idx == m - 1) tile_bins[2 * cur + 1] = (int32_t)m;
    if (idx == 0) continue;
    int32_t prev = (int32_t)(sorted_keys[idx - 1] >> 32);
    if (prev != cur) {
      tile_bins[2 * prev + 1] = (int32_t)idx;
      tile_bins[2 * cur] = (int32_t)idx;
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * Blend forward — CSRC/forward.cu:278-395 (3 channels, FP32) and :159-276 (N-D, binary16 accumulators)
 * half_accum != 0 selects the N-D numerics: pix_out[c] = __hadd(pix_out[c], __float2half(col*vis)).
 * ambiguous [H,W] (nullable, u8): set when a threshold decision (alpha vs 1/255, next_T vs 1e-4, sigma
 * vs 0, alpha vs the 0.999 clamp) was within rel. 1e-5 of flipping, i.e. the pixel is one on which two
 * correct FP32 implementations may legitimately differ by more than rounding noise.
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_rasterize_forward(int img_height, int img_width, int block_width, int channels,
                                   int half_accum, const int32_t *gaussian_ids_sorted,
                                   const int32_t *tile_bins, const float *xys, const float *conics,
                                   const float *colors, const float *opacities,
                                   const float *background, float *out_img, float *final_Ts,
                                   int32_t *final_idx, uint8_t *ambiguous, int tile_row_begin,
                                   int tile_row_end) {
  const int tiles_x = (img_width + block_width - 1) / block_width;
  const int tiles_y = (img_height + block_width - 1) / block_width;
  const float AMB = 1e-5f;
  /* [tile_row_begin, tile_row_end): only these tile rows are rendered (bounded CPU-baseline samples);
   * pass (0, -1) for the whole image */
  if (tile_row_end < 0 || tile_row_end > tiles_y) tile_row_end = tiles_y;
  if (tile_row_begin < 0) tile_row_begin = 0;
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
  for (int ty = tile_row_begin; ty < tile_row_end; ++ty)
    for (int tx = 0; tx < tiles_x; ++tx) {
      int tile = ty * tiles_x + tx;
      int lo = tile_bins[2 * tile], hi = tile_bins[2 * tile + 1];
      for (int i = ty * block_width; i < (ty + 1) * block_width && i < img_height; ++i)
        for (int j = tx * block_width; j < (tx + 1) * block_width && j < img_width; ++j) {
          float px = (float)j, py = (float)i;
          float T = 1.f;
          int last = 0;
          uint8_t amb = 0;
          float acc[32];
          _Float16 acch[32];
          for (int c = 0; c < channels; ++c) { acc[c] = 0.f; acch[c] = (_Float16)0.f; }
          for (int k = lo; k < hi; ++k) {
            int g = gaussian_ids_sorted[k];
            float dx = xys[2 * (size_t)g] - px, dy = xys[2 * (size_t)g + 1] - py;
            const float *co = conics + 3 * (size_t)g;
            float sigma = 0.5f * (co[0] * dx * dx + co[2] * dy * dy) + co[1] * dx * dy;
            float raw = opacities[g] * expf(-sigma);
            float alpha = fminf(0.999f, raw);
            if (fabsf(raw - 1.f / 255.f) < AMB * (1.f / 255.f) || fabsf(sigma) < 1e-7f ||
                fabsf(raw - 0.999f) < AMB)
              amb = 1;
            if (sigma < 0.f || alpha < 1.f / 255.f) continue;
            float next_T = T * (1.f - alpha);
            if (fabsf(next_T - 1e-4f) < AMB * 1e-4f * 10.f) amb = 1;
            if (next_T <= 1e-4f) break;
            float vis = alpha * T;
            if (half_accum) {
              for (int c = 0; c < channels; ++c)
                acch[c] = acch[c] + (_Float16)(colors[(size_t)channels * g + c] * vis);
            } else {
              for (int c = 0; c < channels; ++c) acc[c] = acc[c] + colors[(size_t)channels * g + c] * vis;
            }
            T = next_T;
            last = k;
          }
          size_t pix = (size_t)i * img_width + j;
          final_Ts[pix] = T;
          final_idx[pix] = last;
          if (ambiguous) ambiguous[pix] = amb;
          for (int c = 0; c < channels; ++c) {
            float v = half_accum ? (float)acch[c] : acc[c];
            out_img[pix * channels + c] = v + T * background[c];
          }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Blend backward — CSRC/backward.cu:133-303 (3 channels) and :23-131 (N-D).
 * nd_mode == 0: RGB kernel semantics — contributors are sorted indices k in [lo, final_idx] (inclusive,
 *               backward.cu:215), FP32 running sum S.
 * nd_mode != 0: N-D kernel semantics — contributors k in [lo, final_idx) (last contributor EXCLUDED,
 *               backward.cu:64-65), running sum S kept in binary16 (backward.cu:46-50,105).
 * Alpha clamp is 0.99 here (0.999 in forward) and the clamp's derivative is ignored, as in the source.
 * Per-Gaussian sums accumulate in binary64 via atomics (see header comment).
 * Outputs (binary64, zero-filled here): v_xy [N,2], v_conic [N,3], v_colors [N,C], v_opacity [N].
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_rasterize_backward(int img_height, int img_width, int block_width, int channels,
                                    int nd_mode, int num_points, const int32_t *gaussian_ids_sorted,
                                    const int32_t *tile_bins, const float *xys, const float *conics,
                                    const float *colors, const float *opacities,
                                    const float *background, const float *final_Ts,
                                    const int32_t *final_idx, const float *v_output,
                                    const float *v_output_alpha, double *v_xy, double *v_conic,
                                    double *v_colors, double *v_opacity, int tile_row_begin,
                                    int tile_row_end) {
  const int tiles_x = (img_width + block_width - 1) / block_width;
  const int tiles_y = (img_height + block_width - 1) / block_width;
  if (tile_row_end < 0 || tile_row_end > tiles_y) tile_row_end = tiles_y;
  if (tile_row_begin < 0) tile_row_begin = 0;
  memset(v_xy, 0, sizeof(double) * 2 * (size_t)num_points);
  memset(v_conic, 0, sizeof(double) * 3 * (size_t)num_points);
  memset(v_colors, 0, sizeof(double) * (size_t)channels * (size_t)num_points);
  memset(v_opacity, 0, sizeof(double) * (size_t)num_points);
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
  for (int ty = tile_row_begin; ty < tile_row_end; ++ty)
    for (int tx = 0; tx < tiles_x; ++tx) {
      int tile = ty * tiles_x + tx;
      int lo = tile_bins[2 * tile], hi = tile_bins[2 * tile + 1];
      if (hi <= lo) continue;
      for (int i = ty * block_width; i < (ty + 1) * block_width && i < img_height; ++i)
        for (int j = tx * block_width; j < (tx + 1) * block_width && j < img_width; ++j) {
          size_t pix = (size_t)i * img_width + j;
          float px = (float)j, py = (float)i;
          const float *v_out = v_output + pix * channels;
          float v_out_alpha = v_output_alpha[pix];
          float T_final = final_Ts[pix], T = T_final;
          float S[32];
          _Float16 Sh[32];
          for (int c = 0; c < channels; ++c) { S[c] = 0.f; Sh[c] = (_Float16)0.f; }
          int bin_final = final_idx[pix];
          int k_start = nd_mode ? bin_final - 1 : bin_final;
          if (k_start > hi - 1) k_start = hi - 1;
          for (int k = k_start; k >= lo; --k) {
            int g = gaussian_ids_sorted[k];
            const float *co = conics + 3 * (size_t)g;
            float dx = xys[2 * (size_t)g] - px, dy = xys[2 * (size_t)g + 1] - py;
            float sigma = 0.5f * (co[0] * dx * dx + co[2] * dy * dy) + co[1] * dx * dy;
            float opac = opacities[g];
            float vis = expf(-sigma);
            float alpha = fminf(0.99f, opac * vis);
            if (sigma < 0.f || alpha < 1.f / 255.f) continue;
            float ra = 1.f / (1.f - alpha);
            T *= ra;
            float fac = alpha * T;
            float v_alpha = 0.f;
            const float *rgb = colors + (size_t)channels * g;
            if (nd_mode) {
              for (int c = 0; c < channels; ++c) {
#pragma omp atomic
                v_colors[(size_t)channels * g + c] += (double)(fac * v_out[c]);
                v_alpha += (rgb[c] * T - (float)Sh[c] * ra) * v_out[c];
                v_alpha += -T_final * ra * background[c] * v_out[c];
                Sh[c] = Sh[c] + (_Float16)(rgb[c] * fac);
              }
              v_alpha += T_final * ra * v_out_alpha;
            } else {
              for (int c = 0; c < channels; ++c) {
#pragma omp atomic
                v_colors[(size_t)channels * g + c] += (double)(fac * v_out[c]);
              }
              for (int c = 0; c < channels; ++c) v_alpha += (rgb[c] * T - S[c] * ra) * v_out[c];
              v_alpha += T_final * ra * v_out_alpha;
              for (int c = 0; c < channels; ++c) v_alpha += -T_final * ra * background[c] * v_out[c];
              for (int c = 0; c < channels; ++c) S[c] += rgb[c] * fac;
            }
            float v_sigma = -opac * vis * v_alpha;
            double a0 = (double)(0.5f * v_sigma * dx * dx), a1 = (double)(v_sigma * dx * dy),
                   a2 = (double)(0.5f * v_sigma * dy * dy);
            double b0 = (double)(v_sigma * (co[0] * dx + co[1] * dy)),
                   b1 = (double)(v_sigma * (co[1] * dx + co[2] * dy));
            double o0 = (double)(vis * v_alpha);
#pragma omp atomic
            v_conic[3 * (size_t)g] += a0;
#pragma omp atomic
            v_conic[3 * (size_t)g + 1] += a1;
#pragma omp atomic
            v_conic[3 * (size_t)g + 2] += a2;
#pragma omp atomic
            v_xy[2 * (size_t)g] += b0;
#pragma omp atomic
            v_xy[2 * (size_t)g + 1] += b1;
#pragma omp atomic
            v_opacity[g] += o0;
          }
        }
    }
}
