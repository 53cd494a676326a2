// host_math.cu — TEST INFRASTRUCTURE: runs the product's per-Gaussian maths (csrc/project_math.cuh, csrc/sh_math.cuh — the
// __host__ __device__ functions the fused per-Gaussian kernels call) on the CPU, so that tests/test_math_host.py can
// compare the product SOURCE with the oracle without a GPU.  Built by the test with nvcc; nothing in the product uses it.
#include "project_math.cuh"
#include "sh_math.cuh"

using namespace gsr;

extern "C" {

// project_one over n Gaussians: the arguments of gsr_project_gaussians_forward, quats in (w,x,y,z) order
void host_project_forward(int n, const float *means, const float *scales, float glob_scale, const float *quats,
                          const float *viewmat, const float *projmat, float fx, float fy, float cx, float cy, int img_h,
                          int img_w, int block_width, float clip_thresh, float *cov3d, float *xys, float *depths, int *radii,
                          float *conics, float *comp, int *tiles) {
  const int tiles_x = (img_w + block_width - 1) / block_width, tiles_y = (img_h + block_width - 1) / block_width;
  const float tfx = tan_half_fov((unsigned)img_w, fx), tfy = tan_half_fov((unsigned)img_h, fy);
  for (int i = 0; i < n; ++i) {
    const ProjFwd p = project_one(means[3 * i], means[3 * i + 1], means[3 * i + 2], glob_scale * scales[3 * i],
                                  glob_scale * scales[3 * i + 1], glob_scale * scales[3 * i + 2], quats[4 * i],
                                  quats[4 * i + 1], quats[4 * i + 2], quats[4 * i + 3], viewmat, projmat, fx, fy, cx, cy, tfx,
                                  tfy, img_w, img_h, tiles_x, tiles_y, block_width, clip_thresh);
    for (int k = 0; k < 6; ++k) cov3d[6 * i + k] = p.cov3d[k];
    for (int k = 0; k < 3; ++k) conics[3 * i + k] = p.conic[k];
    xys[2 * i] = p.x;
    xys[2 * i + 1] = p.y;
    depths[i] = p.depth;
    radii[i] = p.radius;
    comp[i] = p.comp;
    tiles[i] = p.tiles;
  }
}

// project_one_vjp over n Gaussians: the arguments of gsr_project_gaussians_backward
void host_project_backward(int n, const float *means, const float *scales, float glob_scale, const float *quats,
                           const float *viewmat, const float *projmat, float fx, float fy, int img_h, int img_w,
                           const float *cov3d, const int *radii, const float *conics, const float *comp, const float *v_xy,
                           const float *v_depth, const float *v_conic, const float *v_comp, float *v_mean, float *v_scale,
                           float *v_quat) {
  for (int i = 0; i < n; ++i) {
    const ProjBwd g = project_one_vjp(radii[i] > 0, means[3 * i], means[3 * i + 1], means[3 * i + 2], scales[3 * i],
                                      scales[3 * i + 1], scales[3 * i + 2], glob_scale, quats[4 * i], quats[4 * i + 1],
                                      quats[4 * i + 2], quats[4 * i + 3], viewmat, projmat, fx, fy, img_w, img_h, cov3d + 6 * i,
                                      conics[3 * i], conics[3 * i + 1], conics[3 * i + 2], comp[i], v_xy[2 * i],
                                      v_xy[2 * i + 1], v_depth[i], v_conic[3 * i], v_conic[3 * i + 1], v_conic[3 * i + 2],
                                      v_comp[i]);
    for (int k = 0; k < 3; ++k) v_mean[3 * i + k] = g.mean[k];
    for (int k = 0; k < 3; ++k) v_scale[3 * i + k] = g.scale[k];
    for (int k = 0; k < 4; ++k) v_quat[4 * i + k] = g.quat[k];
  }
}

// SH colour of n Gaussians from sh_basis_all, summed the way the fused per-Gaussian kernel does (fused.cu)
void host_sh_forward(int n, int K, int deg_use, const float *dirs, const float *coeffs, float *colors) {
  const int Ku = (deg_use + 1) * (deg_use + 1);
  for (int i = 0; i < n; ++i) {
    float Y[25];
    sh_basis_all(deg_use, dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2], Y);
    for (int c = 0; c < 3; ++c) {
      float acc = Y[0] * coeffs[(size_t)i * K * 3 + c];
      for (int k = 1; k < Ku; ++k) acc += Y[k] * coeffs[((size_t)i * K + k) * 3 + c];
      colors[3 * i + c] = acc;
    }
  }
}
}
