// common.cuh — shared host/device helpers of libgsr_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/gsr_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libgsr_b200 is written for sm_100a (B200) only"
#endif

namespace gsr {

// tan(fov / 2) as the reference forms it: `0.5 * img_size.x / fx` evaluated in DOUBLE and rounded once (forward.cu:71-72).
// Computed on the host and passed to the kernels: on the device the two double divisions were ~45 dependent FP64
// instructions per Gaussian (DFMA / DMUL / slow-path CALL) in an otherwise FP32 streaming kernel.
inline float tan_half_fov(unsigned img_size, float focal) { return (float)(0.5 * (double)img_size / (double)focal); }

// thread-local error message (api.cu)
void set_error(const char *fmt, ...);
// process-wide count of own kernel launches (api.cu; read with gsr_launch_count()); with GSR_NVTX=1 every launch also
// leaves an NVTX marker carrying the kernel's name
void count_launch(const char *kernel_name = nullptr);
// NVTX range over a C-ABI entry point (api.cu; active only with GSR_NVTX=1 — one relaxed load otherwise): the ranges and
// markers show up in `ncu --nvtx` / Nsight Systems captures of a process that uses the library
struct TraceScope {
  explicit TraceScope(const char *name);
  ~TraceScope();
  bool on;
};
#define GSR_TRACE_SCOPE(name) gsr::TraceScope _gsr_trace_scope(name)

#define GSR_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      gsr::set_error(__VA_ARGS__);    \
      return (code);                  \
    }                                 \
  } while (0)

#define GSR_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (call);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      gsr::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
      return GSR_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

#define GSR_CHECK_LAUNCH(name)                                                          \
  do {                                                                                  \
    cudaError_t _e = cudaGetLastError();                                                \
    if (_e != cudaSuccess) {                                                            \
      gsr::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));          \
      return GSR_ERR_CUDA;                                                              \
    }                                                                                   \
    gsr::count_launch(name);                                                            \
  } while (0)

static inline unsigned cdiv(unsigned a, unsigned b) { return (a + b - 1) / b; }

// The geometric helpers below (tile boxes, the culling ellipse, the warp-block masks) are __host__ __device__ so that
// tests/test_cull_host.py can run the SAME source on the CPU against a brute-force FP64 evaluation of the reference's
// per-pixel alpha test; the device code is unchanged (fast-math intrinsics on the device, libm on the host).
#define GSR_HD __host__ __device__ __forceinline__
GSR_HD float gsr_logf(float x) {
#ifdef __CUDA_ARCH__
  return __logf(x);
#else
  return logf(x);
#endif
}
GSR_HD float gsr_log2f(float x) {
#ifdef __CUDA_ARCH__
  return __log2f(x);
#else
  return log2f(x);
#endif
}
GSR_HD unsigned gsr_float_as_uint(float x) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(x);
#else
  unsigned u;
  memcpy(&u, &x, sizeof(u));
  return u;
#endif
}
GSR_HD float gsr_int_as_float(int i) {
#ifdef __CUDA_ARCH__
  return __int_as_float(i);
#else
  float f;
  memcpy(&f, &i, sizeof(f));
  return f;
#endif
}

// tile bbox of a projected Gaussian: reference helpers.cuh:11-34 ((int) truncates toward zero,
// clamp to [0, tiles]).
GSR_HD void tile_bbox(float cx, float cy, float radius, int tiles_x, int tiles_y,
                                          int block_width, int &x0, int &y0, int &x1, int &y1) {
  const float bw = (float)block_width;
  const float tcx = cx / bw, tcy = cy / bw, tr = radius / bw;
  x0 = min(max(0, (int)(tcx - tr)), tiles_x);
  x1 = min(max(0, (int)(tcx + tr + 1)), tiles_x);
  y0 = min(max(0, (int)(tcy - tr)), tiles_y);
  y1 = min(max(0, (int)(tcy + tr + 1)), tiles_y);
}

// cov2d -> conic + 3-sigma radius: reference helpers.cuh:36-59
GSR_HD bool cov2d_to_conic_radius(float cxx, float cxy, float cyy, float &ca,
                                                      float &cb, float &cc, float &radius) {
  float det = cxx * cyy - cxy * cxy;
  if (det == 0.f) return false;
  float inv_det = 1.f / det;
  ca = cyy * inv_det;
  cb = -cxy * inv_det;
  cc = cxx * inv_det;
  float b = 0.5f * (cxx + cyy);
  float v1 = b + sqrtf(fmaxf(0.1f, b * b - det));
  float v2 = b - sqrtf(fmaxf(0.1f, b * b - det));
  radius = ceilf(3.f * sqrtf(fmaxf(v1, v2)));
  return true;
}

// (w,x,y,z) quaternion -> row-major rotation matrix (normalises inside): reference helpers.cuh:144-159
GSR_HD void quat_to_rotmat(float qw, float qx, float qy, float qz, float R[9]) {
  // summation order of the reference: its float4 holds (w,x,y,z) in the fields (x,y,z,w) and it adds w*w + x*x + y*y + z*z
  // over the FIELDS, i.e. z^2 first (helpers.cuh:146-147)
  float s = rsqrtf(qz * qz + qw * qw + qx * qx + qy * qy);
  float w = qw * s, x = qx * s, y = qy * s, z = qz * s;
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

// C = A * B, 3x3 row-major, fully unrolled
GSR_HD void mat3_mul(const float A[9], const float B[9], float C[9]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

}  // namespace gsr
