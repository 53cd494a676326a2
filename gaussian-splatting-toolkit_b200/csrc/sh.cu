// sh.cu — spherical-harmonics colour evaluation and its adjoint for sm_100a.
//
// Replaces compute_sh_forward_kernel / compute_sh_backward_kernel of the reference
// (gs_toolkit/gs_components/rasterizer/cuda/csrc/sh.cuh:33-224, launched from bindings.cu:58-103).
//
// Both kernels are pure HBM streaming (216 B per Gaussian at degree 3).  The reference lets every thread
// walk its own 192-byte coefficient row (32 different cache lines per warp-level load).  Here a CTA of
// 128 threads owns 128 consecutive Gaussians = one contiguous 128*K*3-float block, moves it between HBM
// and shared memory with fully coalesced 128-bit accesses, and each thread then reads / writes its own
// row in shared memory with an odd row stride (bank-conflict free).  The backward writes every one of
// the K*3 outputs itself (zeros above degrees_to_use), so no separate zero-fill pass is needed.
#include <stdlib.h>

#include "common.cuh"

namespace gsr {

__constant__ float kSH_C1 = 0.4886025119029199f;
__constant__ float kSH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                -1.0925484305920792f, 0.5462742152960396f};
__constant__ float kSH_C3[7] = {-0.5900435899266435f, 2.890611442640554f,  -0.4570457994644658f,
                                0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                                -0.5900435899266435f};
__constant__ float kSH_C4[9] = {2.5033429417967046f,  -1.7701307697799304f, 0.9461746957575601f,
                                -0.6690465435572892f, 0.10578554691520431f, -0.6690465435572892f,
                                0.47308734787878004f, -1.7701307697799304f, 0.6258357354491761f};
#define GSR_SH_C0 0.28209479177387814f

constexpr int SH_THREADS = 128;

__host__ __device__ inline int num_sh_bases(int degree) {  // sh.cuh:21-31
  return degree == 0 ? 1 : degree == 1 ? 4 : degree == 2 ? 9 : degree == 3 ? 16 : 25;
}

// basis values for the bands 1..deg (Y[0] = C0 is handled by the callers); dir normalised as sh.cuh:44-48
__device__ __forceinline__ void sh_basis(int deg, float dx, float dy, float dz, float *Y) {
  if (deg < 1) return;
  float norm = sqrtf(dx * dx + dy * dy + dz * dz);
  float x = dx / norm, y = dy / norm, z = dz / norm;
  Y[1] = -kSH_C1 * y;
  Y[2] = kSH_C1 * z;
  Y[3] = -kSH_C1 * x;
  if (deg < 2) return;
  float xx = x * x, xy = x * y, xz = x * z, yy = y * y, yz = y * z, zz = z * z;
  Y[4] = kSH_C2[0] * xy;
  Y[5] = kSH_C2[1] * yz;
  Y[6] = kSH_C2[2] * (2.f * zz - xx - yy);
  Y[7] = kSH_C2[3] * xz;
  Y[8] = kSH_C2[4] * (xx - yy);
  if (deg < 3) return;
  Y[9] = kSH_C3[0] * y * (3.f * xx - yy);
  Y[10] = kSH_C3[1] * xy * z;
  Y[11] = kSH_C3[2] * y * (4.f * zz - xx - yy);
  Y[12] = kSH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
  Y[13] = kSH_C3[4] * x * (4.f * zz - xx - yy);
  Y[14] = kSH_C3[5] * z * (xx - yy);
  Y[15] = kSH_C3[6] * x * (xx - 3.f * yy);
  if (deg < 4) return;
  Y[16] = kSH_C4[0] * xy * (xx - yy);
  Y[17] = kSH_C4[1] * yz * (3.f * xx - yy);
  Y[18] = kSH_C4[2] * xy * (7.f * zz - 1.f);
  Y[19] = kSH_C4[3] * yz * (7.f * zz - 3.f);
  Y[20] = kSH_C4[4] * (zz * (35.f * zz - 30.f) + 3.f);
  Y[21] = kSH_C4[5] * xz * (7.f * zz - 3.f);
  Y[22] = kSH_C4[6] * (xx - yy) * (7.f * zz - 1.f);
  Y[23] = kSH_C4[7] * xz * (xx - 3.f * yy);
  Y[24] = kSH_C4[8] * (xx * (xx - 3.f * yy) - yy * (3.f * xx - yy));
}

// smem row stride: odd => row-per-thread accesses are bank-conflict free
__host__ __device__ inline int sh_row_stride(int row_len) { return row_len | 1; }

// Coalesced global -> shared copy of `count` floats laid out [rows][row_len] into rows of `stride`.
__device__ __forceinline__ void stage_in(const float *__restrict__ g, float *s, int count, int row_len,
                                         int stride, bool vec_ok) {
  int tid = threadIdx.x;
  int nvec = vec_ok ? (count >> 2) : 0;
  const float4 *g4 = reinterpret_cast<const float4 *>(g);
  for (int i = tid; i < nvec; i += SH_THREADS) {
    float4 v = __ldg(g4 + i);
    int f = i << 2;
    int r = f / row_len, c = f - r * row_len;
    float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s[r * stride + c] = vv[j];
      if (++c == row_len) { c = 0; ++r; }
    }
  }
  for (int f = (nvec << 2) + tid; f < count; f += SH_THREADS) {
    int r = f / row_len, c = f - r * row_len;
    s[r * stride + c] = __ldg(g + f);
  }
}

// Coalesced shared -> global copy (inverse of stage_in).
__device__ __forceinline__ void stage_out(float *__restrict__ g, const float *s, int count, int row_len,
                                          int stride, bool vec_ok) {
  int tid = threadIdx.x;
  int nvec = vec_ok ? (count >> 2) : 0;
  float4 *g4 = reinterpret_cast<float4 *>(g);
  for (int i = tid; i < nvec; i += SH_THREADS) {
    int f = i << 2;
    int r = f / row_len, c = f - r * row_len;
    float vv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      vv[j] = s[r * stride + c];
      if (++c == row_len) { c = 0; ++r; }
    }
    g4[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
  }
  for (int f = (nvec << 2) + tid; f < count; f += SH_THREADS) {
    int r = f / row_len, c = f - r * row_len;
    g[f] = s[r * stride + c];
  }
}

__global__ void __launch_bounds__(SH_THREADS)
sh_forward_kernel(int n, int K, int deg_use, const float *__restrict__ viewdirs,
                  const float *__restrict__ coeffs, float *__restrict__ colors, int vec_ok) {
  extern __shared__ float smem[];
  const int row_len = 3 * K, stride = sh_row_stride(row_len);
  const int g0 = blockIdx.x * SH_THREADS;
  const int rows = min(SH_THREADS, n - g0);
  stage_in(coeffs + (size_t)g0 * row_len, smem, rows * row_len, row_len, stride, vec_ok != 0);
  __syncthreads();
  const int tid = threadIdx.x;
  if (tid < rows) {
    const int g = g0 + tid;
    float Y[25];
    sh_basis(deg_use, viewdirs[3 * (size_t)g], viewdirs[3 * (size_t)g + 1], viewdirs[3 * (size_t)g + 2], Y);
    const float *c = smem + tid * stride;
    float out[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      float acc = GSR_SH_C0 * c[ch];
      if (deg_use >= 1) acc += Y[1] * c[3 + ch] + Y[2] * c[6 + ch] + Y[3] * c[9 + ch];
      if (deg_use >= 2) {
        float s = 0.f;
#pragma unroll
        for (int k = 4; k < 9; ++k) s += Y[k] * c[3 * k + ch];
        acc += s;
      }
      if (deg_use >= 3) {
        float s = 0.f;
#pragma unroll
        for (int k = 9; k < 16; ++k) s += Y[k] * c[3 * k + ch];
        acc += s;
      }
      if (deg_use >= 4) {
        float s = 0.f;
#pragma unroll
        for (int k = 16; k < 25; ++k) s += Y[k] * c[3 * k + ch];
        acc += s;
      }
      out[ch] = acc;
    }
    colors[3 * (size_t)g] = out[0];
    colors[3 * (size_t)g + 1] = out[1];
    colors[3 * (size_t)g + 2] = out[2];
  }
}

__global__ void __launch_bounds__(SH_THREADS)
sh_backward_kernel(int n, int K, int deg_use, const float *__restrict__ viewdirs,
                   const float *__restrict__ v_colors, float *__restrict__ v_coeffs, int vec_ok) {
  extern __shared__ float smem[];
  const int row_len = 3 * K, stride = sh_row_stride(row_len);
  const int g0 = blockIdx.x * SH_THREADS;
  const int rows = min(SH_THREADS, n - g0);
  const int tid = threadIdx.x;
  const int Ku = num_sh_bases(deg_use);
  if (tid < rows) {
    const int g = g0 + tid;
    float Y[25];
    Y[0] = GSR_SH_C0;
    sh_basis(deg_use, viewdirs[3 * (size_t)g], viewdirs[3 * (size_t)g + 1], viewdirs[3 * (size_t)g + 2], Y);
    const float v0 = v_colors[3 * (size_t)g], v1 = v_colors[3 * (size_t)g + 1], v2 = v_colors[3 * (size_t)g + 2];
    float *o = smem + tid * stride;
#pragma unroll
    for (int k = 0; k < 25; ++k) {
      if (k < K) {
        float y = (k < Ku) ? Y[k] : 0.f;
        o[3 * k] = y * v0;
        o[3 * k + 1] = y * v1;
        o[3 * k + 2] = y * v2;
      }
    }
  }
  __syncthreads();
  stage_out(v_coeffs + (size_t)g0 * row_len, smem, rows * row_len, row_len, stride, vec_ok != 0);
}

// ---- row lengths that are a multiple of 4 floats (K = 4, 16: degrees 1 and 3) -------------------------------------------
// The block of 128 rows is moved with 16-byte cp.async copies (LDGSTS: global -> shared without passing through
// registers, every copy of a thread in flight at once) into rows padded to a stride of 3K + 4 floats: 16-byte aligned and
// = 4 (mod 8) words, so the row-per-thread 128-bit shared-memory accesses of the compute phase are conflict-free (a
// quarter warp covers eight different 16-byte slots of a 128-byte wavefront).  All index arithmetic is compile-time.
__device__ __forceinline__ void sh_cp_async16(float *smem_dst, const float *gmem_src) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(gmem_src) : "memory");
}

template <int K>
__global__ void __launch_bounds__(SH_THREADS)
sh_forward_vec_kernel(int n, int deg_use, const float *__restrict__ viewdirs, const float *__restrict__ coeffs,
                      float *__restrict__ colors) {
  constexpr int RL = 3 * K, ST = (RL % 8 == 4) ? RL : RL + 4;  // = 4 (mod 8) words
  __shared__ __align__(16) float s[SH_THREADS * ST];
  const int tid = threadIdx.x;
  const int g0 = blockIdx.x * SH_THREADS;
  const int rows = min(SH_THREADS, n - g0);
  const float *src = coeffs + (size_t)g0 * RL;
  const int nvec = rows * (RL / 4);
#pragma unroll
  for (int q = 0; q < RL / 4; ++q) {
    const int i = tid + q * SH_THREADS;
    if (i < nvec) {
      const int r = i / (RL / 4), c = i - r * (RL / 4);
      sh_cp_async16(s + r * ST + 4 * c, src + 4 * i);
    }
  }
  asm volatile("cp.async.commit_group;\n" ::: "memory");
  float Y[25];
  const int g = g0 + tid;
  if (tid < rows) sh_basis(deg_use, viewdirs[3 * (size_t)g], viewdirs[3 * (size_t)g + 1], viewdirs[3 * (size_t)g + 2], Y);
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  __syncthreads();
  if (tid < rows) {
    float c[RL];
#pragma unroll
    for (int q = 0; q < RL / 4; ++q) {
      const float4 v = *reinterpret_cast<const float4 *>(s + tid * ST + 4 * q);
      c[4 * q] = v.x; c[4 * q + 1] = v.y; c[4 * q + 2] = v.z; c[4 * q + 3] = v.w;
    }
    float out[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      float acc = GSR_SH_C0 * c[ch];
      if (K >= 4 && deg_use >= 1) acc += Y[1] * c[3 + ch] + Y[2] * c[6 + ch] + Y[3] * c[9 + ch];
      if (K >= 9 && deg_use >= 2) {
        float t = 0.f;
#pragma unroll
        for (int k = 4; k < 9; ++k) t += Y[k] * c[(3 * k + ch) % RL];
        acc += t;
      }
      if (K >= 16 && deg_use >= 3) {
        float t = 0.f;
#pragma unroll
        for (int k = 9; k < 16; ++k) t += Y[k] * c[(3 * k + ch) % RL];
        acc += t;
      }
      out[ch] = acc;
    }
    colors[3 * (size_t)g] = out[0];
    colors[3 * (size_t)g + 1] = out[1];
    colors[3 * (size_t)g + 2] = out[2];
  }
}

// ---- the same forward with TMA bulk copies ---------------------------------------------------------------------------
// One `cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes` per coefficient row (3K floats = 192 B at degree 3,
// 16-byte aligned on both sides) from each of the 128 threads, into the same padded rows; completion is counted in bytes
// on an mbarrier armed by thread 0 (`arrive.expect_tx`), which all threads wait on with `try_wait.parity` after they have
// evaluated the basis.  No register or LSU slot carries data (SASS: UBLKCP.S.G + SYNCS.ARRIVE.TRANS64 / PHASECHK.TRYWAIT;
// the copy instruction takes uniform operands, so the compiler elects the lanes of a warp one after the other).
// Measured at cfg2 on B200 (ncu, per launch): 36.9 us against 33.0 us for the LDGSTS variant above — 128 bulk copies of
// 192 B per CTA cost more TMA requests than 12 coalesced 16-byte copies per thread — and the same 0.060 ms stage inside
// the view, so it is OPT-IN (GSR_SH_TMA=1); one bulk copy of the whole contiguous 24 KB block would need unpadded rows,
// whose row-per-thread 128-bit reads are four-way bank conflicted.
template <int K>
__global__ void __launch_bounds__(SH_THREADS)
sh_forward_tma_kernel(int n, int deg_use, const float *__restrict__ viewdirs, const float *__restrict__ coeffs,
                      float *__restrict__ colors) {
  constexpr int RL = 3 * K, ST = (RL % 8 == 4) ? RL : RL + 4;  // = 4 (mod 8) words
  constexpr unsigned ROW_BYTES = RL * sizeof(float);
  __shared__ __align__(128) float s[SH_THREADS * ST];
  __shared__ __align__(8) unsigned long long mbar;
  const int tid = threadIdx.x;
  const int g0 = blockIdx.x * SH_THREADS;
  const int rows = min(SH_THREADS, n - g0);
  const unsigned mbar_addr = (unsigned)__cvta_generic_to_shared(&mbar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(mbar_addr) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(mbar_addr), "r"(rows * ROW_BYTES) : "memory");
  }
  __syncthreads();
  if (tid < rows) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(s + tid * ST);
    const float *src = coeffs + (size_t)(g0 + tid) * RL;
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                 "l"(src), "r"(ROW_BYTES), "r"(mbar_addr)
                 : "memory");
  }
  float Y[25];
  const int g = g0 + tid;
  if (tid < rows) sh_basis(deg_use, viewdirs[3 * (size_t)g], viewdirs[3 * (size_t)g + 1], viewdirs[3 * (size_t)g + 2], Y);
  {
    unsigned done = 0;
    while (!done) {
      asm volatile(
          "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n"
          : "=r"(done)
          : "r"(mbar_addr)
          : "memory");
    }
  }
  if (tid < rows) {
    float c[RL];
#pragma unroll
    for (int q = 0; q < RL / 4; ++q) {
      const float4 v = *reinterpret_cast<const float4 *>(s + tid * ST + 4 * q);
      c[4 * q] = v.x; c[4 * q + 1] = v.y; c[4 * q + 2] = v.z; c[4 * q + 3] = v.w;
    }
    float out[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      float acc = GSR_SH_C0 * c[ch];
      if (K >= 4 && deg_use >= 1) acc += Y[1] * c[3 + ch] + Y[2] * c[6 + ch] + Y[3] * c[9 + ch];
      if (K >= 9 && deg_use >= 2) {
        float t = 0.f;
#pragma unroll
        for (int k = 4; k < 9; ++k) t += Y[k] * c[(3 * k + ch) % RL];
        acc += t;
      }
      if (K >= 16 && deg_use >= 3) {
        float t = 0.f;
#pragma unroll
        for (int k = 9; k < 16; ++k) t += Y[k] * c[(3 * k + ch) % RL];
        acc += t;
      }
      out[ch] = acc;
    }
    colors[3 * (size_t)g] = out[0];
    colors[3 * (size_t)g + 1] = out[1];
    colors[3 * (size_t)g + 2] = out[2];
  }
}

template <int K>
__global__ void __launch_bounds__(SH_THREADS)
sh_backward_vec_kernel(int n, int deg_use, const float *__restrict__ viewdirs, const float *__restrict__ v_colors,
                       float *__restrict__ v_coeffs) {
  constexpr int RL = 3 * K, ST = (RL % 8 == 4) ? RL : RL + 4;  // = 4 (mod 8) words
  __shared__ __align__(16) float s[SH_THREADS * ST];
  const int tid = threadIdx.x;
  const int g0 = blockIdx.x * SH_THREADS;
  const int rows = min(SH_THREADS, n - g0);
  const int Ku = num_sh_bases(deg_use);
  if (tid < rows) {
    const int g = g0 + tid;
    float Y[25];
    Y[0] = GSR_SH_C0;
    sh_basis(deg_use, viewdirs[3 * (size_t)g], viewdirs[3 * (size_t)g + 1], viewdirs[3 * (size_t)g + 2], Y);
    const float v[3] = {v_colors[3 * (size_t)g], v_colors[3 * (size_t)g + 1], v_colors[3 * (size_t)g + 2]};
    float o[RL];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float y = (k < Ku) ? Y[k] : 0.f;
      o[3 * k] = y * v[0];
      o[3 * k + 1] = y * v[1];
      o[3 * k + 2] = y * v[2];
    }
#pragma unroll
    for (int q = 0; q < RL / 4; ++q)
      *reinterpret_cast<float4 *>(s + tid * ST + 4 * q) = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
  }
  __syncthreads();
  float4 *dst = reinterpret_cast<float4 *>(v_coeffs + (size_t)g0 * RL);
  const int nvec = rows * (RL / 4);
#pragma unroll
  for (int q = 0; q < RL / 4; ++q) {
    const int i = tid + q * SH_THREADS;
    if (i < nvec) {
      const int r = i / (RL / 4), c = i - r * (RL / 4);
      __stcs(dst + i, *reinterpret_cast<const float4 *>(s + r * ST + 4 * c));  // written once, read by the optimizer later
    }
  }
}

// Multi-view SH adjoint: v_coeffs[n,k,c] = sum_v Y_k(means[n] - cam[v]) * v_colors_v[n,c].
// This is the compute half of the view-parallel gradient exchange: instead of all-reducing the
// 3K-float SH gradient of every rank (192 B per Gaussian at degree 3), ranks exchange only the 3-float colour
// gradients (12 B) and every rank evaluates the outer products of ALL views while writing v_coeffs once.  The
// per-view colour gradients are addressed through a pointer table, so they may live in an all-gathered buffer or
// in peer GPUs' memory mapped over NVLink (P2P loads).
constexpr int SH_MAX_VIEWS = 16;
struct ShViews {
  const float *v_colors[SH_MAX_VIEWS];
  const float *cam[SH_MAX_VIEWS];  // each -> 3 floats (camera centre of that view), possibly in peer memory
};

__global__ void __launch_bounds__(SH_THREADS)
sh_backward_multiview_kernel(int n, int K, int deg_use, int num_views, const float *__restrict__ means3d,
                             const ShViews views, float *__restrict__ v_coeffs, int vec_ok, int vc_vec_ok) {
  extern __shared__ float smem[];
  __shared__ float s_cam[SH_MAX_VIEWS * 3];
  // the views' colour gradients of this block's 128 Gaussians: 384 contiguous floats per view.  They may live in PEER
  // memory (NVLink): a thread-per-Gaussian read (three 4-byte loads at a 12-byte stride) fetches every remote sector three
  // times, so the block copies them with coalesced 16-byte loads into shared memory first (round 2: the exchange tail at
  // 8 GPUs was bound by exactly this kernel, 0.30 ms for 84 MB of peer data)
  // (the staging area follows the output rows in the dynamic shared memory: [128][stride] floats, then [views][384])
  if (threadIdx.x < 3 * num_views) s_cam[threadIdx.x] = views.cam[threadIdx.x / 3][threadIdx.x % 3];
  const int row_len = 3 * K, stride = sh_row_stride(row_len);
  const int g0 = blockIdx.x * SH_THREADS;
  const int rows = min(SH_THREADS, n - g0);
  const int tid = threadIdx.x;
  float(*s_vc)[3 * SH_THREADS] = reinterpret_cast<float(*)[3 * SH_THREADS]>(smem + SH_THREADS * stride);
  {
    const int nf = 3 * rows, nvec = vc_vec_ok ? (nf >> 2) : 0;
    for (int idx = tid; idx < num_views * (3 * SH_THREADS / 4); idx += SH_THREADS) {
      const int v = idx / (3 * SH_THREADS / 4), q = idx - v * (3 * SH_THREADS / 4);
      if (q < nvec)
        *reinterpret_cast<float4 *>(&s_vc[v][4 * q]) = *reinterpret_cast<const float4 *>(views.v_colors[v] + 3 * (size_t)g0 + 4 * q);
    }
    for (int v = 0; v < num_views; ++v)
      for (int f = (nvec << 2) + tid; f < nf; f += SH_THREADS) s_vc[v][f] = views.v_colors[v][3 * (size_t)g0 + f];
  }
  __syncthreads();
  const int Ku = num_sh_bases(deg_use);
  if (tid < rows) {
    const int g = g0 + tid;
    const float mx = means3d[3 * (size_t)g], my = means3d[3 * (size_t)g + 1], mz = means3d[3 * (size_t)g + 2];
    // the sum over the views is accumulated in REGISTERS (round 2: a shared-memory read-modify-write of the 3K-float row
    // per view made the kernel 0.18 ms at eight views even with all data local) and written to the row once
    float acc[75];
#pragma unroll
    for (int i = 0; i < 75; ++i) acc[i] = 0.f;
    for (int v = 0; v < num_views; ++v) {
      float Y[25];
      Y[0] = GSR_SH_C0;
      sh_basis(deg_use, mx - s_cam[3 * v], my - s_cam[3 * v + 1], mz - s_cam[3 * v + 2], Y);
      const float v0 = s_vc[v][3 * tid], v1 = s_vc[v][3 * tid + 1], v2 = s_vc[v][3 * tid + 2];
#pragma unroll
      for (int k = 0; k < 25; ++k) {
        if (k < Ku) {
          acc[3 * k] += Y[k] * v0;
          acc[3 * k + 1] += Y[k] * v1;
          acc[3 * k + 2] += Y[k] * v2;
        }
      }
    }
    float *o = smem + tid * stride;
#pragma unroll
    for (int k = 0; k < 25; ++k) {
      if (k < K) {
        o[3 * k] = acc[3 * k];
        o[3 * k + 1] = acc[3 * k + 1];
        o[3 * k + 2] = acc[3 * k + 2];
      }
    }
  }
  __syncthreads();
  stage_out(v_coeffs + (size_t)g0 * row_len, smem, rows * row_len, row_len, stride, vec_ok != 0);
}

}  // namespace gsr

extern "C" {

GSR_API int gsr_compute_sh_backward_multiview(int num_points, int degree, int degrees_to_use, int num_views,
                                              const float *means3d, const float *cam_positions,
                                              const float *const *v_colors_views_host, float *v_coeffs,
                                              void *stream) {
  gsr::TraceScope _gsr_trace_scope("gsr_compute_sh_backward_multiview");
  // contiguous camera array -> pointer table
  if (num_views < 1 || num_views > gsr::SH_MAX_VIEWS || cam_positions == nullptr) {
    gsr::set_error("compute_sh_backward_multiview: num_views %d not in [1,%d] or null cam_positions", num_views,
                   gsr::SH_MAX_VIEWS);
    return GSR_ERR_INVALID_ARGUMENT;
  }
  const float *cams[gsr::SH_MAX_VIEWS];
  for (int v = 0; v < num_views; ++v) cams[v] = cam_positions + 3 * v;
  return gsr_compute_sh_backward_multiview_ptrs(num_points, degree, degrees_to_use, num_views, means3d, cams,
                                                v_colors_views_host, v_coeffs, stream);
}

GSR_API int gsr_compute_sh_backward_multiview_ptrs(int num_points, int degree, int degrees_to_use, int num_views,
                                                   const float *means3d, const float *const *cam_views_host,
                                                   const float *const *v_colors_views_host, float *v_coeffs,
                                                   void *stream) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_compute_sh_backward_multiview_ptrs");
  const float *const *cam_positions = cam_views_host;
  GSR_REQUIRE(num_points >= 0, GSR_ERR_INVALID_ARGUMENT, "compute_sh_backward_multiview: num_points < 0");
  GSR_REQUIRE(degree >= 0 && degree <= 4, GSR_ERR_UNSUPPORTED, "compute_sh_backward_multiview: degree %d not in [0,4]", degree);
  GSR_REQUIRE(degrees_to_use >= 0 && degrees_to_use <= degree, GSR_ERR_INVALID_ARGUMENT,
              "compute_sh_backward_multiview: degrees_to_use %d not in [0,%d]", degrees_to_use, degree);
  GSR_REQUIRE(num_views >= 1 && num_views <= SH_MAX_VIEWS, GSR_ERR_UNSUPPORTED,
              "compute_sh_backward_multiview: num_views %d not in [1,%d]", num_views, SH_MAX_VIEWS);
  if (num_points == 0) return GSR_OK;
  GSR_REQUIRE(means3d && cam_positions && v_colors_views_host && v_coeffs, GSR_ERR_INVALID_ARGUMENT,
              "compute_sh_backward_multiview: null pointer");
  ShViews views;
  for (int v = 0; v < num_views; ++v) {
    GSR_REQUIRE(v_colors_views_host[v] != nullptr, GSR_ERR_INVALID_ARGUMENT, "compute_sh_backward_multiview: null view pointer");
    GSR_REQUIRE(cam_views_host[v] != nullptr, GSR_ERR_INVALID_ARGUMENT, "compute_sh_backward_multiview: null camera pointer");
    views.v_colors[v] = v_colors_views_host[v];
    views.cam[v] = cam_views_host[v];
  }
  const int K = num_sh_bases(degree);
  const size_t smem = ((size_t)SH_THREADS * sh_row_stride(3 * K) + (size_t)num_views * 3 * SH_THREADS) * sizeof(float);
  if (smem > 48 * 1024) {
    static size_t configured = 0;  // opt in to more than 48 KB of dynamic shared memory (many views at degree 4)
    if (smem > configured) {
      GSR_CUDA(cudaFuncSetAttribute(sh_backward_multiview_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = smem;
    }
  }
  const int vec_ok = ((uintptr_t)v_coeffs % 16 == 0) ? 1 : 0;
  int vc_vec_ok = 1;
  for (int v = 0; v < num_views; ++v)
    if ((uintptr_t)views.v_colors[v] % 16 != 0) vc_vec_ok = 0;
  sh_backward_multiview_kernel<<<cdiv(num_points, SH_THREADS), SH_THREADS, smem, (cudaStream_t)stream>>>(
      num_points, K, degrees_to_use, num_views, means3d, views, v_coeffs, vec_ok, vc_vec_ok);
  GSR_CHECK_LAUNCH("sh_backward_multiview_kernel");
  return GSR_OK;
}


GSR_API int gsr_compute_sh_forward(int num_points, int degree, int degrees_to_use, const float *viewdirs,
                                   const float *coeffs, float *colors, void *stream) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_compute_sh_forward");
  GSR_REQUIRE(num_points >= 0, GSR_ERR_INVALID_ARGUMENT, "compute_sh_forward: num_points < 0");
  GSR_REQUIRE(degree >= 0 && degree <= 4, GSR_ERR_UNSUPPORTED, "compute_sh_forward: degree %d not in [0,4]", degree);
  GSR_REQUIRE(degrees_to_use >= 0 && degrees_to_use <= degree, GSR_ERR_INVALID_ARGUMENT,
              "compute_sh_forward: degrees_to_use %d not in [0,%d]", degrees_to_use, degree);
  if (num_points == 0) return GSR_OK;
  GSR_REQUIRE(viewdirs && coeffs && colors, GSR_ERR_INVALID_ARGUMENT, "compute_sh_forward: null pointer");
  const int K = num_sh_bases(degree);
  const size_t smem = (size_t)SH_THREADS * sh_row_stride(3 * K) * sizeof(float);
  // every CTA's block starts at g0*3K floats: 16-byte aligned iff the base is and 128*3K*4 % 16 == 0 (always)
  const int vec_ok = ((uintptr_t)coeffs % 16 == 0) ? 1 : 0;
  static const bool use_tma = [] {
    const char *e = getenv("GSR_SH_TMA");
    return e && e[0] == '1';
  }();
  if (vec_ok && K == 16 && use_tma)
    sh_forward_tma_kernel<16><<<cdiv(num_points, SH_THREADS), SH_THREADS, 0, (cudaStream_t)stream>>>(
        num_points, degrees_to_use, viewdirs, coeffs, colors);
  else if (vec_ok && K == 16)
    sh_forward_vec_kernel<16><<<cdiv(num_points, SH_THREADS), SH_THREADS, 0, (cudaStream_t)stream>>>(
        num_points, degrees_to_use, viewdirs, coeffs, colors);
  else if (vec_ok && K == 4)
    sh_forward_vec_kernel<4><<<cdiv(num_points, SH_THREADS), SH_THREADS, 0, (cudaStream_t)stream>>>(
        num_points, degrees_to_use, viewdirs, coeffs, colors);
  else
    sh_forward_kernel<<<cdiv(num_points, SH_THREADS), SH_THREADS, smem, (cudaStream_t)stream>>>(
        num_points, K, degrees_to_use, viewdirs, coeffs, colors, vec_ok);
  GSR_CHECK_LAUNCH("sh_forward_kernel");
  return GSR_OK;
}

GSR_API int gsr_compute_sh_backward(int num_points, int degree, int degrees_to_use, const float *viewdirs,
                                    const float *v_colors, float *v_coeffs, void *stream) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_compute_sh_backward");
  GSR_REQUIRE(num_points >= 0, GSR_ERR_INVALID_ARGUMENT, "compute_sh_backward: num_points < 0");
  GSR_REQUIRE(degree >= 0 && degree <= 4, GSR_ERR_UNSUPPORTED, "compute_sh_backward: degree %d not in [0,4]", degree);
  GSR_REQUIRE(degrees_to_use >= 0 && degrees_to_use <= degree, GSR_ERR_INVALID_ARGUMENT,
              "compute_sh_backward: degrees_to_use %d not in [0,%d]", degrees_to_use, degree);
  if (num_points == 0) return GSR_OK;
  GSR_REQUIRE(viewdirs && v_colors && v_coeffs, GSR_ERR_INVALID_ARGUMENT, "compute_sh_backward: null pointer");
  const int K = num_sh_bases(degree);
  const size_t smem = (size_t)SH_THREADS * sh_row_stride(3 * K) * sizeof(float);
  const int vec_ok = ((uintptr_t)v_coeffs % 16 == 0) ? 1 : 0;
  if (vec_ok && K == 16)
    sh_backward_vec_kernel<16><<<cdiv(num_points, SH_THREADS), SH_THREADS, 0, (cudaStream_t)stream>>>(
        num_points, degrees_to_use, viewdirs, v_colors, v_coeffs);
  else if (vec_ok && K == 4)
    sh_backward_vec_kernel<4><<<cdiv(num_points, SH_THREADS), SH_THREADS, 0, (cudaStream_t)stream>>>(
        num_points, degrees_to_use, viewdirs, v_colors, v_coeffs);
  else
    sh_backward_kernel<<<cdiv(num_points, SH_THREADS), SH_THREADS, smem, (cudaStream_t)stream>>>(
        num_points, K, degrees_to_use, viewdirs, v_colors, v_coeffs, vec_ok);
  GSR_CHECK_LAUNCH("sh_backward_kernel");
  return GSR_OK;
}
}
