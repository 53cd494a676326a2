// optim.cu — multi-tensor Adam step + opacity reset for the Gaussian parameter groups, sm_100a.
//
// Replaces, per training iteration, the six `torch.optim.Adam(lr=..., eps=1e-15).step()` calls the reference issues
// through `Optimizers.optimizer_step_all` (gs_toolkit/engine/optimizers.py:173-180; groups and learning rates
// configs/method_configs.py:98-125: means, features_dc, features_rest, opacities, scales, quats).  torch's default
// CUDA implementation is the "foreach" one: ~12 elementwise kernels per group, i.e. ~70 launches and ~10 passes over
// 59·N floats of parameters + 2·59·N floats of state.  Here ALL groups are stepped by ONE launch that touches every
// array once: 16 B read + 12 B written per parameter element (p, g, m, v -> p, m, v), HBM-bound.
//
// Arithmetic = torch/optim/adam.py `_multi_tensor_adam` / `_single_tensor_adam` (amsgrad=False, weight_decay=0,
// maximize=False, capturable=False), the scalars evaluated by the host in binary64 as torch does:
//     m   <- m + (g - m) * (1 - beta1)                      (Tensor.lerp_)
//     v   <- v * beta2 + (1 - beta2) * g * g                (mul_ + addcmul_)
//     p   <- p - step_size * m / (sqrt(v) / bias2_sqrt + eps),  step_size = lr / (1 - beta1^t), bias2_sqrt = sqrt(1 - beta2^t)
// This translation unit is compiled WITHOUT --use_fast_math (IEEE division and square root; see csrc/Makefile).
#include "common.cuh"

namespace gsr {

constexpr int ADAM_THREADS = 256;
constexpr int ADAM_VEC_PER_THREAD = 2;                                // float4 per array per thread per chunk
constexpr int ADAM_CHUNK = ADAM_THREADS * ADAM_VEC_PER_THREAD * 4;    // 2048 floats per CTA iteration

struct AdamSegment {
  float *p;
  const float *g;
  float *m;
  float *v;
  long long numel;
  float step_size;        // lr / (1 - beta1^t)
  float bias2_sqrt;       // sqrt(1 - beta2^t)
  int vec_ok;             // all four base pointers 16-byte aligned
};

struct AdamLaunch {
  AdamSegment seg[GSR_ADAM_MAX_SEGMENTS];
  unsigned chunk_end[GSR_ADAM_MAX_SEGMENTS];  // running total of chunks
  int num_segments;
  float beta1, beta2, one_minus_beta1, one_minus_beta2, eps, grad_scale;
};

__device__ __forceinline__ void adam_elem(float &p, float g, float &m, float &v, const AdamLaunch &L,
                                          const AdamSegment &S) {
  g *= L.grad_scale;
  m = m + (g - m) * L.one_minus_beta1;
  v = v * L.beta2 + (L.one_minus_beta2 * g) * g;
  const float denom = sqrtf(v) / S.bias2_sqrt + L.eps;
  p = p - S.step_size * (m / denom);
}

__global__ void __launch_bounds__(ADAM_THREADS)
adam_multi_kernel(const __grid_constant__ AdamLaunch L, unsigned total_chunks) {
  for (unsigned chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
    int k = 0;
    while (k + 1 < L.num_segments && chunk >= L.chunk_end[k]) ++k;
    const AdamSegment &S = L.seg[k];
    const long long base = (long long)(chunk - (k ? L.chunk_end[k - 1] : 0u)) * ADAM_CHUNK;
    const long long left = S.numel - base;
    if (S.vec_ok && left >= ADAM_CHUNK) {
      float4 p[ADAM_VEC_PER_THREAD], g[ADAM_VEC_PER_THREAD], m[ADAM_VEC_PER_THREAD], v[ADAM_VEC_PER_THREAD];
#pragma unroll
      for (int u = 0; u < ADAM_VEC_PER_THREAD; ++u) {
        const long long i = base / 4 + u * ADAM_THREADS + threadIdx.x;
        p[u] = reinterpret_cast<const float4 *>(S.p)[i];
        g[u] = __ldcs(reinterpret_cast<const float4 *>(S.g) + i);   // gradients are dead after the step: stream them
        m[u] = reinterpret_cast<const float4 *>(S.m)[i];
        v[u] = reinterpret_cast<const float4 *>(S.v)[i];
      }
#pragma unroll
      for (int u = 0; u < ADAM_VEC_PER_THREAD; ++u) {
        adam_elem(p[u].x, g[u].x, m[u].x, v[u].x, L, S);
        adam_elem(p[u].y, g[u].y, m[u].y, v[u].y, L, S);
        adam_elem(p[u].z, g[u].z, m[u].z, v[u].z, L, S);
        adam_elem(p[u].w, g[u].w, m[u].w, v[u].w, L, S);
        const long long i = base / 4 + u * ADAM_THREADS + threadIdx.x;
        reinterpret_cast<float4 *>(S.p)[i] = p[u];
        reinterpret_cast<float4 *>(S.m)[i] = m[u];
        reinterpret_cast<float4 *>(S.v)[i] = v[u];
      }
    } else {
      const int n = (int)(left < ADAM_CHUNK ? left : ADAM_CHUNK);
      for (int j = threadIdx.x; j < n; j += ADAM_THREADS) {
        const long long i = base + j;
        float p = S.p[i], m = S.m[i], v = S.v[i];
        adam_elem(p, S.g[i], m, v, L, S);
        S.p[i] = p;
        S.m[i] = m;
        S.v[i] = v;
      }
    }
  }
}

// opacities <- min(opacities, max_logit); Adam moments of the group <- 0  (vanilla_gs.py:472-489)
__global__ void __launch_bounds__(256)
opacity_reset_kernel(int n, float max_logit, float *__restrict__ opac, float *__restrict__ m, float *__restrict__ v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  opac[i] = fminf(opac[i], max_logit);
  if (m) m[i] = 0.f;
  if (v) v[i] = 0.f;
}

}  // namespace gsr

extern "C" {

GSR_API int gsr_adam_step_multi(int num_segments, float *const *params_host, const float *const *grads_host,
                                float *const *exp_avg_host, float *const *exp_avg_sq_host,
                                const int64_t *numels_host, const double *lrs_host, const int64_t *steps_host,
                                double beta1, double beta2, double eps, float grad_scale, void *stream) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_adam_step_multi");
  GSR_REQUIRE(num_segments >= 1 && num_segments <= GSR_ADAM_MAX_SEGMENTS, GSR_ERR_INVALID_ARGUMENT,
              "adam_step_multi: num_segments must be in [1,%d] (got %d)", GSR_ADAM_MAX_SEGMENTS, num_segments);
  GSR_REQUIRE(params_host && grads_host && exp_avg_host && exp_avg_sq_host && numels_host && lrs_host && steps_host,
              GSR_ERR_INVALID_ARGUMENT, "adam_step_multi: null table pointer");
  AdamLaunch L;
  L.num_segments = 0;
  L.beta1 = (float)beta1;
  L.beta2 = (float)beta2;
  L.one_minus_beta1 = (float)(1.0 - beta1);
  L.one_minus_beta2 = (float)(1.0 - beta2);
  L.eps = (float)eps;
  L.grad_scale = grad_scale;
  unsigned long long total = 0;
  for (int k = 0; k < num_segments; ++k) {
    GSR_REQUIRE(numels_host[k] >= 0, GSR_ERR_INVALID_ARGUMENT, "adam_step_multi: negative numel in segment %d", k);
    GSR_REQUIRE(steps_host[k] >= 1, GSR_ERR_INVALID_ARGUMENT, "adam_step_multi: step of segment %d must be >= 1", k);
    if (numels_host[k] == 0) continue;
    GSR_REQUIRE(params_host[k] && grads_host[k] && exp_avg_host[k] && exp_avg_sq_host[k], GSR_ERR_INVALID_ARGUMENT,
                "adam_step_multi: null pointer in segment %d", k);
    AdamSegment &S = L.seg[L.num_segments];
    S.p = params_host[k];
    S.g = grads_host[k];
    S.m = exp_avg_host[k];
    S.v = exp_avg_sq_host[k];
    S.numel = numels_host[k];
    // the scalars exactly as torch evaluates them (python floats = binary64), then rounded to binary32
    const double bc1 = 1.0 - pow(beta1, (double)steps_host[k]);
    const double bc2 = 1.0 - pow(beta2, (double)steps_host[k]);
    S.step_size = (float)(lrs_host[k] / bc1);
    S.bias2_sqrt = (float)sqrt(bc2);
    const uintptr_t bits = (uintptr_t)S.p | (uintptr_t)S.g | (uintptr_t)S.m | (uintptr_t)S.v;
    S.vec_ok = (bits & 15u) == 0;
    total += (unsigned long long)((S.numel + ADAM_CHUNK - 1) / ADAM_CHUNK);
    GSR_REQUIRE(total < 0xffffffffull, GSR_ERR_UNSUPPORTED, "adam_step_multi: too many elements");
    L.chunk_end[L.num_segments] = (unsigned)total;
    ++L.num_segments;
  }
  if (L.num_segments == 0) return GSR_OK;
  int dev = 0, sms = 148;
  GSR_CUDA(cudaGetDevice(&dev));
  GSR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const unsigned grid = (unsigned)(total < (unsigned long long)sms * 8 ? total : (unsigned long long)sms * 8);
  adam_multi_kernel<<<grid, ADAM_THREADS, 0, (cudaStream_t)stream>>>(L, (unsigned)total);
  GSR_CHECK_LAUNCH("adam_multi_kernel");
  return GSR_OK;
}

GSR_API int gsr_opacity_reset(int num_points, float max_logit, float *opacities_raw, float *exp_avg /*nullable*/,
                              float *exp_avg_sq /*nullable*/, void *stream) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_opacity_reset");
  GSR_REQUIRE(num_points >= 0, GSR_ERR_INVALID_ARGUMENT, "opacity_reset: negative num_points");
  if (num_points == 0) return GSR_OK;
  GSR_REQUIRE(opacities_raw, GSR_ERR_INVALID_ARGUMENT, "opacity_reset: null pointer");
  opacity_reset_kernel<<<cdiv(num_points, 256), 256, 0, (cudaStream_t)stream>>>(num_points, max_logit, opacities_raw,
                                                                               exp_avg, exp_avg_sq);
  GSR_CHECK_LAUNCH("opacity_reset_kernel");
  return GSR_OK;
}
}
