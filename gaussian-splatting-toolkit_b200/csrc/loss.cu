// loss.cu — fused photometric loss of the reference model, forward and adjoint, for sm_100a.
//
// Replaces, per training view, gs_toolkit/models/vanilla_gs.py:926-934:
//     Ll1     = torch.abs(gt_img - pred_img).mean()
//     simloss = 1 - self.ssim(gt_img.permute(2,0,1)[None], pred_img.permute(2,0,1)[None])
//     loss    = (1 - ssim_lambda) * Ll1 + ssim_lambda * simloss
// where self.ssim = pytorch_msssim.SSIM(data_range=1.0, size_average=True, channel=3) (vanilla_gs.py:226; the package
// is a third-party dependency that is not vendored in the reference tree — its published algorithm is restated here:
// 11-tap Gaussian window, sigma 1.5, separable, VALID padding, K1 = 0.01, K2 = 0.03,
// ssim = mean over (H-10) x (W-10) x 3 of ((2 mu1 mu2 + C1)(2 s12 + C2)) / ((mu1^2 + mu2^2 + C1)(s1 + s2 + C2))).
// In torch this is 10 grouped convolutions forward and as many backward plus ~30 elementwise kernels over full-size
// images; here it is ONE stencil kernel each way, working directly on the HWC images the rasterizer produces:
//   forward : a CTA loads a (16+10)^2 x 3 patch of both images with coalesced row reads, runs the separable window
//             through shared memory for the five moments, evaluates SSIM and the three derivative maps
//             A = ds/dmu1 - 2 mu1 ds/ds1 - mu2 ds/ds12, B = ds/ds1, C = ds/ds12, and reduces SSIM and L1 partial sums
//             (one partial per CTA => deterministic final sum);
//   backward: d loss / d pred(x) = (1-l) sign(pred-gt)/n_px - l/n_ssim [ (w*A)(x) + 2 pred(x) (w*B)(x) + gt(x) (w*C)(x) ]
//             with w* the transposed ("full") window applied to the maps, again separably through shared memory.
// HBM-bound stencils: forward reads 24 P and writes 36 P' bytes, backward reads 36 P' + 24 P and writes 12 P.
#include "common.cuh"

namespace gsr {

constexpr int LT = 16;           // output tile edge
constexpr int LW = 11;           // window taps
constexpr int LH = LT + LW - 1;  // 26: patch edge
constexpr int LOSS_THREADS = 256;

__constant__ float kWin[LW] = {0.00102838036f, 0.00759875821f, 0.0360007733f, 0.109360687f, 0.213005528f, 0.266011715f, 0.213005528f, 0.109360687f, 0.0360007733f, 0.00759875821f, 0.00102838036f};

// pred, gt: [H,W,3]; maps: [3][H-10][W-10][3] (A, B, C); partials: [num_ctas][2] = {sum ssim, sum |pred-gt|}
__global__ void __launch_bounds__(LOSS_THREADS)
l1_ssim_forward_kernel(int H, int W, const float *__restrict__ pred, const float *__restrict__ gt,
                       float *__restrict__ maps, float *__restrict__ partials) {
  __shared__ float s_p[LH][LH * 3 + 1];
  __shared__ float s_g[LH][LH * 3 + 1];
  __shared__ float s_h[5][LH][LT * 3 + 1];
  __shared__ float s_red[2][LOSS_THREADS / 32];

  const int Ho = H - (LW - 1), Wo = W - (LW - 1);
  const int ox0 = blockIdx.x * LT, oy0 = blockIdx.y * LT;  // first output (= first input) pixel of the tile
  const int tid = threadIdx.x;
  const bool last_x = (blockIdx.x == gridDim.x - 1), last_y = (blockIdx.y == gridDim.y - 1);

  // load the patch row by row (a row = 26 px x 3 ch = 78 contiguous floats: warp w takes rows w, w+8, ..; lanes take
  // columns lane, lane+32, lane+64) and accumulate the L1 of the pixels this tile owns
  float l1 = 0.f;
  {
    const int lane = tid & 31, wrp = tid >> 5;
    const int ncols = min(LH, W - ox0) * 3;  // valid floats of a patch row
    for (int r = wrp; r < LH; r += LOSS_THREADS / 32) {
      const int y = oy0 + r;
      const bool own_y = (r < LT) || last_y;
      const size_t row = ((size_t)y * W + ox0) * 3;
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        const int c = lane + 32 * cc;
        if (c < LH * 3) {
          float p = 0.f, g = 0.f;
          if (y < H && c < ncols) {
            p = pred[row + c];
            g = gt[row + c];
            if (own_y && (c < LT * 3 || last_x)) l1 += fabsf(p - g);
          }
          s_p[r][c] = p;
          s_g[r][c] = g;
        }
      }
    }
  }
  __syncthreads();
  // horizontal pass, register-tiled: one item = 4 consecutive output pixels of one (row, channel); the 14 inputs are
  // loaded once and their products formed once (28 LDS for 20 outputs instead of 88)
  for (int i = tid; i < LH * (LT / 4) * 3; i += LOSS_THREADS) {
    const int ch = i % 3, pg = (i / 3) % (LT / 4), r = i / (3 * (LT / 4));
    float p[14], g[14];
#pragma unroll
    for (int j = 0; j < 14; ++j) {
      p[j] = s_p[r][(4 * pg + j) * 3 + ch];
      g[j] = s_g[r][(4 * pg + j) * 3 + ch];
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
      for (int k = 0; k < LW; ++k) {
        const float w = kWin[k], pv = p[o + k], gv = g[o + k];
        m1 += w * pv;
        m2 += w * gv;
        e11 += w * (pv * pv);
        e22 += w * (gv * gv);
        e12 += w * (pv * gv);
      }
      const int c = (4 * pg + o) * 3 + ch;
      s_h[0][r][c] = m1;
      s_h[1][r][c] = m2;
      s_h[2][r][c] = e11;
      s_h[3][r][c] = e22;
      s_h[4][r][c] = e12;
    }
  }
  __syncthreads();
  // vertical pass, register-tiled: one item = 4 consecutive output rows of one (pixel, channel) column
  const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
  const size_t plane = (size_t)Ho * Wo * 3;
  float ssim_sum = 0.f;
  for (int i = tid; i < (LT / 4) * LT * 3; i += LOSS_THREADS) {
    const int c = i % (LT * 3), rg = i / (LT * 3);
    const int ox = ox0 + c / 3;
    if (ox >= Wo || oy0 + 4 * rg >= Ho) continue;
    float acc[4][5];
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
      for (int q = 0; q < 5; ++q) acc[o][q] = 0.f;
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      float v[14];
#pragma unroll
      for (int j = 0; j < 14; ++j) v[j] = s_h[q][4 * rg + j][c];
#pragma unroll
      for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int k = 0; k < LW; ++k) acc[o][q] += kWin[k] * v[o + k];
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const int oy = oy0 + 4 * rg + o;
      if (oy >= Ho) continue;
      const float m1 = acc[o][0], m2 = acc[o][1];
      const float s1 = acc[o][2] - m1 * m1, s2 = acc[o][3] - m2 * m2, s12 = acc[o][4] - m1 * m2;
      const float a1 = 2.f * m1 * m2 + C1, a2 = 2.f * s12 + C2;
      const float b1 = m1 * m1 + m2 * m2 + C1, b2 = s1 + s2 + C2;
      const float inv = 1.f / (b1 * b2);
      const float ssim = a1 * a2 * inv;
      ssim_sum += ssim;
      // d ssim / d(mu1, s1, s12) at fixed (mu2, s2)
      const float d_mu1 = (2.f * m2 * a2 * inv) - ssim * (2.f * m1 / b1);
      const float d_s1 = -ssim / b2;
      const float d_s12 = 2.f * a1 * inv;
      const size_t oidx = ((size_t)oy * Wo + ox) * 3 + (c % 3);
      maps[oidx] = d_mu1 - 2.f * m1 * d_s1 - m2 * d_s12;
      maps[plane + oidx] = d_s1;
      maps[2 * plane + oidx] = d_s12;
    }
  }
  // CTA reduction of the two sums -> one partial per CTA
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ssim_sum += __shfl_xor_sync(0xffffffffu, ssim_sum, o);
    l1 += __shfl_xor_sync(0xffffffffu, l1, o);
  }
  if ((tid & 31) == 0) {
    s_red[0][tid >> 5] = ssim_sum;
    s_red[1][tid >> 5] = l1;
  }
  __syncthreads();
  if (tid == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < LOSS_THREADS / 32; ++w) {
      a += s_red[0][w];
      b += s_red[1][w];
    }
    const int cta = blockIdx.y * gridDim.x + blockIdx.x;
    partials[2 * cta] = a;
    partials[2 * cta + 1] = b;
  }
}

// v_pred [H,W,3] = scale_l1 * sign(pred - gt) + scale_ssim * [ (w*A) + 2 pred (w*B) + gt (w*C) ]
__global__ void __launch_bounds__(LOSS_THREADS)
l1_ssim_backward_kernel(int H, int W, const float *__restrict__ pred, const float *__restrict__ gt,
                        const float *__restrict__ maps, float scale_l1, float scale_ssim,
                        const float *__restrict__ v_loss, float *__restrict__ v_pred) {
  __shared__ float s_m[3][LH][LH * 3 + 1];
  __shared__ float s_h[3][LH][LT * 3 + 1];
  const int Ho = H - (LW - 1), Wo = W - (LW - 1);
  const int x0 = blockIdx.x * LT, y0 = blockIdx.y * LT;  // first input pixel of the tile
  const int tid = threadIdx.x;
  const size_t plane = (size_t)Ho * Wo * 3;
  // outputs o in [x-10, x] contribute to input x: patch of the maps starting at (x0-10, y0-10), zero outside;
  // loaded row by row (warp per row, lanes over the 78 floats of a row)
  {
    const int lane = tid & 31, wrp = tid >> 5;
    const int oxs = x0 - (LW - 1);
    for (int r = wrp; r < LH; r += LOSS_THREADS / 32) {
      const int oy = y0 - (LW - 1) + r;
      const bool row_ok = (oy >= 0 && oy < Ho);
      const long long row = ((long long)oy * Wo + oxs) * 3;
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        const int c = lane + 32 * cc;
        if (c < LH * 3) {
          const int ox = oxs + c / 3;
          float a = 0.f, b = 0.f, c2 = 0.f;
          if (row_ok && ox >= 0 && ox < Wo) {
            a = maps[row + c];
            b = maps[plane + row + c];
            c2 = maps[2 * plane + row + c];
          }
          s_m[0][r][c] = a;
          s_m[1][r][c] = b;
          s_m[2][r][c] = c2;
        }
      }
    }
  }
  __syncthreads();
  // horizontal pass (transposed window), register-tiled like the forward
  for (int i = tid; i < LH * (LT / 4) * 3; i += LOSS_THREADS) {
    const int ch = i % 3, pg = (i / 3) % (LT / 4), r = i / (3 * (LT / 4));
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      float v[14];
#pragma unroll
      for (int j = 0; j < 14; ++j) v[j] = s_m[q][r][(4 * pg + j) * 3 + ch];
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < LW; ++k) acc += kWin[LW - 1 - k] * v[o + k];
        s_h[q][r][(4 * pg + o) * 3 + ch] = acc;
      }
    }
  }
  __syncthreads();
  const float up = v_loss ? v_loss[0] : 1.f;
  for (int i = tid; i < (LT / 4) * LT * 3; i += LOSS_THREADS) {
    const int c = i % (LT * 3), rg = i / (LT * 3);
    const int x = x0 + c / 3;
    if (x >= W || y0 + 4 * rg >= H) continue;
    float acc[4][3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      float v[14];
#pragma unroll
      for (int j = 0; j < 14; ++j) v[j] = s_h[q][4 * rg + j][c];
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < LW; ++k) t += kWin[LW - 1 - k] * v[o + k];
        acc[o][q] = t;
      }
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const int y = y0 + 4 * rg + o;
      if (y >= H) continue;
      const size_t idx = ((size_t)y * W + x) * 3 + (c % 3);
      const float p = pred[idx], g = gt[idx];
      const float d = p - g;
      const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
      v_pred[idx] = up * (scale_l1 * sgn + scale_ssim * (acc[o][0] + 2.f * p * acc[o][1] + g * acc[o][2]));
    }
  }
}

// one CTA: deterministic FP64 reduction of the per-CTA partials -> out = {loss, L1, SSIM}
__global__ void __launch_bounds__(1024)
l1_ssim_finalize_kernel(int n, const float *__restrict__ partials, double inv_n_ssim, double inv_n_px,
                        double ssim_lambda, float *__restrict__ out) {
  __shared__ double s_a[1024], s_b[1024];
  double a = 0.0, b = 0.0;
  const float2 *p2 = reinterpret_cast<const float2 *>(partials);
#pragma unroll 4
  for (int i = threadIdx.x; i < n; i += 1024) {
    const float2 v = __ldg(p2 + i);
    a += (double)v.x;
    b += (double)v.y;
  }
  s_a[threadIdx.x] = a;
  s_b[threadIdx.x] = b;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s_a[threadIdx.x] += s_a[threadIdx.x + o];
      s_b[threadIdx.x] += s_b[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double ssim = s_a[0] * inv_n_ssim, l1 = s_b[0] * inv_n_px;
    out[0] = (float)((1.0 - ssim_lambda) * l1 + ssim_lambda * (1.0 - ssim));
    out[1] = (float)l1;
    out[2] = (float)ssim;
  }
}

}  // namespace gsr

extern "C" {

GSR_API int gsr_l1_ssim_num_partials(unsigned img_height, unsigned img_width) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_l1_ssim_num_partials");
  if (img_height < (unsigned)LW || img_width < (unsigned)LW) return 0;
  return (int)(cdiv(img_width - (LW - 1), LT) * cdiv(img_height - (LW - 1), LT));
}

GSR_API int gsr_l1_ssim_forward(unsigned img_height, unsigned img_width, float ssim_lambda, const float *pred,
                                const float *gt, float *maps, float *partials, float *loss_l1_ssim, void *stream) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_l1_ssim_forward");
  GSR_REQUIRE(img_height >= (unsigned)LW && img_width >= (unsigned)LW, GSR_ERR_INVALID_ARGUMENT,
              "l1_ssim_forward: image must be at least 11 x 11 (got %u x %u)", img_height, img_width);
  GSR_REQUIRE(pred && gt && maps && partials && loss_l1_ssim, GSR_ERR_INVALID_ARGUMENT, "l1_ssim_forward: null pointer");
  const dim3 grid(cdiv(img_width - (LW - 1), LT), cdiv(img_height - (LW - 1), LT), 1);
  l1_ssim_forward_kernel<<<grid, LOSS_THREADS, 0, (cudaStream_t)stream>>>((int)img_height, (int)img_width, pred, gt, maps,
                                                                          partials);
  GSR_CHECK_LAUNCH("l1_ssim_forward_kernel");
  const double n_px = 3.0 * img_height * img_width;
  const double n_ssim = 3.0 * (img_height - (LW - 1)) * (double)(img_width - (LW - 1));
  l1_ssim_finalize_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>((int)(grid.x * grid.y), partials, 1.0 / n_ssim, 1.0 / n_px,
                                                              (double)ssim_lambda, loss_l1_ssim);
  GSR_CHECK_LAUNCH("l1_ssim_finalize_kernel");
  return GSR_OK;
}

GSR_API int gsr_l1_ssim_backward(unsigned img_height, unsigned img_width, float ssim_lambda, const float *pred,
                                 const float *gt, const float *maps, const float *v_loss /*nullable: 1.0*/,
                                 float *v_pred, void *stream) {
  using namespace gsr;
  GSR_TRACE_SCOPE("gsr_l1_ssim_backward");
  GSR_REQUIRE(img_height >= (unsigned)LW && img_width >= (unsigned)LW, GSR_ERR_INVALID_ARGUMENT,
              "l1_ssim_backward: image must be at least 11 x 11 (got %u x %u)", img_height, img_width);
  GSR_REQUIRE(pred && gt && maps && v_pred, GSR_ERR_INVALID_ARGUMENT, "l1_ssim_backward: null pointer");
  const double n_px = 3.0 * img_height * img_width;
  const double n_ssim = 3.0 * (img_height - (LW - 1)) * (double)(img_width - (LW - 1));
  const dim3 grid(cdiv(img_width, LT), cdiv(img_height, LT), 1);
  l1_ssim_backward_kernel<<<grid, LOSS_THREADS, 0, (cudaStream_t)stream>>>(
      (int)img_height, (int)img_width, pred, gt, maps, (float)((1.0 - ssim_lambda) / n_px), (float)(-ssim_lambda / n_ssim),
      v_loss, v_pred);
  GSR_CHECK_LAUNCH("l1_ssim_backward_kernel");
  return GSR_OK;
}
}
