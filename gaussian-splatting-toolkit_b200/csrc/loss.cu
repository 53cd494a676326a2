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

  // load the patch (rows of 26 px x 3 ch = 78 contiguous floats) and accumulate the L1 of the pixels this tile owns
  float l1 = 0.f;
  for (int i = tid; i < LH * LH * 3; i += LOSS_THREADS) {
    const int r = i / (LH * 3), c = i - r * (LH * 3);
    const int y = oy0 + r, x = ox0 + c / 3;
    float p = 0.f, g = 0.f;
    if (y < H && x < W) {
      const size_t a = ((size_t)y * W + x) * 3 + (c % 3);
      p = pred[a];
      g = gt[a];
      const bool own_x = (c / 3 < LT) || last_x, own_y = (r < LT) || last_y;
      if (own_x && own_y) l1 += fabsf(p - g);
    }
    s_p[r][c] = p;
    s_g[r][c] = g;
  }
  __syncthreads();
  // horizontal pass: 26 rows x (16 px x 3 ch)
  for (int i = tid; i < LH * LT * 3; i += LOSS_THREADS) {
    const int r = i / (LT * 3), c = i - r * (LT * 3);
    float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
    for (int k = 0; k < LW; ++k) {
      const float w = kWin[k], p = s_p[r][c + 3 * k], g = s_g[r][c + 3 * k];
      m1 += w * p;
      m2 += w * g;
      e11 += w * p * p;
      e22 += w * g * g;
      e12 += w * p * g;
    }
    s_h[0][r][c] = m1;
    s_h[1][r][c] = m2;
    s_h[2][r][c] = e11;
    s_h[3][r][c] = e22;
    s_h[4][r][c] = e12;
  }
  __syncthreads();
  // vertical pass + SSIM + derivative maps
  const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
  float ssim_sum = 0.f;
  for (int i = tid; i < LT * LT * 3; i += LOSS_THREADS) {
    const int r = i / (LT * 3), c = i - r * (LT * 3);
    const int oy = oy0 + r, ox = ox0 + c / 3;
    if (oy >= Ho || ox >= Wo) continue;
    float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
    for (int k = 0; k < LW; ++k) {
      const float w = kWin[k];
      m1 += w * s_h[0][r + k][c];
      m2 += w * s_h[1][r + k][c];
      e11 += w * s_h[2][r + k][c];
      e22 += w * s_h[3][r + k][c];
      e12 += w * s_h[4][r + k][c];
    }
    const float s1 = e11 - m1 * m1, s2 = e22 - m2 * m2, s12 = e12 - m1 * m2;
    const float a1 = 2.f * m1 * m2 + C1, a2 = 2.f * s12 + C2;
    const float b1 = m1 * m1 + m2 * m2 + C1, b2 = s1 + s2 + C2;
    const float inv = 1.f / (b1 * b2);
    const float ssim = a1 * a2 * inv;
    ssim_sum += ssim;
    // d ssim / d(mu1, s1, s12) at fixed (mu2, s2)
    const float d_mu1 = (2.f * m2 * a2 * inv) - ssim * (2.f * m1 / b1);
    const float d_s1 = -ssim / b2;
    const float d_s12 = 2.f * a1 * inv;
    const size_t o = ((size_t)oy * Wo + ox) * 3 + (c % 3);
    const size_t plane = (size_t)Ho * Wo * 3;
    maps[o] = d_mu1 - 2.f * m1 * d_s1 - m2 * d_s12;
    maps[plane + o] = d_s1;
    maps[2 * plane + o] = d_s12;
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
  // outputs o in [x-10, x] contribute to input x: patch of the maps starting at (x0-10, y0-10), zero outside
  for (int i = tid; i < LH * LH * 3; i += LOSS_THREADS) {
    const int r = i / (LH * 3), c = i - r * (LH * 3);
    const int oy = y0 - (LW - 1) + r, ox = x0 - (LW - 1) + c / 3;
    float a = 0.f, b = 0.f, cc = 0.f;
    if (oy >= 0 && oy < Ho && ox >= 0 && ox < Wo) {
      const size_t o = ((size_t)oy * Wo + ox) * 3 + (c % 3);
      a = maps[o];
      b = maps[plane + o];
      cc = maps[2 * plane + o];
    }
    s_m[0][r][c] = a;
    s_m[1][r][c] = b;
    s_m[2][r][c] = cc;
  }
  __syncthreads();
  for (int i = tid; i < LH * LT * 3; i += LOSS_THREADS) {
    const int r = i / (LT * 3), c = i - r * (LT * 3);
    float a = 0.f, b = 0.f, cc = 0.f;
#pragma unroll
    for (int k = 0; k < LW; ++k) {
      const float w = kWin[LW - 1 - k];  // transposed window (symmetric, kept explicit)
      a += w * s_m[0][r][c + 3 * k];
      b += w * s_m[1][r][c + 3 * k];
      cc += w * s_m[2][r][c + 3 * k];
    }
    s_h[0][r][c] = a;
    s_h[1][r][c] = b;
    s_h[2][r][c] = cc;
  }
  __syncthreads();
  const float up = v_loss ? v_loss[0] : 1.f;
  for (int i = tid; i < LT * LT * 3; i += LOSS_THREADS) {
    const int r = i / (LT * 3), c = i - r * (LT * 3);
    const int y = y0 + r, x = x0 + c / 3;
    if (y >= H || x >= W) continue;
    float a = 0.f, b = 0.f, cc = 0.f;
#pragma unroll
    for (int k = 0; k < LW; ++k) {
      const float w = kWin[LW - 1 - k];
      a += w * s_h[0][r + k][c];
      b += w * s_h[1][r + k][c];
      cc += w * s_h[2][r + k][c];
    }
    const size_t idx = ((size_t)y * W + x) * 3 + (c % 3);
    const float p = pred[idx], g = gt[idx];
    const float d = p - g;
    const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    v_pred[idx] = up * (scale_l1 * sgn + scale_ssim * (a + 2.f * p * b + g * cc));
  }
}

}  // namespace gsr

extern "C" {

GSR_API int gsr_l1_ssim_num_partials(unsigned img_height, unsigned img_width) {
  using namespace gsr;
  if (img_height < (unsigned)LW || img_width < (unsigned)LW) return 0;
  return (int)(cdiv(img_width - (LW - 1), LT) * cdiv(img_height - (LW - 1), LT));
}

GSR_API int gsr_l1_ssim_forward(unsigned img_height, unsigned img_width, const float *pred, const float *gt,
                                float *maps, float *partials, void *stream) {
  using namespace gsr;
  GSR_REQUIRE(img_height >= (unsigned)LW && img_width >= (unsigned)LW, GSR_ERR_INVALID_ARGUMENT,
              "l1_ssim_forward: image must be at least 11 x 11 (got %u x %u)", img_height, img_width);
  GSR_REQUIRE(pred && gt && maps && partials, GSR_ERR_INVALID_ARGUMENT, "l1_ssim_forward: null pointer");
  const dim3 grid(cdiv(img_width - (LW - 1), LT), cdiv(img_height - (LW - 1), LT), 1);
  l1_ssim_forward_kernel<<<grid, LOSS_THREADS, 0, (cudaStream_t)stream>>>((int)img_height, (int)img_width, pred, gt, maps,
                                                                          partials);
  GSR_CHECK_LAUNCH("l1_ssim_forward_kernel");
  return GSR_OK;
}

GSR_API int gsr_l1_ssim_backward(unsigned img_height, unsigned img_width, float ssim_lambda, const float *pred,
                                 const float *gt, const float *maps, const float *v_loss /*nullable: 1.0*/,
                                 float *v_pred, void *stream) {
  using namespace gsr;
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
