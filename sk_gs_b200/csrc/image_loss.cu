// image_loss.cu - photometric loss of the training step, forward and backward in two launches.
//
//   loss = w_img * mean|I - G|  (or mean (I-G)^2)  +  w_ssim * (1 - mean SSIM(I, G))
//
// Reference: networks/losses/image_loss.py:8-33 (nn.L1Loss / nn.MSELoss, reduction 'mean'),
// networks/losses/ssim.py:9-18 (11-tap Gaussian window, sigma 1.5), :46-62 (_ssim: five zero-padded depthwise
// convolutions, C1 = 0.01^2, C2 = 0.03^2), :39-41 (1 - mean), weights exps/default.yaml:83-84 (0.8 / 0.2),
// call site networks/sk_gs.py:1524-1529.  The reference runs 5 conv2d + ~20 element-wise kernels forward and the
// autograd mirror backward; here
//   ssim_stats_kernel : window statistics (mu1, mu2, E[xx+yy], E[xy]) by a separable convolution in shared memory,
//                       SSIM map, both loss sums, and the three partial-derivative maps dS/dmu1, dS/dE[xx], dS/dE[xy];
//   ssim_grad_kernel  : the adjoint (same symmetric window) convolution of those maps -> dL/dI, and the loss scalars.
// The rendered image arrives channel-major [3,H,W] (what composite_fwd_kernel writes); the target may be channel-major
// or pixel-major [H,W,3|4] (datasets hand out HWC / RGBA, sk_gs.py:1525).
#include "common.cuh"

namespace skgs {
namespace {

constexpr int LT = 32;           // output tile edge
constexpr int LR = 5;            // window radius (window_size 11)
constexpr int LW = 2 * LR + 1;   // taps
constexpr int LH = LT + 2 * LR;  // tile + halo
constexpr int LS = 8;            // outputs per thread in the horizontal pass
constexpr int LV = 4;            // outputs per thread in the vertical pass
constexpr int L_THREADS = 256;
constexpr int L_LOADS = (LH * LH + L_THREADS - 1) / L_THREADS;  // tile + halo elements per thread
static_assert(LT * LT == L_THREADS * LV, "vertical pass covers the tile");
static_assert(LH * (LT / LS) <= L_THREADS, "horizontal pass fits one round");

struct Window {
  float w[LW];
};

__device__ __forceinline__ float target_at(const float* __restrict__ tgt, int pix_stride, int c, int y, int x, int H,
                                           int W) {
  return pix_stride ? tgt[((size_t)y * W + x) * pix_stride + c] : tgt[((size_t)c * H + y) * W + x];
}

// One thread: LS consecutive outputs of one halo row, NQ maps at once.  in[q] points at the first tap.
template <int NQ>
__device__ __forceinline__ void hpass(const Window& win, float (&v)[NQ][LS + LW - 1], float (&o)[NQ][LS]) {
#pragma unroll
  for (int j = 0; j < LS; ++j) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) o[q][j] = 0.f;
#pragma unroll
    for (int k = 0; k < LW; ++k) {
#pragma unroll
      for (int q = 0; q < NQ; ++q) o[q][j] = fmaf(win.w[k], v[q][j + k], o[q][j]);
    }
  }
}

__global__ void __launch_bounds__(L_THREADS)
    ssim_stats_kernel(int H, int W, const float* __restrict__ img, const float* __restrict__ tgt, int tgt_pix_stride,
                      int mse, Window win, float* __restrict__ dmaps, double* __restrict__ sums) {
  __shared__ float sX[LH][LH + 1];
  __shared__ float sY[LH][LH + 1];
  __shared__ float sHz[4][LH][LT + 1];
  __shared__ float sRed[2][L_THREADS / 32];
  const int tid = threadIdx.x;
  const int c = blockIdx.z;
  const int x0 = blockIdx.x * LT, y0 = blockIdx.y * LT;

  // all global loads of the thread are issued before the first shared-memory store (the tile load is pure latency)
  {
    float rx[L_LOADS], ry[L_LOADS];
#pragma unroll
    for (int it = 0; it < L_LOADS; ++it) {
      const int i = tid + it * L_THREADS;
      const int r = i / LH, col = i - r * LH;
      const int gy = y0 + r - LR, gx = x0 + col - LR;
      const bool in = i < LH * LH && gy >= 0 && gy < H && gx >= 0 && gx < W;  // zero padding (F.conv2d padding=5, ssim.py:47)
      rx[it] = in ? img[((size_t)c * H + gy) * W + gx] : 0.f;
      ry[it] = in ? target_at(tgt, tgt_pix_stride, c, gy, gx, H, W) : 0.f;
    }
#pragma unroll
    for (int it = 0; it < L_LOADS; ++it) {
      const int i = tid + it * L_THREADS;
      if (i < LH * LH) {
        const int r = i / LH, col = i - r * LH;
        sX[r][col] = rx[it];
        sY[r][col] = ry[it];
      }
    }
  }
  __syncthreads();

  if (tid < LH * (LT / LS)) {
    const int r = tid / (LT / LS), s = (tid % (LT / LS)) * LS;
    // four maps, not five: sigma1^2 and sigma2^2 only ever appear as their sum, so x^2 + y^2 is filtered once
    float v[4][LS + LW - 1], o[4][LS];
#pragma unroll
    for (int k = 0; k < LS + LW - 1; ++k) {
      const float a = sX[r][s + k], b = sY[r][s + k];
      v[0][k] = a;
      v[1][k] = b;
      v[2][k] = fmaf(a, a, b * b);
      v[3][k] = a * b;
    }
    hpass<4>(win, v, o);
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int j = 0; j < LS; ++j) sHz[q][r][s + j] = o[q][j];
  }
  __syncthreads();

  const int col = tid & 31, r0 = (tid >> 5) * LV;
  float st[4][LV];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float v[LV + LW - 1];
#pragma unroll
    for (int k = 0; k < LV + LW - 1; ++k) v[k] = sHz[q][r0 + k][col];
#pragma unroll
    for (int j = 0; j < LV; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < LW; ++k) acc = fmaf(win.w[k], v[j + k], acc);
      st[q][j] = acc;
    }
  }

  const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;  // ssim.py:58-59
  float sum_pix = 0.f, sum_ssim = 0.f;
  const int gx = x0 + col;
#pragma unroll
  for (int j = 0; j < LV; ++j) {
    const int gy = y0 + r0 + j;
    if (gy < H && gx < W) {
      const float mu1 = st[0][j], mu2 = st[1][j];
      const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
      const float s12 = st[3][j] - mu12;
      const float A1 = 2.f * mu12 + C1, A2 = 2.f * s12 + C2;
      const float B1 = mu1_sq + mu2_sq + C1, B2 = (st[2][j] - mu1_sq - mu2_sq) + C2;  // sigma1^2 + sigma2^2 + C2
      const float iB1 = 1.f / B1, iB2 = 1.f / B2;
      const float S = A1 * A2 * iB1 * iB2;  // ssim.py:61
      sum_ssim += S;
      const float d = sX[r0 + j + LR][col + LR] - sY[r0 + j + LR][col + LR];
      sum_pix += mse ? d * d : fabsf(d);
      if (dmaps) {
        // partial derivatives of S w.r.t. the window statistics of the rendered image (mu1, E[xx], E[xy])
        const size_t n = (size_t)3 * H * W, at = ((size_t)c * H + gy) * W + gx;
        dmaps[at] = 2.f * mu2 * (A2 - A1) * iB1 * iB2 - 2.f * mu1 * S * (iB1 - iB2);
        dmaps[n + at] = -S * iB2;
        dmaps[2 * n + at] = 2.f * A1 * iB1 * iB2;
      }
    }
  }
  sum_pix = warp_sum(sum_pix);
  sum_ssim = warp_sum(sum_ssim);
  if ((tid & 31) == 0) {
    sRed[0][tid >> 5] = sum_pix;
    sRed[1][tid >> 5] = sum_ssim;
  }
  __syncthreads();
  if (tid < 2) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < L_THREADS / 32; ++w) t += (double)sRed[tid][w];
    atomicAdd(&sums[tid], t);
  }
}

__device__ __forceinline__ void write_terms(const double* sums, double n, float w_img, float w_ssim, float* terms) {
  const double pix = sums[0] / n, ssim = 1.0 - sums[1] / n;
  terms[0] = (float)pix;
  terms[1] = (float)ssim;
  terms[2] = (float)((double)w_img * pix + (double)w_ssim * ssim);
}

__global__ void loss_terms_kernel(const double* __restrict__ sums, double n, float w_img, float w_ssim,
                                  float* __restrict__ terms) {
  write_terms(sums, n, w_img, w_ssim, terms);
}

__global__ void __launch_bounds__(L_THREADS)
    ssim_grad_kernel(int H, int W, const float* __restrict__ img, const float* __restrict__ tgt, int tgt_pix_stride,
                     int mse, Window win, const float* __restrict__ dmaps, const double* __restrict__ sums, float w_img,
                     float w_ssim, float grad_scale, float* __restrict__ dL_dimg, float* __restrict__ terms) {
  __shared__ float sD[3][LH][LH + 1];
  __shared__ float sHz[3][LH][LT + 1];
  const int tid = threadIdx.x;
  const int c = blockIdx.z;
  const int x0 = blockIdx.x * LT, y0 = blockIdx.y * LT;
  const size_t n = (size_t)3 * H * W;
  if (blockIdx.x == 0 && blockIdx.y == 0 && c == 0 && tid == 0) write_terms(sums, (double)n, w_img, w_ssim, terms);

  {
    float rd[3][L_LOADS];
#pragma unroll
    for (int it = 0; it < L_LOADS; ++it) {
      const int i = tid + it * L_THREADS;
      const int r = i / LH, col = i - r * LH;
      const int gy = y0 + r - LR, gx = x0 + col - LR;
      const bool in = i < LH * LH && gy >= 0 && gy < H && gx >= 0 && gx < W;  // the adjoint only sums over pixels that exist
      const size_t at = in ? ((size_t)c * H + gy) * W + gx : 0;
#pragma unroll
      for (int q = 0; q < 3; ++q) rd[q][it] = in ? dmaps[q * n + at] : 0.f;
    }
#pragma unroll
    for (int it = 0; it < L_LOADS; ++it) {
      const int i = tid + it * L_THREADS;
      if (i < LH * LH) {
        const int r = i / LH, col = i - r * LH;
#pragma unroll
        for (int q = 0; q < 3; ++q) sD[q][r][col] = rd[q][it];
      }
    }
  }
  __syncthreads();

  if (tid < LH * (LT / LS)) {
    const int r = tid / (LT / LS), s = (tid % (LT / LS)) * LS;
    float v[3][LS + LW - 1], o[3][LS];
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
      for (int k = 0; k < LS + LW - 1; ++k) v[q][k] = sD[q][r][s + k];
    hpass<3>(win, v, o);
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
      for (int j = 0; j < LS; ++j) sHz[q][r][s + j] = o[q][j];
  }
  __syncthreads();

  const int col = tid & 31, r0 = (tid >> 5) * LV;
  float g[3][LV];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    float v[LV + LW - 1];
#pragma unroll
    for (int k = 0; k < LV + LW - 1; ++k) v[k] = sHz[q][r0 + k][col];
#pragma unroll
    for (int j = 0; j < LV; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < LW; ++k) acc = fmaf(win.w[k], v[j + k], acc);
      g[q][j] = acc;
    }
  }

  const float inv_n = (float)(1.0 / (double)n);
  const float k_img = grad_scale * w_img * inv_n, k_ssim = -grad_scale * w_ssim * inv_n;
  const int gx = x0 + col;
#pragma unroll
  for (int j = 0; j < LV; ++j) {
    const int gy = y0 + r0 + j;
    if (gy < H && gx < W) {
      const size_t at = ((size_t)c * H + gy) * W + gx;
      const float x = img[at], y = target_at(tgt, tgt_pix_stride, c, gy, gx, H, W);
      const float d = x - y;
      const float dpix = mse ? 2.f * d : (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
      const float dS = g[0][j] + 2.f * x * g[1][j] + y * g[2][j];
      dL_dimg[at] = k_img * dpix + k_ssim * dS;
    }
  }
}

Window make_window() {
  // ssim.py:9-11: exp(-(x - 5)^2 / (2 * 1.5^2)) rounded to fp32, divided by their fp32 sum
  Window w;
  float s = 0.f;
  for (int i = 0; i < LW; ++i) {
    w.w[i] = (float)exp(-(double)((i - LR) * (i - LR)) / (2.0 * 1.5 * 1.5));
    s += w.w[i];
  }
  for (int i = 0; i < LW; ++i) w.w[i] /= s;
  return w;
}

}  // namespace
}  // namespace skgs

using namespace skgs;

extern "C" {

size_t skgs_image_loss_workspace_bytes(int32_t H, int32_t W) {
  if (H <= 0 || W <= 0) return 256;
  return 256 + align_up((size_t)9 * H * W * sizeof(float), 256);
}

int skgs_image_loss(int32_t H, int32_t W, const float* image, const float* target, int32_t target_pixel_stride,
                    int32_t method, float w_image, float w_ssim, float grad_scale, void* workspace, float* loss_terms,
                    float* dL_dimage, void* stream) {
  SKGS_CHECK_ARG(H > 0 && W > 0, "image_loss: empty image %dx%d", H, W);
  SKGS_CHECK_ARG(image && target && workspace && loss_terms, "image_loss: null pointer");
  SKGS_CHECK_ARG(target_pixel_stride == 0 || target_pixel_stride == 3 || target_pixel_stride == 4,
                 "image_loss: target_pixel_stride must be 0 (channel-major), 3 or 4, got %d", target_pixel_stride);
  SKGS_CHECK_ARG(method == 0 || method == 1, "image_loss: method must be 0 (l1) or 1 (mse), got %d", method);
  cudaStream_t st = (cudaStream_t)stream;
  static const Window win = make_window();
  double* sums = (double*)workspace;
  float* dmaps = dL_dimage ? (float*)((char*)workspace + 256) : nullptr;
  SKGS_CUDA(cudaMemsetAsync(sums, 0, 2 * sizeof(double), st));
  const dim3 grid((W + LT - 1) / LT, (H + LT - 1) / LT, 3);
  {
    ProfScope prof_("ssim_stats_kernel", st);
    ssim_stats_kernel<<<grid, L_THREADS, 0, st>>>(H, W, image, target, target_pixel_stride, method, win, dmaps, sums);
  }
  SKGS_CHECK_LAUNCH("ssim_stats_kernel");
  if (dL_dimage) {
    ProfScope prof_("ssim_grad_kernel", st);
    ssim_grad_kernel<<<grid, L_THREADS, 0, st>>>(H, W, image, target, target_pixel_stride, method, win, dmaps, sums,
                                                 w_image, w_ssim, grad_scale, dL_dimage, loss_terms);
  } else {
    loss_terms_kernel<<<1, 1, 0, st>>>(sums, 3.0 * H * W, w_image, w_ssim, loss_terms);
  }
  SKGS_CHECK_LAUNCH("ssim_grad_kernel");
  return SKGS_OK;
}

}  // extern "C"
