// adam.cu - one-launch Adam over all parameter tensors of the path.
//
// Reference: torch.optim.Adam as configured by networks/gaussian_splatting.py:445-453 (one param group per Gaussian
// attribute, each with its own lr) and exps/default.yaml:121-125 (eps 1e-15, betas (0.9, 0.999), no weight decay).
// torch runs ~8 foreach kernels per step over the parameter list; here every tensor is a segment of one grid
// (multi-tensor apply): 28 B of HBM traffic per parameter (p, g, m, v in; p, m, v out), nothing else.
// The skinning-weight table sp_W [P,M] receives only K gradients per row (sk_gs.py:767-768): its segment reads the
// compact [P,K] gradient + KNN indices instead of a dense, mostly zero [P,M] gradient (24 instead of 28 B/parameter and
// no scatter kernel), while moments and parameters are still updated densely like torch does.
#include "common.cuh"

namespace skgs {
namespace {

constexpr int AD_THREADS = 256;
constexpr int AD_VEC = 4;
#ifndef SKGS_AD_ILP
#define SKGS_AD_ILP 2
#endif
#ifndef SKGS_AD_MINB
#define SKGS_AD_MINB 4
#endif
constexpr int AD_ILP = SKGS_AD_ILP;
constexpr int AD_CHUNK = AD_THREADS * AD_VEC * AD_ILP;  // parameters per CTA

struct AdamArgs {
  skgs_adam_tensor t[SKGS_ADAM_MAX_TENSORS];
  float step_size[SKGS_ADAM_MAX_TENSORS];   // lr / (1 - beta1^step)
  float step_size2[SKGS_ADAM_MAX_TENSORS];  // lr2 / (1 - beta1^step)
  int block_start[SKGS_ADAM_MAX_TENSORS + 1];
  int slot[SKGS_ADAM_MAX_TENSORS];         // index in the caller's table (empty tensors are dropped)
  const float* dyn;                        // optional device table {bc2_sqrt, (step_size, step_size2)[...]}
  const uint32_t* skip;                    // optional device flag: non-zero -> the whole step is a no-op
  int count;
  float w1;         // 1 - beta1
  float beta2;
  float w2;         // 1 - beta2
  float bc2_sqrt;   // sqrt(1 - beta2^step)
  float eps;
  float grad_scale;
};

struct Hyper {
  float w1, beta2, w2, inv_bc2_sqrt, eps, step_size, step_size2;
  uint32_t period, split;
  __device__ __forceinline__ float ss(uint32_t phase) const {  // phase = element index mod period
    return period != 0 && phase >= split ? step_size2 : step_size;
  }
};

__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, const Hyper& h, float step_size) {
  // moments: same operation order as torch's foreach implementation (lerp, mul + addcmul).  The parameter update uses
  // the 1-ulp hardware sqrt and the 2-ulp fast division instead of the IEEE sequences: entries that never receive a
  // gradient (v == 0: most of the skinning table) would otherwise run the slow-path subroutines on every step and make
  // the kernel instruction-bound; the effect on p is <= 3e-7 of one update, far inside the 2e-6 test tolerance.
  m = m + h.w1 * (g - m);
  v = v * h.beta2 + (h.w2 * g) * g;
  const float denom = sqrt_approx(v) * h.inv_bc2_sqrt + h.eps;
  p = p + (-step_size) * __fdividef(m, denom);
}

// gradient of element (row, col) of the skinning table from its compact [rows, K] form
__device__ __forceinline__ float knn_grad(const skgs_adam_tensor& T, int64_t row, int col) {
  const int64_t* __restrict__ idx = T.knn_indices + row * T.K;
  const float* __restrict__ gk = T.grad + row * T.K;
  float g = 0.f;
  for (int k = 0; k < T.K; ++k) g += ((int)__ldg(idx + k) == col) ? __ldg(gk + k) : 0.f;
  return g;
}

// All loads of a thread's AD_ILP x 4 parameters are issued before the first dependent instruction: the kernel is a pure
// stream (28 B per parameter) and lives on memory-level parallelism.
template <bool KNN>
__device__ __forceinline__ void adam_chunk(const skgs_adam_tensor& T, const Hyper& h, float grad_scale, int64_t base,
                                           int64_t end) {
  const bool vec_ok = ((((uintptr_t)T.param) | ((uintptr_t)T.exp_avg) | ((uintptr_t)T.exp_avg_sq) |
                        (KNN ? (uintptr_t)0 : (uintptr_t)T.grad)) & 15) == 0;
  if (vec_ok && end - base == AD_CHUNK && (!KNN || (T.cols & 3) == 0)) {
    float4 p[AD_ILP], g[AD_ILP], m[AD_ILP], v[AD_ILP];
    uint32_t phase[AD_ILP];
#pragma unroll
    for (int u = 0; u < AD_ILP; ++u) {
      const int64_t i = base + (u * AD_THREADS + threadIdx.x) * AD_VEC;
      p[u] = *(const float4*)(T.param + i);
      m[u] = *(const float4*)(T.exp_avg + i);
      v[u] = *(const float4*)(T.exp_avg_sq + i);
      if (!KNN) g[u] = __ldg((const float4*)(T.grad + i));
    }
#pragma unroll
    for (int u = 0; u < AD_ILP; ++u) {
      const int64_t i = base + (u * AD_THREADS + threadIdx.x) * AD_VEC;
      phase[u] = h.period ? (uint32_t)(i % h.period) : 0u;
      if (KNN) {
        // cols % 4 == 0 here (checked by the caller of this path): the four parameters share one row, so the row's K
        // (index, gradient) pairs are read once and routed to the lane they belong to
        const int64_t row = (T.numel >> 31) == 0 ? (int64_t)((uint32_t)i / (uint32_t)T.cols) : i / T.cols;
        const int col = (int)(i - row * T.cols);
        const int* __restrict__ idx = (const int*)(T.knn_indices + row * T.K);  // low words (little endian, < 2^31)
        const float* __restrict__ gk = T.grad + row * T.K;
        float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
        for (int k = 0; k < T.K; ++k) {
          const int d = __ldg(idx + 2 * k) - col;
          const float gv = __ldg(gk + k);
          g0 += d == 0 ? gv : 0.f;
          g1 += d == 1 ? gv : 0.f;
          g2 += d == 2 ? gv : 0.f;
          g3 += d == 3 ? gv : 0.f;
        }
        g[u] = make_float4(g0, g1, g2, g3);
      }
    }
#pragma unroll
    for (int u = 0; u < AD_ILP; ++u) {
      const uint32_t ph = phase[u], per = h.period;
      auto wrap = [&](uint32_t x) {
        while (per != 0 && x >= per) x -= per;
        return x;
      };
      adam_update(p[u].x, g[u].x * grad_scale, m[u].x, v[u].x, h, h.ss(ph));
      adam_update(p[u].y, g[u].y * grad_scale, m[u].y, v[u].y, h, h.ss(wrap(ph + 1)));
      adam_update(p[u].z, g[u].z * grad_scale, m[u].z, v[u].z, h, h.ss(wrap(ph + 2)));
      adam_update(p[u].w, g[u].w * grad_scale, m[u].w, v[u].w, h, h.ss(wrap(ph + 3)));
    }
#pragma unroll
    for (int u = 0; u < AD_ILP; ++u) {
      const int64_t i = base + (u * AD_THREADS + threadIdx.x) * AD_VEC;
      *(float4*)(T.param + i) = p[u];
      *(float4*)(T.exp_avg + i) = m[u];
      *(float4*)(T.exp_avg_sq + i) = v[u];
    }
    return;
  }
  for (int64_t i = base + threadIdx.x; i < end; i += AD_THREADS) {
    float g;
    if (KNN) {
      const int64_t row = i / T.cols;
      g = knn_grad(T, row, (int)(i - row * T.cols));
    } else {
      g = T.grad[i];
    }
    float p = T.param[i], m = T.exp_avg[i], v = T.exp_avg_sq[i];
    adam_update(p, g * grad_scale, m, v, h, h.ss(h.period ? (uint32_t)(i % h.period) : 0u));
    T.param[i] = p;
    T.exp_avg[i] = m;
    T.exp_avg_sq[i] = v;
  }
}

__global__ void __launch_bounds__(AD_THREADS, SKGS_AD_MINB) adam_kernel(const __grid_constant__ AdamArgs a) {
  // the gradients of a render whose binning arena overflowed are garbage: leave parameters and moments untouched
  if (a.skip != nullptr && __ldg(a.skip) != 0u) return;
  int ti = 0;
  while (ti + 1 < a.count && (int)blockIdx.x >= a.block_start[ti + 1]) ++ti;
  const skgs_adam_tensor& T = a.t[ti];
  Hyper h;
  h.w1 = a.w1;
  h.beta2 = a.beta2;
  h.w2 = a.w2;
  h.eps = a.eps;
  h.inv_bc2_sqrt = 1.f / (a.dyn ? __ldg(a.dyn) : a.bc2_sqrt);
  h.step_size = a.dyn ? __ldg(a.dyn + 1 + 2 * a.slot[ti]) : a.step_size[ti];
  h.step_size2 = a.dyn ? __ldg(a.dyn + 2 + 2 * a.slot[ti]) : a.step_size2[ti];
  h.period = (uint32_t)T.period;
  h.split = (uint32_t)T.split;
  const int64_t base = (int64_t)(blockIdx.x - a.block_start[ti]) * AD_CHUNK;
  const int64_t end = min(T.numel, base + AD_CHUNK);
  if (T.knn_indices)
    adam_chunk<true>(T, h, a.grad_scale, base, end);
  else
    adam_chunk<false>(T, h, a.grad_scale, base, end);
}

}  // namespace
}  // namespace skgs

using namespace skgs;

extern "C" int skgs_adam_step(const skgs_adam_tensor* tensors, int32_t count, int32_t step, double beta1,
                              double beta2, double eps, float grad_scale, const float* dynamic_hyper,
                              const uint32_t* skip_if_nonzero, void* stream) {
  SKGS_CHECK_ARG(count >= 0 && count <= SKGS_ADAM_MAX_TENSORS, "adam_step: count %d outside [0, %d]", count,
                 SKGS_ADAM_MAX_TENSORS);
  SKGS_CHECK_ARG(count == 0 || tensors, "adam_step: null tensor table");
  SKGS_CHECK_ARG(step >= 1, "adam_step: step must be >= 1 (torch increments before the first update), got %d", step);
  SKGS_CHECK_ARG(beta1 >= 0. && beta1 < 1. && beta2 >= 0. && beta2 < 1. && eps >= 0.,
                 "adam_step: invalid hyper-parameters beta1 %g beta2 %g eps %g", beta1, beta2, eps);
  AdamArgs a;
  a.count = 0;
  a.block_start[0] = 0;
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  for (int i = 0; i < count; ++i) {
    const skgs_adam_tensor& t = tensors[i];
    SKGS_CHECK_ARG(t.numel >= 0, "adam_step: tensor %d has negative numel", i);
    if (t.numel == 0) continue;
    SKGS_CHECK_ARG(t.param && t.grad && t.exp_avg && t.exp_avg_sq, "adam_step: tensor %d has a null pointer", i);
    if (t.knn_indices)
      SKGS_CHECK_ARG(t.cols > 0 && t.K > 0 && t.K <= t.cols && t.numel % t.cols == 0,
                     "adam_step: tensor %d compact gradient needs 0 < K <= cols and numel %% cols == 0", i);
    const int64_t blocks = (t.numel + AD_CHUNK - 1) / AD_CHUNK;
    SKGS_CHECK_ARG(a.block_start[a.count] + blocks < (int64_t)1 << 31, "adam_step: too many parameters");
    a.t[a.count] = t;
    a.slot[a.count] = i;
    SKGS_CHECK_ARG(t.period >= 0 && (t.period == 0 || (t.split > 0 && t.split < t.period)),
                   "adam_step: tensor %d needs 0 < split < period", i);
    a.step_size[a.count] = (float)(t.lr / bc1);
    a.step_size2[a.count] = (float)(t.lr2 / bc1);
    a.block_start[a.count + 1] = a.block_start[a.count] + (int)blocks;
    ++a.count;
  }
  if (a.count == 0) return SKGS_OK;
  a.w1 = (float)(1.0 - beta1);
  a.beta2 = (float)beta2;
  a.w2 = (float)(1.0 - beta2);
  a.bc2_sqrt = (float)sqrt(bc2);
  a.eps = (float)eps;
  a.grad_scale = grad_scale;
  a.dyn = dynamic_hyper;
  a.skip = skip_if_nonzero;
  cudaStream_t st = (cudaStream_t)stream;
  {
    ProfScope prof_("adam_kernel", st);
    adam_kernel<<<a.block_start[a.count], AD_THREADS, 0, st>>>(a);
  }
  SKGS_CHECK_LAUNCH("adam_kernel");
  return SKGS_OK;
}
