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
constexpr int AD_ILP = 4;
constexpr int AD_CHUNK = AD_THREADS * AD_VEC * AD_ILP;  // parameters per CTA

struct AdamArgs {
  skgs_adam_tensor t[SKGS_ADAM_MAX_TENSORS];
  float step_size[SKGS_ADAM_MAX_TENSORS];   // lr / (1 - beta1^step)
  float step_size2[SKGS_ADAM_MAX_TENSORS];  // lr2 / (1 - beta1^step)
  int block_start[SKGS_ADAM_MAX_TENSORS + 1];
  int slot[SKGS_ADAM_MAX_TENSORS];         // index in the caller's table (empty tensors are dropped)
  const float* dyn;                        // optional device table {bc2_sqrt, (step_size, step_size2)[...]}
  int count;
  float w1;         // 1 - beta1
  float beta2;
  float w2;         // 1 - beta2
  float bc2_sqrt;   // sqrt(1 - beta2^step)
  float eps;
  float grad_scale;
};

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, const AdamArgs& a, float step_size,
                                            float bc2_sqrt) {
  // same operation order as torch's foreach implementation (lerp, mul + addcmul, sqrt / div / add, addcdiv)
  m = m + a.w1 * (g - m);
  v = v * a.beta2 + (a.w2 * g) * g;
  const float denom = sqrtf(v) / bc2_sqrt + a.eps;
  p = p + (-step_size) * (m / denom);
}

__global__ void __launch_bounds__(AD_THREADS) adam_kernel(const __grid_constant__ AdamArgs a) {
  int ti = 0;
  while (ti + 1 < a.count && (int)blockIdx.x >= a.block_start[ti + 1]) ++ti;
  const skgs_adam_tensor& T = a.t[ti];
  const float step_size = a.dyn ? __ldg(a.dyn + 1 + 2 * a.slot[ti]) : a.step_size[ti];
  const float step_size2 = a.dyn ? __ldg(a.dyn + 2 + 2 * a.slot[ti]) : a.step_size2[ti];
  const int period = T.period, split = T.split;
  auto ss = [&](int64_t i) { return period > 0 && (int)(i % period) >= split ? step_size2 : step_size; };
  const float bc2_sqrt = a.dyn ? __ldg(a.dyn) : a.bc2_sqrt;
  const int64_t base = (int64_t)(blockIdx.x - a.block_start[ti]) * AD_CHUNK;
  const int64_t end = min(T.numel, base + AD_CHUNK);

  if (T.knn_indices) {
    for (int64_t i = base + threadIdx.x; i < end; i += AD_THREADS) {
      const int64_t row = i / T.cols;
      const int col = (int)(i - row * T.cols);
      float g = 0.f;
      for (int k = 0; k < T.K; ++k)
        if ((int)T.knn_indices[row * T.K + k] == col) g += T.grad[row * T.K + k];
      g *= a.grad_scale;
      float p = T.param[i], m = T.exp_avg[i], v = T.exp_avg_sq[i];
      adam_update(p, g, m, v, a, ss(i), bc2_sqrt);
      T.param[i] = p;
      T.exp_avg[i] = m;
      T.exp_avg_sq[i] = v;
    }
    return;
  }

  const bool vec_ok =
      ((((uintptr_t)T.param) | ((uintptr_t)T.grad) | ((uintptr_t)T.exp_avg) | ((uintptr_t)T.exp_avg_sq)) & 15) == 0;
  if (vec_ok && end - base == AD_CHUNK) {
    float4 p[AD_ILP], g[AD_ILP], m[AD_ILP], v[AD_ILP];
#pragma unroll
    for (int u = 0; u < AD_ILP; ++u) {
      const int64_t i = base + ((int64_t)u * AD_THREADS + threadIdx.x) * AD_VEC;
      p[u] = *(const float4*)(T.param + i);
      g[u] = __ldg((const float4*)(T.grad + i));
      m[u] = *(const float4*)(T.exp_avg + i);
      v[u] = *(const float4*)(T.exp_avg_sq + i);
    }
#pragma unroll
    for (int u = 0; u < AD_ILP; ++u) {
      const int64_t i = base + ((int64_t)u * AD_THREADS + threadIdx.x) * AD_VEC;
      adam_update(p[u].x, g[u].x * a.grad_scale, m[u].x, v[u].x, a, ss(i), bc2_sqrt);
      adam_update(p[u].y, g[u].y * a.grad_scale, m[u].y, v[u].y, a, ss(i + 1), bc2_sqrt);
      adam_update(p[u].z, g[u].z * a.grad_scale, m[u].z, v[u].z, a, ss(i + 2), bc2_sqrt);
      adam_update(p[u].w, g[u].w * a.grad_scale, m[u].w, v[u].w, a, ss(i + 3), bc2_sqrt);
    }
#pragma unroll
    for (int u = 0; u < AD_ILP; ++u) {
      const int64_t i = base + ((int64_t)u * AD_THREADS + threadIdx.x) * AD_VEC;
      *(float4*)(T.param + i) = p[u];
      *(float4*)(T.exp_avg + i) = m[u];
      *(float4*)(T.exp_avg_sq + i) = v[u];
    }
    return;
  }
  for (int64_t i = base + threadIdx.x; i < end; i += AD_THREADS) {
    float p = T.param[i], m = T.exp_avg[i], v = T.exp_avg_sq[i];
    adam_update(p, T.grad[i] * a.grad_scale, m, v, a, ss(i), bc2_sqrt);
    T.param[i] = p;
    T.exp_avg[i] = m;
    T.exp_avg_sq[i] = v;
  }
}

}  // namespace
}  // namespace skgs

using namespace skgs;

extern "C" int skgs_adam_step(const skgs_adam_tensor* tensors, int32_t count, int32_t step, double beta1,
                              double beta2, double eps, float grad_scale, const float* dynamic_hyper,
                              void* stream) {
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
  cudaStream_t st = (cudaStream_t)stream;
  {
    ProfScope prof_("adam_kernel", st);
    adam_kernel<<<a.block_start[a.count], AD_THREADS, 0, st>>>(a);
  }
  SKGS_CHECK_LAUNCH("adam_kernel");
  return SKGS_OK;
}
