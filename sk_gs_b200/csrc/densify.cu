// densify.cu - densification bookkeeping of the Gaussian set: what makes the captured step a
// training loop.  Reference: networks/gaussian_splatting.py
//   :503-513  add_densification_stats  (per step: |viewspace grad|, visibility count; :672-674 max radii)
//   :515-563  change_optimizer         (parameters + Adam moments re-created by cat / boolean indexing, per tensor)
//   :565-587  prune_points / densification_postfix
//   :589-651  densify_and_split / densify_and_clone / densify
//   :653-660  prune            :662-665 reset_opacity          :667-703 adaptive_control (the schedule; host side)
// The reference runs clone, split, the pruning of the split originals and the opacity/size pruning as FOUR passes, each
// re-allocating every parameter and both Adam moments through torch.cat / mask indexing (~12 full copies of the state).
// Here one PLAN pass decides, per existing Gaussian, what it turns into (kept / clone / 2 split samples, each possibly
// pruned right away) and an exclusive scan gives every survivor its slot in the reference's final order
//   [ kept originals (index order) | clones (index order) | split samples n = 0 | split samples n = 1 ];
// one APPLY pass then gathers all parameter tensors and their moments into the new arrays: every byte is read once and
// written once.
#include "common.cuh"

namespace skgs {
namespace {

constexpr int DN_THREADS = 256;
constexpr int DN_ITEMS = 4;
constexpr int DN_TILE = DN_THREADS * DN_ITEMS;

struct Fate {      // what Gaussian i contributes to the new set
  uint32_t keep;   // 1: the original survives
  uint32_t clone;  // 1: a clone is appended
  uint32_t split;  // 1: two samples are appended (the original is dropped, :650-651)
  uint32_t sel;    // 1: selected for splitting (before pruning; indexes the noise rows)
};

__device__ __forceinline__ float sigmoid_ref(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ Fate decide(const skgs_densify_config& c, int i, const float* __restrict__ scaling,
                                       const float* __restrict__ opacity, const float* __restrict__ grad_accum,
                                       const float* __restrict__ denom, const float* __restrict__ max_radii2D) {
  const float s0 = expf(scaling[3 * (size_t)i]), s1 = expf(scaling[3 * (size_t)i + 1]),
              s2 = expf(scaling[3 * (size_t)i + 2]);
  const float smax = fmaxf(s0, fmaxf(s1, s2));
  bool clone = false, split = false;
  if (c.do_densify) {
    float g = grad_accum[i] / denom[i];  // :641-642
    if (g != g) g = 0.f;
    const bool hot = g >= c.grad_threshold;
    clone = hot && smax <= c.densify_extent;  // :626-630
    split = hot && smax > c.densify_extent;   // :595-599
  }
  bool prune_orig = false, prune_split = false;
  if (c.do_prune) {
    const bool low = sigmoid_ref(opacity[i]) < c.min_opacity;  // :654
    prune_orig = prune_split = low;
    if (c.max_screen_size > 0.f) {  // :655-658; densification_postfix zeroed max_radii2D if it ran (:586)
      const float r = c.do_densify ? 0.f : max_radii2D[i];
      prune_orig = prune_orig || r > c.max_screen_size || smax > c.prune_extent;
      // the samples carry log(exp(s) / (0.8 N)) (:611): evaluate the activation of that value like the reference does
      const float d = 0.8f * 2.0f;
      const float t0 = expf(logf(s0 / d)), t1 = expf(logf(s1 / d)), t2 = expf(logf(s2 / d));
      prune_split = prune_split || fmaxf(t0, fmaxf(t1, t2)) > c.prune_extent;
    }
  }
  Fate f;
  f.sel = split ? 1u : 0u;
  f.keep = (!split && !prune_orig) ? 1u : 0u;
  f.clone = (clone && !prune_orig) ? 1u : 0u;
  f.split = (split && !prune_split) ? 1u : 0u;
  return f;
}

// pass 1: per-tile counts of the four categories
__global__ void __launch_bounds__(DN_THREADS)
densify_count_kernel(skgs_densify_config c, int P, const float* __restrict__ scaling, const float* __restrict__ opacity,
                     const float* __restrict__ grad_accum, const float* __restrict__ denom,
                     const float* __restrict__ max_radii2D, uint4* __restrict__ tile_counts) {
  __shared__ uint32_t s_cnt[4];
  if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  uint32_t k = 0, cl = 0, sp = 0, se = 0;
  const int base = blockIdx.x * DN_TILE;
#pragma unroll
  for (int u = 0; u < DN_ITEMS; u++) {
    const int i = base + u * DN_THREADS + threadIdx.x;
    if (i < P) {
      const Fate f = decide(c, i, scaling, opacity, grad_accum, denom, max_radii2D);
      k += f.keep; cl += f.clone; sp += f.split; se += f.sel;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    k += __shfl_xor_sync(0xffffffffu, k, o);
    cl += __shfl_xor_sync(0xffffffffu, cl, o);
    sp += __shfl_xor_sync(0xffffffffu, sp, o);
    se += __shfl_xor_sync(0xffffffffu, se, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&s_cnt[0], k); atomicAdd(&s_cnt[1], cl); atomicAdd(&s_cnt[2], sp); atomicAdd(&s_cnt[3], se);
  }
  __syncthreads();
  if (threadIdx.x == 0) tile_counts[blockIdx.x] = make_uint4(s_cnt[0], s_cnt[1], s_cnt[2], s_cnt[3]);
}

// pass 2: every tile sums the counts of the tiles before it (a few hundred uint4), ranks its own Gaussians in index
// order and writes the gather plan: src[d] = source Gaussian of slot d, kind[d] = 0 kept / 1 clone / 2,3 split sample
// n = 0,1, noise_row[d] = row of the standard-normal table the sample uses (n * n_selected + rank among the selected,
// the layout of `samples` at :601-603)
__global__ void __launch_bounds__(DN_THREADS)
densify_plan_kernel(skgs_densify_config c, int P, int tiles, const float* __restrict__ scaling,
                    const float* __restrict__ opacity, const float* __restrict__ grad_accum,
                    const float* __restrict__ denom, const float* __restrict__ max_radii2D,
                    const uint4* __restrict__ tile_counts, int32_t* __restrict__ src, uint8_t* __restrict__ kind,
                    int32_t* __restrict__ noise_row, skgs_densify_counts* __restrict__ counts) {
  __shared__ uint32_t s_tot[4], s_before[4];
  __shared__ uint32_t s_warp[DN_THREADS / 32][4];
  if (threadIdx.x < 4) s_tot[threadIdx.x] = s_before[threadIdx.x] = 0;
  __syncthreads();
  {
    uint32_t t[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
    for (int j = threadIdx.x; j < tiles; j += DN_THREADS) {
      const uint4 v = tile_counts[j];
      t[0] += v.x; t[1] += v.y; t[2] += v.z; t[3] += v.w;
      if (j < (int)blockIdx.x) { b[0] += v.x; b[1] += v.y; b[2] += v.z; b[3] += v.w; }
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        t[q] += __shfl_xor_sync(0xffffffffu, t[q], o);
        b[q] += __shfl_xor_sync(0xffffffffu, b[q], o);
      }
      if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_tot[q], t[q]);
        atomicAdd(&s_before[q], b[q]);
      }
    }
  }
  __syncthreads();
  const uint32_t n_keep = s_tot[0], n_clone = s_tot[1], n_split = s_tot[2], n_sel = s_tot[3];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    counts->n_keep = n_keep; counts->n_clone = n_clone; counts->n_split = n_split; counts->n_selected = n_sel;
    counts->n_new = n_keep + n_clone + 2u * n_split;
  }
  // running offsets of this tile; items are taken in index order: thread-major blocks of DN_ITEMS consecutive items
  uint32_t run[4] = {s_before[0], s_before[1], s_before[2], s_before[3]};
  const int base = blockIdx.x * DN_TILE;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Fate f[DN_ITEMS];
  uint32_t mine[4] = {0, 0, 0, 0};
#pragma unroll
  for (int u = 0; u < DN_ITEMS; u++) {
    const int i = base + threadIdx.x * DN_ITEMS + u;
    f[u] = Fate{0, 0, 0, 0};
    if (i < P) f[u] = decide(c, i, scaling, opacity, grad_accum, denom, max_radii2D);
    mine[0] += f[u].keep; mine[1] += f[u].clone; mine[2] += f[u].split; mine[3] += f[u].sel;
  }
  uint32_t excl[4];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    uint32_t incl = mine[q];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    excl[q] = incl - mine[q];
    if (lane == 31) s_warp[warp][q] = incl;
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 4; q++) {
    uint32_t w = 0;
    for (int j = 0; j < warp; j++) w += s_warp[j][q];
    run[q] += w + excl[q];
  }
#pragma unroll
  for (int u = 0; u < DN_ITEMS; u++) {
    const int i = base + threadIdx.x * DN_ITEMS + u;
    if (i >= P) break;
    if (f[u].keep) {
      const uint32_t d = run[0]++;
      src[d] = i; kind[d] = 0; noise_row[d] = -1;
    }
    if (f[u].clone) {
      const uint32_t d = n_keep + run[1]++;
      src[d] = i; kind[d] = 1; noise_row[d] = -1;
    }
    if (f[u].split) {
      const uint32_t r = run[2]++;
      const uint32_t d0 = n_keep + n_clone + r, d1 = d0 + n_split;
      src[d0] = i; kind[d0] = 2; noise_row[d0] = (int32_t)run[3];
      src[d1] = i; kind[d1] = 3; noise_row[d1] = (int32_t)(n_sel + run[3]);
    }
    run[3] += f[u].sel;
  }
}

struct ApplyArgs {
  skgs_densify_tensor t[SKGS_DENSIFY_MAX_TENSORS];
  int count;
  const float* scaling;   // source log-scales [P][3]
  const float* rotation;  // source rotations [P][4] xyzw
  const float* noise;     // [2 n_selected][3] standard normal
};

// gather of every tensor (blockIdx.y) and its Adam moments into the new arrays; one thread per output float
__global__ void __launch_bounds__(DN_THREADS)
densify_apply_kernel(ApplyArgs a, int P_new, const int32_t* __restrict__ src, const uint8_t* __restrict__ kind,
                     const int32_t* __restrict__ noise_row) {
  const skgs_densify_tensor& T = a.t[blockIdx.y];
  const int W = T.width;
  const size_t total = (size_t)P_new * W;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int d = (int)(e / W), col = (int)(e - (size_t)d * W);
    const int i = src[d];
    const int k = kind[d];
    float v = T.in[(size_t)i * W + col];
    if (k >= 2 && T.role == SKGS_DENSIFY_ROLE_XYZ) {
      // R(normalize(q)) (noise * exp(scaling)) + xyz   (:600-610; quaternion.toR my_ext/ops_3d/quaternion.py:162-172)
      const float* q = a.rotation + 4 * (size_t)i;
      float x = q[0], y = q[1], z = q[2], w = q[3];
      const float n = fmaxf(sqrtf(x * x + y * y + z * z + w * w), 1e-12f);
      x /= n; y /= n; z /= n; w /= n;
      const float* nz = a.noise + 3 * (size_t)noise_row[d];
      const float* ls = a.scaling + 3 * (size_t)i;
      const float sx = nz[0] * expf(ls[0]), sy = nz[1] * expf(ls[1]), sz = nz[2] * expf(ls[2]);
      float r0, r1, r2;
      if (col == 0) { r0 = 1.f - 2.f * y * y - 2.f * z * z; r1 = 2.f * x * y - 2.f * w * z; r2 = 2.f * w * y + 2.f * x * z; }
      else if (col == 1) { r0 = 2.f * x * y + 2.f * w * z; r1 = 1.f - 2.f * x * x - 2.f * z * z; r2 = 2.f * y * z - 2.f * w * x; }
      else { r0 = 2.f * x * z - 2.f * w * y; r1 = 2.f * w * x + 2.f * y * z; r2 = 1.f - 2.f * x * x - 2.f * y * y; }
      v = (r0 * sx + r1 * sy + r2 * sz) + v;
    } else if (k >= 2 && T.role == SKGS_DENSIFY_ROLE_SCALING) {
      v = logf(expf(v) / (0.8f * 2.0f));  // :611
    }
    T.out[e] = v;
    // Adam moments: kept Gaussians carry theirs, new ones start at zero (:548-552)
    if (T.m_out != nullptr) T.m_out[e] = k == 0 ? T.m_in[(size_t)i * W + col] : 0.f;
    if (T.v_out != nullptr) T.v_out[e] = k == 0 ? T.v_in[(size_t)i * W + col] : 0.f;
  }
}

// per-step statistics (:503-513, :672-674)
__global__ void __launch_bounds__(DN_THREADS)
densify_stats_kernel(int P, const int32_t* __restrict__ radii, const float* __restrict__ vs_grad, int vs_stride,
                     float* __restrict__ max_radii2D, float* __restrict__ grad_accum, float* __restrict__ denom,
                     const uint32_t* __restrict__ skip) {
  pdl_wait();
  pdl_trigger();
  if (skip != nullptr && *skip != 0u) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
    const int r = radii[i];
    if (r > 0) {
      max_radii2D[i] = fmaxf(max_radii2D[i], (float)r);
      const float gx = vs_grad[(size_t)i * vs_stride], gy = vs_grad[(size_t)i * vs_stride + 1];
      grad_accum[i] += sqrtf(gx * gx + gy * gy);
      denom[i] += 1.0f;
    }
  }
}

// reset_opacity (:662-665): logit(min(sigmoid(o), 0.01)), moments zeroed (change_optimizer op='replace', :553-555)
__global__ void __launch_bounds__(DN_THREADS)
opacity_reset_kernel(int P, float* __restrict__ opacity, float* __restrict__ m, float* __restrict__ v, float cap) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
    const float o = fminf(sigmoid_ref(opacity[i]), cap);
    opacity[i] = logf(o / (1.0f - o));
    if (m != nullptr) m[i] = 0.f;
    if (v != nullptr) v[i] = 0.f;
  }
}

int grid_for(size_t n, int per_block) {
  size_t g = (n + per_block - 1) / per_block;
  return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

}  // namespace
}  // namespace skgs

using namespace skgs;

extern "C" {

int skgs_densify_stats(int32_t P, const int32_t* radii, const float* viewspace_grad, int32_t grad_stride,
                       float* max_radii2D, float* grad_accum, float* denom, const uint32_t* skip_if_nonzero,
                       void* stream) {
  SKGS_CHECK_ARG(P >= 0, "P < 0");
  if (P == 0) return SKGS_OK;
  SKGS_CHECK_ARG(radii && viewspace_grad && max_radii2D && grad_accum && denom, "NULL buffer");
  SKGS_CHECK_ARG(grad_stride >= 2, "grad_stride=%d must be >= 2", grad_stride);
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof_("densify_stats_kernel", st);
  SKGS_CUDA(launch_pdl(densify_stats_kernel, dim3(grid_for((size_t)P, DN_THREADS)), dim3(DN_THREADS), 0, st, (int)P,
                       radii, viewspace_grad, (int)grad_stride, max_radii2D, grad_accum, denom, skip_if_nonzero));
  SKGS_CHECK_LAUNCH("densify_stats_kernel");
  return SKGS_OK;
}

size_t skgs_densify_workspace_bytes(int32_t P) {
  const size_t tiles = ((size_t)(P > 0 ? P : 1) + DN_TILE - 1) / DN_TILE;
  return tiles * sizeof(uint4) + 64;
}

int skgs_densify_plan(const skgs_densify_config* cfg, int32_t P, const float* scaling, const float* opacity,
                      const float* grad_accum, const float* denom, const float* max_radii2D, int32_t* src,
                      uint8_t* kind, int32_t* noise_row, skgs_densify_counts* counts, void* workspace, void* stream) {
  SKGS_CHECK_ARG(cfg != nullptr && counts != nullptr && workspace != nullptr, "NULL config / counts / workspace");
  SKGS_CHECK_ARG(P >= 1, "P=%d must be >= 1", P);
  SKGS_CHECK_ARG(scaling && opacity && src && kind && noise_row, "NULL buffer");
  SKGS_CHECK_ARG(!cfg->do_densify || (grad_accum && denom), "densification needs grad_accum / denom");
  SKGS_CHECK_ARG(!(cfg->do_prune && cfg->max_screen_size > 0.f && !cfg->do_densify) || max_radii2D,
                 "screen-size pruning needs max_radii2D");
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles = (P + DN_TILE - 1) / DN_TILE;
  uint4* tc = reinterpret_cast<uint4*>(workspace);
  {
    ProfScope prof_("densify_count_kernel", st);
    densify_count_kernel<<<tiles, DN_THREADS, 0, st>>>(*cfg, P, scaling, opacity, grad_accum, denom, max_radii2D, tc);
    SKGS_CHECK_LAUNCH("densify_count_kernel");
  }
  {
    ProfScope prof_("densify_plan_kernel", st);
    densify_plan_kernel<<<tiles, DN_THREADS, 0, st>>>(*cfg, P, tiles, scaling, opacity, grad_accum, denom, max_radii2D,
                                                      tc, src, kind, noise_row, counts);
    SKGS_CHECK_LAUNCH("densify_plan_kernel");
  }
  return SKGS_OK;
}

int skgs_densify_apply(const skgs_densify_tensor* tensors, int32_t count, int32_t P_new, const int32_t* src,
                       const uint8_t* kind, const int32_t* noise_row, const float* scaling, const float* rotation,
                       const float* noise, void* stream) {
  SKGS_CHECK_ARG(count >= 0 && count <= SKGS_DENSIFY_MAX_TENSORS, "count=%d out of range", count);
  SKGS_CHECK_ARG(P_new >= 0, "P_new < 0");
  if (count == 0 || P_new == 0) return SKGS_OK;
  SKGS_CHECK_ARG(tensors && src && kind && noise_row, "NULL buffer");
  ApplyArgs a{};
  a.count = count;
  a.scaling = scaling; a.rotation = rotation; a.noise = noise;
  size_t widest = 1;
  for (int j = 0; j < count; j++) {
    const skgs_densify_tensor& t = tensors[j];
    SKGS_CHECK_ARG(t.in && t.out && t.width >= 1, "tensor %d: NULL in/out or width < 1", j);
    SKGS_CHECK_ARG((t.m_in == nullptr) == (t.m_out == nullptr) && (t.v_in == nullptr) == (t.v_out == nullptr),
                   "tensor %d: moments need both in and out", j);
    SKGS_CHECK_ARG(t.role == SKGS_DENSIFY_ROLE_COPY || t.role == SKGS_DENSIFY_ROLE_XYZ ||
                   t.role == SKGS_DENSIFY_ROLE_SCALING, "tensor %d: unknown role %d", j, t.role);
    SKGS_CHECK_ARG(t.role == SKGS_DENSIFY_ROLE_COPY || t.width == 3, "tensor %d: xyz / scaling must be [P][3]", j);
    SKGS_CHECK_ARG(t.role != SKGS_DENSIFY_ROLE_XYZ || (scaling && rotation && noise),
                   "splitting needs the source scaling / rotation and the noise table");
    a.t[j] = t;
    if ((size_t)t.width > widest) widest = t.width;
  }
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof_("densify_apply_kernel", st);
  dim3 grid(grid_for((size_t)P_new * widest, DN_THREADS * 4), count);
  densify_apply_kernel<<<grid, DN_THREADS, 0, st>>>(a, P_new, src, kind, noise_row);
  SKGS_CHECK_LAUNCH("densify_apply_kernel");
  return SKGS_OK;
}

int skgs_opacity_reset(int32_t P, float* opacity, float* exp_avg, float* exp_avg_sq, float cap, void* stream) {
  SKGS_CHECK_ARG(P >= 0, "P < 0");
  if (P == 0) return SKGS_OK;
  SKGS_CHECK_ARG(opacity != nullptr && cap > 0.f && cap < 1.f, "NULL opacity or cap outside (0, 1)");
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof_("opacity_reset_kernel", st);
  opacity_reset_kernel<<<grid_for((size_t)P, DN_THREADS), DN_THREADS, 0, st>>>(P, opacity, exp_avg, exp_avg_sq, cap);
  SKGS_CHECK_LAUNCH("opacity_reset_kernel");
  return SKGS_OK;
}

}  // extern "C"
