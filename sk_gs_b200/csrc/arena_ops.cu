// arena_ops.cu - element-wise helpers of a multi-view training step: the reference loops over the views
// of a step and lets autograd SUM their gradients (networks/sk_gs.py:1220, framework backward), and tracks the MAX of
// the screen radii over views (networks/gaussian_splatting.py:638-640 `radii.amax(dim=0)`).  Here every view's backward
// writes its gradients into a flat fp32 arena (sk_gs_b200.dist.GradArena); the arenas of the second, third ... view of a
// rank are folded into the first with skgs_accumulate_f32 (one vectorised stream: 12 B per element).
#include "common.cuh"

namespace skgs {

__global__ void __launch_bounds__(256) accumulate_f32_kernel(float* __restrict__ dst, const float* __restrict__ src,
                                                             size_t n) {
  const size_t nvec = n / 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  float4* d4 = reinterpret_cast<float4*>(dst);
  const float4* s4 = reinterpret_cast<const float4*>(src);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += 2 * stride) {
    const size_t j = i + stride;
    float4 a = d4[i];
    const float4 b = __ldg(s4 + i);
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f), e = c;
    if (j < nvec) {
      c = d4[j];
      e = __ldg(s4 + j);
    }
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    d4[i] = a;
    if (j < nvec) {
      c.x += e.x; c.y += e.y; c.z += e.z; c.w += e.w;
      d4[j] = c;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) dst[nvec * 4 + threadIdx.x] += src[nvec * 4 + threadIdx.x];
}

__global__ void __launch_bounds__(256) max_i32_kernel(int32_t* __restrict__ dst, const int32_t* __restrict__ src,
                                                      size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = max(dst[i], __ldg(src + i));
}

static int stream_grid(size_t work_items) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  size_t g = (work_items + 255) / 256;
  const size_t cap = (size_t)sms * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace skgs

using namespace skgs;

extern "C" int skgs_accumulate_f32(float* dst, const float* src, int64_t numel, void* stream) {
  SKGS_CHECK_ARG(numel >= 0, "accumulate: numel < 0");
  if (numel == 0) return SKGS_OK;
  SKGS_CHECK_ARG(dst && src, "accumulate: null pointer");
  SKGS_CHECK_ARG((((uintptr_t)dst | (uintptr_t)src) & 15) == 0, "accumulate: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  {
    ProfScope prof_("accumulate_f32_kernel", st);
    accumulate_f32_kernel<<<stream_grid((size_t)numel / 8 + 1), 256, 0, st>>>(dst, src, (size_t)numel);
    SKGS_CHECK_LAUNCH("accumulate_f32_kernel");
  }
  return SKGS_OK;
}

extern "C" int skgs_max_i32(int32_t* dst, const int32_t* src, int64_t numel, void* stream) {
  SKGS_CHECK_ARG(numel >= 0, "max_i32: numel < 0");
  if (numel == 0) return SKGS_OK;
  SKGS_CHECK_ARG(dst && src, "max_i32: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  {
    ProfScope prof_("max_i32_kernel", st);
    max_i32_kernel<<<stream_grid((size_t)numel), 256, 0, st>>>(dst, src, (size_t)numel);
    SKGS_CHECK_LAUNCH("max_i32_kernel");
  }
  return SKGS_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// debug: a chain of n dependent trivial kernels, with or without programmatic dependent launch (tools/pdl_probe.py
// measures what a kernel boundary costs on this box, eagerly and inside a captured graph)
// ------------------------------------------------------------------------------------------------------------------
namespace skgs {
__global__ void pdl_probe_kernel(uint32_t* p) {
  pdl_wait();
  pdl_trigger();
  if (threadIdx.x == 0 && blockIdx.x == 0) p[0] += 1u;
}
}  // namespace skgs

extern "C" __attribute__((visibility("default"))) int skgs_debug_pdl_chain(uint32_t* p, int n, int use_pdl, int grid,
                                                                          void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  for (int i = 0; i < n; i++) {
    if (use_pdl) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid);
      cfg.blockDim = dim3(128);
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      SKGS_CUDA(cudaLaunchKernelEx(&cfg, skgs::pdl_probe_kernel, p));
    } else {
      skgs::pdl_probe_kernel<<<grid, 128, 0, st>>>(p);
    }
  }
  return SKGS_OK;
}
