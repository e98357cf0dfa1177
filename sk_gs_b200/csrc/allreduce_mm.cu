// allreduce_mm.cu - in-switch (NVLS) all-reduce of the flat gradient arena for view-sharded data parallelism.
//
// The arena is SYMMETRIC memory: the same allocation on every GPU of the NVSwitch box, bound to one multicast address.
// Rank r owns the r-th slice of the arena: it reads the slice with `multimem.ld_reduce` (the switch sums the N copies
// on the fly - 1/N of the arena crosses each GPU's links instead of 2(N-1)/N for a ring) and writes the sum back with
// `multimem.st`, which the switch broadcasts to all N copies.  One pass, no staging buffers, no NCCL protocol latency.
//
// skgs_multimem_allreduce is the bare kernel; the caller brackets it with cross-GPU barriers on the same stream
// (torch symmetric-memory signal pads).  A variant with both barriers inside the launch was tried and dropped
// (tools/experiments/allreduce_mm_synced.cu).
#include <cstdlib>

#include "common.cuh"

namespace skgs {

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

__device__ __forceinline__ void reduce_slice(float* __restrict__ mc, size_t nvec, int rank, int world) {
  const size_t per = (nvec + world - 1) / world;
  const size_t beg = (size_t)rank * per, end = beg + per < nvec ? beg + per : nvec;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = beg + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += 4 * stride) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (i + u * stride < end) v[u] = multimem_ld_reduce_add(mc + 4 * (i + u * stride));
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (i + u * stride < end) multimem_st(mc + 4 * (i + u * stride), v[u]);
  }
}

__global__ void __launch_bounds__(512)
multimem_allreduce_kernel(float* __restrict__ mc, size_t nvec, int rank, int world) {
  pdl_wait();
  pdl_trigger();
  reduce_slice(mc, nvec, rank, world);
}

static int mm_blocks(size_t per, int dflt_cap) {
  int blocks = (int)((per + 512 * 4 - 1) / (512 * 4));
  static int max_blocks = 0;
  if (max_blocks == 0) {
    const char* e = getenv("SKGS_MM_BLOCKS");
    max_blocks = e ? atoi(e) : 0;
    if (max_blocks < 1) max_blocks = -1;
  }
  const int cap = max_blocks > 0 ? max_blocks : dflt_cap;
  return blocks < 1 ? 1 : (blocks > cap ? cap : blocks);
}

}  // namespace skgs

using namespace skgs;

extern "C" int skgs_multimem_allreduce(void* multicast_ptr, int64_t numel, int32_t rank, int32_t world, void* stream) {
  SKGS_CHECK_ARG(multicast_ptr != nullptr, "multicast pointer is NULL (no NVLS multicast support?)");
  SKGS_CHECK_ARG(numel >= 0 && numel % 4 == 0, "numel=%lld must be a multiple of 4", (long long)numel);
  SKGS_CHECK_ARG(((uintptr_t)multicast_ptr & 15) == 0, "multicast pointer must be 16-byte aligned");
  SKGS_CHECK_ARG(world >= 1 && rank >= 0 && rank < world, "bad rank/world %d/%d", rank, world);
  if (numel == 0) return SKGS_OK;
  const size_t nvec = (size_t)numel / 4;
  const size_t per = (nvec + world - 1) / world;
  cudaStream_t st = (cudaStream_t)stream;
  {
    ProfScope prof_("multimem_allreduce_kernel", st);
    multimem_allreduce_kernel<<<mm_blocks(per, 296), 512, 0, st>>>((float*)multicast_ptr, nvec, rank, world);
    SKGS_CHECK_LAUNCH("multimem_allreduce_kernel");
  }
  return SKGS_OK;
}
