// allreduce_mm_synced.cu - NOT part of the product build.  One-launch variant of the in-switch all-reduce: both
// cross-GPU barriers inside the reduce kernel (epoch flags through the symmetric-memory signal pads).  Correct in
// tools/mm_test.py (eager and captured), but (a) no faster than barrier-kernel + reduce + barrier-kernel at N = 2
// (101.5 vs 103.4 us; the two barrier kernels cost 13 us together) and (b) the second range of bench.py's split exchange
// came out wrong inside the step graph (exactly 2x: the reduce saw the previous step's values) - root cause not found.
// The product keeps the bare kernel between torch's signal-pad barriers, which bench.py's self-check verifies every run.
// allreduce_mm.cu - in-switch (NVLS) all-reduce of the flat gradient arena for view-sharded data parallelism.
//
// The arena is SYMMETRIC memory: the same allocation on every GPU of the NVSwitch box, bound to one multicast address.
// Rank r owns the r-th slice of the arena: it reads the slice with `multimem.ld_reduce` (the switch sums the N copies
// on the fly - 1/N of the arena crosses each GPU's links instead of 2(N-1)/N for a ring) and writes the sum back with
// `multimem.st`, which the switch broadcasts to all N copies.  One pass, no staging buffers, no NCCL protocol latency.
//
// Two entry points:
//   skgs_multimem_allreduce         the bare kernel; the caller brackets it with cross-GPU barriers on the same stream
//   skgs_multimem_allreduce_synced  ONE launch that also does both barriers: block 0 exchanges epoch flags with every
//                                   peer through the symmetric-memory signal pads (st.release.sys / ld.acquire.sys over
//                                   NVLink) before the other blocks start, the last block to finish does the same after
//                                   its stores are fenced.  The epoch lives in device memory, so a captured CUDA graph
//                                   can be replayed.  Replaces barrier-kernel + reduce-kernel + barrier-kernel
//                                   (three launches, ~20 us of launch gaps and single-CTA kernels per exchange).
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
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
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

// control words of the synced kernel (device memory of this rank, zero-initialised by the owner of the arena)
struct MmCtrl {
  uint32_t done_epoch;  // epoch of the last completed call (written once, by the last block, at the very end)
  uint32_t go;          // epoch whose entry barrier has completed (block 0 -> all blocks)
  uint32_t arrive;      // blocks that have finished their stores
  uint32_t error;       // a spin ran into its cap (a peer never arrived): results are invalid
};
constexpr uint32_t SPIN_CAP = 1u << 27;  // ~1 s of polling: never hang the GPU on a lost peer

// signal every peer on `channel` and wait for every peer's signal (threads 0..world-1 of one block)
__device__ __forceinline__ void pad_barrier(uint32_t* const* pads, int rank, int world, int channel, uint32_t epoch,
                                            MmCtrl* ctrl) {
  if ((int)threadIdx.x < world) {
    const int peer = (int)threadIdx.x;
    st_release_sys(pads[peer] + (size_t)channel * world + rank, epoch);
    const uint32_t* mine = pads[rank] + (size_t)channel * world + peer;
    uint32_t spins = 0;
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
      if (++spins > SPIN_CAP) {
        ctrl->error = 1u;
        break;
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(512)
multimem_allreduce_synced_kernel(float* __restrict__ mc, size_t nvec, int rank, int world, uint32_t* const* pads,
                                 int channel, MmCtrl* ctrl, int exit_barrier) {
  __shared__ uint32_t s_last;
  pdl_wait();
  pdl_trigger();
  const uint32_t epoch = ld_acquire_gpu(&ctrl->done_epoch) + 1u;  // stable until the last block's final store
  if (blockIdx.x == 0) {
    // entry: every rank's producers of this range have finished (they precede this kernel in stream order)
    pad_barrier(pads, rank, world, channel, epoch, ctrl);
    if (threadIdx.x == 0) st_release_gpu(&ctrl->go, epoch);
  } else {
    if (threadIdx.x == 0) {
      uint32_t spins = 0;
      while (ld_acquire_gpu(&ctrl->go) != epoch) {
        if (++spins > SPIN_CAP) {
          ctrl->error = 1u;
          break;
        }
      }
    }
    __syncthreads();
  }
  reduce_slice(mc, nvec, rank, world);
  // exit: this block's multimem.st are ordered before its arrival; the last block signals the peers
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&ctrl->arrive, 1u) == gridDim.x - 1u) ? 1u : 0u;
  __syncthreads();
  if (s_last == 0u) return;
  __threadfence_system();
  if (threadIdx.x == 0) ctrl->arrive = 0u;
  if (exit_barrier) pad_barrier(pads, rank, world, channel + 1, epoch, ctrl);
  if (threadIdx.x == 0) st_release_gpu(&ctrl->done_epoch, epoch);
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

extern "C" int skgs_multimem_allreduce_synced(void* multicast_ptr, int64_t numel, int32_t rank, int32_t world,
                                              void* const* signal_pads_dev, int32_t channel, void* ctrl,
                                              int32_t exit_barrier, int32_t max_blocks, void* stream) {
  SKGS_CHECK_ARG(multicast_ptr != nullptr, "multicast pointer is NULL (no NVLS multicast support?)");
  SKGS_CHECK_ARG(numel >= 0 && numel % 4 == 0, "numel=%lld must be a multiple of 4", (long long)numel);
  SKGS_CHECK_ARG(((uintptr_t)multicast_ptr & 15) == 0, "multicast pointer must be 16-byte aligned");
  SKGS_CHECK_ARG(world >= 1 && world <= 32 && rank >= 0 && rank < world, "bad rank/world %d/%d", rank, world);
  SKGS_CHECK_ARG(signal_pads_dev != nullptr && ctrl != nullptr, "signal pads / control block are NULL");
  SKGS_CHECK_ARG(channel >= 0, "channel < 0");
  const size_t nvec = (size_t)numel / 4;
  const size_t per = (nvec + world - 1) / world;
  cudaStream_t st = (cudaStream_t)stream;
  {
    ProfScope prof_("multimem_allreduce_synced_kernel", st);
    multimem_allreduce_synced_kernel<<<mm_blocks(per, max_blocks > 0 ? max_blocks : 296), 512, 0, st>>>(
        (float*)multicast_ptr, nvec, rank, world, reinterpret_cast<uint32_t* const*>(signal_pads_dev), channel,
        reinterpret_cast<MmCtrl*>(ctrl), exit_barrier);
    SKGS_CHECK_LAUNCH("multimem_allreduce_synced_kernel");
  }
  return SKGS_OK;
}
