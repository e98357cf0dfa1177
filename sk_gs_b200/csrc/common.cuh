// common.cuh - shared device helpers of libskgs_b200 (sm_100a only).
//
// Arithmetic contract: everything that decides an integer result (radii, tile rects, keys) or the set of contributing
// (pixel, Gaussian) pairs is written with explicit round-to-nearest intrinsics (__fmul_rn / __fadd_rn / __fmaf_rn),
// which nvcc never contracts or reorders, so the results are bit-identical to the CPU oracle (oracle/raster_oracle.c,
// built with -ffp-contract=off) and identical between the forward and the backward kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <utility>

#include "../../include/skgs_b200.h"

namespace skgs {

constexpr int TILE = 16;            // BLOCK_X == BLOCK_Y of the reference (include/gaussian_render.h:29-30)
constexpr int TILE_PIX = TILE * TILE;

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// Optional per-kernel device timing (skgs_profile_enable): a CUDA-event pair around every launch on the launching
// stream.  Off by default; bench.py turns it on only for its roofline pass, never for the throughput pass.
bool prof_enabled();
void prof_begin(const char* name, cudaStream_t st);
void prof_end(cudaStream_t st);
struct ProfScope {
  cudaStream_t st;
  bool on;
  ProfScope(const char* name, cudaStream_t s) : st(s), on(prof_enabled()) {
    if (on) prof_begin(name, st);
  }
  ~ProfScope() {
    if (on) prof_end(st);
  }
};

#define SKGS_CHECK_ARG(cond, ...)            \
  do {                                       \
    if (!(cond)) {                           \
      skgs::set_error(__VA_ARGS__);          \
      return SKGS_ERR_INVALID_ARG;           \
    }                                        \
  } while (0)

// same convention as the reference's CHECK_CUDA_ERROR (include/util.cuh:31-35): cudaGetLastError after launch, no sync
#define SKGS_CHECK_LAUNCH(name)                                                          \
  do {                                                                                   \
    cudaError_t e_ = cudaGetLastError();                                                 \
    if (e_ != cudaSuccess) {                                                             \
      skgs::set_error("%s: %s", name, cudaGetErrorString(e_));                           \
      return SKGS_ERR_CUDA;                                                              \
    }                                                                                    \
    skgs::count_launch();                                                                \
  } while (0)

#define SKGS_CUDA(call)                                                                  \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      skgs::set_error("%s: %s", #call, cudaGetErrorString(e_));                          \
      return SKGS_ERR_CUDA;                                                              \
    }                                                                                    \
  } while (0)

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------------------------------
// Programmatic dependent launch (sm_90+): a kernel launched with launch_pdl() may begin launching while its stream
// predecessor is still draining; it must execute pdl_wait() before touching ANY global memory (griddepcontrol.wait
// returns once every prerequisite grid has completed and flushed), and it releases its own successor with
// pdl_trigger().  Used on the 20-kernel chain of one training step, where every kernel boundary otherwise costs a full
// launch latency inside the CUDA graph.  SKGS_PDL=0 in the environment turns the attribute off (plain stream order).
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

// ------------------------------------------------------------------------------------------------------------------
// skgs_exp: the fully specified exp for x <= 0 shared (as a specification) with the oracle's orc_exp():
// t = x*log2e (one rounding), n = rint(t) via the 1.5*2^23 trick, f = t - n (exact), 2^f by a degree-6 minimax
// polynomial in Horner form with explicit FMAs, exponent patched by an integer add.  12 instructions, <= 3 ulp.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float skgs_exp(float x) {
  float t = __fmul_rn(x, 1.4426950408889634f);
  t = fmaxf(t, -120.0f);
  const float r = __fadd_rn(t, 12582912.0f);
  const float nf = __fadd_rn(r, -12582912.0f);
  const float f = __fadd_rn(t, -nf);
  float p = 0.00015345810970757157f;
  p = __fmaf_rn(p, f, 0.0013399930903688073f);
  p = __fmaf_rn(p, f, 0.009618489071726799f);
  p = __fmaf_rn(p, f, 0.05550328642129898f);
  p = __fmaf_rn(p, f, 0.24022646248340607f);
  p = __fmaf_rn(p, f, 0.6931471824645996f);
  p = __fmaf_rn(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
}

// power = -1/2 (A dx^2 + C dy^2) - B dx dy, evaluated as fma(dx, fma(A', dx, B'*dy), (C'*dy)*dy) with
// A' = -A/2, B' = -B, C' = -C/2 (exact scalings).  Row-shared terms bdy = B'*dy and cdy2 = (C'*dy)*dy are hoisted.
__device__ __forceinline__ float pair_power(float Ap, float dx, float bdy, float cdy2) {
  return __fmaf_rn(dx, __fmaf_rn(Ap, dx, bdy), cdy2);
}

// reference: include/gaussian_render.h:42-47
__device__ __forceinline__ void get_rect(float px, float py, int max_radius, int gx, int gy, int& x0, int& y0, int& x1,
                                         int& y1) {
  const float r = (float)max_radius;
  x0 = min(gx, max(0, (int)(__fadd_rn(px, -r) / (float)TILE)));
  y0 = min(gy, max(0, (int)(__fadd_rn(py, -r) / (float)TILE)));
  x1 = min(gx, max(0, (int)(__fadd_rn(__fadd_rn(px, r), (float)(TILE - 1)) / (float)TILE)));
  y1 = min(gy, max(0, (int)(__fadd_rn(__fadd_rn(py, r), (float)(TILE - 1)) / (float)TILE)));
}

// reference: gaussian_rasterizer_forward.cu:30-42
inline uint32_t higher_msb(uint32_t n) {
  uint32_t msb = sizeof(n) * 4;
  uint32_t step = msb;
  while (step > 1) {
    step /= 2;
    if (n >> msb)
      msb += step;
    else
      msb -= step;
  }
  if (n >> msb) msb++;
  return msb;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// vector reduction into global memory (sm_90+): one 16-byte RED instead of four scalar ones
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct RasterParams {
  int P, M, D, W, H, gx, gy;
  float tanfovx, tanfovy, fx, fy, mod;
  int quat_wxyz;
  const float *view, *proj, *campos, *bg;
};

int launch_preprocess_scan(const RasterParams& rp, const float* means3D, const float* shs, const float* colors_precomp,
                           const float* opacities, const float* scales, const float* rotations,
                           const float* cov3D_precomp, char* geom, const skgs_raster_layout& lay, int32_t* radii,
                           uint32_t* num_rendered_host, char* binning, char* img, int64_t R_cap, cudaStream_t st);
int launch_binning(const RasterParams& rp, char* geom, char* binning, char* img, const skgs_raster_layout& lay,
                   const int32_t* radii, int64_t R_cap, int64_t R_hint, bool emit, uint32_t* num_rendered_host,
                   cudaStream_t st);
int launch_deform_preprocess(const RasterParams& rp, const skgs_skeleton* sk, const float* table, const float* xyz,
                             const float* scaling, const float* rotation, const float* opacity_logit, const float* shs,
                             float* points, float* scales, float* rotations, float* opacities, float* d_rot,
                             float* weights, int64_t* indices, char* geom, const skgs_raster_layout& lay,
                             int32_t* radii, char* binning, char* img, int64_t R_cap, cudaStream_t st);
int launch_fk_table(const skgs_skeleton* sk, float* sk_T, float* table, cudaStream_t st);
int launch_tile_order(const RasterParams& rp, char* img, const skgs_raster_layout& lay, cudaStream_t st);
// binning by tile-segmented sort (tile_sort.cu)
int tile_cell_stride();
int launch_tile_binning(const RasterParams& rp, char* geom, char* binning, char* img, const skgs_raster_layout& lay,
                        int64_t R_cap, int64_t R_hint, cudaStream_t st);
int launch_composite_fwd(const RasterParams& rp, char* geom, char* binning, char* img, const skgs_raster_layout& lay,
                         float* out_color, float* out_depth, float* out_alpha, cudaStream_t st);
int launch_composite_bwd(const RasterParams& rp, char* geom, const char* binning, const char* img,
                         const skgs_raster_layout& lay, const float* dL_dcolor, const float* dL_ddepth,
                         const float* dL_dalpha, int tfinal_via_opacity, cudaStream_t st);
int launch_preprocess_bwd(const RasterParams& rp, const float* means3D, const float* shs, const float* colors_precomp,
                          const float* scales, const float* rotations, const float* cov3D_precomp,
                          const int32_t* radii, char* geom, const skgs_raster_layout& lay, uint32_t* bwd_ticket,
                          float* dL_dmeans3D, float* dL_dmeans2D, float* dL_dsh, float* dL_dcolors, float* dL_dopacity,
                          float* dL_dscales, float* dL_drotations, float* dL_dcov3D, const float* const* assemble_in,
                          float* const* assemble_out, cudaStream_t st);

}  // namespace skgs
