// raster_fwd.cu - the per-Gaussian forward kernels of the tile rasterizer for sm_100a:
//   preprocess_scan_kernel   : projection, EWA cov2D, conic, radius, tile rect, SH->RGB, footprint-culling record
//                              + fused decoupled-look-back prefix sum of tiles_touched (no scan kernel, no host sync
//                              for R) + (EMIT) warp-cooperative emission of (tile << 32 | depth bits, id) and the tile
//                              rectangle of every Gaussian added to the difference grid the per-tile sort is planned from
//   deform_preprocess_kernel : the same behind K-nearest-joint skinning + output assembly (deform.cuh): the whole
//                              per-Gaussian forward of the SK_GS step in one kernel
//   duplicate_keys_kernel    : the emission alone, from stored geometry (split API / capacity retry only)
//   (binning continues in tile_sort.cu, compositing in composite.cu)
// Semantics follow SURVEY.md App. A.4-A.6 (reference: my_ext/_C/src/nerf/gaussian_preprocess_colmap.cu:155-224,
// gaussian_rasterizer_forward.cu:45-94,203-241, gaussian_render.cu:16-112).  This file MUST be compiled with
// -fmad=false (and without fast-math): plain '*' and '+' below are separately rounded, exactly like the CPU oracle and
// like the reference extension built with -fmad=false (oracle/build_ref.sh, the bit-exact parity target).
#ifndef SKGS_NO_FMAD
#error "raster_fwd.cu must be built with -fmad=false -DSKGS_NO_FMAD (bit-exact radii/keys depend on it)"
#endif
#include <cooperative_groups.h>

#include "deform.cuh"

namespace skgs {

__device__ __constant__ float c_SH_C0 = 0.28209479177387814f;
__device__ __constant__ float c_SH_C1 = 0.4886025119029199f;
__device__ __constant__ float c_SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                            -1.0925484305920792f, 0.5462742152960396f};
__device__ __constant__ float c_SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                            0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                            -0.5900435899266435f};

// ------------------------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------------------------
constexpr int PRE_THREADS = 256;
constexpr uint64_t SCAN_FLAG_AGG = 1ull << 62;
constexpr uint64_t SCAN_FLAG_INC = 2ull << 62;
constexpr uint64_t SCAN_VAL_MASK = (1ull << 62) - 1;
constexpr uint32_t FULL = 0xffffffffu;

__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_volatile_v4(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u64(uint64_t* p, uint64_t v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Everything a Gaussian hands to the binning stage.
struct Emit {
  uint32_t cnt;    // tiles touched (0: culled)
  int x0, y0, w;   // tile rect origin and width
  uint32_t dbits;  // depth as sortable bits (positive float)
};

struct BinningOut {            // where the emitting kernel writes; all NULL in the geometry-only variant
  uint64_t* keys;              // (tile << 32 | depth bits) in emission order
  uint32_t* vals;              // Gaussian ids
  int* grid;                   // [(gy+1)][(gx+1)] x cell_stride: 2-D difference grid of the tile rectangles; its prefix
                               // sum (tile_plan_kernel) is the number of keys per tile
  uint32_t R_cap;
  int cell_stride;             // ints between two cells (one 128-byte line per cell: same-line atomics serialise in L2)
};

// ------------------------------------------------------------------------------------------------------------------
// Per-Gaussian preprocessing (reference: gaussian_preprocess_colmap.cu:155-224 + computeColorFromSH,
// gaussian_rasterizer_forward.cu:97-137).  Plain '*' and '+' are separately rounded (this TU is built with -fmad=false)
// and follow the association of the reference's source text, glm products included (sum of three products, left to
// right), so that radii, rects and depth bits equal those of the reference built without FMA contraction.
// ------------------------------------------------------------------------------------------------------------------
struct GeomOut {
  int32_t* radii;
  float2* means2D;
  float* depths;
  float* cov3Ds;
  float4* conic_opacity;
  float4* rgbd;
  float4* cull;
  uint8_t* clamped;
};

__device__ __forceinline__ Emit preprocess_gaussian(const RasterParams& rp, const float* V, const float* Pm,
                                                    const float* cam, int i, float x, float y, float z, float opacity,
                                                    const float* c6_pre, float s0, float s1, float s2, float qx,
                                                    float qy, float qz, float qr, const float* __restrict__ shs,
                                                    const float* __restrict__ colors_precomp, const GeomOut& o) {
  Emit e;
  e.cnt = 0; e.x0 = 0; e.y0 = 0; e.w = 0; e.dbits = 0;
  int my_rad = 0;
  const float pvx = V[0] * x + V[4] * y + V[8] * z + V[12];
  const float pvy = V[1] * x + V[5] * y + V[9] * z + V[13];
  const float pvz = V[2] * x + V[6] * y + V[10] * z + V[14];
  if (pvz > 0.2f) {
    const float hx = Pm[0] * x + Pm[4] * y + Pm[8] * z + Pm[12];
    const float hy = Pm[1] * x + Pm[5] * y + Pm[9] * z + Pm[13];
    const float hw = Pm[3] * x + Pm[7] * y + Pm[11] * z + Pm[15];
    const float pw = 1.0f / (hw + 0.0000001f);
    const float ppx = hx * pw, ppy = hy * pw;
    float c6[6];
    if (c6_pre != nullptr) {
#pragma unroll
      for (int k = 0; k < 6; k++) c6[k] = c6_pre[k];
    } else {
      float R[3][3];
      R[0][0] = 1.f - 2.f * (qy * qy + qz * qz);
      R[0][1] = 2.f * (qx * qy - qr * qz);
      R[0][2] = 2.f * (qx * qz + qr * qy);
      R[1][0] = 2.f * (qx * qy + qr * qz);
      R[1][1] = 1.f - 2.f * (qx * qx + qz * qz);
      R[1][2] = 2.f * (qy * qz - qr * qx);
      R[2][0] = 2.f * (qx * qz - qr * qy);
      R[2][1] = 2.f * (qy * qz + qr * qx);
      R[2][2] = 1.f - 2.f * (qx * qx + qy * qy);
      float Mm[3][3];
#pragma unroll
      for (int a = 0; a < 3; a++) {
        Mm[0][a] = s0 * R[a][0];
        Mm[1][a] = s1 * R[a][1];
        Mm[2][a] = s2 * R[a][2];
      }
      c6[0] = Mm[0][0] * Mm[0][0] + Mm[1][0] * Mm[1][0] + Mm[2][0] * Mm[2][0];
      c6[1] = Mm[0][0] * Mm[0][1] + Mm[1][0] * Mm[1][1] + Mm[2][0] * Mm[2][1];
      c6[2] = Mm[0][0] * Mm[0][2] + Mm[1][0] * Mm[1][2] + Mm[2][0] * Mm[2][2];
      c6[3] = Mm[0][1] * Mm[0][1] + Mm[1][1] * Mm[1][1] + Mm[2][1] * Mm[2][1];
      c6[4] = Mm[0][1] * Mm[0][2] + Mm[1][1] * Mm[1][2] + Mm[2][1] * Mm[2][2];
      c6[5] = Mm[0][2] * Mm[0][2] + Mm[1][2] * Mm[1][2] + Mm[2][2] * Mm[2][2];
#pragma unroll
      for (int k = 0; k < 6; k++) o.cov3Ds[6 * (size_t)i + k] = c6[k];
    }
    // EWA: rows of A = J * Rv
    const float limx = 1.3f * rp.tanfovx, limy = 1.3f * rp.tanfovy;
    const float txtz = pvx / pvz, tytz = pvy / pvz;
    const float tx = fminf(limx, fmaxf(-limx, txtz)) * pvz;
    const float ty = fminf(limy, fmaxf(-limy, tytz)) * pvz;
    const float j00 = rp.fx / pvz, j02 = -(rp.fx * tx) / (pvz * pvz);
    const float j11 = rp.fy / pvz, j12 = -(rp.fy * ty) / (pvz * pvz);
    const float a00 = j00 * V[0] + j02 * V[2], a01 = j00 * V[4] + j02 * V[6], a02 = j00 * V[8] + j02 * V[10];
    const float a10 = j11 * V[1] + j12 * V[2], a11 = j11 * V[5] + j12 * V[6], a12 = j11 * V[9] + j12 * V[10];
    const float u00 = c6[0] * a00 + c6[1] * a01 + c6[2] * a02;
    const float u01 = c6[1] * a00 + c6[3] * a01 + c6[4] * a02;
    const float u02 = c6[2] * a00 + c6[4] * a01 + c6[5] * a02;
    const float u10 = c6[0] * a10 + c6[1] * a11 + c6[2] * a12;
    const float u11 = c6[1] * a10 + c6[3] * a11 + c6[4] * a12;
    const float u12 = c6[2] * a10 + c6[4] * a11 + c6[5] * a12;
    const float c00 = (a00 * u00 + a01 * u01 + a02 * u02) + 0.3f;
    const float c01 = a00 * u10 + a01 * u11 + a02 * u12;
    const float c11 = (a10 * u10 + a11 * u11 + a12 * u12) + 0.3f;
    const float det = c00 * c11 - c01 * c01;
    if (det != 0.0f) {
      const float det_inv = 1.f / det;
      const float con_x = c11 * det_inv, con_y = -c01 * det_inv, con_z = c00 * det_inv;
      const float mid = 0.5f * (c00 + c11);
      const float disc = sqrtf(fmaxf(0.1f, mid * mid - det));
      const float lambda1 = mid + disc, lambda2 = mid - disc;
      const float my_radius = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));
      // ndc2Pix is double arithmetic in the reference (gaussian_preprocess_colmap.cu:26)
      const float pix_x = (float)__dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn((double)ppx, 1.0), (double)rp.W), -1.0), 0.5);
      const float pix_y = (float)__dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn((double)ppy, 1.0), (double)rp.H), -1.0), 0.5);
      const int irad = my_radius > 2.0e9f ? 2000000000 : (int)my_radius;
      int x0, y0, x1, y1;
      get_rect(pix_x, pix_y, irad, rp.gx, rp.gy, x0, y0, x1, y1);
      const int cnt = (x1 - x0) * (y1 - y0);
      if (cnt != 0) {
        float rgb[3];
        uint8_t cl = 0;
        if (colors_precomp == nullptr) {
          const float dirx = x - cam[0], diry = y - cam[1], dirz = z - cam[2];
          const float len = sqrtf(dirx * dirx + diry * diry + dirz * dirz);
          const float dx = dirx / len, dy = diry / len, dz = dirz / len;
          // 48 floats per Gaussian, 16-byte aligned: 12 vector loads
          float sh[48];
          const float4* sp = reinterpret_cast<const float4*>(shs + (size_t)i * rp.M * 3);
          const int nvec = (rp.M * 3) / 4;  // M is 1,4,9,16 -> handle the general case below
          if ((rp.M * 3) % 4 == 0) {
#pragma unroll
            for (int k = 0; k < 12; k++)
              if (k < nvec) {
                const float4 v = __ldg(sp + k);
                sh[4 * k] = v.x; sh[4 * k + 1] = v.y; sh[4 * k + 2] = v.z; sh[4 * k + 3] = v.w;
              }
          } else {
            const float* sf = shs + (size_t)i * rp.M * 3;
#pragma unroll
            for (int k = 0; k < 48; k++)
              if (k < rp.M * 3) sh[k] = __ldg(sf + k);
          }
          const int deg = rp.D;
#pragma unroll
          for (int c = 0; c < 3; c++) {
            float res = c_SH_C0 * sh[c];
            if (deg > 0) {
              res = res - c_SH_C1 * dy * sh[3 + c] + c_SH_C1 * dz * sh[6 + c] - c_SH_C1 * dx * sh[9 + c];
              if (deg > 1) {
                const float xx = dx * dx, yy = dy * dy, zz = dz * dz, xy = dx * dy, yz = dy * dz, xz = dx * dz;
                res = res + c_SH_C2[0] * xy * sh[12 + c] + c_SH_C2[1] * yz * sh[15 + c] +
                      c_SH_C2[2] * (2.0f * zz - xx - yy) * sh[18 + c] + c_SH_C2[3] * xz * sh[21 + c] +
                      c_SH_C2[4] * (xx - yy) * sh[24 + c];
                if (deg > 2) {
                  res = res + c_SH_C3[0] * dy * (3.0f * xx - yy) * sh[27 + c] + c_SH_C3[1] * xy * dz * sh[30 + c] +
                        c_SH_C3[2] * dy * (4.0f * zz - xx - yy) * sh[33 + c] +
                        c_SH_C3[3] * dz * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[36 + c] +
                        c_SH_C3[4] * dx * (4.0f * zz - xx - yy) * sh[39 + c] +
                        c_SH_C3[5] * dz * (xx - yy) * sh[42 + c] + c_SH_C3[6] * dx * (xx - 3.0f * yy) * sh[45 + c];
                }
              }
            }
            res += 0.5f;
            if (res < 0.f) cl |= (uint8_t)(1u << c);
            rgb[c] = res < 0.f ? 0.f : res;
          }
        } else {
          rgb[0] = colors_precomp[3 * (size_t)i];
          rgb[1] = colors_precomp[3 * (size_t)i + 1];
          rgb[2] = colors_precomp[3 * (size_t)i + 2];
        }
        my_rad = irad;
        e.cnt = (uint32_t)cnt;
        e.x0 = x0; e.y0 = y0; e.w = x1 - x0;
        e.dbits = __float_as_uint(pvz);
        o.depths[i] = pvz;
        o.means2D[i] = make_float2(pix_x, pix_y);
        o.conic_opacity[i] = make_float4(con_x, con_y, con_z, opacity);
        o.rgbd[i] = make_float4(rgb[0], rgb[1], rgb[2], pvz);
        o.clamped[i] = cl;
        // Footprint-culling record of the compositing kernels (composite.cu): pmin = the exponent below which
        // alpha = o*exp(power) < 1/255 can never be reached (conservative: margin 1e-4 on log), and the slopes of the
        // conditional minimisers of f = A dx^2 + 2B dx dy + C dy^2 (dy* = -B/C dx on a vertical edge, dx* = -B/A dy on a
        // horizontal one).  A conic that is not positive definite (fp32 cancellation in det for absurdly large
        // Gaussians) gets pmin = -inf: never culled, the exact per-pixel tests decide.
        const bool pd = con_x > 0.f && con_z > 0.f && (con_x * con_z - con_y * con_y) > 0.f;
        float pmin = 1.0f;  // opacity < 1/255: no pixel can reach alpha >= 1/255 (power <= 0)
        if (opacity >= (1.0f / 255.0f)) pmin = -logf(255.0f * opacity) - 1e-4f;
        if (!pd) pmin = -__int_as_float(0x7f800000);
        o.cull[i] = make_float4(pmin, pd ? -con_y / con_z : 0.f, pd ? -con_y / con_x : 0.f, 0.f);
      }
    }
  }
  o.radii[i] = my_rad;
  return e;
}

// ------------------------------------------------------------------------------------------------------------------
// Warp-cooperative emission of (tile << 32 | depth bits, Gaussian id) for the 32 Gaussians of a warp.
// The warp's output slots [wbase, wbase + wtotal) are consecutive: lane l writes slots l, l+32, ... (coalesced 8- and
// 4-byte stores; the owner of a slot is found by a 5-step binary search over the warp's inclusive prefix).  Every
// Gaussian also adds its tile rectangle to the difference grid the per-tile sort is planned from (reference:
// duplicateWithKeys, gaussian_rasterizer_forward.cu:45-73, emission order y-major within the rect, Gaussians in index
// order).
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void emit_warp(int lane, const Emit& e, uint32_t id, uint32_t excl_in_warp, uint32_t wbase,
                                          uint32_t wtotal, int gx, const BinningOut& b, skgs_raster_header* hdr) {
  if (wtotal == 0) return;
  if ((uint64_t)wbase + wtotal > (uint64_t)b.R_cap) {  // arena too small: flag, keep what fits out of bounds-safe
    if (lane == 0) hdr->overflow = 1;
  }
  if (e.cnt > 0) {
    // +1 on the rectangle [x0, x0+w) x [y0, y0+h) as four corner updates: 4 atomics per Gaussian instead of one per
    // key, and a tile that thousands of Gaussians cover is not one hot address (their corners are spread out)
    const int gs = gx + 1, h = (int)e.cnt / e.w, cs = b.cell_stride;
    atomicAdd(&b.grid[(e.y0 * gs + e.x0) * cs], 1);
    atomicAdd(&b.grid[(e.y0 * gs + e.x0 + e.w) * cs], -1);
    atomicAdd(&b.grid[((e.y0 + h) * gs + e.x0) * cs], -1);
    atomicAdd(&b.grid[((e.y0 + h) * gs + e.x0 + e.w) * cs], 1);
  }
  const uint32_t incl = excl_in_warp + e.cnt;
  for (uint32_t s0 = 0; s0 < wtotal; s0 += 32) {
    const uint32_t s = s0 + lane;
    int g = 0;  // owner = number of lanes whose inclusive prefix is <= s
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
      const uint32_t v = __shfl_sync(FULL, incl, g + step - 1);
      if (v <= s) g += step;
    }
    const uint32_t g_excl = __shfl_sync(FULL, excl_in_warp, g);
    const int g_x0 = __shfl_sync(FULL, e.x0, g), g_y0 = __shfl_sync(FULL, e.y0, g), g_w = __shfl_sync(FULL, e.w, g);
    const uint32_t g_d = __shfl_sync(FULL, e.dbits, g), g_id = __shfl_sync(FULL, id, g);
    if (s < wtotal) {
      const uint32_t k = s - g_excl;
      const uint32_t row = k / (uint32_t)g_w;
      const uint32_t tile = (uint32_t)((g_y0 + (int)row) * gx + g_x0 + (int)(k - row * (uint32_t)g_w));
      const uint32_t slot = wbase + s;
      if (slot < b.R_cap) {
        b.keys[slot] = ((uint64_t)tile << 32) | g_d;
        b.vals[slot] = g_id;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Second half of the per-Gaussian forward kernels: block scan of the tile counts, decoupled look-back across CTAs (one
// warp inspects 32 predecessors per step), then - when EMIT - the warp-cooperative key emission and the radix plan.
// ------------------------------------------------------------------------------------------------------------------
template <bool EMIT>
__device__ __forceinline__ void scan_and_emit(const RasterParams& rp, const Emit& e, int i, int bid, int num_blocks,
                                              uint32_t* s_warp_sum, uint32_t* s_excl_p,
                                              uint32_t* __restrict__ tiles_touched,
                                              uint32_t* __restrict__ point_offsets, uint64_t* __restrict__ scan_state,
                                              skgs_raster_header* __restrict__ hdr, const BinningOut& bo) {
  const int tid = threadIdx.x;
  if (i < rp.P) tiles_touched[i] = e.cnt;
  const uint32_t touched = e.cnt;

  // ---- block-inclusive scan of `touched`
  const int lane = tid & 31, warp = tid >> 5;
  uint32_t incl = touched;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t n = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) s_warp_sum[warp] = incl;
  const uint32_t vis_ballot = __ballot_sync(FULL, touched > 0);
  if (lane == 0 && vis_ballot) atomicAdd(&hdr->num_visible, (uint32_t)__popc(vis_ballot));
  __syncthreads();
  uint32_t warp_off = 0, block_total = 0;
#pragma unroll
  for (int w = 0; w < PRE_THREADS / 32; w++) {
    const uint32_t s = s_warp_sum[w];
    if (w < warp) warp_off += s;
    block_total += s;
  }
  // ---- decoupled look-back across blocks, one warp inspects 32 predecessors per step
  if (warp == 0) {
    uint64_t excl = 0;
    if (bid == 0) {
      if (lane == 0) st_volatile_u64(&scan_state[0], SCAN_FLAG_INC | (uint64_t)block_total);
    } else {
      if (lane == 0) st_volatile_u64(&scan_state[bid], SCAN_FLAG_AGG | (uint64_t)block_total);
      int j = bid - 1;
      while (true) {
        const int jj = j - lane;
        uint64_t w = SCAN_FLAG_INC;  // lanes before block 0 contribute an inclusive 0
        if (jj >= 0) {
          do {
            w = ld_volatile_u64(&scan_state[jj]);
          } while ((w >> 62) == 0);
        }
        const uint32_t inc_mask = __ballot_sync(FULL, (w >> 62) == 2);
        // sum values of lanes up to and including the first inclusive one
        const int first_inc = inc_mask ? (__ffs(inc_mask) - 1) : 32;
        uint64_t v = (lane <= first_inc) ? (w & SCAN_VAL_MASK) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        excl += v;
        if (inc_mask) break;
        j -= 32;
      }
      if (lane == 0) st_volatile_u64(&scan_state[bid], SCAN_FLAG_INC | (excl + block_total));
    }
    if (lane == 0) {
      *s_excl_p = (uint32_t)excl;
      if (bid == num_blocks - 1) st_volatile_u32(&hdr->num_rendered, (uint32_t)(excl + block_total));
    }
  }
  __syncthreads();
  const uint32_t s_excl = *s_excl_p;
  if (i < rp.P) point_offsets[i] = s_excl + warp_off + incl;
  if (EMIT) {
    const uint32_t wtotal = __shfl_sync(FULL, incl, 31);
    emit_warp(lane, e, (uint32_t)i, incl - touched, s_excl + warp_off, wtotal, rp.gx, bo, hdr);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K1: preprocess + fused decoupled-look-back prefix sum (+ key emission and digit histograms when EMIT)
// ------------------------------------------------------------------------------------------------------------------
template <bool EMIT>
__global__ void __launch_bounds__(PRE_THREADS)
preprocess_scan_kernel(RasterParams rp, const float* __restrict__ means3D, const float* __restrict__ shs,
                       const float* __restrict__ colors_precomp, const float* __restrict__ opacities,
                       const float* __restrict__ scales, const float* __restrict__ rotations,
                       const float* __restrict__ cov3D_precomp, GeomOut go, uint32_t* __restrict__ tiles_touched,
                       uint32_t* __restrict__ point_offsets, uint64_t* __restrict__ scan_state,
                       float4* __restrict__ ggrad, skgs_raster_header* __restrict__ hdr, int num_blocks, BinningOut bo) {
  __shared__ int s_bid;
  __shared__ uint32_t s_warp_sum[PRE_THREADS / 32];
  __shared__ uint32_t s_excl;
  __shared__ float s_V[16], s_P[16], s_cam[3];
  const int tid = threadIdx.x;
  pdl_wait();
  pdl_trigger();
  // dynamic block id: the look-back below requires that block b only ever waits on blocks that already started
  if (tid == 0) s_bid = (int)atomicAdd(&hdr->scan_ticket, 1u);
  if (tid < 16) {
    s_V[tid] = rp.view[tid];
    s_P[tid] = rp.proj[tid];
  }
  if (tid < 3) s_cam[tid] = rp.campos[tid];
  __syncthreads();
  const int bid = s_bid;
  const int i = bid * PRE_THREADS + tid;

  Emit e;
  e.cnt = 0; e.x0 = 0; e.y0 = 0; e.w = 0; e.dbits = 0;
  if (i < rp.P) {
    // backward accumulators start at zero (composite_bwd adds into them, preprocess_bwd consumes and re-zeroes them)
    ggrad[3 * (size_t)i] = ggrad[3 * (size_t)i + 1] = ggrad[3 * (size_t)i + 2] = make_float4(0.f, 0.f, 0.f, 0.f);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, qx = 0.f, qy = 0.f, qz = 0.f, qr = 1.f;
    if (cov3D_precomp == nullptr) {
      const float4 q = *reinterpret_cast<const float4*>(rotations + 4 * (size_t)i);
      if (rp.quat_wxyz) {
        qr = q.x; qx = q.y; qy = q.z; qz = q.w;
      } else {
        qx = q.x; qy = q.y; qz = q.z; qr = q.w;
      }
      s0 = rp.mod * scales[3 * (size_t)i];
      s1 = rp.mod * scales[3 * (size_t)i + 1];
      s2 = rp.mod * scales[3 * (size_t)i + 2];
    }
    e = preprocess_gaussian(rp, s_V, s_P, s_cam, i, means3D[3 * (size_t)i], means3D[3 * (size_t)i + 1],
                            means3D[3 * (size_t)i + 2], opacities[i],
                            cov3D_precomp ? cov3D_precomp + 6 * (size_t)i : nullptr, s0, s1, s2, qx, qy, qz, qr, shs,
                            colors_precomp, go);
  }
  scan_and_emit<EMIT>(rp, e, i, bid, num_blocks, s_warp_sum, &s_excl, tiles_touched, point_offsets, scan_state, hdr,
                      bo);
}

// ------------------------------------------------------------------------------------------------------------------
// K1f: the whole per-Gaussian forward in ONE kernel - K nearest joints + skinning weights + linear blend (deform.cuh),
// output assembly, preprocess, prefix sum, key emission.  Replaces lbs_fwd_kernel -> assemble_fwd_kernel ->
// preprocess_scan_kernel (two 40 B / Gaussian round trips through HBM and two kernel boundaries); the joint table comes
// from fk_table_kernel.  Same device functions, same flags: bit-identical to the three-kernel path.
// Outputs the backward needs are still written: weights / indices (LBS), d_rot (assembly), points / scales / rotations
// (rasterizer), opacities (API parity).
// ------------------------------------------------------------------------------------------------------------------
struct DeformIO {
  int M, mode;
  float temperature;
  const float* table;                                   // joint table (fk_table_kernel)
  const float *xyz, *scaling, *rotation, *opacity, *sp_W;  // canonical parameters
  float *points, *scales, *rotations, *opacities;       // assembled Gaussians
  float *d_rot, *weights;                               // kept for the backward
  int64_t* indices;
};

#ifndef SKGS_DP_MINBLOCKS
#define SKGS_DP_MINBLOCKS 3   // 3 CTAs x 256 threads per SM: the whole grid of c2 is resident at once (A/B: -15 us)
#endif
template <int KT>
__global__ void __launch_bounds__(PRE_THREADS, SKGS_DP_MINBLOCKS)
deform_preprocess_kernel(RasterParams rp, DeformIO io, const float* __restrict__ shs, GeomOut go,
                         uint32_t* __restrict__ tiles_touched, uint32_t* __restrict__ point_offsets,
                         uint64_t* __restrict__ scan_state, float4* __restrict__ ggrad,
                         skgs_raster_header* __restrict__ hdr, int num_blocks, BinningOut bo) {
  extern __shared__ float s_table[];
  __shared__ int s_bid;
  __shared__ uint32_t s_warp_sum[PRE_THREADS / 32];
  __shared__ uint32_t s_excl;
  __shared__ float s_V[16], s_P[16], s_cam[3];
  const int tid = threadIdx.x;
  pdl_wait();
  pdl_trigger();
  if (tid == 0) s_bid = (int)atomicAdd(&hdr->scan_ticket, 1u);
  if (tid < 16) {
    s_V[tid] = rp.view[tid];
    s_P[tid] = rp.proj[tid];
  }
  if (tid < 3) s_cam[tid] = rp.campos[tid];
  load_joint_table(s_table, io.table, io.M);
  __syncthreads();
  const JointTable jt = joint_table_view(s_table, io.M);
  const int bid = s_bid;
  const int i = bid * PRE_THREADS + tid;
  Emit e;
  e.cnt = 0; e.x0 = 0; e.y0 = 0; e.w = 0; e.dbits = 0;
  if (i < rp.P) {
    ggrad[3 * (size_t)i] = ggrad[3 * (size_t)i + 1] = ggrad[3 * (size_t)i + 2] = make_float4(0.f, 0.f, 0.f, 0.f);
    const size_t i3 = 3 * (size_t)i;
    const float x0 = io.xyz[i3], y0 = io.xyz[i3 + 1], z0 = io.xyz[i3 + 2];
    LbsOut<KT> o;
    lbs_gaussian<KT>(jt, io.M, io.mode, io.temperature, io.sp_W ? io.sp_W + (size_t)i * io.M : nullptr, x0, y0, z0, o);
#pragma unroll
    for (int k = 0; k < KT; k++) {
      io.weights[(size_t)i * KT + k] = o.w[k];
      io.indices[(size_t)i * KT + k] = (int64_t)o.idx[k];
    }
    const float4 d_rot = make_float4(o.r0, o.r1, o.r2, o.r3);
    *reinterpret_cast<float4*>(io.d_rot + 4 * (size_t)i) = d_rot;
    const Assembled a = assemble_gaussian(x0, y0, z0, io.scaling[i3], io.scaling[i3 + 1], io.scaling[i3 + 2],
                                          *reinterpret_cast<const float4*>(io.rotation + 4 * (size_t)i), io.opacity[i],
                                          o.dx, o.dy, o.dz, d_rot, o.s0, o.s1, o.s2);
    io.points[i3] = a.px; io.points[i3 + 1] = a.py; io.points[i3 + 2] = a.pz;
    io.scales[i3] = a.sx; io.scales[i3 + 1] = a.sy; io.scales[i3 + 2] = a.sz;
    *reinterpret_cast<float4*>(io.rotations + 4 * (size_t)i) = make_float4(a.qx, a.qy, a.qz, a.qw);
    io.opacities[i] = a.opacity;
    e = preprocess_gaussian(rp, s_V, s_P, s_cam, i, a.px, a.py, a.pz, a.opacity, nullptr, rp.mod * a.sx, rp.mod * a.sy,
                            rp.mod * a.sz, a.qx, a.qy, a.qz, a.qw, shs, nullptr, go);
  }
  scan_and_emit<true>(rp, e, i, bid, num_blocks, s_warp_sum, &s_excl, tiles_touched, point_offsets, scan_state, hdr,
                      bo);
}

// ------------------------------------------------------------------------------------------------------------------
// K2 (split API only: skgs_raster_forward_geometry + _render): duplicate with keys from the stored geometry
// ------------------------------------------------------------------------------------------------------------------
constexpr int DUP_THREADS = 256;

__global__ void __launch_bounds__(DUP_THREADS)
duplicate_keys_kernel(int P, int gx, int gy, const int32_t* __restrict__ radii, const float2* __restrict__ means2D,
                      const float* __restrict__ depths, const uint32_t* __restrict__ point_offsets,
                      skgs_raster_header* __restrict__ hdr, BinningOut bo) {
  const int tid = threadIdx.x, lane = tid & 31;
  pdl_wait();
  pdl_trigger();
  const int i = blockIdx.x * DUP_THREADS + tid;
  Emit e;
  e.cnt = 0; e.x0 = 0; e.y0 = 0; e.w = 0; e.dbits = 0;
  uint32_t incl_g = 0;
  if (i < P) {
    incl_g = point_offsets[i];
    const int rad = radii[i];
    if (rad > 0) {
      const float2 p = means2D[i];
      int x0, y0, x1, y1;
      get_rect(p.x, p.y, rad, gx, gy, x0, y0, x1, y1);
      e.cnt = (uint32_t)((x1 - x0) * (y1 - y0));
      e.x0 = x0; e.y0 = y0; e.w = x1 - x0;
      e.dbits = __float_as_uint(depths[i]);
    }
  }
  // offsets inside the warp from the global inclusive prefix: lanes past P carry cnt = 0 and the last valid prefix
  const uint32_t excl_g = incl_g - e.cnt;
  const uint32_t wbase = __shfl_sync(FULL, excl_g, 0);
  uint32_t incl_w = e.cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t n = __shfl_up_sync(FULL, incl_w, o);
    if (lane >= o) incl_w += n;
  }
  const uint32_t wtotal = __shfl_sync(FULL, incl_w, 31);
  emit_warp(lane, e, (uint32_t)i, incl_w - e.cnt, wbase, wtotal, gx, bo, hdr);
}

// ------------------------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------------------------
static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

static GeomOut geom_out(char* geom, const skgs_raster_layout& lay, int32_t* radii) {
  GeomOut o;
  o.radii = radii;
  o.means2D = reinterpret_cast<float2*>(geom + lay.means2D);
  o.depths = reinterpret_cast<float*>(geom + lay.depths);
  o.cov3Ds = reinterpret_cast<float*>(geom + lay.cov3D);
  o.conic_opacity = reinterpret_cast<float4*>(geom + lay.conic_opacity);
  o.rgbd = reinterpret_cast<float4*>(geom + lay.rgbd);
  o.cull = reinterpret_cast<float4*>(geom + lay.cull);
  o.clamped = reinterpret_cast<uint8_t*>(geom + lay.clamped);
  return o;
}

static void binning_out(const RasterParams& rp, char* binning, const skgs_raster_layout& lay, int64_t R_cap,
                        BinningOut& b) {
  b.keys = reinterpret_cast<uint64_t*>(binning + lay.keys);
  b.vals = reinterpret_cast<uint32_t*>(binning + lay.vals);
  b.grid = reinterpret_cast<int*>(binning + lay.tile_grid);
  b.cell_stride = tile_cell_stride();
  b.R_cap = (uint32_t)R_cap;
}

// the difference grid and the per-tile fill cursors (adjacent in the arena) start every forward at zero
static cudaError_t reset_tile_counters(char* binning, const skgs_raster_layout& lay, cudaStream_t st) {
  return cudaMemsetAsync(binning + lay.tile_grid, 0, lay.binning_bytes - lay.tile_grid, st);
}

// preprocess + scan; with a binning arena (binning != NULL, R_cap > 0) the keys are emitted by the same kernel
int launch_preprocess_scan(const RasterParams& rp, const float* means3D, const float* shs, const float* colors_precomp,
                           const float* opacities, const float* scales, const float* rotations,
                           const float* cov3D_precomp, char* geom, const skgs_raster_layout& lay, int32_t* radii,
                           uint32_t* num_rendered_host, char* binning, char* img, int64_t R_cap, cudaStream_t st) {
  auto* hdr = reinterpret_cast<skgs_raster_header*>(geom + lay.header);
  const int nblocks = (rp.P + PRE_THREADS - 1) / PRE_THREADS;
  const bool emit = binning != nullptr && R_cap > 0;
  // header and scan_state are adjacent in the arena: one memset resets the tickets, the flags and the counters
  SKGS_CUDA(cudaMemsetAsync(geom + lay.header, 0, lay.means2D - lay.header, st));
  BinningOut bo = {};
  if (emit) {
    binning_out(rp, binning, lay, R_cap, bo);
    SKGS_CUDA(reset_tile_counters(binning, lay, st));
  }
  if (rp.P > 0) {
    ProfScope prof_("preprocess_scan_kernel", st);
    GeomOut go = geom_out(geom, lay, radii);
    auto* tt = reinterpret_cast<uint32_t*>(geom + lay.tiles_touched);
    auto* po = reinterpret_cast<uint32_t*>(geom + lay.point_offsets);
    auto* ss = reinterpret_cast<uint64_t*>(geom + lay.scan_state);
    auto* gg = reinterpret_cast<float4*>(geom + lay.geom_grads);
    if (emit)
      SKGS_CUDA(launch_pdl(preprocess_scan_kernel<true>, dim3(nblocks), dim3(PRE_THREADS), 0, st, rp, means3D, shs,
                           colors_precomp, opacities, scales, rotations, cov3D_precomp, go, tt, po, ss, gg, hdr,
                           nblocks, bo));
    else
      SKGS_CUDA(launch_pdl(preprocess_scan_kernel<false>, dim3(nblocks), dim3(PRE_THREADS), 0, st, rp, means3D, shs,
                           colors_precomp, opacities, scales, rotations, cov3D_precomp, go, tt, po, ss, gg, hdr,
                           nblocks, bo));
    SKGS_CHECK_LAUNCH("preprocess_scan_kernel");
  }
  if (num_rendered_host)
    SKGS_CUDA(cudaMemcpyAsync(num_rendered_host, hdr, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  return SKGS_OK;
}


// the fused per-Gaussian forward (K1f); the joint table must have been produced on the same stream (launch_fk_table)
int launch_deform_preprocess(const RasterParams& rp, const skgs_skeleton* sk, const float* table, const float* xyz,
                             const float* scaling, const float* rotation, const float* opacity_logit, const float* shs,
                             float* points, float* scales, float* rotations, float* opacities, float* d_rot,
                             float* weights, int64_t* indices, char* geom, const skgs_raster_layout& lay,
                             int32_t* radii, char* binning, char* img, int64_t R_cap, cudaStream_t st) {
  auto* hdr = reinterpret_cast<skgs_raster_header*>(geom + lay.header);
  const int nblocks = (rp.P + PRE_THREADS - 1) / PRE_THREADS;
  SKGS_CHECK_ARG(binning != nullptr && R_cap > 0 && rp.P > 0, "fused forward needs P > 0 and a binning arena");
  SKGS_CUDA(cudaMemsetAsync(geom + lay.header, 0, lay.means2D - lay.header, st));
  BinningOut bo = {};
  binning_out(rp, binning, lay, R_cap, bo);
  SKGS_CUDA(reset_tile_counters(binning, lay, st));
  DeformIO io;
  io.M = sk->M; io.mode = sk->mode; io.temperature = sk->temperature; io.table = table;
  io.xyz = xyz; io.scaling = scaling; io.rotation = rotation; io.opacity = opacity_logit; io.sp_W = sk->sp_W;
  io.points = points; io.scales = scales; io.rotations = rotations; io.opacities = opacities;
  io.d_rot = d_rot; io.weights = weights; io.indices = indices;
  GeomOut go = geom_out(geom, lay, radii);
  auto* tt = reinterpret_cast<uint32_t*>(geom + lay.tiles_touched);
  auto* po = reinterpret_cast<uint32_t*>(geom + lay.point_offsets);
  auto* ss = reinterpret_cast<uint64_t*>(geom + lay.scan_state);
  auto* gg = reinterpret_cast<float4*>(geom + lay.geom_grads);
  const size_t smem = (size_t)sk->M * JT_FLOATS * sizeof(float);
  {
    ProfScope prof_("deform_preprocess_kernel", st);
#define SKGS_DP_CASE(KK)                                                                                             \
  case KK: {                                                                                                         \
    static size_t smem_set = 0;                                                                                      \
    if (smem + 12 * 1024 > 48 * 1024 && smem > smem_set) {                                                           \
      SKGS_CUDA(cudaFuncSetAttribute(deform_preprocess_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                     (int)smem));                                                                    \
      smem_set = smem;                                                                                               \
    }                                                                                                                \
    SKGS_CUDA(launch_pdl(deform_preprocess_kernel<KK>, dim3(nblocks), dim3(PRE_THREADS), smem, st, rp, io, shs, go, \
                         tt, po, ss, gg, hdr, nblocks, bo));                                                         \
  } break;
    switch (sk->K) {
      SKGS_DP_CASE(1) SKGS_DP_CASE(2) SKGS_DP_CASE(3) SKGS_DP_CASE(4) SKGS_DP_CASE(5) SKGS_DP_CASE(6) SKGS_DP_CASE(7)
      SKGS_DP_CASE(8)
      default: SKGS_CHECK_ARG(false, "K=%d out of range", sk->K);
    }
#undef SKGS_DP_CASE
    SKGS_CHECK_LAUNCH("deform_preprocess_kernel");
  }
  return SKGS_OK;
}

// binning: (key emission from the stored geometry when `emit`: the split API and the capacity-retry path, then) the
// tile-segmented sort of tile_sort.cu, which also plans the work order of the compositing kernels
int launch_binning(const RasterParams& rp, char* geom, char* binning, char* img, const skgs_raster_layout& lay,
                   const int32_t* radii, int64_t R_cap, int64_t R_hint, bool emit, uint32_t* num_rendered_host,
                   cudaStream_t st) {
  auto* hdr = reinterpret_cast<skgs_raster_header*>(geom + lay.header);
  const int tiles = rp.gx * rp.gy;
  if (rp.P == 0 || R_cap <= 0) {  // nothing to bin: empty ranges, launch_tile_order builds the (all empty) schedule
    SKGS_CUDA(cudaMemsetAsync(img + lay.ranges, 0, (size_t)tiles * sizeof(uint2), st));
    SKGS_CUDA(cudaMemsetAsync(img + lay.work_counters, 0, 8 * sizeof(uint32_t), st));
    return launch_tile_order(rp, img, lay, st);
  }
  if (emit) {
    BinningOut bo = {};
    binning_out(rp, binning, lay, R_cap, bo);
    // reset what a previous render stage on the same geometry may have left: the overflow flag, the tile counters
    SKGS_CUDA(cudaMemsetAsync(&hdr->overflow, 0, sizeof(uint32_t), st));
    SKGS_CUDA(reset_tile_counters(binning, lay, st));
    ProfScope prof_("duplicate_keys_kernel", st);
    SKGS_CUDA(launch_pdl(duplicate_keys_kernel, dim3((rp.P + DUP_THREADS - 1) / DUP_THREADS), dim3(DUP_THREADS), 0, st,
                         rp.P, rp.gx, rp.gy, radii, reinterpret_cast<const float2*>(geom + lay.means2D),
                         reinterpret_cast<const float*>(geom + lay.depths),
                         reinterpret_cast<const uint32_t*>(geom + lay.point_offsets), hdr, bo));
    SKGS_CHECK_LAUNCH("duplicate_keys_kernel");
  }
  if (num_rendered_host)
    SKGS_CUDA(cudaMemcpyAsync(num_rendered_host, hdr, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  return launch_tile_binning(rp, geom, binning, img, lay, R_cap, R_hint, st);
}

}  // namespace skgs
