// raster_fwd.cu - forward half of the tile rasterizer for sm_100a:
//   K1 preprocess_scan_kernel : projection, EWA cov2D, conic, radius, tile rect, SH->RGB, footprint-culling record
//                               + fused decoupled-look-back prefix sum of tiles_touched (no scan kernel, no host sync
//                               for R) + warp-cooperative emission of (tile << 32 | depth bits, id) with all radix digit
//                               histograms + the plan of the radix passes (last CTA done)
//   K2 duplicate_keys_kernel  : the same emission from stored geometry (split API / capacity retry only)
//   K3 onesweep_pass_kernel   : one kernel per 8-bit digit, chained scan across key tiles with a warp-parallel
//                               look-back, stable warp-level multi-split ranking (match.any), smem-staged coalesced
//                               scatter; constant digits are skipped on the device; the last pass writes the tile ranges
//   (tile order + compositing kernels live in composite.cu)
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
constexpr int MAX_PASSES = 8;
constexpr uint64_t SCAN_FLAG_AGG = 1ull << 62;
constexpr uint64_t SCAN_FLAG_INC = 2ull << 62;
constexpr uint64_t SCAN_VAL_MASK = (1ull << 62) - 1;
constexpr uint32_t FULL = 0xffffffffu;
constexpr uint32_t RANGE_UNSET = 0xffffffffu;  // ranges[t].x before the last radix pass has seen tile t

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
  uint64_t* keys;              // buffer a
  uint32_t* vals;
  uint32_t* hist;              // [passes][256] (radix binning)
  uint32_t* tile_counts;       // [(gy+1)][(gx+1)] 2-D difference grid of the tile rectangles (tile-segmented binning:
                               // its prefix sum is the number of keys per tile); NULL selects the radix bookkeeping
  uint2* ranges;               // [tiles] reset to (RANGE_UNSET, 0)
  uint32_t* counters;          // [2] compositing work tickets, reset to 0
  uint32_t R_cap;
  int passes;
  int tiles;
  int cell_stride;             // ints between two cells of the difference grid
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
// 4-byte stores; the owner of a slot is found by a 5-step binary search over the warp's inclusive prefix).  The digit
// histograms of the radix sort are accumulated on the way (reference: duplicateWithKeys,
// gaussian_rasterizer_forward.cu:45-73, emission order y-major within the rect, Gaussians in index order).
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void emit_warp(int lane, const Emit& e, uint32_t id, uint32_t excl_in_warp, uint32_t wbase,
                                          uint32_t wtotal, int gx, const BinningOut& b, uint32_t* s_hist,
                                          skgs_raster_header* hdr) {
  if (wtotal == 0) return;
  if ((uint64_t)wbase + wtotal > (uint64_t)b.R_cap) {  // arena too small: flag, keep what fits out of bounds-safe
    if (lane == 0) hdr->overflow = 1;
  }
  const bool by_tile = b.tile_counts != nullptr;
  if (by_tile && e.cnt > 0) {
    // +1 on the rectangle [x0, x0+w) x [y0, y0+h) as four corner updates: 4 atomics per Gaussian instead of one per
    // key, and a tile that thousands of Gaussians cover is not one hot address (their corners are spread out)
    int* dg = reinterpret_cast<int*>(b.tile_counts);
    const int gs = gx + 1, h = (int)e.cnt / e.w, cs = b.cell_stride;
    atomicAdd(&dg[(e.y0 * gs + e.x0) * cs], 1);
    atomicAdd(&dg[(e.y0 * gs + e.x0 + e.w) * cs], -1);
    atomicAdd(&dg[((e.y0 + h) * gs + e.x0) * cs], -1);
    atomicAdd(&dg[((e.y0 + h) * gs + e.x0 + e.w) * cs], 1);
  }
  if (!by_tile && e.cnt > 0) {  // depth digits: the same for all of this Gaussian's keys
#pragma unroll
    for (int p = 0; p < 4; p++) atomicAdd(&s_hist[p * 256 + ((e.dbits >> (8 * p)) & 255u)], e.cnt);
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
    const bool valid = s < wtotal;
    uint32_t tile = 0;
    if (valid) {
      const uint32_t k = s - g_excl;
      const uint32_t row = k / (uint32_t)g_w;
      tile = (uint32_t)((g_y0 + (int)row) * gx + g_x0 + (int)(k - row * (uint32_t)g_w));
      const uint32_t slot = wbase + s;
      if (slot < b.R_cap) {
        b.keys[slot] = ((uint64_t)tile << 32) | g_d;
        b.vals[slot] = g_id;
      }
    }
    if (by_tile) continue;
    // tile digits: neighbouring slots mostly share the high digit -> aggregate equal digits before the atomic
    const uint32_t act = __ballot_sync(FULL, valid);
    for (int p = 4; p < b.passes; p++) {
      const uint32_t d = valid ? ((tile >> (8 * (p - 4))) & 255u) : 0xffffffffu;
      const uint32_t m = __match_any_sync(FULL, d) & act;
      if (valid && lane == __ffs(m) - 1) atomicAdd(&s_hist[p * 256 + d], (uint32_t)__popc(m));
    }
  }
}

// Flush the CTA's digit histograms; the LAST CTA of the emitting kernel to get here writes the plan of the radix
// passes: a pass whose digit is identical for all n keys is skipped (never the last one, which also produces the tile
// ranges), the others ping-pong a -> b -> a ...; final_buf = where the sorted lists end up.
__device__ __forceinline__ void finish_emission(int tid, int nthreads, uint32_t* s_hist, const BinningOut& b,
                                                skgs_raster_header* hdr, uint32_t num_ctas, uint32_t* s_flag) {
  if (b.tile_counts != nullptr) return;  // tile-segmented binning: the per-tile counts are all the next stage needs
  __syncthreads();
  for (int k = tid; k < b.passes * 256; k += nthreads) {
    const uint32_t c = s_hist[k];
    if (c) atomicAdd(&b.hist[k], c);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) *s_flag = (atomicAdd(&hdr->emit_done, 1u) == num_ctas - 1u) ? 1u : 0u;
  __syncthreads();
  if (*s_flag == 0u) return;
  __threadfence();
  const uint32_t R = ld_volatile_u32(&hdr->num_rendered);
  const uint32_t n = min(R, b.R_cap);
  uint32_t constant_mask = 0;  // bit p: pass p has one digit holding all n keys
  for (int p = 0; p < b.passes; p++) {
    bool full = false;
    for (int d = tid; d < 256; d += nthreads) full |= (n > 0 && ld_volatile_u32(&b.hist[p * 256 + d]) == n);
    if (__syncthreads_or(full)) constant_mask |= 1u << p;
  }
  if (tid == 0) {
    uint32_t buf = 0;
    for (int p = 0; p < b.passes; p++) {
      const bool skip = ((constant_mask >> p) & 1u) && p != b.passes - 1;
      hdr->sort_plan[p] = (skip ? 1u : 0u) | (buf << 1);
      if (!skip) buf ^= 1u;
    }
    hdr->final_buf = buf;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Second half of the per-Gaussian forward kernels: block scan of the tile counts, decoupled look-back across CTAs (one
// warp inspects 32 predecessors per step), then - when EMIT - the warp-cooperative key emission and the radix plan.
// ------------------------------------------------------------------------------------------------------------------
template <bool EMIT>
__device__ __forceinline__ void scan_and_emit(const RasterParams& rp, const Emit& e, int i, int bid, int num_blocks,
                                              uint32_t* s_warp_sum, uint32_t* s_excl_p, uint32_t* s_flag,
                                              uint32_t* s_hist, uint32_t* __restrict__ tiles_touched,
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
    emit_warp(lane, e, (uint32_t)i, incl - touched, s_excl + warp_off, wtotal, rp.gx, bo, s_hist, hdr);
    finish_emission(tid, PRE_THREADS, s_hist, bo, hdr, (uint32_t)num_blocks, s_flag);
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
  __shared__ uint32_t s_flag;
  __shared__ float s_V[16], s_P[16], s_cam[3];
  __shared__ uint32_t s_hist[EMIT ? MAX_PASSES * 256 : 1];
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
  if (EMIT)
    for (int k = tid; k < bo.passes * 256; k += PRE_THREADS) s_hist[k] = 0;
  __syncthreads();
  const int bid = s_bid;
  const int i = bid * PRE_THREADS + tid;
  if (EMIT) {  // this CTA's share of the per-forward resets: tile ranges, compositing tickets
    const int per = (bo.tiles + num_blocks - 1) / num_blocks;
    for (int t = bid * per + tid; t < min(bo.tiles, (bid + 1) * per); t += PRE_THREADS)
      bo.ranges[t] = make_uint2(RANGE_UNSET, 0u);
    if (bid == 0 && tid < 2) bo.counters[tid] = 0u;
  }

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
  scan_and_emit<EMIT>(rp, e, i, bid, num_blocks, s_warp_sum, &s_excl, &s_flag, s_hist, tiles_touched, point_offsets,
                      scan_state, hdr, bo);
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
  __shared__ uint32_t s_flag;
  __shared__ float s_V[16], s_P[16], s_cam[3];
  __shared__ uint32_t s_hist[MAX_PASSES * 256];
  const int tid = threadIdx.x;
  pdl_wait();
  pdl_trigger();
  if (tid == 0) s_bid = (int)atomicAdd(&hdr->scan_ticket, 1u);
  if (tid < 16) {
    s_V[tid] = rp.view[tid];
    s_P[tid] = rp.proj[tid];
  }
  if (tid < 3) s_cam[tid] = rp.campos[tid];
  for (int k = tid; k < bo.passes * 256; k += PRE_THREADS) s_hist[k] = 0;
  load_joint_table(s_table, io.table, io.M);
  __syncthreads();
  const JointTable jt = joint_table_view(s_table, io.M);
  const int bid = s_bid;
  const int i = bid * PRE_THREADS + tid;
  {  // this CTA's share of the per-forward resets: tile ranges, compositing tickets
    const int per = (bo.tiles + num_blocks - 1) / num_blocks;
    for (int t = bid * per + tid; t < min(bo.tiles, (bid + 1) * per); t += PRE_THREADS)
      bo.ranges[t] = make_uint2(RANGE_UNSET, 0u);
    if (bid == 0 && tid < 2) bo.counters[tid] = 0u;
  }
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
  scan_and_emit<true>(rp, e, i, bid, num_blocks, s_warp_sum, &s_excl, &s_flag, s_hist, tiles_touched, point_offsets,
                      scan_state, hdr, bo);
}

// ------------------------------------------------------------------------------------------------------------------
// K2 (split API only: skgs_raster_forward_geometry + _render): duplicate with keys from the stored geometry
// ------------------------------------------------------------------------------------------------------------------
constexpr int DUP_THREADS = 256;

__global__ void __launch_bounds__(DUP_THREADS)
duplicate_keys_kernel(int P, int gx, int gy, const int32_t* __restrict__ radii, const float2* __restrict__ means2D,
                      const float* __restrict__ depths, const uint32_t* __restrict__ point_offsets,
                      skgs_raster_header* __restrict__ hdr, BinningOut bo) {
  __shared__ uint32_t s_hist[MAX_PASSES * 256];
  __shared__ uint32_t s_flag;
  const int tid = threadIdx.x, lane = tid & 31;
  pdl_wait();
  pdl_trigger();
  for (int k = tid; k < bo.passes * 256; k += DUP_THREADS) s_hist[k] = 0;
  {
    const int per = (bo.tiles + (int)gridDim.x - 1) / (int)gridDim.x;
    for (int t = blockIdx.x * per + tid; t < min(bo.tiles, ((int)blockIdx.x + 1) * per); t += DUP_THREADS)
      bo.ranges[t] = make_uint2(RANGE_UNSET, 0u);
    if (blockIdx.x == 0 && tid < 2) bo.counters[tid] = 0u;
  }
  __syncthreads();
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
  emit_warp(lane, e, (uint32_t)i, incl_w - e.cnt, wbase, wtotal, gx, bo, s_hist, hdr);
  finish_emission(tid, DUP_THREADS, s_hist, bo, hdr, gridDim.x, &s_flag);
}

// ------------------------------------------------------------------------------------------------------------------
// K3: onesweep radix pass (8-bit digit), stable.  Status word: [31:29] pass tag, [28:27] flag, [26:0] count.
//   * a pass the plan marks as skipped returns at once (its digit is the same for every key: identity permutation);
//   * chained scan with a WARP-PARALLEL look-back: warp w owns digits 32w..32w+31, lane l fetches the 128-byte status
//     slab of predecessor tile-1-l (eight 16-byte volatile loads), the 32 x 32 words are transposed through shared
//     memory and lane k walks the 32 predecessors of digit 32w+k: 32 predecessors per L2 round trip instead of one (with
//     every CTA of a pass resident at once, a serial walk costs tiles/2 dependent round trips - 19 us of a 20 us pass);
//   * the LAST pass also produces the tile ranges: inside a CTA the keys of one digit run are fully sorted, so tile
//     boundaries are visible locally; the first / last entry of every (CTA, tile) run does an atomicMin / atomicMax on
//     ranges[tile] (identifyTileRanges of the reference, gaussian_rasterizer_forward.cu:77-94, without a kernel).
// ------------------------------------------------------------------------------------------------------------------
#ifndef SKGS_OS_THREADS
#define SKGS_OS_THREADS 512
#endif
#ifndef SKGS_OS_ITEMS
#define SKGS_OS_ITEMS 12   // 512 x 12 = 6144 keys per CTA tile, 16 warps: every phase is latency bound, warps hide it
#endif
constexpr int OS_THREADS = SKGS_OS_THREADS;
constexpr int OS_ITEMS = SKGS_OS_ITEMS;
constexpr int OS_TILE = OS_THREADS * OS_ITEMS;  // keys per CTA tile
constexpr int OS_WARPS = OS_THREADS / 32;
constexpr int OS_DIGITS = 256;                  // threads 0..255 also own one digit each
constexpr int OS_DWARPS = OS_DIGITS / 32;
constexpr uint32_t OS_FLAG_AGG = 1u, OS_FLAG_INC = 2u;
constexpr uint32_t OS_VAL_MASK = (1u << 27) - 1;
static_assert(OS_THREADS >= OS_DIGITS && OS_THREADS % 32 == 0, "one thread per digit");
static_assert(OS_TILE >= 2048, "api.cu sizes the look-back words for tiles of at least 2048 keys");
static_assert(OS_TILE * 12 >= OS_DWARPS * 32 * 33 * 4, "the look-back slabs alias the key + value staging area");

// debug: per-tile phase timestamps of one radix pass (tools/sort_trace.py)
__device__ unsigned long long* g_os_trace = nullptr;
__device__ __forceinline__ void os_trace(uint32_t tile, int phase) {
  if (g_os_trace != nullptr && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_os_trace[(size_t)tile * 8 + phase] = t;
  }
}

struct OnesweepSmem {
  uint64_t keys[OS_TILE];   // reorder staging; during the look-back keys + vals hold OS_DWARPS slabs of 32 x 33 words
  uint32_t vals[OS_TILE];
  uint32_t whist[OS_WARPS][OS_DIGITS];
  uint32_t texcl[OS_DIGITS];   // exclusive prefix of this tile's digit counts
  uint32_t goff[OS_DIGITS];    // global output offset of digit d minus texcl[d]
  uint32_t gbase[OS_DIGITS];   // exclusive prefix of the global digit histogram
  uint32_t warp_tot[OS_DWARPS];
  uint32_t tile;
};

// exclusive scan over the 256 digits, one value per thread of the first 8 warps (all threads must call: barriers)
__device__ __forceinline__ uint32_t digit_exclusive_scan(uint32_t v, int tid, uint32_t* warp_tot) {
  const int lane = tid & 31, warp = tid >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += t;
  }
  if (tid < OS_DIGITS && lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  uint32_t woff = 0;
  if (tid < OS_DIGITS)
    for (int w = 0; w < warp; w++) woff += warp_tot[w];
  __syncthreads();
  return woff + incl - v;
}

__global__ void __launch_bounds__(OS_THREADS)
onesweep_pass_kernel(uint64_t* __restrict__ keys_a, uint32_t* __restrict__ vals_a, uint64_t* __restrict__ keys_b,
                     uint32_t* __restrict__ vals_b, skgs_raster_header* __restrict__ hdr, uint32_t R_cap,
                     const uint32_t* __restrict__ hist, uint32_t* __restrict__ status, int pass, int is_last,
                     uint2* __restrict__ ranges) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  OnesweepSmem& S = *reinterpret_cast<OnesweepSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool dthread = tid < OS_DIGITS;  // owns digit `tid`
  pdl_wait();
  pdl_trigger();
  if (hdr->overflow) return;
  const uint32_t plan = hdr->sort_plan[pass];
  if (plan & 1u) return;  // constant digit: nothing to do, the next pass reads the same buffer
  const bool from_b = (plan >> 1) & 1u;
  const uint64_t* __restrict__ kin = from_b ? keys_b : keys_a;
  const uint32_t* __restrict__ vin = from_b ? vals_b : vals_a;
  uint64_t* __restrict__ kout = from_b ? keys_a : keys_b;
  uint32_t* __restrict__ vout = from_b ? vals_a : vals_b;
  const int shift = 8 * pass;
  const uint32_t tag = (uint32_t)pass;
  uint32_t* ticket = &hdr->sort_ticket[pass];
  const uint32_t n = min(hdr->num_rendered, R_cap);
  const uint32_t num_tiles = (n + OS_TILE - 1) / OS_TILE;
  const uint32_t lanemask_lt = (1u << lane) - 1u;

  // exclusive scan of the global digit histogram
  {
    const uint32_t c = dthread ? hist[tid] : 0u;
    const uint32_t ex = digit_exclusive_scan(c, tid, S.warp_tot);
    if (dthread) S.gbase[tid] = ex;
  }

  while (true) {
    if (tid == 0) S.tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = S.tile;
    if (tile >= num_tiles) break;
    const uint32_t base = tile * OS_TILE;
    const uint32_t cnt = min((uint32_t)OS_TILE, n - base);
    os_trace(tile, 0);

    uint64_t key[OS_ITEMS];
    uint32_t val[OS_ITEMS];
    uint16_t pos[OS_ITEMS];
#pragma unroll
    for (int i = 0; i < OS_ITEMS; i++) {
      const uint32_t idx = warp * (32 * OS_ITEMS) + i * 32 + lane;
      key[i] = idx < cnt ? kin[base + idx] : ~0ull;
    }
#pragma unroll
    for (int i = 0; i < OS_ITEMS; i++) {  // values travel with the keys: issued now, consumed after the look-back
      const uint32_t idx = warp * (32 * OS_ITEMS) + i * 32 + lane;
      val[i] = idx < cnt ? vin[base + idx] : 0u;
    }
    for (int k = tid; k < OS_WARPS * OS_DIGITS; k += OS_THREADS) (&S.whist[0][0])[k] = 0;
    __syncthreads();
    os_trace(tile, 1);
    // ---- stable per-warp ranking (items are warp-striped: item-major, then lane)
#pragma unroll
    for (int i = 0; i < OS_ITEMS; i++) {
      const uint32_t idx = warp * (32 * OS_ITEMS) + i * 32 + lane;
      const bool valid = idx < cnt;
      const uint32_t d = valid ? (uint32_t)((key[i] >> shift) & 255ull) : 0xffffffffu;
      const uint32_t m = __match_any_sync(FULL, d);
      const int leader = __ffs(m) - 1;
      uint32_t old = 0;
      if (valid && lane == leader) {
        old = S.whist[warp][d];
        S.whist[warp][d] = old + __popc(m);
      }
      old = __shfl_sync(FULL, old, leader);
      pos[i] = (uint16_t)(old + __popc(m & lanemask_lt));
      __syncwarp();
    }
    __syncthreads();
    os_trace(tile, 2);
    // ---- per digit: cross-warp exclusive prefix, tile totals, publish, look back
    uint32_t total = 0;
    const int d = tid;
    uint32_t* my = status + (size_t)tile * OS_DIGITS + (dthread ? d : 0);
    if (dthread) {
#pragma unroll
      for (int w = 0; w < OS_WARPS; w++) {
        const uint32_t c = S.whist[w][d];
        S.whist[w][d] = total;
        total += c;
      }
      st_volatile_u32(my, (tag << 29) | ((tile == 0 ? OS_FLAG_INC : OS_FLAG_AGG) << 27) | total);
    }
    {  // exclusive scan of totals over digits
      const uint32_t ex = digit_exclusive_scan(total, tid, S.warp_tot);
      if (dthread) S.texcl[d] = ex;
    }
    os_trace(tile, 3);
    uint32_t excl = 0;
    if (tile > 0 && dthread) {
      // warp-parallel look-back: warp w (< 8) owns digits 32w .. 32w+31 (thread tid owns digit tid)
      uint32_t* slab = reinterpret_cast<uint32_t*>(S.keys) + warp * (32 * 33);
      const uint32_t sentinel = (tag << 29) | (OS_FLAG_INC << 27);  // "before tile 0": inclusive prefix 0
      bool done = false;
      int j0 = (int)tile - 1;
      while (true) {
        const int jj = j0 - lane;
        uint32_t w[32];
        if (jj >= 0) {
          const uint4* row = reinterpret_cast<const uint4*>(status + (size_t)jj * OS_DIGITS + warp * 32);
          bool ready;
          do {
            ready = true;
#pragma unroll
            for (int q = 0; q < 8; q++) {
              const uint4 v = ld_volatile_v4(row + q);
              w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
            }
            // [31:29] tag, [28:27] flag: ready <=> (w >> 27) is tag*4 + 1 (aggregate) or tag*4 + 2 (inclusive)
#pragma unroll
            for (int k = 0; k < 32; k++) ready &= ((w[k] >> 27) - (tag * 4u + 1u)) < 2u;
          } while (!ready);
        } else {
#pragma unroll
          for (int k = 0; k < 32; k++) w[k] = sentinel;
        }
#pragma unroll
        for (int k = 0; k < 32; k++) slab[lane * 33 + k] = w[k];
        __syncwarp();
        {  // lane k walks the 32 predecessors of its digit in order, branch-free: every read is in flight at once
          uint32_t alive = done ? 0u : 1u;
#pragma unroll
          for (int l = 0; l < 32; l++) {
            const uint32_t x = slab[l * 33 + lane];
            excl += alive ? (x & OS_VAL_MASK) : 0u;
            alive &= (((x >> 27) & 3u) == OS_FLAG_INC) ? 0u : 1u;
          }
          done = alive == 0u;
        }
        __syncwarp();
        if (__all_sync(FULL, done)) break;
        j0 -= 32;
      }
      st_volatile_u32(my, (tag << 29) | (OS_FLAG_INC << 27) | (excl + total));
    }
    if (dthread) S.goff[d] = S.gbase[d] + excl - S.texcl[d];
    __syncthreads();  // look-back slabs (aliasing S.keys) are dead from here on
    os_trace(tile, 4);
    // ---- reorder through shared memory, then coalesced scatter
#pragma unroll
    for (int i = 0; i < OS_ITEMS; i++) {
      const uint32_t idx = warp * (32 * OS_ITEMS) + i * 32 + lane;
      if (idx < cnt) {
        const uint32_t dd = (uint32_t)((key[i] >> shift) & 255ull);
        const uint32_t p = S.texcl[dd] + S.whist[warp][dd] + pos[i];
        S.keys[p] = key[i];
        S.vals[p] = val[i];
      }
    }
    __syncthreads();
    // fixed trip count: the shared-memory reads of all of a thread's keys are in flight together
#pragma unroll
    for (int i = 0; i < OS_ITEMS; i++) {
      const uint32_t k = tid + i * OS_THREADS;
      key[i] = k < cnt ? S.keys[k] : 0ull;
      val[i] = k < cnt ? S.vals[k] : 0u;
    }
#pragma unroll
    for (int i = 0; i < OS_ITEMS; i++) {
      const uint32_t k = tid + i * OS_THREADS;
      if (k < cnt) {
        const uint64_t kk = key[i];
        const uint32_t dd = (uint32_t)((kk >> shift) & 255ull);
        const uint32_t o = S.goff[dd] + k;
        kout[o] = kk;
        vout[o] = val[i];
        if (is_last) {
          // inside one digit run of this CTA the keys are fully sorted and land on consecutive output slots
          const uint32_t t = (uint32_t)(kk >> 32);
          if (k == 0 || (uint32_t)(S.keys[k - 1] >> 32) != t) atomicMin(&ranges[t].x, o);
          if (k + 1 == cnt || (uint32_t)(S.keys[k + 1] >> 32) != t) atomicMax(&ranges[t].y, o + 1u);
        }
      }
    }
    __syncthreads();
    os_trace(tile, 5);
  }
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

int sort_passes(int gx, int gy) { return (32 + (int)higher_msb((uint32_t)(gx * gy)) + 7) / 8; }

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

static int binning_out(const RasterParams& rp, char* binning, char* img, const skgs_raster_layout& lay, int64_t R_cap,
                       BinningOut& b) {
  b.passes = sort_passes(rp.gx, rp.gy);
  SKGS_CHECK_ARG(b.passes <= MAX_PASSES, "tile grid too large for the 64-bit key layout");
  b.keys = reinterpret_cast<uint64_t*>(binning + lay.keys_a);
  b.vals = reinterpret_cast<uint32_t*>(binning + lay.vals_a);
  b.hist = reinterpret_cast<uint32_t*>(binning + lay.sort_hist);
  b.tile_counts = tile_sort_enabled() ? reinterpret_cast<uint32_t*>(binning + lay.tile_counts) : nullptr;
  b.cell_stride = tile_cell_stride();
  b.ranges = reinterpret_cast<uint2*>(img + lay.ranges);
  b.counters = reinterpret_cast<uint32_t*>(img + lay.work_counters);
  b.R_cap = (uint32_t)R_cap;
  b.tiles = rp.gx * rp.gy;
  return SKGS_OK;
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
    int rc = binning_out(rp, binning, img, lay, R_cap, bo);
    if (rc) return rc;
    // digit histograms + look-back words of the radix passes (adjacent): one memset
    SKGS_CUDA(cudaMemsetAsync(binning + lay.sort_hist, 0, lay.binning_bytes - lay.sort_hist, st));
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
  } else if (emit) {  // P == 0: no kernel ran, the consumers still expect initialised ranges / tickets
    SKGS_CUDA(cudaMemsetAsync(img + lay.ranges, 0, (size_t)bo.tiles * sizeof(uint2), st));
    SKGS_CUDA(cudaMemsetAsync(img + lay.work_counters, 0, 2 * sizeof(uint32_t), st));
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
  int rc = binning_out(rp, binning, img, lay, R_cap, bo);
  if (rc) return rc;
  SKGS_CUDA(cudaMemsetAsync(binning + lay.sort_hist, 0, lay.binning_bytes - lay.sort_hist, st));
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

// radix passes (+ key emission from the stored geometry when `emit`: the split API and the capacity-retry path)
int launch_binning(const RasterParams& rp, char* geom, char* binning, char* img, const skgs_raster_layout& lay,
                   const int32_t* radii, int64_t R_cap, int64_t R_hint, bool emit, uint32_t* num_rendered_host,
                   cudaStream_t st) {
  auto* hdr = reinterpret_cast<skgs_raster_header*>(geom + lay.header);
  const int tiles = rp.gx * rp.gy;
  if (rp.P == 0 || R_cap <= 0) {
    SKGS_CUDA(cudaMemsetAsync(img + lay.ranges, 0, (size_t)tiles * sizeof(uint2), st));
    SKGS_CUDA(cudaMemsetAsync(img + lay.work_counters, 0, 2 * sizeof(uint32_t), st));
    return SKGS_OK;
  }
  BinningOut bo = {};
  int rc = binning_out(rp, binning, img, lay, R_cap, bo);
  if (rc) return rc;
  if (emit) {
    // reset what a previous render stage on the same geometry may have left: overflow flag, tickets, plan, histograms
    SKGS_CUDA(cudaMemsetAsync(&hdr->overflow, 0, sizeof(uint32_t) * (1 + 8 + 8 + 2), st));
    SKGS_CUDA(cudaMemsetAsync(binning + lay.sort_hist, 0, lay.binning_bytes - lay.sort_hist, st));
    ProfScope prof_("duplicate_keys_kernel", st);
    SKGS_CUDA(launch_pdl(duplicate_keys_kernel, dim3((rp.P + DUP_THREADS - 1) / DUP_THREADS), dim3(DUP_THREADS), 0, st,
                         rp.P, rp.gx, rp.gy, radii, reinterpret_cast<const float2*>(geom + lay.means2D),
                         reinterpret_cast<const float*>(geom + lay.depths),
                         reinterpret_cast<const uint32_t*>(geom + lay.point_offsets), hdr, bo));
    SKGS_CHECK_LAUNCH("duplicate_keys_kernel");
  }
  if (num_rendered_host)
    SKGS_CUDA(cudaMemcpyAsync(num_rendered_host, hdr, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  if (bo.tile_counts != nullptr) return launch_tile_binning(rp, geom, binning, img, lay, R_cap, R_hint, st);
  static bool attr_set = false;
  if (!attr_set) {
    SKGS_CUDA(cudaFuncSetAttribute(onesweep_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)sizeof(OnesweepSmem)));
    attr_set = true;
  }
  static int ctas_per_sm = 0;
  if (ctas_per_sm == 0) {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, onesweep_pass_kernel, OS_THREADS, sizeof(OnesweepSmem));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
  }
  const int64_t hint = R_hint > 0 ? (R_hint < R_cap ? R_hint : R_cap) : R_cap;
  int grid = (int)((hint + OS_TILE - 1) / OS_TILE);
  const int cap = ctas_per_sm * num_sms();
  grid = grid < 1 ? 1 : (grid > cap ? cap : grid);
  uint32_t* status = reinterpret_cast<uint32_t*>(binning + lay.sort_status);
  for (int p = 0; p < bo.passes; p++) {
    static const char* kPassName[MAX_PASSES] = {"onesweep_pass0", "onesweep_pass1", "onesweep_pass2", "onesweep_pass3",
                                                "onesweep_pass4", "onesweep_pass5", "onesweep_pass6", "onesweep_pass7"};
    ProfScope prof_(kPassName[p], st);
    SKGS_CUDA(launch_pdl(onesweep_pass_kernel, dim3(grid), dim3(OS_THREADS), sizeof(OnesweepSmem), st,
                         reinterpret_cast<uint64_t*>(binning + lay.keys_a),
                         reinterpret_cast<uint32_t*>(binning + lay.vals_a),
                         reinterpret_cast<uint64_t*>(binning + lay.keys_b),
                         reinterpret_cast<uint32_t*>(binning + lay.vals_b), hdr, (uint32_t)R_cap, bo.hist + p * 256,
                         status, p, p == bo.passes - 1 ? 1 : 0, bo.ranges));
    SKGS_CHECK_LAUNCH("onesweep_pass_kernel");
  }
  return SKGS_OK;
}

}  // namespace skgs

extern "C" __attribute__((visibility("default"))) void skgs_debug_set_sort_trace(void* p) {
  unsigned long long* q = reinterpret_cast<unsigned long long*>(p);
  cudaMemcpyToSymbol(skgs::g_os_trace, &q, sizeof(q));
}
