// raster_fwd.cu - forward half of the tile rasterizer for sm_100a:
//   K1 preprocess_scan_kernel : projection, EWA cov2D, conic, radius, tile rect, SH->RGB  +  fused decoupled-look-back
//                               prefix sum of tiles_touched (no separate scan kernel, no host sync for R)
//   K2 duplicate_keys_kernel  : (tile << 32 | depth bits, id) emission, warp-cooperative for large rects,
//                               + all radix digit histograms of the sort, accumulated while emitting
//   K3 onesweep_pass_kernel   : one kernel per 8-bit digit, chained-scan (decoupled look-back) across key tiles,
//                               stable warp-level multi-split ranking (match.any), smem-staged coalesced scatter
//   K4 tile_ranges_kernel     : [start, end) of each screen tile in the sorted list
//   (K5/K6 compositing kernels live in composite.cu)
// Semantics follow SURVEY.md App. A.4-A.6 (reference: my_ext/_C/src/nerf/gaussian_preprocess_colmap.cu:155-224,
// gaussian_rasterizer_forward.cu:45-94,203-241, gaussian_render.cu:16-112).  This file MUST be compiled with
// -fmad=false (and without fast-math): plain '*' and '+' below are separately rounded, exactly like the CPU oracle.
#ifndef SKGS_NO_FMAD
#error "raster_fwd.cu must be built with -fmad=false -DSKGS_NO_FMAD (bit-exact radii/keys depend on it)"
#endif
#include <cooperative_groups.h>

#include "common.cuh"

namespace skgs {

__device__ __constant__ float c_SH_C0 = 0.28209479177387814f;
__device__ __constant__ float c_SH_C1 = 0.4886025119029199f;
__device__ __constant__ float c_SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                            -1.0925484305920792f, 0.5462742152960396f};
__device__ __constant__ float c_SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                            0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                            -0.5900435899266435f};

// ------------------------------------------------------------------------------------------------------------------
// K1: preprocess + fused prefix sum
// ------------------------------------------------------------------------------------------------------------------
constexpr int PRE_THREADS = 256;
constexpr uint64_t SCAN_FLAG_AGG = 1ull << 62;
constexpr uint64_t SCAN_FLAG_INC = 2ull << 62;
constexpr uint64_t SCAN_VAL_MASK = (1ull << 62) - 1;

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
__device__ __forceinline__ void st_volatile_u64(uint64_t* p, uint64_t v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(PRE_THREADS)
preprocess_scan_kernel(RasterParams rp, const float* __restrict__ means3D, const float* __restrict__ shs,
                       const float* __restrict__ colors_precomp, const float* __restrict__ opacities,
                       const float* __restrict__ scales, const float* __restrict__ rotations,
                       const float* __restrict__ cov3D_precomp, int32_t* __restrict__ radii,
                       float2* __restrict__ means2D, float* __restrict__ depths, float* __restrict__ cov3Ds,
                       float4* __restrict__ conic_opacity, float4* __restrict__ rgbd, uint8_t* __restrict__ clamped,
                       uint32_t* __restrict__ tiles_touched, uint32_t* __restrict__ point_offsets,
                       uint64_t* __restrict__ scan_state, skgs_raster_header* __restrict__ hdr, int num_blocks,
                       uint32_t* __restrict__ tile_count) {
  __shared__ int s_bid;
  __shared__ uint32_t s_warp_sum[PRE_THREADS / 32];
  __shared__ uint32_t s_excl;
  __shared__ float s_V[16], s_P[16], s_cam[3];
  const int tid = threadIdx.x;
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
  const float* V = s_V;
  const float* Pm = s_P;

  uint32_t touched = 0;
  int rx0 = 0, ry0 = 0, rx1 = 0;
  if (i < rp.P) {
    int my_rad = 0;
    const float x = means3D[3 * i], y = means3D[3 * i + 1], z = means3D[3 * i + 2];
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
      if (cov3D_precomp != nullptr) {
#pragma unroll
        for (int k = 0; k < 6; k++) c6[k] = cov3D_precomp[6 * i + k];
      } else {
        float qr, qx, qy, qz;
        const float4 q = *reinterpret_cast<const float4*>(rotations + 4 * i);
        if (rp.quat_wxyz) {
          qr = q.x; qx = q.y; qy = q.z; qz = q.w;
        } else {
          qx = q.x; qy = q.y; qz = q.z; qr = q.w;
        }
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
        const float s0 = rp.mod * scales[3 * i], s1 = rp.mod * scales[3 * i + 1], s2 = rp.mod * scales[3 * i + 2];
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
        for (int k = 0; k < 6; k++) cov3Ds[6 * i + k] = c6[k];
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
          rx0 = x0; ry0 = y0; rx1 = x1;
          float rgb[3];
          uint8_t cl = 0;
          if (colors_precomp == nullptr) {
            const float dirx = x - s_cam[0], diry = y - s_cam[1], dirz = z - s_cam[2];
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
            rgb[0] = colors_precomp[3 * i];
            rgb[1] = colors_precomp[3 * i + 1];
            rgb[2] = colors_precomp[3 * i + 2];
          }
          my_rad = irad;
          touched = (uint32_t)cnt;
          depths[i] = pvz;
          means2D[i] = make_float2(pix_x, pix_y);
          conic_opacity[i] = make_float4(con_x, con_y, con_z, opacities[i]);
          rgbd[i] = make_float4(rgb[0], rgb[1], rgb[2], pvz);
          clamped[i] = cl;
        }
      }
    }
    radii[i] = my_rad;
    tiles_touched[i] = touched;
  }
  // ---- per-tile list lengths for the bucketed binning (one atomic per (Gaussian, tile) pair; rects with more than 32
  //      tiles are walked by the whole warp)
  if (tile_count != nullptr) {
    const int lane_ = tid & 31;
    const int w = rx1 - rx0;
    if (touched > 0 && touched <= 32u)
      for (uint32_t k = 0; k < touched; k++) atomicAdd(&tile_count[(ry0 + (int)k / w) * rp.gx + rx0 + (int)k % w], 1u);
    uint32_t big = __ballot_sync(0xffffffffu, touched > 32u);
    while (big) {
      const int src = __ffs(big) - 1;
      big &= big - 1;
      const int bx0 = __shfl_sync(0xffffffffu, rx0, src), by0 = __shfl_sync(0xffffffffu, ry0, src);
      const int bw = __shfl_sync(0xffffffffu, w, src);
      const uint32_t bc = __shfl_sync(0xffffffffu, touched, src);
      for (uint32_t k = lane_; k < bc; k += 32) atomicAdd(&tile_count[(by0 + (int)k / bw) * rp.gx + bx0 + (int)k % bw], 1u);
    }
  }

  // ---- block-inclusive scan of `touched`
  const int lane = tid & 31, warp = tid >> 5;
  uint32_t incl = touched;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) s_warp_sum[warp] = incl;
  const uint32_t vis_ballot = __ballot_sync(0xffffffffu, touched > 0);
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
        const uint32_t inc_mask = __ballot_sync(0xffffffffu, (w >> 62) == 2);
        // sum values of lanes up to and including the first inclusive one
        const int first_inc = inc_mask ? (__ffs(inc_mask) - 1) : 32;
        uint64_t v = (lane <= first_inc) ? (w & SCAN_VAL_MASK) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        excl += v;
        if (inc_mask) break;
        j -= 32;
      }
      if (lane == 0) st_volatile_u64(&scan_state[bid], SCAN_FLAG_INC | (excl + block_total));
    }
    if (lane == 0) {
      s_excl = (uint32_t)excl;
      if (bid == num_blocks - 1) hdr->num_rendered = (uint32_t)(excl + block_total);
    }
  }
  __syncthreads();
  if (i < rp.P) point_offsets[i] = s_excl + warp_off + incl;
}

// ------------------------------------------------------------------------------------------------------------------
// K2: duplicate with keys (+ radix digit histograms)
// ------------------------------------------------------------------------------------------------------------------
constexpr int DUP_THREADS = 256;
constexpr int MAX_PASSES = 8;

__global__ void __launch_bounds__(DUP_THREADS)
duplicate_keys_kernel(int P, int gx, int gy, const int32_t* __restrict__ radii, const float2* __restrict__ means2D,
                      const float* __restrict__ depths, const uint32_t* __restrict__ point_offsets,
                      uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t* __restrict__ hist,
                      skgs_raster_header* __restrict__ hdr, uint32_t R_cap, int passes) {
  __shared__ uint32_t s_hist[MAX_PASSES * 256];
  const int tid = threadIdx.x, lane = tid & 31;
  for (int k = tid; k < passes * 256; k += DUP_THREADS) s_hist[k] = 0;
  __syncthreads();
  const int i = blockIdx.x * DUP_THREADS + tid;
  int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
  uint32_t off = 0, dbits = 0;
  int cnt = 0;
  if (i < P) {
    const int rad = radii[i];
    if (rad > 0) {
      const float2 p = means2D[i];
      get_rect(p.x, p.y, rad, gx, gy, x0, y0, x1, y1);
      cnt = (x1 - x0) * (y1 - y0);
      off = (i == 0) ? 0u : point_offsets[i - 1];
      dbits = __float_as_uint(depths[i]);
      if ((uint64_t)off + (uint64_t)cnt > (uint64_t)R_cap) {  // arena too small: flag and emit nothing
        hdr->overflow = 1;
        cnt = 0;
      }
    }
  }
  if (cnt > 0) {
#pragma unroll
    for (int p = 0; p < 4; p++) atomicAdd(&s_hist[p * 256 + ((dbits >> (8 * p)) & 255u)], (uint32_t)cnt);
  }
  constexpr int COOP = 32;  // rects with more tiles than this are emitted by the whole warp
  if (cnt > 0 && cnt <= COOP) {
    const int w = x1 - x0;
    for (int k = 0; k < cnt; k++) {
      const int ty = y0 + k / w, tx = x0 + k % w;
      const uint32_t tile = (uint32_t)(ty * gx + tx);
      keys[off + k] = ((uint64_t)tile << 32) | dbits;
      vals[off + k] = (uint32_t)i;
      for (int p = 4; p < passes; p++) atomicAdd(&s_hist[p * 256 + ((tile >> (8 * (p - 4))) & 255u)], 1u);
    }
  }
  uint32_t big = __ballot_sync(0xffffffffu, cnt > COOP);
  while (big) {
    const int src = __ffs(big) - 1;
    big &= big - 1;
    const int bx0 = __shfl_sync(0xffffffffu, x0, src), by0 = __shfl_sync(0xffffffffu, y0, src);
    const int bx1 = __shfl_sync(0xffffffffu, x1, src);
    const int bcnt = __shfl_sync(0xffffffffu, cnt, src);
    const uint32_t boff = __shfl_sync(0xffffffffu, off, src), bd = __shfl_sync(0xffffffffu, dbits, src);
    const int bi = __shfl_sync(0xffffffffu, i, src);
    const int w = bx1 - bx0;
    for (int k = lane; k < bcnt; k += 32) {
      const int ty = by0 + k / w, tx = bx0 + k % w;
      const uint32_t tile = (uint32_t)(ty * gx + tx);
      keys[boff + k] = ((uint64_t)tile << 32) | bd;
      vals[boff + k] = (uint32_t)bi;
      for (int p = 4; p < passes; p++) atomicAdd(&s_hist[p * 256 + ((tile >> (8 * (p - 4))) & 255u)], 1u);
    }
  }
  __syncthreads();
  for (int k = tid; k < passes * 256; k += DUP_THREADS) {
    const uint32_t c = s_hist[k];
    if (c) atomicAdd(&hist[k], c);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K3: onesweep radix pass (8-bit digit), stable.  Status word: [31:29] pass tag, [28:27] flag, [26:0] count.
// ------------------------------------------------------------------------------------------------------------------
#ifndef SKGS_OS_THREADS
#define SKGS_OS_THREADS 256
#endif
#ifndef SKGS_OS_ITEMS
#define SKGS_OS_ITEMS 24
#endif
constexpr int OS_THREADS = SKGS_OS_THREADS;
constexpr int OS_ITEMS = SKGS_OS_ITEMS;
constexpr int OS_TILE = OS_THREADS * OS_ITEMS;  // 4096 keys per tile
constexpr int OS_WARPS = OS_THREADS / 32;
constexpr uint32_t OS_FLAG_AGG = 1u, OS_FLAG_INC = 2u;
constexpr uint32_t OS_VAL_MASK = (1u << 27) - 1;

struct OnesweepSmem {
  uint64_t keys[OS_TILE];
  uint32_t vals[OS_TILE];
  uint32_t whist[OS_WARPS][256];
  uint32_t texcl[256];   // exclusive prefix of this tile's digit counts
  uint32_t goff[256];    // global output offset of digit d minus texcl[d]
  uint32_t gbase[256];   // exclusive prefix of the global digit histogram
  uint32_t warp_tot[OS_WARPS];
  uint32_t tile;
};

__global__ void __launch_bounds__(OS_THREADS)
onesweep_pass_kernel(const uint64_t* __restrict__ kin, const uint32_t* __restrict__ vin, uint64_t* __restrict__ kout,
                     uint32_t* __restrict__ vout, const skgs_raster_header* __restrict__ hdr, uint32_t R_cap,
                     const uint32_t* __restrict__ hist, uint32_t* __restrict__ status, uint32_t* __restrict__ ticket,
                     int shift, uint32_t tag) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  OnesweepSmem& S = *reinterpret_cast<OnesweepSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (hdr->overflow) return;
  const uint32_t n = min(hdr->num_rendered, R_cap);
  const uint32_t num_tiles = (n + OS_TILE - 1) / OS_TILE;
  const uint32_t lanemask_lt = (1u << lane) - 1u;

  // a digit that is constant over all keys (typically the sign/exponent byte of the depth) makes the pass the
  // identity permutation: copy the tiles straight through, no ranking, no look-back
  if (__syncthreads_or(n > 0 && hist[tid] == n)) {
    while (true) {
      if (tid == 0) S.tile = atomicAdd(ticket, 1u);
      __syncthreads();
      const uint32_t tile = S.tile;
      __syncthreads();
      if (tile >= num_tiles) return;
      const uint32_t base = tile * OS_TILE;
      const uint32_t cnt = min((uint32_t)OS_TILE, n - base);
      for (uint32_t k0 = tid; k0 < cnt; k0 += 4 * OS_THREADS) {  // four independent loads in flight per thread
        uint64_t kk[4];
        uint32_t vv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const uint32_t k = k0 + u * OS_THREADS;
          if (k < cnt) {
            kk[u] = kin[base + k];
            vv[u] = vin[base + k];
          }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const uint32_t k = k0 + u * OS_THREADS;
          if (k < cnt) {
            kout[base + k] = kk[u];
            vout[base + k] = vv[u];
          }
        }
      }
    }
  }
  // exclusive scan of the global digit histogram (256 values, one per thread)
  {
    const uint32_t c = hist[tid];
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) S.warp_tot[warp] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < warp; w++) woff += S.warp_tot[w];
    S.gbase[tid] = woff + incl - c;
    __syncthreads();
  }

  while (true) {
    if (tid == 0) S.tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = S.tile;
    if (tile >= num_tiles) break;
    const uint32_t base = tile * OS_TILE;
    const uint32_t cnt = min((uint32_t)OS_TILE, n - base);

    uint64_t key[OS_ITEMS];
    uint16_t pos[OS_ITEMS];
#pragma unroll
    for (int i = 0; i < OS_ITEMS; i++) {
      const uint32_t idx = warp * (32 * OS_ITEMS) + i * 32 + lane;
      key[i] = idx < cnt ? kin[base + idx] : ~0ull;
    }
    for (int k = tid; k < OS_WARPS * 256; k += OS_THREADS) (&S.whist[0][0])[k] = 0;
    __syncthreads();
    // ---- stable per-warp ranking (items are warp-striped: item-major, then lane)
#pragma unroll
    for (int i = 0; i < OS_ITEMS; i++) {
      const uint32_t idx = warp * (32 * OS_ITEMS) + i * 32 + lane;
      const bool valid = idx < cnt;
      const uint32_t d = valid ? (uint32_t)((key[i] >> shift) & 255ull) : 0xffffffffu;
      const uint32_t m = __match_any_sync(0xffffffffu, d);
      const int leader = __ffs(m) - 1;
      uint32_t old = 0;
      if (valid && lane == leader) {
        old = S.whist[warp][d];
        S.whist[warp][d] = old + __popc(m);
      }
      old = __shfl_sync(0xffffffffu, old, leader);
      pos[i] = (uint16_t)(old + __popc(m & lanemask_lt));
      __syncwarp();
    }
    __syncthreads();
    // ---- per digit: cross-warp exclusive prefix, tile totals, publish, look back
    uint32_t total = 0;
    {
      const int d = tid;
#pragma unroll
      for (int w = 0; w < OS_WARPS; w++) {
        const uint32_t c = S.whist[w][d];
        S.whist[w][d] = total;
        total += c;
      }
      uint32_t* my = status + (size_t)tile * 256 + d;
      st_volatile_u32(my, (tag << 29) | ((tile == 0 ? OS_FLAG_INC : OS_FLAG_AGG) << 27) | total);
      // exclusive scan of totals over digits
      uint32_t incl = total;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) S.warp_tot[warp] = incl;
      __syncthreads();
      uint32_t woff = 0;
      for (int w = 0; w < warp; w++) woff += S.warp_tot[w];
      const uint32_t texcl = woff + incl - total;
      S.texcl[d] = texcl;
      uint32_t excl = 0;
      if (tile > 0) {
        int j = (int)tile - 1;
        while (true) {
          uint32_t w;
          do {
            w = ld_volatile_u32(status + (size_t)j * 256 + d);
          } while ((w >> 29) != tag || ((w >> 27) & 3u) == 0);
          excl += w & OS_VAL_MASK;
          if (((w >> 27) & 3u) == OS_FLAG_INC) break;
          j--;
        }
        st_volatile_u32(my, (tag << 29) | (OS_FLAG_INC << 27) | (excl + total));
      }
      S.goff[d] = S.gbase[d] + excl - texcl;
    }
    __syncthreads();
    // ---- reorder through shared memory, then coalesced scatter
#pragma unroll
    for (int i = 0; i < OS_ITEMS; i++) {
      const uint32_t idx = warp * (32 * OS_ITEMS) + i * 32 + lane;
      if (idx < cnt) {
        const uint32_t d = (uint32_t)((key[i] >> shift) & 255ull);
        const uint32_t p = S.texcl[d] + S.whist[warp][d] + pos[i];
        pos[i] = (uint16_t)p;
        S.keys[p] = key[i];
        S.vals[p] = vin[base + idx];
      }
    }
    __syncthreads();
    for (uint32_t k = tid; k < cnt; k += OS_THREADS) {
      const uint64_t kk = S.keys[k];
      const uint32_t d = (uint32_t)((kk >> shift) & 255ull);
      const uint32_t o = S.goff[d] + k;
      kout[o] = kk;
      vout[o] = S.vals[k];
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K4: tile ranges
// ------------------------------------------------------------------------------------------------------------------
__global__ void tile_ranges_kernel(const uint64_t* __restrict__ keys, const skgs_raster_header* __restrict__ hdr,
                                   uint32_t R_cap, uint2* __restrict__ ranges) {
  if (hdr->overflow) return;  // arena too small: keys are incomplete, leave every range empty
  const uint32_t n = min(hdr->num_rendered, R_cap);
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
    const uint32_t cur = (uint32_t)(keys[idx] >> 32);
    if (idx == 0)
      ranges[cur].x = 0;
    else {
      const uint32_t prev = (uint32_t)(keys[idx - 1] >> 32);
      if (cur != prev) {
        ranges[prev].y = idx;
        ranges[cur].x = idx;
      }
    }
    if (idx == n - 1) ranges[cur].y = n;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------------------------
int launch_preprocess_scan(const RasterParams& rp, const float* means3D, const float* shs, const float* colors_precomp,
                           const float* opacities, const float* scales, const float* rotations,
                           const float* cov3D_precomp, char* geom, const skgs_raster_layout& lay, int32_t* radii,
                           uint32_t* num_rendered_host, bool count_tiles, cudaStream_t st) {
  auto* hdr = reinterpret_cast<skgs_raster_header*>(geom + lay.header);
  const int nblocks = (rp.P + PRE_THREADS - 1) / PRE_THREADS;
  // header and scan_state are adjacent in the arena: one memset resets the ticket, the flags and the counters
  SKGS_CUDA(cudaMemsetAsync(geom + lay.header, 0, lay.means2D - lay.header, st));
  if (rp.P > 0) {
    {
      ProfScope prof_("preprocess_scan_kernel", st);
      preprocess_scan_kernel<<<nblocks, PRE_THREADS, 0, st>>>(
        rp, means3D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp, radii,
        reinterpret_cast<float2*>(geom + lay.means2D), reinterpret_cast<float*>(geom + lay.depths),
        reinterpret_cast<float*>(geom + lay.cov3D), reinterpret_cast<float4*>(geom + lay.conic_opacity),
        reinterpret_cast<float4*>(geom + lay.rgbd), reinterpret_cast<uint8_t*>(geom + lay.clamped),
        reinterpret_cast<uint32_t*>(geom + lay.tiles_touched), reinterpret_cast<uint32_t*>(geom + lay.point_offsets),
        reinterpret_cast<uint64_t*>(geom + lay.scan_state), hdr, nblocks,
        count_tiles ? reinterpret_cast<uint32_t*>(geom + lay.tile_count) : nullptr);
    SKGS_CHECK_LAUNCH("preprocess_scan_kernel");
    }
  }
  if (num_rendered_host)
    SKGS_CUDA(cudaMemcpyAsync(num_rendered_host, hdr, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  return SKGS_OK;
}

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

int launch_binning(const RasterParams& rp, char* geom, char* binning, char* img, const skgs_raster_layout& lay,
                   const int32_t* radii, int64_t R_cap, int64_t R_hint, uint32_t* num_rendered_host,
                   cudaStream_t st) {
  auto* hdr = reinterpret_cast<skgs_raster_header*>(geom + lay.header);
  const int tiles = rp.gx * rp.gy;
  const int passes = sort_passes(rp.gx, rp.gy);
  SKGS_CHECK_ARG(passes <= MAX_PASSES, "tile grid too large for the 64-bit key layout");
  SKGS_CUDA(cudaMemsetAsync(img + lay.ranges, 0, (size_t)tiles * sizeof(uint2), st));
  if (rp.P == 0 || R_cap <= 0) return SKGS_OK;
  // sort_hist and sort_status are adjacent: one memset
  SKGS_CUDA(cudaMemsetAsync(binning + lay.sort_hist, 0, lay.binning_bytes - lay.sort_hist, st));
  uint64_t* kA = reinterpret_cast<uint64_t*>(binning + lay.keys_unsorted);
  uint32_t* vA = reinterpret_cast<uint32_t*>(binning + lay.vals_unsorted);
  uint64_t* kB = reinterpret_cast<uint64_t*>(binning + lay.keys_sorted);
  uint32_t* vB = reinterpret_cast<uint32_t*>(binning + lay.point_list);
  uint32_t* hist = reinterpret_cast<uint32_t*>(binning + lay.sort_hist);
  uint32_t* status = reinterpret_cast<uint32_t*>(binning + lay.sort_status);
  {
    ProfScope prof_("duplicate_keys_kernel", st);
    duplicate_keys_kernel<<<(rp.P + DUP_THREADS - 1) / DUP_THREADS, DUP_THREADS, 0, st>>>(
      rp.P, rp.gx, rp.gy, radii, reinterpret_cast<const float2*>(geom + lay.means2D),
      reinterpret_cast<const float*>(geom + lay.depths), reinterpret_cast<const uint32_t*>(geom + lay.point_offsets),
      kA, vA, hist, hdr, (uint32_t)R_cap, passes);
  SKGS_CHECK_LAUNCH("duplicate_keys_kernel");
  }
  if (num_rendered_host)
    SKGS_CUDA(cudaMemcpyAsync(num_rendered_host, hdr, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  static bool attr_set = false;
  if (!attr_set) {
    SKGS_CUDA(cudaFuncSetAttribute(onesweep_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)sizeof(OnesweepSmem)));
    attr_set = true;
  }
  const int64_t hint = R_hint > 0 ? (R_hint < R_cap ? R_hint : R_cap) : R_cap;
  int grid = (int)((hint + OS_TILE - 1) / OS_TILE);
  grid = grid < 1 ? 1 : (grid > 2 * num_sms() ? 2 * num_sms() : grid);
  uint64_t *kin = kA, *kout = kB;
  uint32_t *vin = vA, *vout = vB;
  for (int p = 0; p < passes; p++) {
    {
      static const char* kPassName[MAX_PASSES] = {"onesweep_pass0", "onesweep_pass1", "onesweep_pass2", "onesweep_pass3",
                                                  "onesweep_pass4", "onesweep_pass5", "onesweep_pass6", "onesweep_pass7"};
      ProfScope prof_(kPassName[p], st);
      onesweep_pass_kernel<<<grid, OS_THREADS, sizeof(OnesweepSmem), st>>>(kin, vin, kout, vout, hdr, (uint32_t)R_cap,
                                                                          hist + p * 256, status,
                                                                          &hdr->sort_ticket[p], 8 * p, (uint32_t)p);
    SKGS_CHECK_LAUNCH("onesweep_pass_kernel");
    }
    uint64_t* tk = kin; kin = kout; kout = tk;
    uint32_t* tv = vin; vin = vout; vout = tv;
  }
  // after the swap `kin` holds the sorted keys.  With an even number of passes that is buffer A again:
  // skgs_raster_layout_query already reports keys_sorted/point_list at the physical location of the final result.
  int rgrid = (int)((hint + 255) / 256);
  rgrid = rgrid < 1 ? 1 : (rgrid > 8 * num_sms() ? 8 * num_sms() : rgrid);
  {
    ProfScope prof_("tile_ranges_kernel", st);
    tile_ranges_kernel<<<rgrid, 256, 0, st>>>(kin, hdr, (uint32_t)R_cap, reinterpret_cast<uint2*>(img + lay.ranges));
  SKGS_CHECK_LAUNCH("tile_ranges_kernel");
  }
  return SKGS_OK;
}

}  // namespace skgs
