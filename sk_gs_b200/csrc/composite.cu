// composite.cu - per-tile alpha compositing, forward and backward, for sm_100a.
//
// Design (measured motivation in profiles/r1_composite_v1.md): the first version ran one CTA per 16x16 tile and was
// latency / tail bound (42 % issue utilisation, SMs idle 40 % of the kernel because a few long tiles finish last).
// This version is WARP-granular and persistent:
//   * work item  = half a tile (16 x 8 pixels), one warp, 4 horizontally adjacent pixels per lane (row terms of the
//                  exponent are shared by the 4 pixels, so a (pixel, Gaussian) pair costs 3 FP instructions);
//   * scheduling = a global ticket hands out work items in order of DEcreasing tile length (tile_order_kernel), so the
//                  long tiles start first and the tail is made of short ones; no __syncthreads anywhere - a warp that
//                  finishes (all its pixels saturated) immediately takes the next item;
//   * staging    = 32 Gaussians per batch in the warp's own shared-memory slice, the NEXT batch is gathered into
//                  registers while the current one is composited (the dependent point_list -> attribute gathers are
//                  hidden behind ~2k cycles of math);
//   * culling    = conservative exponent cut-off pmin = -log(255 o) - 1e-4 per Gaussian: pairs below it can never
//                  reach alpha >= 1/255 and skip exp(); finished pixels are parked at x = +huge so they fail the same
//                  single comparison;
//   * backward   = per-lane sums over its 4 pixels, 5-step shuffle reduction, lane j keeps the totals of batch entry j,
//                  and each lane flushes its entry with three 16-byte vector REDs (red.global.add.v4.f32) - one RED
//                  set per (half tile, Gaussian) instead of the reference's 9-13 scalar atomics per (pixel, Gaussian)
//                  (my_ext/_C/src/nerf/gaussian_render.cu:295-338).
// Semantics: SURVEY.md App. A.6 / A.7 (reference gaussian_render.cu:16-112, 182-341 + bg / depth / alpha terms).
// Every operation that decides WHICH pairs contribute uses the contraction-proof helpers of common.cuh.
#include "common.cuh"

namespace skgs {

#ifndef SKGS_CW_WARPS
#define SKGS_CW_WARPS 4
#endif
#ifndef SKGS_BWD_MINBLOCKS
#define SKGS_BWD_MINBLOCKS 6
#endif
#ifndef SKGS_RSLOTS
#define SKGS_RSLOTS 6
#endif
constexpr int CW_WARPS = SKGS_CW_WARPS;     // warps per CTA (independent workers)
constexpr int CW_THREADS = CW_WARPS * 32;
constexpr int NGRAD = 12;                   // packed per-Gaussian accumulators: mx my ca cb | cc op z - | r g b -
constexpr float PARKED = 1.0e18f;           // x coordinate of a finished pixel: its exponent is -inf

// ------------------------------------------------------------------------------------------------------------------
// tile order: tiles sorted by decreasing list length (coarse: 8 sub-steps per octave), one CTA
// ------------------------------------------------------------------------------------------------------------------
constexpr int TO_THREADS = 1024;
constexpr int TO_BINS = 8 * 33;

__device__ __forceinline__ int length_bin(uint32_t len) {
  if (len == 0) return 0;
  const int e = 31 - __clz(len);                       // floor(log2 len)
  const int m = e >= 3 ? (int)((len >> (e - 3)) & 7u) : (int)((len << (3 - e)) & 7u);
  return 1 + e * 8 + m;                                // monotone in len
}

__global__ void __launch_bounds__(TO_THREADS)
tile_order_kernel(const uint2* __restrict__ ranges, int tiles, uint32_t* __restrict__ order,
                  uint32_t* __restrict__ counters) {
  __shared__ uint32_t s_hist[TO_BINS];
  __shared__ uint32_t s_base[TO_BINS];
  for (int k = threadIdx.x; k < TO_BINS; k += TO_THREADS) s_hist[k] = 0;
  if (threadIdx.x < 2) counters[threadIdx.x] = 0;  // work tickets of the forward / backward compositing kernels
  __syncthreads();
  for (int t = threadIdx.x; t < tiles; t += TO_THREADS) {
    const uint2 r = ranges[t];
    atomicAdd(&s_hist[length_bin(r.y - r.x)], 1u);
  }
  __syncthreads();
  // descending exclusive prefix (largest bin first): block scan over the reversed histogram
  {
    __shared__ uint32_t s_wsum[TO_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t v = tid < TO_BINS ? s_hist[TO_BINS - 1 - tid] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < warp; w++) woff += s_wsum[w];
    if (tid < TO_BINS) s_base[TO_BINS - 1 - tid] = woff + incl - v;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < tiles; t += TO_THREADS) {
    const uint2 r = ranges[t];
    const uint32_t p = atomicAdd(&s_base[length_bin(r.y - r.x)], 1u);
    order[p] = (uint32_t)t;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// shared pieces
// ------------------------------------------------------------------------------------------------------------------
// Raw gathered attributes of one list entry.  NOTHING is computed from them until they are committed to shared memory
// one batch later: an in-order warp stalls at the first USE of a pending load, not at its issue.
struct Staged {
  float2 m;    // pixel-space mean
  float4 co;   // conic a, b, c, opacity
  float4 c;    // r, g, b, depth
  uint32_t g;  // Gaussian id
};

__device__ __forceinline__ void gather_id(uint32_t g, const float2* __restrict__ means2D,
                                          const float4* __restrict__ conic_opacity, const float4* __restrict__ rgbd,
                                          Staged& s) {
  s.g = g;
  s.m = __ldg(means2D + g);
  s.co = __ldg(conic_opacity + g);
  s.c = __ldg(rgbd + g);
}

__device__ __forceinline__ void gather(const uint32_t* __restrict__ point_list, const float2* __restrict__ means2D,
                                       const float4* __restrict__ conic_opacity, const float4* __restrict__ rgbd,
                                       uint32_t pos, Staged& s) {
  gather_id(__ldg(point_list + pos), means2D, conic_opacity, rgbd, s);
}

// commit a gathered entry to the warp's staging slot: g0 = (gx, gy, A', B'), g1 = (C', opacity, pmin, id), c
__device__ __forceinline__ void commit(const Staged& s, float4* g0s, float4* g1s, float4* cs, int lane) {
  // alpha = o*exp(power) >= 1/255 needs power >= -log(255 o): conservative cut-off (margin 1e-4); everything that
  // passes it still takes the exact test on alpha
  const float pmin = s.co.w >= (1.0f / 255.0f) ? (-__logf(255.0f * s.co.w) - 1e-4f) : 1.0f;
  g0s[lane] = make_float4(s.m.x, s.m.y, -0.5f * s.co.x, -s.co.y);
  g1s[lane] = make_float4(-0.5f * s.co.z, s.co.w, pmin, __uint_as_float(s.g));
  cs[lane] = s.c;
}

// Conservative footprint test (phase 1 of every batch, one staged Gaussian per lane): can ANY pixel centre of the
// rectangle [X0,X1] x [Y0,Y1] reach power >= pmin?  power = -f/2 with f = A dx^2 + 2 B dx dy + C dy^2 convex, so its
// minimum over the box of offsets is 0 if the centre is inside, else attained on one of the four edges at the clamped
// 1-D minimiser.  The margin covers the fp32 rounding of both this test and pair_power() (relative 1e-6 of the largest
// term magnitude over the box), so a pair that passes the exact per-pixel test is never culled here.
__device__ __forceinline__ bool footprint_may_hit(const float4 g0, const float4 g1, float X0, float X1, float Y0,
                                                  float Y1) {
  const float A = -2.0f * g0.z, B = -g0.w, C = -2.0f * g1.x, pmin = g1.z;
  const float dxlo = g0.x - X1, dxhi = g0.x - X0, dylo = g0.y - Y1, dyhi = g0.y - Y0;
  const bool in_x = dxlo <= 0.f && dxhi >= 0.f, in_y = dylo <= 0.f && dyhi >= 0.f;
  float fmin = 0.f;
  if (!(in_x && in_y)) {
    const float invC = 1.0f / C, invA = 1.0f / A;
    float f = __int_as_float(0x7f800000);
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const float ex = e ? dxhi : dxlo;
      const float yy = fminf(fmaxf(-B * ex * invC, dylo), dyhi);
      f = fminf(f, A * ex * ex + 2.0f * B * ex * yy + C * yy * yy);
      const float ey = e ? dyhi : dylo;
      const float xx = fminf(fmaxf(-B * ey * invA, dxlo), dxhi);
      f = fminf(f, A * xx * xx + 2.0f * B * xx * ey + C * ey * ey);
    }
    fmin = f;
  }
  const float DX = fmaxf(fabsf(dxlo), fabsf(dxhi)), DY = fmaxf(fabsf(dylo), fabsf(dyhi));
  const float eps = 1e-6f * (fabsf(A) * DX * DX + 2.0f * fabsf(B) * DX * DY + fabsf(C) * DY * DY) + 1e-3f;
  return !(-0.5f * fmin < pmin - eps);  // NaN-safe: anything unordered is kept
}

// ------------------------------------------------------------------------------------------------------------------
// forward.  Work item = 8 x 4 pixels (one pixel per lane), 8 items per tile.
// ------------------------------------------------------------------------------------------------------------------
constexpr int ITEM_W = 8, ITEM_H = 4, ITEMS_PER_TILE = (TILE / ITEM_W) * (TILE / ITEM_H);
constexpr int RSLOTS = SKGS_RSLOTS, RVALS = 10, RSTRIDE = 33;  // backward reduction staging (see composite_bwd_kernel)

__global__ void __launch_bounds__(CW_THREADS)
composite_fwd_kernel(int W, int H, int gx, int tiles, const uint32_t* __restrict__ order,
                     uint32_t* __restrict__ ticket, const uint2* __restrict__ ranges,
                     const uint32_t* __restrict__ point_list, const float2* __restrict__ means2D,
                     const float4* __restrict__ conic_opacity, const float4* __restrict__ rgbd,
                     const float* __restrict__ bg, float* __restrict__ out_color, float* __restrict__ out_depth,
                     float* __restrict__ out_alpha, uint32_t* __restrict__ n_contrib, float* __restrict__ final_T,
                     unsigned long long* __restrict__ stats) {
  __shared__ float4 s_g0[CW_WARPS][32];
  __shared__ float4 s_g1[CW_WARPS][32];
  __shared__ float4 s_c[CW_WARPS][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4* g0s = s_g0[warp];
  float4* g1s = s_g1[warp];
  float4* cs = s_c[warp];
  const float bg0 = bg ? bg[0] : 0.f, bg1 = bg ? bg[1] : 0.f, bg2 = bg ? bg[2] : 0.f;
  const size_t HW = (size_t)H * W;
  const uint32_t num_items = (uint32_t)tiles * ITEMS_PER_TILE;

  while (true) {
    uint32_t item = 0;
    if (lane == 0) item = atomicAdd(ticket, 1u);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= num_items) break;
    const int tile = (int)order[item / ITEMS_PER_TILE];
    const int sub = (int)(item % ITEMS_PER_TILE);
    const int tx = tile % gx, ty = tile / gx;
    const int X0 = tx * TILE + (sub & 1) * ITEM_W, Y0 = ty * TILE + (sub >> 1) * ITEM_H;
    const int px = X0 + (lane & 7), py = Y0 + (lane >> 3);
    const bool inside = px < W && py < H;
    const float pyf = (float)py;
    const float fx0 = (float)X0, fy0 = (float)Y0;
    const uint2 range = ranges[tile];
    const int total = (int)(range.y - range.x);

    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, Dp = 0.f;
    float pxf = inside ? (float)px : PARKED;
    uint32_t last = 0;
    bool live = inside;
    unsigned long long t_start = 0;
    uint32_t n_batches = 0, n_surv = 0, n_hit = 0;
    if (stats) t_start = clock64();
    // two-stage prefetch: the point_list index of batch b+2 and the attributes of batch b+1 are in flight while
    // batch b is composited (a dependent load would otherwise stall the in-order warp at its first use)
    uint32_t id_nxt = 0;
    Staged nxt;
    if (lane < total) gather(point_list, means2D, conic_opacity, rgbd, range.x + lane, nxt);
    if (32 + lane < total) id_nxt = __ldg(point_list + range.x + 32 + lane);
    for (int b0 = 0; b0 < total; b0 += 32) {
      if (__all_sync(0xffffffffu, !live)) break;
      n_batches++;
      __syncwarp();
      if (b0 + lane < total) commit(nxt, g0s, g1s, cs, lane);
      __syncwarp();
      if (b0 + 32 + lane < total) gather_id(id_nxt, means2D, conic_opacity, rgbd, nxt);
      if (b0 + 64 + lane < total) id_nxt = __ldg(point_list + range.x + b0 + 64 + lane);
      const int nb = min(32, total - b0);
      // phase 1: lane j tests staged Gaussian j against this warp's footprint
      const bool may = lane < nb && footprint_may_hit(g0s[lane], g1s[lane], fx0, fx0 + (float)(ITEM_W - 1), fy0,
                                                        fy0 + (float)(ITEM_H - 1));
      uint32_t todo = __ballot_sync(0xffffffffu, may);
      n_surv += __popc(todo);
      // phase 2: front-to-back over the survivors
      while (todo) {
        const int j = __ffs(todo) - 1;
        todo &= todo - 1;
        const float4 g0 = g0s[j];
        const float4 g1 = g1s[j];
        const float dy = __fsub_rn(g0.y, pyf);
        const float bdy = __fmul_rn(g0.w, dy);
        const float cdy2 = __fmul_rn(__fmul_rn(g1.x, dy), dy);
        const float pw = pair_power(g0.z, __fsub_rn(g0.x, pxf), bdy, cdy2);
        if (!(pw >= g1.z) || pw > 0.0f) continue;
        n_hit++;
        const float alpha = fminf(0.99f, __fmul_rn(g1.y, skgs_exp(pw)));
        if (alpha < 1.0f / 255.0f) continue;
        const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
        if (test_T < 0.0001f) {  // this Gaussian is NOT blended; the pixel is finished
          pxf = PARKED;
          live = false;
          continue;
        }
        const float4 c = cs[j];
        const float w = __fmul_rn(alpha, T);
        C0 = __fmaf_rn(c.x, w, C0);
        C1 = __fmaf_rn(c.y, w, C1);
        C2 = __fmaf_rn(c.z, w, C2);
        Dp = __fmaf_rn(c.w, w, Dp);
        T = test_T;
        last = (uint32_t)(b0 + j + 1);
      }
    }
    if (stats) {
      const unsigned long long dt = clock64() - t_start;
      const uint32_t any_hit = __reduce_add_sync(0xffffffffu, n_hit);
      if (lane == 0) {
        unsigned long long* o = stats + (size_t)item * 6;
        o[0] = (unsigned long long)tile; o[1] = (unsigned long long)total; o[2] = n_batches; o[3] = n_surv;
        o[4] = any_hit; o[5] = dt;
      }
    }
    if (inside) {
      const size_t pid = (size_t)py * W + px;
      out_color[pid] = __fmaf_rn(T, bg0, C0);
      out_color[HW + pid] = __fmaf_rn(T, bg1, C1);
      out_color[2 * HW + pid] = __fmaf_rn(T, bg2, C2);
      out_depth[pid] = Dp;
      out_alpha[pid] = __fsub_rn(1.0f, T);
      n_contrib[pid] = last;
      final_T[pid] = T;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// backward.  Same work items as the forward.
// ------------------------------------------------------------------------------------------------------------------
// Sum the parked rows (two per lane), leave each row total in the row's pad word, then lane s < n flushes slot s with
// three 16-byte vector REDs: one RED set per (item, Gaussian).
__device__ __forceinline__ void flush_slots(float* part, const uint32_t* slot_id, int n, int lane, float ddelx_dx,
                                            float ddely_dy, float* __restrict__ ggrad) {
  __syncwarp();
  for (int r = lane; r < n * RVALS; r += 32) {
    const float* row = part + r * RSTRIDE;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int k = 0; k < 32; k += 4) {
      s0 += row[k]; s1 += row[k + 1]; s2 += row[k + 2]; s3 += row[k + 3];
    }
    part[r * RSTRIDE + 32] = (s0 + s1) + (s2 + s3);
  }
  __syncwarp();
  if (lane < n) {
    const float* t = part + lane * (RVALS * RSTRIDE) + 32;
    const float mx = t[0 * RSTRIDE], my = t[1 * RSTRIDE], ca = t[2 * RSTRIDE], cb = t[3 * RSTRIDE];
    const float cc = t[4 * RSTRIDE], op = t[5 * RSTRIDE], z = t[6 * RSTRIDE];
    const float r = t[7 * RSTRIDE], g = t[8 * RSTRIDE], b = t[9 * RSTRIDE];
    float* dst = ggrad + (size_t)slot_id[lane] * NGRAD;
    red_add_v4(dst, mx * ddelx_dx, my * ddely_dy, -0.5f * ca, -0.5f * cb);
    red_add_v4(dst + 4, -0.5f * cc, op, z, 0.f);
    red_add_v4(dst + 8, r, g, b, 0.f);
  }
  __syncwarp();
}

__global__ void __launch_bounds__(CW_THREADS, SKGS_BWD_MINBLOCKS)
composite_bwd_kernel(int W, int H, int gx, int tiles, const uint32_t* __restrict__ order,
                     uint32_t* __restrict__ ticket, const uint2* __restrict__ ranges,
                     const uint32_t* __restrict__ point_list, const float2* __restrict__ means2D,
                     const float4* __restrict__ conic_opacity, const float4* __restrict__ rgbd,
                     const float* __restrict__ bg, const uint32_t* __restrict__ n_contrib,
                     const float* __restrict__ final_T, const float* __restrict__ dL_dpix,
                     const float* __restrict__ dL_ddepth, const float* __restrict__ dL_dalpha_map,
                     float* __restrict__ ggrad) {
  __shared__ float4 s_g0[CW_WARPS][32];
  __shared__ float4 s_g1[CW_WARPS][32];
  __shared__ float4 s_c[CW_WARPS][32];
  // cross-lane reduction through shared memory: every lane parks its 10 partial sums of up to RSLOTS surviving
  // Gaussians (rows of 32 + 1 pad word, conflict free), then the 10*RSLOTS rows are summed two per lane - about 35
  // instructions per surviving Gaussian instead of 100 for ten 5-step shuffle reductions
  __shared__ float s_part[CW_WARPS][RSLOTS * RVALS * RSTRIDE];
  __shared__ uint32_t s_slot_id[CW_WARPS][RSLOTS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4* g0s = s_g0[warp];
  float4* g1s = s_g1[warp];
  float4* cs = s_c[warp];
  float* part = s_part[warp];
  uint32_t* slot_id = s_slot_id[warp];
  const float bg0 = bg ? bg[0] : 0.f, bg1 = bg ? bg[1] : 0.f, bg2 = bg ? bg[2] : 0.f;
  const size_t HW = (size_t)H * W;
  const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
  const uint32_t num_items = (uint32_t)tiles * ITEMS_PER_TILE;

  while (true) {
    uint32_t item = 0;
    if (lane == 0) item = atomicAdd(ticket, 1u);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= num_items) break;
    const int tile = (int)order[item / ITEMS_PER_TILE];
    const int sub = (int)(item % ITEMS_PER_TILE);
    const int tx = tile % gx, ty = tile / gx;
    const int X0 = tx * TILE + (sub & 1) * ITEM_W, Y0 = ty * TILE + (sub >> 1) * ITEM_H;
    const int px = X0 + (lane & 7), py = Y0 + (lane >> 3);
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const float fx0 = (float)X0, fy0 = (float)Y0;
    const uint2 range = ranges[tile];
    const size_t pid = (size_t)py * W + px;

    const uint32_t last = inside ? n_contrib[pid] : 0u;
    const float Tfin = inside ? final_T[pid] : 0.f;
    float T = Tfin;
    const float dp0 = inside ? dL_dpix[pid] : 0.f;
    const float dp1 = inside ? dL_dpix[HW + pid] : 0.f;
    const float dp2 = inside ? dL_dpix[2 * HW + pid] : 0.f;
    const float dD = (inside && dL_ddepth) ? dL_ddepth[pid] : 0.f;
    const float dA = (inside && dL_dalpha_map) ? dL_dalpha_map[pid] : 0.f;
    const float tail = bg0 * dp0 + bg1 * dp1 + bg2 * dp2 - dA;
    float ac0 = 0.f, ac1 = 0.f, ac2 = 0.f, acd = 0.f, la = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, ld = 0.f;
    uint32_t mymax = last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mymax = max(mymax, __shfl_xor_sync(0xffffffffu, mymax, o));
    // positions >= mymax contribute to no pixel of this item: start there and walk to the front
    int nslots = 0;
    uint32_t id_nxt = 0;
    Staged nxt;
    if ((int)mymax - 1 - lane >= 0) gather(point_list, means2D, conic_opacity, rgbd, range.x + mymax - 1 - lane, nxt);
    if ((int)mymax - 33 - lane >= 0) id_nxt = __ldg(point_list + range.x + mymax - 33 - lane);
    for (int top = (int)mymax; top > 0; top -= 32) {
      __syncwarp();
      if (top - 1 - lane >= 0) commit(nxt, g0s, g1s, cs, lane);
      __syncwarp();
      if (top - 33 - lane >= 0) gather_id(id_nxt, means2D, conic_opacity, rgbd, nxt);
      if (top - 65 - lane >= 0) id_nxt = __ldg(point_list + range.x + top - 65 - lane);
      const int nb = min(32, top);
      const bool may = lane < nb && footprint_may_hit(g0s[lane], g1s[lane], fx0, fx0 + (float)(ITEM_W - 1), fy0,
                                                        fy0 + (float)(ITEM_H - 1));
      uint32_t todo = __ballot_sync(0xffffffffu, may);
      while (todo) {
        const int j = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t posn = (uint32_t)(top - 1 - j);
        const float4 g0 = g0s[j];
        const float4 g1 = g1s[j];
        const float dx = __fsub_rn(g0.x, pxf), dy = __fsub_rn(g0.y, pyf);
        const float bdy = __fmul_rn(g0.w, dy);
        const float cdy2 = __fmul_rn(__fmul_rn(g1.x, dy), dy);
        const float pw = pair_power(g0.z, dx, bdy, cdy2);
        float G = 0.f, alpha = 0.f;
        bool valid = (pw >= g1.z) && (posn < last) && !(pw > 0.0f);
        if (valid) {
          G = skgs_exp(pw);
          alpha = fminf(0.99f, __fmul_rn(g1.y, G));
          valid = !(alpha < 1.0f / 255.0f);
        }
        if (!__any_sync(0xffffffffu, valid)) continue;
        float a_mx = 0.f, a_my = 0.f, a_ca = 0.f, a_cb = 0.f, a_cc = 0.f, a_op = 0.f, a_r = 0.f, a_g = 0.f, a_b = 0.f,
              a_z = 0.f;
        if (valid) {
          const float4 c = cs[j];
          const float A = -2.0f * g0.z, B = -g0.w, Cc = -2.0f * g1.x, o = g1.y;
          const float inv = __fdividef(1.0f, 1.0f - alpha);
          T = T * inv;
          const float w = alpha * T;
          ac0 = fmaf(la, lc0 - ac0, ac0);
          ac1 = fmaf(la, lc1 - ac1, ac1);
          ac2 = fmaf(la, lc2 - ac2, ac2);
          acd = fmaf(la, ld - acd, acd);
          lc0 = c.x; lc1 = c.y; lc2 = c.z; ld = c.w;
          float dL_dalpha = (c.x - ac0) * dp0 + (c.y - ac1) * dp1 + (c.z - ac2) * dp2 + (c.w - acd) * dD;
          a_r = w * dp0; a_g = w * dp1; a_b = w * dp2; a_z = w * dD;
          dL_dalpha *= T;
          la = alpha;
          dL_dalpha = fmaf(-Tfin * inv, tail, dL_dalpha);
          const float dL_dG = o * dL_dalpha;
          const float gdx = G * dx, gdy = G * dy;
          a_mx = dL_dG * (-gdx * A - gdy * B);
          a_my = dL_dG * (-gdy * Cc - gdx * B);
          a_ca = gdx * dx * dL_dG;
          a_cb = gdx * dy * dL_dG;
          a_cc = gdy * dy * dL_dG;
          a_op = G * dL_dalpha;
        }
        // park the partial sums of this Gaussian in slot `nslots`
        {
          float* row = part + nslots * (RVALS * RSTRIDE) + lane;
          row[0 * RSTRIDE] = a_mx; row[1 * RSTRIDE] = a_my; row[2 * RSTRIDE] = a_ca; row[3 * RSTRIDE] = a_cb;
          row[4 * RSTRIDE] = a_cc; row[5 * RSTRIDE] = a_op; row[6 * RSTRIDE] = a_z; row[7 * RSTRIDE] = a_r;
          row[8 * RSTRIDE] = a_g; row[9 * RSTRIDE] = a_b;
          if (lane == 0) slot_id[nslots] = __float_as_uint(g1.w);
        }
        nslots++;
        if (nslots == RSLOTS) {
          flush_slots(part, slot_id, nslots, lane, ddelx_dx, ddely_dy, ggrad);
          nslots = 0;
        }
      }
    }
    if (nslots > 0) {
      flush_slots(part, slot_id, nslots, lane, ddelx_dx, ddely_dy, ggrad);
      nslots = 0;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------------------------
unsigned long long* g_item_stats = nullptr;  // debug: per-item counters (skgs_debug_set_item_stats)

static int persistent_grid(const void* kernel) {
  int dev = 0, sms = 0, occ = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, CW_THREADS, 0);
  if (sms <= 0) sms = 148;
  if (occ <= 0) occ = 1;
  return sms * occ;
}

int launch_tile_order(const RasterParams& rp, char* img, const skgs_raster_layout& lay, cudaStream_t st) {
  const int tiles = rp.gx * rp.gy;
  if (tiles == 0) return SKGS_OK;
  {
    ProfScope prof_("tile_order_kernel", st);
    tile_order_kernel<<<1, TO_THREADS, 0, st>>>(reinterpret_cast<const uint2*>(img + lay.ranges), tiles,
                                                reinterpret_cast<uint32_t*>(img + lay.tile_order),
                                                reinterpret_cast<uint32_t*>(img + lay.work_counters));
    SKGS_CHECK_LAUNCH("tile_order_kernel");
  }
  return SKGS_OK;
}

int launch_composite_fwd(const RasterParams& rp, char* geom, char* binning, char* img, const skgs_raster_layout& lay,
                         float* out_color, float* out_depth, float* out_alpha, cudaStream_t st) {
  const int tiles = rp.gx * rp.gy;
  if (tiles == 0) return SKGS_OK;
  static int grid_cap = 0;
  if (grid_cap == 0) grid_cap = persistent_grid((const void*)composite_fwd_kernel);
  const int want = (tiles * ITEMS_PER_TILE + CW_WARPS - 1) / CW_WARPS;
  const int grid = want < grid_cap ? want : grid_cap;
  {
    ProfScope prof_("composite_fwd_kernel", st);
    composite_fwd_kernel<<<grid, CW_THREADS, 0, st>>>(
        rp.W, rp.H, rp.gx, tiles, reinterpret_cast<const uint32_t*>(img + lay.tile_order),
        reinterpret_cast<uint32_t*>(img + lay.work_counters), reinterpret_cast<const uint2*>(img + lay.ranges),
        reinterpret_cast<const uint32_t*>(binning + lay.point_list),
        reinterpret_cast<const float2*>(geom + lay.means2D), reinterpret_cast<const float4*>(geom + lay.conic_opacity),
        reinterpret_cast<const float4*>(geom + lay.rgbd), rp.bg, out_color, out_depth, out_alpha,
        reinterpret_cast<uint32_t*>(img + lay.n_contrib), reinterpret_cast<float*>(img + lay.final_T),
        g_item_stats);
    SKGS_CHECK_LAUNCH("composite_fwd_kernel");
  }
  return SKGS_OK;
}

int launch_composite_bwd(const RasterParams& rp, char* geom, const char* binning, const char* img,
                         const skgs_raster_layout& lay, const float* dL_dcolor, const float* dL_ddepth,
                         const float* dL_dalpha, cudaStream_t st) {
  float* ggrad = reinterpret_cast<float*>(geom + lay.geom_grads);
  SKGS_CUDA(cudaMemsetAsync(ggrad, 0, (size_t)rp.P * NGRAD * sizeof(float), st));
  const int tiles = rp.gx * rp.gy;
  if (tiles == 0 || rp.P == 0) return SKGS_OK;
  uint32_t* counters = reinterpret_cast<uint32_t*>(const_cast<char*>(img) + lay.work_counters);
  SKGS_CUDA(cudaMemsetAsync(counters + 1, 0, sizeof(uint32_t), st));  // backward may be re-run on the same state
  static int grid_cap = 0;
  if (grid_cap == 0) grid_cap = persistent_grid((const void*)composite_bwd_kernel);
  const int want = (tiles * ITEMS_PER_TILE + CW_WARPS - 1) / CW_WARPS;
  const int grid = want < grid_cap ? want : grid_cap;
  {
    ProfScope prof_("composite_bwd_kernel", st);
    composite_bwd_kernel<<<grid, CW_THREADS, 0, st>>>(
        rp.W, rp.H, rp.gx, tiles, reinterpret_cast<const uint32_t*>(img + lay.tile_order), counters + 1,
        reinterpret_cast<const uint2*>(img + lay.ranges), reinterpret_cast<const uint32_t*>(binning + lay.point_list),
        reinterpret_cast<const float2*>(geom + lay.means2D), reinterpret_cast<const float4*>(geom + lay.conic_opacity),
        reinterpret_cast<const float4*>(geom + lay.rgbd), rp.bg, reinterpret_cast<const uint32_t*>(img + lay.n_contrib),
        reinterpret_cast<const float*>(img + lay.final_T), dL_dcolor, dL_ddepth, dL_dalpha, ggrad);
    SKGS_CHECK_LAUNCH("composite_bwd_kernel");
  }
  return SKGS_OK;
}

}  // namespace skgs

extern "C" __attribute__((visibility("default"))) void skgs_debug_set_item_stats(void* p) {
  skgs::g_item_stats = reinterpret_cast<unsigned long long*>(p);
}
