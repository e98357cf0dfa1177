// composite.cu - per-tile alpha compositing, forward and backward, for sm_100a.
//
// Design (measured motivation: profiles/r1_composite_v1.txt -> r1_final_composite_onesweep.txt -> r2_*): one CTA per
// 16x16 tile is latency / tail bound (42 % issue utilisation, SMs idle 40 % of the kernel because a few long tiles finish
// last).  These kernels are WARP-granular and persistent:
//   * work item  = 8 x 4 pixels of a tile (8 items per tile), one warp, ONE pixel per lane;
//   * scheduling = a global ticket hands out work items in order of DEcreasing tile length (the order tile_plan_kernel
//                  of tile_sort.cu writes: one 16-byte descriptor per tile = tile id + list range), so the long tiles
//                  start first and the tail is made of short ones; no __syncthreads anywhere - a warp that finishes
//                  (all its pixels saturated) immediately takes the next item (prefetching the next item's ticket was
//                  measured and removed: profiles/r2_composite_notes.txt);
//   * staging    = 32 Gaussians per batch in the warp's own shared-memory slice; the point_list index of batch b+2 and
//                  the attributes of batch b+1 are in flight while batch b is composited;
//   * culling    = phase 1 of every batch, lane j <-> staged Gaussian j: can any pixel centre of the 8x4 footprint reach
//                  power >= pmin (= -log(255 o) - margin, precomputed per Gaussian by preprocess together with the two
//                  slopes -B/C, -B/A)?  The minimum of the convex quadratic over the box lies on one of the (at most
//                  two) box faces that look at the Gaussian's centre: two clamped 1-D minimisations, no division.  Only
//                  the ballot's set bits are visited by the serial loop; finished pixels are parked at x = +huge;
//   * backward   = every lane parks 9 partial sums (5 moments of q = o G dL/dalpha, opacity, colour; 10 with a depth
//                  cotangent) of up to 6 surviving Gaussians in padded shared-memory rows, the rows are summed two per
//                  lane, lane s turns slot s into the conic / mean gradients and flushes it with three 16-byte vector
//                  REDs (red.global.add.v4.f32): one RED set per (item, Gaussian) instead of the reference's 9-13 scalar
//                  atomics per (pixel, Gaussian) (my_ext/_C/src/nerf/gaussian_render.cu:295-338).
// Semantics: SURVEY.md App. A.6 / A.7 (reference gaussian_render.cu:16-112, 182-341 + bg / depth / alpha terms).
// Every operation that decides WHICH pairs contribute uses the contraction-proof helpers of common.cuh.
#include <cstdlib>

#include "common.cuh"

namespace skgs {

#ifndef SKGS_CW_WARPS
#define SKGS_CW_WARPS 4
#endif
#ifndef SKGS_BWD_MINBLOCKS
#define SKGS_BWD_MINBLOCKS 5
#endif
#ifndef SKGS_RSLOTS
#define SKGS_RSLOTS 6
#endif
#ifndef SKGS_FWD_ILP2
#define SKGS_FWD_ILP2 0
#endif
#ifndef SKGS_FWD_CTAS
#define SKGS_FWD_CTAS 0   // 0: as many CTAs per SM as fit
#endif
#ifndef SKGS_BWD_CTAS
#define SKGS_BWD_CTAS 0
#endif
constexpr int CW_WARPS = SKGS_CW_WARPS;     // warps per CTA (independent workers)
constexpr int CW_THREADS = CW_WARPS * 32;
constexpr int NGRAD = 12;                   // packed per-Gaussian accumulators: mx my ca cb | cc op z - | r g b -
constexpr float PARKED = 1.0e18f;           // x coordinate of a finished pixel: its exponent is -inf
constexpr uint32_t FULL = 0xffffffffu;

// ------------------------------------------------------------------------------------------------------------------
// tile order for a forward WITHOUT binning (P == 0 or no list capacity: every range is (0, 0), the image is the
// background): tiles sorted by decreasing list length (coarse: 8 sub-steps per octave), one CTA; resets the compositing
// tickets.  With binning the same order is produced by tile_plan_kernel (tile_sort.cu).
// ------------------------------------------------------------------------------------------------------------------
constexpr int TO_THREADS = 1024;
constexpr int TO_BINS = 8 * 33;
constexpr uint32_t RANGE_UNSET = 0xffffffffu;

__device__ __forceinline__ int length_bin(uint32_t len) {
  if (len == 0) return 0;
  const int e = 31 - __clz(len);                       // floor(log2 len)
  const int m = e >= 3 ? (int)((len >> (e - 3)) & 7u) : (int)((len << (3 - e)) & 7u);
  return 1 + e * 8 + m;                                // monotone in len
}

__global__ void __launch_bounds__(TO_THREADS)
tile_order_kernel(uint2* __restrict__ ranges, int tiles, uint4* __restrict__ order,
                  uint32_t* __restrict__ counters) {
  __shared__ uint32_t s_hist[TO_BINS];
  __shared__ uint32_t s_base[TO_BINS];
  pdl_wait();
  pdl_trigger();
  for (int k = threadIdx.x; k < TO_BINS; k += TO_THREADS) s_hist[k] = 0;
  if (threadIdx.x < 2) counters[threadIdx.x] = 0;  // work tickets of the forward / backward compositing kernels
  __syncthreads();
  for (int t = threadIdx.x; t < tiles; t += TO_THREADS) {
    uint2 r = ranges[t];
    if (r.x == RANGE_UNSET || r.y <= r.x) {  // tile without list entries (or a forward that bailed out on overflow)
      r = make_uint2(0u, 0u);
      ranges[t] = r;
    }
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
      const uint32_t t = __shfl_up_sync(FULL, incl, o);
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
    const uint2 r = ranges[t];  // this thread's own write above
    const uint32_t p = atomicAdd(&s_base[length_bin(r.y - r.x)], 1u);
    order[p] = make_uint4((uint32_t)t, r.x, r.y, 0u);  // one 16-byte descriptor per work unit: tile, [start, end)
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
  float4 k;    // culling record: pmin, -B/C, -B/A, -
  uint32_t g;  // Gaussian id
};

struct GeomIn {
  const float2* means2D;
  const float4* conic_opacity;
  const float4* rgbd;
  const float4* cull;
};

__device__ __forceinline__ void gather_id(uint32_t g, const GeomIn& G, Staged& s) {
  s.g = g;
  s.m = __ldg(G.means2D + g);
  s.co = __ldg(G.conic_opacity + g);
  s.c = __ldg(G.rgbd + g);
  s.k = __ldg(G.cull + g);
}

// commit a gathered entry to the warp's staging slot: g0 = (gx, gy, A', B'), g1 = (C', opacity, pmin, id), c
// with A' = -A/2, B' = -B, C' = -C/2 (exact scalings)
__device__ __forceinline__ void commit(const Staged& s, float4* g0s, float4* g1s, float4* cs, int lane) {
  g0s[lane] = make_float4(s.m.x, s.m.y, -0.5f * s.co.x, -s.co.y);
  g1s[lane] = make_float4(-0.5f * s.co.z, s.co.w, s.k.x, __uint_as_float(s.g));
  cs[lane] = s.c;
}

// Conservative footprint test (phase 1 of every batch, one staged Gaussian per lane, from registers): can ANY pixel
// centre of the rectangle [X0,X1] x [Y0,Y1] reach power >= pmin?  power = -f/2 with f = A dx^2 + 2 B dx dy + C dy^2
// convex (d = centre - pixel), so its minimum over the box of offsets is 0 if the centre is inside; otherwise it lies
// on a box face that looks at the centre - the x-face nearest to dx = 0 (if 0 is outside the dx range) or the y-face
// nearest to dy = 0 - at the clamped 1-D minimiser dy* = (-B/C) dx resp. dx* = (-B/A) dy (every segment from a box
// point to the centre leaves the box through one of those faces, and f decreases along it).  The margin covers the
// fp32 rounding of both this test and pair_power() (relative 1e-6 of the largest term magnitude over the box), so a
// pair that passes the exact per-pixel test is never culled here.  Non-positive-definite conics carry pmin = -inf.
__device__ __forceinline__ bool footprint_may_hit(const Staged& s, float X0, float X1, float Y0, float Y1) {
  const float A = s.co.x, B = s.co.y, C = s.co.z, pmin = s.k.x;
  const float dxlo = s.m.x - X1, dxhi = s.m.x - X0, dylo = s.m.y - Y1, dyhi = s.m.y - Y0;
  const bool in_x = dxlo <= 0.f && dxhi >= 0.f, in_y = dylo <= 0.f && dyhi >= 0.f;
  const float ex = dxlo > 0.f ? dxlo : dxhi;  // x-face nearest to the centre (meaningful when !in_x)
  const float ey = dylo > 0.f ? dylo : dyhi;
  // face dx = ex
  const float ty = fminf(fmaxf(s.k.y * ex, dylo), dyhi);
  const float fxe = ex * (A * ex + 2.0f * B * ty) + C * ty * ty;
  // face dy = ey
  const float tx = fminf(fmaxf(s.k.z * ey, dxlo), dxhi);
  const float fye = ey * (C * ey + 2.0f * B * tx) + A * tx * tx;
  const float inf = __int_as_float(0x7f800000);
  float fmin = fminf(in_x ? inf : fxe, in_y ? inf : fye);
  if (in_x && in_y) fmin = 0.f;
  const float DX = fmaxf(fabsf(dxlo), fabsf(dxhi)), DY = fmaxf(fabsf(dylo), fabsf(dyhi));
  const float eps = 1e-6f * (fabsf(A) * DX * DX + 2.0f * fabsf(B) * DX * DY + fabsf(C) * DY * DY) + 1e-3f;
  return !(-0.5f * fmin < pmin - eps);  // NaN-safe: anything unordered is kept
}

// ------------------------------------------------------------------------------------------------------------------
// work distribution: one global ticket per item; the descriptor (tile, start, end) of ticket / 8 comes from the list
// tile_order_kernel sorted by decreasing tile length.  (Claiming the NEXT item while the current one runs was tried -
// profiles/r2_composite_prefetch.txt: it binds the second-longest items to the warps that are busy with the longest
// ones and the tail of the kernel grows by 10 %.)
// ------------------------------------------------------------------------------------------------------------------
constexpr int ITEM_W = 8, ITEM_H = 4, ITEMS_PER_TILE = (TILE / ITEM_W) * (TILE / ITEM_H);

struct Item {
  uint32_t id;     // ticket (>= num_items: none)
  int tile;
  uint2 range;
};

__device__ __forceinline__ Item claim_item(uint32_t* ticket, uint32_t num_items, const uint4* __restrict__ order,
                                           int lane) {
  Item it;
  uint32_t t = 0;
  if (lane == 0) t = atomicAdd(ticket, 1u);
  it.id = __shfl_sync(FULL, t, 0);
  it.tile = 0;
  it.range = make_uint2(0u, 0u);
  if (it.id < num_items) {
    const uint4 w = __ldg(order + it.id / ITEMS_PER_TILE);
    it.tile = (int)w.x;
    it.range = make_uint2(w.y, w.z);
  }
  return it;
}

// ------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CW_THREADS)
composite_fwd_kernel(int W, int H, int gx, int tiles, const uint4* __restrict__ order,
                     uint32_t* __restrict__ ticket, const uint32_t* __restrict__ point_list,
                     const skgs_raster_header* __restrict__ hdr, GeomIn G, const float* __restrict__ bg,
                     float* __restrict__ out_color, float* __restrict__ out_depth, float* __restrict__ out_alpha,
                     uint32_t* __restrict__ n_contrib, float* __restrict__ final_T) {
  __shared__ float4 s_g0[CW_WARPS][32];
  __shared__ float4 s_g1[CW_WARPS][32];
  __shared__ float4 s_c[CW_WARPS][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4* g0s = s_g0[warp];
  float4* g1s = s_g1[warp];
  float4* cs = s_c[warp];
  pdl_wait();
  pdl_trigger();
  const float bg0 = bg ? bg[0] : 0.f, bg1 = bg ? bg[1] : 0.f, bg2 = bg ? bg[2] : 0.f;
  const size_t HW = (size_t)H * W;
  const uint32_t num_items = (uint32_t)tiles * ITEMS_PER_TILE;

  while (true) {
    const Item it = claim_item(ticket, num_items, order, lane);
    if (it.id >= num_items) break;
    const int sub = (int)(it.id % ITEMS_PER_TILE);
    const int tx = it.tile % gx, ty = it.tile / gx;
    const int X0 = tx * TILE + (sub & 1) * ITEM_W, Y0 = ty * TILE + (sub >> 1) * ITEM_H;
    const int px = X0 + (lane & 7), py = Y0 + (lane >> 3);
    const bool inside = px < W && py < H;
    const float pyf = (float)py;
    const float fx0 = (float)X0, fy0 = (float)Y0;
    const uint2 range = it.range;
    const int total = (int)(range.y - range.x);

    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, Dp = 0.f;
    float pxf = inside ? (float)px : PARKED;
    uint32_t last = 0;
    bool live = inside;
    // two-stage prefetch: the point_list index of batch b+2 and the attributes of batch b+1 are in flight while
    // batch b is composited (a dependent load would otherwise stall the in-order warp at its first use)
    uint32_t id_nxt = 0;
    Staged nxt;
    if (lane < total) gather_id(__ldg(point_list + range.x + lane), G, nxt);
    if (32 + lane < total) id_nxt = __ldg(point_list + range.x + 32 + lane);
    for (int b0 = 0; b0 < total; b0 += 32) {
      if (__all_sync(FULL, !live)) break;
      const int nb = min(32, total - b0);
      // phase 1: lane j tests staged Gaussian j (still in its registers) against this warp's footprint
      const bool may = lane < nb && footprint_may_hit(nxt, fx0, fx0 + (float)(ITEM_W - 1), fy0,
                                                        fy0 + (float)(ITEM_H - 1));
      uint32_t todo = __ballot_sync(FULL, may);
      __syncwarp();
      if (may) commit(nxt, g0s, g1s, cs, lane);
      __syncwarp();
      if (b0 + 32 + lane < total) gather_id(id_nxt, G, nxt);
      if (b0 + 64 + lane < total) id_nxt = __ldg(point_list + range.x + b0 + 64 + lane);
#if SKGS_FWD_ILP2
      // phase 2: front-to-back over the survivors, TWO per iteration: the gathers from shared memory, the exponents
      // and the exp() of both are independent and overlap; only the blend (T, colour) is applied in list order.  A
      // warp runs one item alone and the longest items bound the kernel, so per-warp latency matters, not just issue.
      while (todo) {
        const int ja = __ffs(todo) - 1;
        todo &= todo - 1;
        const bool two = todo != 0u;
        const int jb = two ? __ffs(todo) - 1 : ja;
        todo &= todo - 1;
        const float4 g0a = g0s[ja], g1a = g1s[ja];
        const float4 g0b = g0s[jb], g1b = g1s[jb];
        const float dya = __fsub_rn(g0a.y, pyf), dyb = __fsub_rn(g0b.y, pyf);
        const float pwa = pair_power(g0a.z, __fsub_rn(g0a.x, pxf), __fmul_rn(g0a.w, dya),
                                     __fmul_rn(__fmul_rn(g1a.x, dya), dya));
        const float pwb = pair_power(g0b.z, __fsub_rn(g0b.x, pxf), __fmul_rn(g0b.w, dyb),
                                     __fmul_rn(__fmul_rn(g1b.x, dyb), dyb));
        const bool ha = (pwa >= g1a.z) && !(pwa > 0.0f);
        const bool hb = two && (pwb >= g1b.z) && !(pwb > 0.0f);
        if (!(ha || hb)) continue;
        const float aa = fminf(0.99f, __fmul_rn(g1a.y, skgs_exp(pwa)));
        const float ab = fminf(0.99f, __fmul_rn(g1b.y, skgs_exp(pwb)));
        if (ha && !(aa < 1.0f / 255.0f)) {
          const float test_T = __fmul_rn(T, __fsub_rn(1.0f, aa));
          if (test_T < 0.0001f) {  // this Gaussian is NOT blended; the pixel is finished
            pxf = PARKED;
            live = false;
          } else {
            const float4 c = cs[ja];
            const float w = __fmul_rn(aa, T);
            C0 = __fmaf_rn(c.x, w, C0);
            C1 = __fmaf_rn(c.y, w, C1);
            C2 = __fmaf_rn(c.z, w, C2);
            Dp = __fmaf_rn(c.w, w, Dp);
            T = test_T;
            last = (uint32_t)(b0 + ja + 1);
          }
        }
        if (hb && live && !(ab < 1.0f / 255.0f)) {
          const float test_T = __fmul_rn(T, __fsub_rn(1.0f, ab));
          if (test_T < 0.0001f) {
            pxf = PARKED;
            live = false;
          } else {
            const float4 c = cs[jb];
            const float w = __fmul_rn(ab, T);
            C0 = __fmaf_rn(c.x, w, C0);
            C1 = __fmaf_rn(c.y, w, C1);
            C2 = __fmaf_rn(c.z, w, C2);
            Dp = __fmaf_rn(c.w, w, Dp);
            T = test_T;
            last = (uint32_t)(b0 + jb + 1);
          }
        }
      }
    }
#else
      // phase 2: front-to-back over the survivors.  (Two survivors per iteration - staging reads, exponents and exp()
      // of both overlapped - was measured: +13 % instructions for no gain, the kernel is issue-bound with 2.5 eligible
      // warps per scheduler, profiles/r2_composite_ilp.txt; SKGS_FWD_ILP2=1 builds that variant.)
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
#endif
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
constexpr int RSLOTS = SKGS_RSLOTS, RSTRIDE = 36;  // reduction staging: rows of 32 + 4 pad words (16-byte aligned rows)

// Sum the parked rows (two per lane, eight 16-byte reads each: with a row stride of 36 words the quarter-warps of an
// LDS.128 hit disjoint banks), leave each row total in the row's first pad word, then lane s < n turns slot s into
// gradients and flushes it with three 16-byte vector REDs: one RED set per (item, Gaussian).
// Row order of a slot: S1x S1y Sxx Sxy Syy op r g b [z] with S.. the moments of q = o G dL/dalpha over the pixels:
//   dL/dmean2D = -(A S1x + B S1y, C S1y + B S1x) * (W/2, H/2),   dL/dconic = -1/2 (Sxx, Sxy, Syy)   (App. A.7)
template <int RV>
__device__ __forceinline__ void flush_slots(float* part, const uint32_t* slot_id, const float4* slot_abc, int n,
                                            int lane, float ddelx_dx, float ddely_dy, float* __restrict__ ggrad) {
  __syncwarp();
  for (int r = lane; r < n * RV; r += 32) {
    const float4* row = reinterpret_cast<const float4*>(part + r * RSTRIDE);
    float4 a = row[0], b = row[1];
#pragma unroll
    for (int k = 2; k < 8; k += 2) {
      const float4 c = row[k], d = row[k + 1];
      a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w;
      b.x += d.x; b.y += d.y; b.z += d.z; b.w += d.w;
    }
    part[r * RSTRIDE + 32] = ((a.x + b.x) + (a.y + b.y)) + ((a.z + b.z) + (a.w + b.w));
  }
  __syncwarp();
  if (lane < n) {
    const float* t = part + lane * (RV * RSTRIDE) + 32;
    const float S1x = t[0 * RSTRIDE], S1y = t[1 * RSTRIDE], Sxx = t[2 * RSTRIDE], Sxy = t[3 * RSTRIDE];
    const float Syy = t[4 * RSTRIDE], op = t[5 * RSTRIDE];
    const float r = t[6 * RSTRIDE], g = t[7 * RSTRIDE], b = t[8 * RSTRIDE];
    const float z = RV > 9 ? t[9 * RSTRIDE] : 0.f;
    const float4 abc = slot_abc[lane];  // conic A, B, C of the slot's Gaussian
    float* dst = ggrad + (size_t)slot_id[lane] * NGRAD;
    red_add_v4(dst, -(abc.x * S1x + abc.y * S1y) * ddelx_dx, -(abc.z * S1y + abc.y * S1x) * ddely_dy, -0.5f * Sxx,
               -0.5f * Sxy);
    red_add_v4(dst + 4, -0.5f * Syy, op, z, 0.f);
    red_add_v4(dst + 8, r, g, b, 0.f);
  }
  __syncwarp();
}

// AUX: a depth and/or alpha cotangent is present (upstream boundary B1 hands them over; the SK_GS training step does
// not, and then the depth channel - two accumulators, one reduced value - is not carried at all)
template <bool AUX>
__global__ void __launch_bounds__(CW_THREADS, SKGS_BWD_MINBLOCKS)
composite_bwd_kernel(int W, int H, int gx, int tiles, const uint4* __restrict__ order,
                     uint32_t* __restrict__ ticket, const uint32_t* __restrict__ point_list,
                     const skgs_raster_header* __restrict__ hdr, GeomIn G, const float* __restrict__ bg,
                     const uint32_t* __restrict__ n_contrib, const float* __restrict__ final_T,
                     const float* __restrict__ dL_dpix, const float* __restrict__ dL_ddepth,
                     const float* __restrict__ dL_dalpha_map, float* __restrict__ ggrad, int tfinal_via_opacity) {
  constexpr int RV = AUX ? 10 : 9;
  __shared__ float4 s_g0[CW_WARPS][32];
  __shared__ float4 s_g1[CW_WARPS][32];
  __shared__ float4 s_c[CW_WARPS][32];
  // cross-lane reduction through shared memory: every lane parks its RV partial sums of up to RSLOTS surviving
  // Gaussians (rows of 32 + 4 pad words), then the RV*RSLOTS rows are summed two per lane - about 25 instructions per
  // surviving Gaussian instead of 100 for ten 5-step shuffle reductions
  __shared__ __align__(16) float s_part[CW_WARPS][RSLOTS * RV * RSTRIDE];
  __shared__ uint32_t s_slot_id[CW_WARPS][RSLOTS];
  __shared__ float4 s_slot_abc[CW_WARPS][RSLOTS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4* g0s = s_g0[warp];
  float4* g1s = s_g1[warp];
  float4* cs = s_c[warp];
  float* part = s_part[warp];
  uint32_t* slot_id = s_slot_id[warp];
  float4* slot_abc = s_slot_abc[warp];
  pdl_wait();
  pdl_trigger();
  const float bg0 = bg ? bg[0] : 0.f, bg1 = bg ? bg[1] : 0.f, bg2 = bg ? bg[2] : 0.f;
  const size_t HW = (size_t)H * W;
  const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
  const uint32_t num_items = (uint32_t)tiles * ITEMS_PER_TILE;

  while (true) {
    const Item it = claim_item(ticket, num_items, order, lane);
    if (it.id >= num_items) break;
    const int sub = (int)(it.id % ITEMS_PER_TILE);
    const int tx = it.tile % gx, ty = it.tile / gx;
    const int X0 = tx * TILE + (sub & 1) * ITEM_W, Y0 = ty * TILE + (sub >> 1) * ITEM_H;
    const int px = X0 + (lane & 7), py = Y0 + (lane >> 3);
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const float fx0 = (float)X0, fy0 = (float)Y0;
    const uint2 range = it.range;
    const size_t pid = (size_t)py * W + px;

    const uint32_t last = inside ? n_contrib[pid] : 0u;
    float Tfin = inside ? final_T[pid] : 0.f;
    // test hook (settings.debug bit 2): recover T_final the way the reference's in-tree extension does, from its stored
    // opacity 1 - T (gaussian_render.cu:215 `T_final = 1.0f - out_opacity`): the cancellation costs up to 1e-4 of
    // relative accuracy in every gradient.  Off by default: the exact final_T (upstream semantics) is used.
    if (tfinal_via_opacity) Tfin = __fsub_rn(1.0f, __fsub_rn(1.0f, Tfin));
    float T = Tfin;
    const float dp0 = inside ? dL_dpix[pid] : 0.f;
    const float dp1 = inside ? dL_dpix[HW + pid] : 0.f;
    const float dp2 = inside ? dL_dpix[2 * HW + pid] : 0.f;
    const float dD = (AUX && inside && dL_ddepth) ? dL_ddepth[pid] : 0.f;
    const float dA = (AUX && inside && dL_dalpha_map) ? dL_dalpha_map[pid] : 0.f;
    const float tail = bg0 * dp0 + bg1 * dp1 + bg2 * dp2 - dA;
    float ac0 = 0.f, ac1 = 0.f, ac2 = 0.f, acd = 0.f, la = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, ld = 0.f;
    uint32_t mymax = last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mymax = max(mymax, __shfl_xor_sync(FULL, mymax, o));
    // positions >= mymax contribute to no pixel of this item: start there and walk to the front
    int nslots = 0;
    uint32_t id_nxt = 0;
    Staged nxt;
    if ((int)mymax - 1 - lane >= 0) gather_id(__ldg(point_list + range.x + mymax - 1 - lane), G, nxt);
    if ((int)mymax - 33 - lane >= 0) id_nxt = __ldg(point_list + range.x + mymax - 33 - lane);
    for (int top = (int)mymax; top > 0; top -= 32) {
      const int nb = min(32, top);
      const bool may = lane < nb && footprint_may_hit(nxt, fx0, fx0 + (float)(ITEM_W - 1), fy0,
                                                        fy0 + (float)(ITEM_H - 1));
      uint32_t todo = __ballot_sync(FULL, may);
      __syncwarp();
      if (may) commit(nxt, g0s, g1s, cs, lane);
      __syncwarp();
      if (top - 33 - lane >= 0) gather_id(id_nxt, G, nxt);
      if (top - 65 - lane >= 0) id_nxt = __ldg(point_list + range.x + top - 65 - lane);
      // the sequential part of one surviving Gaussian: T, the back-to-front colour recurrence, its 9 (10) partial sums
      auto apply = [&](bool valid, const float4& g0, const float4& g1, const float4& c, float dx, float dy, float G_,
                       float alpha) {
        float a_1x = 0.f, a_1y = 0.f, a_xx = 0.f, a_xy = 0.f, a_yy = 0.f, a_op = 0.f, a_r = 0.f, a_g = 0.f, a_b = 0.f,
              a_z = 0.f;
        if (valid) {
          const float inv = __fdividef(1.0f, 1.0f - alpha);
          T = T * inv;
          const float w = alpha * T;
          ac0 = fmaf(la, lc0 - ac0, ac0);
          ac1 = fmaf(la, lc1 - ac1, ac1);
          ac2 = fmaf(la, lc2 - ac2, ac2);
          lc0 = c.x; lc1 = c.y; lc2 = c.z;
          float dL_dalpha = (c.x - ac0) * dp0 + (c.y - ac1) * dp1 + (c.z - ac2) * dp2;
          a_r = w * dp0; a_g = w * dp1; a_b = w * dp2;
          if (AUX) {
            acd = fmaf(la, ld - acd, acd);
            ld = c.w;
            dL_dalpha = fmaf(c.w - acd, dD, dL_dalpha);
            a_z = w * dD;
          }
          dL_dalpha *= T;
          la = alpha;
          dL_dalpha = fmaf(-Tfin * inv, tail, dL_dalpha);
          a_op = G_ * dL_dalpha;
          const float q = g1.y * a_op;  // o G dL/dalpha
          const float qx = q * dx, qy = q * dy;
          a_1x = qx; a_1y = qy;
          a_xx = qx * dx; a_xy = qx * dy; a_yy = qy * dy;
        }
        // park the partial sums of this Gaussian in slot `nslots`
        float* row = part + nslots * (RV * RSTRIDE) + lane;
        row[0 * RSTRIDE] = a_1x; row[1 * RSTRIDE] = a_1y; row[2 * RSTRIDE] = a_xx; row[3 * RSTRIDE] = a_xy;
        row[4 * RSTRIDE] = a_yy; row[5 * RSTRIDE] = a_op; row[6 * RSTRIDE] = a_r; row[7 * RSTRIDE] = a_g;
        row[8 * RSTRIDE] = a_b;
        if (AUX) row[9 * RSTRIDE] = a_z;
        if (lane == 0) {
          slot_id[nslots] = __float_as_uint(g1.w);
          slot_abc[nslots] = make_float4(-2.0f * g0.z, -g0.w, -2.0f * g1.x, 0.f);
        }
        nslots++;
        if (nslots == RSLOTS) {
          flush_slots<RV>(part, slot_id, slot_abc, nslots, lane, ddelx_dx, ddely_dy, ggrad);
          nslots = 0;
        }
      };
      // back to front over the survivors, TWO per iteration: staging reads, exponents and exp() of both overlap, the
      // recurrences are applied in list order (a warp runs an item alone; the longest items bound the kernel)
      while (todo) {
        const int ja = __ffs(todo) - 1;
        todo &= todo - 1;
        const bool two = todo != 0u;
        const int jb = two ? __ffs(todo) - 1 : ja;
        todo &= todo - 1;
        const uint32_t posa = (uint32_t)(top - 1 - ja), posb = (uint32_t)(top - 1 - jb);
        const float4 g0a = g0s[ja], g1a = g1s[ja], ca = cs[ja];
        const float4 g0b = g0s[jb], g1b = g1s[jb], cb = cs[jb];
        const float dxa = __fsub_rn(g0a.x, pxf), dya = __fsub_rn(g0a.y, pyf);
        const float dxb = __fsub_rn(g0b.x, pxf), dyb = __fsub_rn(g0b.y, pyf);
        const float pwa = pair_power(g0a.z, dxa, __fmul_rn(g0a.w, dya), __fmul_rn(__fmul_rn(g1a.x, dya), dya));
        const float pwb = pair_power(g0b.z, dxb, __fmul_rn(g0b.w, dyb), __fmul_rn(__fmul_rn(g1b.x, dyb), dyb));
        bool va = (pwa >= g1a.z) && (posa < last) && !(pwa > 0.0f);
        bool vb = two && (pwb >= g1b.z) && (posb < last) && !(pwb > 0.0f);
        float Ga = 0.f, Gb = 0.f, aa = 0.f, ab = 0.f;
        if (va || vb) {
          Ga = skgs_exp(pwa);
          Gb = skgs_exp(pwb);
          aa = fminf(0.99f, __fmul_rn(g1a.y, Ga));
          ab = fminf(0.99f, __fmul_rn(g1b.y, Gb));
          va = va && !(aa < 1.0f / 255.0f);
          vb = vb && !(ab < 1.0f / 255.0f);
        }
        const bool anya = __any_sync(FULL, va), anyb = __any_sync(FULL, vb);
        if (anya) apply(va, g0a, g1a, ca, dxa, dya, Ga, aa);
        if (anyb) apply(vb, g0b, g1b, cb, dxb, dyb, Gb, ab);
      }
    }
    if (nslots > 0) {
      flush_slots<RV>(part, slot_id, slot_abc, nslots, lane, ddelx_dx, ddely_dy, ggrad);
      nslots = 0;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------------------------
// Persistent grid = SMs x resident CTAs.  A warp runs one work item alone, and the longest items bound the kernel: their
// speed is the warp's share of its scheduler's issue slots, so FEWER resident warps can be faster than full occupancy.
// `max_ctas` (> 0) caps the CTAs per SM by padding the launch with dynamic shared memory (env SKGS_FWD_CTAS / _BWD_CTAS
// override the built-in choice; tuning: profiles/r2_composite_occupancy.txt).
struct PersistentCfg {
  int grid_cap;
  size_t dyn_smem;
};

static PersistentCfg persistent_cfg(const void* kernel, int max_ctas) {
  int dev = 0, sms = 0, occ = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, CW_THREADS, 0);
  if (sms <= 0) sms = 148;
  if (occ <= 0) occ = 1;
  PersistentCfg c{sms * occ, 0};
  if (max_ctas > 0 && max_ctas < occ) {
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kernel);
    int smem_sm = 0;
    cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    const long per_cta = smem_sm / max_ctas - 1024 - (long)fa.sharedSizeBytes - 256;  // 1 KB is reserved per CTA
    if (per_cta > 0) {
      c.dyn_smem = (size_t)per_cta / 128 * 128;
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.dyn_smem);
      int occ2 = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, kernel, CW_THREADS, c.dyn_smem);
      if (occ2 > 0) c.grid_cap = sms * occ2;
    }
  }
  return c;
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

static GeomIn geom_in(const char* geom, const skgs_raster_layout& lay) {
  GeomIn G;
  G.means2D = reinterpret_cast<const float2*>(geom + lay.means2D);
  G.conic_opacity = reinterpret_cast<const float4*>(geom + lay.conic_opacity);
  G.rgbd = reinterpret_cast<const float4*>(geom + lay.rgbd);
  G.cull = reinterpret_cast<const float4*>(geom + lay.cull);
  return G;
}

int launch_tile_order(const RasterParams& rp, char* img, const skgs_raster_layout& lay, cudaStream_t st) {
  const int tiles = rp.gx * rp.gy;
  if (tiles == 0) return SKGS_OK;
  {
    ProfScope prof_("tile_order_kernel", st);
    SKGS_CUDA(launch_pdl(tile_order_kernel, dim3(1), dim3(TO_THREADS), 0, st, reinterpret_cast<uint2*>(img + lay.ranges),
                         tiles, reinterpret_cast<uint4*>(img + lay.tile_order),
                         reinterpret_cast<uint32_t*>(img + lay.work_counters)));
    SKGS_CHECK_LAUNCH("tile_order_kernel");
  }
  return SKGS_OK;
}

int launch_composite_fwd(const RasterParams& rp, char* geom, char* binning, char* img, const skgs_raster_layout& lay,
                         float* out_color, float* out_depth, float* out_alpha, cudaStream_t st) {
  const int tiles = rp.gx * rp.gy;
  if (tiles == 0) return SKGS_OK;
  static PersistentCfg cfg{0, 0};
  if (cfg.grid_cap == 0) cfg = persistent_cfg((const void*)composite_fwd_kernel, env_int("SKGS_FWD_CTAS", SKGS_FWD_CTAS));
  const int want = (tiles * ITEMS_PER_TILE + CW_WARPS - 1) / CW_WARPS;
  const int grid = want < cfg.grid_cap ? want : cfg.grid_cap;
  {
    ProfScope prof_("composite_fwd_kernel", st);
    SKGS_CUDA(launch_pdl(composite_fwd_kernel, dim3(grid), dim3(CW_THREADS), cfg.dyn_smem, st, rp.W, rp.H, rp.gx, tiles,
                         reinterpret_cast<const uint4*>(img + lay.tile_order),
                         reinterpret_cast<uint32_t*>(img + lay.work_counters),
                         reinterpret_cast<const uint32_t*>(binning + lay.vals),
                         reinterpret_cast<const skgs_raster_header*>(geom + lay.header), geom_in(geom, lay), rp.bg,
                         out_color, out_depth, out_alpha, reinterpret_cast<uint32_t*>(img + lay.n_contrib),
                         reinterpret_cast<float*>(img + lay.final_T)));
    SKGS_CHECK_LAUNCH("composite_fwd_kernel");
  }
  return SKGS_OK;
}

// The per-Gaussian accumulators (geom_grads) are zero on entry: the forward zeroes them, preprocess_bwd re-zeroes them
// (and the backward ticket) after consuming them - no memset between the loss and this kernel.
int launch_composite_bwd(const RasterParams& rp, char* geom, const char* binning, const char* img,
                         const skgs_raster_layout& lay, const float* dL_dcolor, const float* dL_ddepth,
                         const float* dL_dalpha, int tfinal_via_opacity, cudaStream_t st) {
  float* ggrad = reinterpret_cast<float*>(geom + lay.geom_grads);
  const int tiles = rp.gx * rp.gy;
  if (tiles == 0 || rp.P == 0) return SKGS_OK;
  uint32_t* counters = reinterpret_cast<uint32_t*>(const_cast<char*>(img) + lay.work_counters);
  const bool aux = dL_ddepth != nullptr || dL_dalpha != nullptr;
  static PersistentCfg cfg[2] = {{0, 0}, {0, 0}};
  if (cfg[aux].grid_cap == 0)
    cfg[aux] = persistent_cfg(aux ? (const void*)composite_bwd_kernel<true> : (const void*)composite_bwd_kernel<false>,
                              env_int("SKGS_BWD_CTAS", SKGS_BWD_CTAS));
  const int want = (tiles * ITEMS_PER_TILE + CW_WARPS - 1) / CW_WARPS;
  const int grid = want < cfg[aux].grid_cap ? want : cfg[aux].grid_cap;
  {
    ProfScope prof_("composite_bwd_kernel", st);
    auto kern = aux ? composite_bwd_kernel<true> : composite_bwd_kernel<false>;
    SKGS_CUDA(launch_pdl(kern, dim3(grid), dim3(CW_THREADS), cfg[aux].dyn_smem, st, rp.W, rp.H, rp.gx, tiles,
                         reinterpret_cast<const uint4*>(img + lay.tile_order), counters + 1,
                         reinterpret_cast<const uint32_t*>(binning + lay.vals),
                         reinterpret_cast<const skgs_raster_header*>(geom + lay.header), geom_in(geom, lay), rp.bg,
                         reinterpret_cast<const uint32_t*>(img + lay.n_contrib),
                         reinterpret_cast<const float*>(img + lay.final_T), dL_dcolor, dL_ddepth, dL_dalpha, ggrad,
                         tfinal_via_opacity));
    SKGS_CHECK_LAUNCH("composite_bwd_kernel");
  }
  return SKGS_OK;
}

}  // namespace skgs
