// tile_sort.cu - binning by TILE-SEGMENTED sort: the per-tile depth-ordered lists of the rasterizer without a
// device-wide radix sort.
//
// The reference sorts all R (tile << 32 | depth) keys with cub::DeviceRadixSort::SortPairs
// (gaussian_rasterizer_forward.cu:226-229; 6 passes of 8 bits at 800x800), then finds the tile boundaries
// (identifyTileRanges, :77-94).  A device-wide pass reads and scatters 12 B per key and chains a prefix across every
// CTA; five of them were ~105 us of the 436 us step.  The key's upper half is a small integer whose histogram is known
// the moment the keys are emitted, so the sort is split at that boundary:
//   (emitting kernel)    every Gaussian adds its tile rectangle to a 2-D difference grid (4 atomics per Gaussian);
//   tile_plan_kernel     one CTA: 2-D prefix sum = keys per tile, exclusive scan = the tile ranges (no boundary search),
//                        tiles ordered by decreasing length (the schedule of the sort and of the compositing kernels);
//   tile_scatter_kernel  ONE pass over the keys: entry -> its tile's segment.  A CTA counts the tiles of a 4096-entry
//                        chunk in shared memory and reserves one range per (chunk, tile) with a single global atomic
//                        (a hot tile sees one atomic per chunk, not one per entry);
//   tile_sort_kernel     every segment is sorted on its own, entirely in shared memory: LSD radix sort over the depth
//                        bytes (stable warp-synchronous ranking as in the device-wide pass it replaces, but the digit
//                        prefix is local - no look-back, no global round trip between passes; bytes that are constant
//                        over the segment are skipped), bitonic network for short segments.
// Result: bit-identical lists.  The reference's sort is stable and one Gaussian emits at most one key per tile, in
// Gaussian order: its order inside a tile is (depth, Gaussian id) ascending.  The scatter does not preserve emission
// order, so after the depth passes a segment with EQUAL depths out of id order (rare: exact float ties) is re-sorted
// with the id bytes as the least significant digits.
#include "common.cuh"

namespace skgs {

namespace {

constexpr uint32_t FULLM = 0xffffffffu;
constexpr int TP_THREADS = 1024;
constexpr int TP_BINS = 8 * 33;
constexpr int TP_SMEM_CELLS = 11 * 1024;  // difference grids up to this many cells are summed in shared memory

constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 16;
constexpr int SC_CHUNK = SC_THREADS * SC_ITEMS;
constexpr int SC_MAX_TILES = 12 * 1024;   // tile counters of a chunk in shared memory (48 KB)

constexpr int BITONIC_MAX = 512;          // segments up to this length take the comparison network
constexpr int RS_DIGITS = 256;
constexpr int TIE_RUN_MAX = 16;          // equal-depth runs up to this length are fixed by insertion

__device__ __forceinline__ int length_bin(uint32_t len) {  // the bins of the compositing schedule
  if (len == 0) return 0;
  const int e = 31 - __clz(len);
  const int m = e >= 3 ? (int)((len >> (e - 3)) & 7u) : (int)((len << (3 - e)) & 7u);
  return 1 + e * 8 + m;
}

// ------------------------------------------------------------------------------------------------------------------
// plan: difference grid -> keys per tile -> ranges (reference form: [start, end), (0, 0) for an empty tile), work
// order, tickets.  counters: [0] forward compositing ticket, [1] backward, [2] large-sort ticket, [3] number of large
// tiles (len >= large_len), [4] regular-sort ticket.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TP_THREADS)
tile_plan_kernel(int* __restrict__ grid, int gx, int gy, const skgs_raster_header* __restrict__ hdr,
                 uint2* __restrict__ ranges, uint4* __restrict__ order, uint32_t* __restrict__ counters,
                 uint32_t large_len, int cell_stride) {
  extern __shared__ int s_grid[];
  __shared__ uint32_t s_hist[TP_BINS];
  __shared__ uint32_t s_base[TP_BINS];
  __shared__ uint32_t s_wsum[TP_THREADS / 32];
  __shared__ uint32_t s_nlarge;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gs = gx + 1, cells = gs * (gy + 1), tiles = gx * gy;
  pdl_wait();
  pdl_trigger();
  for (int k = tid; k < TP_BINS; k += TP_THREADS) s_hist[k] = 0;
  if (tid == 0) s_nlarge = 0;
  // an overflowed emission left incomplete lists: hand out empty ranges (the image is invalid, the flag says so)
  const bool dead = hdr->overflow != 0;
  // ---- 2-D inclusive prefix of the difference grid: rows, then columns
  // small grids are summed in shared memory, large ones in place in global memory (cell k lives at k * gstride)
  const bool in_smem = cells <= TP_SMEM_CELLS;
  int* G = in_smem ? s_grid : grid;
  const int gstride = in_smem ? 1 : cell_stride;
  if (in_smem)
    for (int k = tid; k < cells; k += TP_THREADS) G[k] = grid[(size_t)k * cell_stride];
  __syncthreads();
  for (int y = warp; y <= gy; y += TP_THREADS / 32) {  // rows: one warp per row, 32 cells per shuffle scan
    int carry = 0;
    for (int x0 = 0; x0 <= gx; x0 += 32) {
      const int x = x0 + lane;
      int v = x <= gx ? G[(size_t)(y * gs + x) * gstride] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULLM, v, o);
        if (lane >= o) v += t;
      }
      v += carry;
      if (x <= gx) G[(size_t)(y * gs + x) * gstride] = v;
      carry = __shfl_sync(FULLM, v, 31);
    }
  }
  __syncthreads();
  for (int x = warp; x <= gx; x += TP_THREADS / 32) {  // columns
    int carry = 0;
    for (int y0 = 0; y0 <= gy; y0 += 32) {
      const int y = y0 + lane;
      int v = y <= gy ? G[(size_t)(y * gs + x) * gstride] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULLM, v, o);
        if (lane >= o) v += t;
      }
      v += carry;
      if (y <= gy) G[(size_t)(y * gs + x) * gstride] = v;
      carry = __shfl_sync(FULLM, v, 31);
    }
  }
  __syncthreads();
  // ---- exclusive scan over the tiles: thread t owns `per` consecutive tiles
  const int per = (tiles + TP_THREADS - 1) / TP_THREADS;
  const int t0 = min(tiles, tid * per), t1 = min(tiles, t0 + per);
  auto count_of = [&](int t) -> uint32_t {
    return dead ? 0u : (uint32_t)G[(size_t)((t / gx) * gs + (t % gx)) * gstride];
  };
  uint32_t mine = 0;
  for (int t = t0; t < t1; t++) mine += count_of(t);
  uint32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(FULLM, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_wsum[warp] = incl;
  __syncthreads();
  uint32_t run = incl - mine;
  for (int w = 0; w < warp; w++) run += s_wsum[w];
  const uint32_t run0 = run;
  const uint32_t lanemask_lt = (1u << lane) - 1u;
  // every lane walks its `per` tiles in lock-step; equal length bins inside a warp (hundreds of empty tiles share bin
  // 0) are aggregated before the shared-memory atomic
  for (int j = 0; j < per; j++) {
    const int t = t0 + j;
    const bool valid = t < t1;
    const uint32_t c = valid ? count_of(t) : 0u;
    if (valid) ranges[t] = c ? make_uint2(run, run + c) : make_uint2(0u, 0u);
    run += c;
    const int bin = valid ? length_bin(c) : -1;
    const uint32_t m = __match_any_sync(FULLM, bin);
    if (valid && lane == __ffs(m) - 1) atomicAdd(&s_hist[bin], (uint32_t)__popc(m));
  }
  __syncthreads();
  // ---- descending exclusive prefix over the length bins (largest first)
  {
    const uint32_t v = tid < TP_BINS ? s_hist[TP_BINS - 1 - tid] : 0u;
    if (tid < TP_BINS && TP_BINS - 1 - tid >= length_bin(large_len) && v) atomicAdd(&s_nlarge, v);
    uint32_t in2 = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(FULLM, in2, o);
      if (lane >= o) in2 += t;
    }
    __syncthreads();  // s_wsum is reused
    if (lane == 31) s_wsum[warp] = in2;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < warp; w++) woff += s_wsum[w];
    if (tid < TP_BINS) s_base[TP_BINS - 1 - tid] = woff + in2 - v;
  }
  __syncthreads();
  run = run0;
  for (int j = 0; j < per; j++) {
    const int t = t0 + j;
    const bool valid = t < t1;
    const uint32_t c = valid ? count_of(t) : 0u;
    const int bin = valid ? length_bin(c) : -1;
    const uint32_t m = __match_any_sync(FULLM, bin);
    const int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (valid && lane == leader) base = atomicAdd(&s_base[bin], (uint32_t)__popc(m));
    base = __shfl_sync(FULLM, base, leader);
    if (valid) {
      const uint32_t p = base + (uint32_t)__popc(m & lanemask_lt);
      order[p] = c ? make_uint4((uint32_t)t, run, run + c, 0u) : make_uint4((uint32_t)t, 0u, 0u, 0u);
    }
    run += c;
  }
  if (tid < 8) counters[tid] = tid == 3 ? s_nlarge : 0u;
}

// ------------------------------------------------------------------------------------------------------------------
// scatter: emission order -> tile segments, stored as (depth bits << 32 | Gaussian id)
// ------------------------------------------------------------------------------------------------------------------
template <bool AGG>
__global__ void __launch_bounds__(SC_THREADS)
tile_scatter_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                    const skgs_raster_header* __restrict__ hdr, uint32_t R_cap, int tiles,
                    const uint2* __restrict__ ranges, uint32_t* __restrict__ cursors, uint64_t* __restrict__ pairs) {
  extern __shared__ uint32_t s_cnt[];  // AGG: [tiles] entries of this chunk per tile, then the chunk's base in the tile
  const int tid = threadIdx.x;
  pdl_wait();
  pdl_trigger();
  if (hdr->overflow) return;
  const uint32_t n = min(hdr->num_rendered, R_cap);
  const uint32_t chunks = (n + SC_CHUNK - 1) / SC_CHUNK;
  for (uint32_t c = blockIdx.x; c < chunks; c += gridDim.x) {
    const uint32_t base = c * SC_CHUNK;
    uint64_t key[SC_ITEMS];
    uint32_t val[SC_ITEMS];
#pragma unroll
    for (int i = 0; i < SC_ITEMS; i++) {
      const uint32_t s = base + i * SC_THREADS + tid;
      key[i] = s < n ? keys[s] : 0ull;
      val[i] = s < n ? vals[s] : 0u;
    }
    if (AGG) {
      uint32_t rank[SC_ITEMS];
      __syncthreads();  // the previous chunk's bases are no longer needed
      for (int t = tid; t < tiles; t += SC_THREADS) s_cnt[t] = 0;
      __syncthreads();
#pragma unroll
      for (int i = 0; i < SC_ITEMS; i++) {
        const uint32_t s = base + i * SC_THREADS + tid;
        rank[i] = s < n ? atomicAdd(&s_cnt[(uint32_t)(key[i] >> 32)], 1u) : 0u;
      }
      __syncthreads();
      // one reservation per (chunk, tile); four tiles per thread and round so that the atomics' round trips overlap
      for (int tb = tid; tb < tiles; tb += 4 * SC_THREADS) {
        uint32_t k[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int t = tb + u * SC_THREADS;
          k[u] = t < tiles ? s_cnt[t] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int t = tb + u * SC_THREADS;
          b[u] = k[u] ? __ldg(&ranges[t].x) + atomicAdd(&cursors[t], k[u]) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (k[u]) s_cnt[tb + u * SC_THREADS] = b[u];
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < SC_ITEMS; i++) {
        const uint32_t s = base + i * SC_THREADS + tid;
        if (s < n) pairs[s_cnt[(uint32_t)(key[i] >> 32)] + rank[i]] = (key[i] << 32) | (uint64_t)val[i];
      }
    } else {  // tile grids too large for the shared-memory counters: one atomic per entry
#pragma unroll
      for (int i = 0; i < SC_ITEMS; i++) {
        const uint32_t s = base + i * SC_THREADS + tid;
        if (s < n) {
          const uint32_t tile = (uint32_t)(key[i] >> 32);
          pairs[__ldg(&ranges[tile].x) + atomicAdd(&cursors[tile], 1u)] = (key[i] << 32) | (uint64_t)val[i];
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// per-tile sort
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ce(uint64_t* s, int i, int j) {
  const uint64_t a = s[i], b = s[j];
  if (a > b) {
    s[i] = b;
    s[j] = a;
  }
}

// barrier between two sub-steps of the network: in every sub-step of distance <= 32 a warp's 32 compare-exchanges cover
// one aligned block of 64 elements - the same block in all of them - so between two such sub-steps the data a warp
// reads was written by itself and a warp barrier is enough
__device__ __forceinline__ void step_sync(int done_distance, int next_distance) {
  if (done_distance > 32 || next_distance > 32)
    __syncthreads();
  else
    __syncwarp();
}

// ascending bitonic network in its direction-free form (first sub-step of every merge stage mirrored), n <= N = 2^k:
// the virtual elements n..N-1 are +inf and never move, so compare-exchanges whose upper partner is >= n are skipped.
// All THREADS threads of the CTA must call it.
template <int THREADS>
__device__ __forceinline__ void bitonic_sort_smem(uint64_t* s, int n) {
  int N = 2;
  while (N < n) N <<= 1;
  const int tid = threadIdx.x;
  const int half_n = N >> 1;
  for (int k = 2; k <= N; k <<= 1) {
    const int half = k >> 1;
    for (int p = tid; p < half_n; p += THREADS) {
      const int r = p & (half - 1);
      const int base = (p - r) << 1;  // (p / half) * k
      const int i = base + r, j = base + k - 1 - r;
      if (j < n) ce(s, i, j);
    }
    step_sync(half, half > 1 ? (half >> 1) : k);  // next: distance half/2, or the next stage's mirror (k)
    for (int d = half >> 1; d > 0; d >>= 1) {
      for (int p = tid; p < half_n; p += THREADS) {
        const int r = p & (d - 1);
        const int i = ((p - r) << 1) | r, j = i + d;
        if (j < n) ce(s, i, j);
      }
      step_sync(d, d > 1 ? (d >> 1) : k);
    }
  }
  __syncthreads();
}

// merge the sorted runs A = src[0, la) and B = src[la, la + lb) into dst[0, la + lb): every thread produces a contiguous
// slice of the output, its starting split found by a merge-path binary search (the words are unique)
template <int THREADS>
__device__ __forceinline__ void merge_runs(const uint64_t* __restrict__ src, uint64_t* __restrict__ dst, int la, int lb) {
  const int total = la + lb;
  const int per = (total + THREADS - 1) / THREADS;
  const int o0 = min(total, (int)threadIdx.x * per), o1 = min(total, o0 + per);
  if (o0 >= o1) return;
  const uint64_t* A = src;
  const uint64_t* B = src + la;
  int lo = max(0, o0 - lb), hi = min(o0, la);
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (A[mid] <= B[o0 - 1 - mid])
      lo = mid + 1;
    else
      hi = mid;
  }
  int ia = lo, ib = o0 - lo;
  for (int o = o0; o < o1; o++) {
    const bool take_a = ib >= lb || (ia < la && A[ia] <= B[ib]);
    dst[o] = take_a ? A[ia++] : B[ib++];
  }
}

template <int THREADS, int ITEMS>
struct RadixSmem {
  static constexpr int CAP = THREADS * ITEMS;
  static constexpr int WARPS = THREADS / 32;
  uint64_t w[CAP];                    // the segment: (depth << 32 | id)
  uint32_t whist[WARPS][RS_DIGITS];   // per-warp digit counts, then exclusive prefix over the warps
  uint32_t texcl[RS_DIGITS];          // exclusive prefix of the segment's digit counts
  uint32_t warp_tot[RS_DIGITS / 32];
  uint32_t item;
};

// stable LSD radix passes over bytes [byte_lo, byte_hi) of the 64-bit words of s.w[0, n), n <= CAP, in place
template <int THREADS, int ITEMS>
__device__ __forceinline__ void radix_passes_smem(RadixSmem<THREADS, ITEMS>& S, int n, int byte_lo, int byte_hi) {
  constexpr int WARPS = THREADS / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lanemask_lt = (1u << lane) - 1u;
  for (int b = byte_lo; b < byte_hi; b++) {
    const int shift = 8 * b;
    // ---- registers <- shared, warp-striped (item-major, then lane): the order the ranking below is stable in
    uint64_t word[ITEMS];
    const uint32_t first = (uint32_t)(S.w[0] >> shift) & 255u;
    bool same = true;
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
      const int idx = warp * (32 * ITEMS) + i * 32 + lane;
      word[i] = idx < n ? S.w[idx] : ~0ull;
      if (idx < n) same &= ((uint32_t)(word[i] >> shift) & 255u) == first;
    }
    for (int k = tid; k < WARPS * RS_DIGITS; k += THREADS) (&S.whist[0][0])[k] = 0;
    if (__syncthreads_and(same)) continue;  // this byte is the same for the whole segment: identity permutation
    uint16_t pos[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
      const int idx = warp * (32 * ITEMS) + i * 32 + lane;
      const bool valid = idx < n;
      const uint32_t d = valid ? ((uint32_t)(word[i] >> shift) & 255u) : 0xffffffffu;
      const uint32_t m = __match_any_sync(FULLM, d);
      const int leader = __ffs(m) - 1;
      uint32_t old = 0;
      if (valid && lane == leader) {
        old = S.whist[warp][d];
        S.whist[warp][d] = old + __popc(m);
      }
      old = __shfl_sync(FULLM, old, leader);
      pos[i] = (uint16_t)(old + __popc(m & lanemask_lt));
      __syncwarp();
    }
    __syncthreads();
    // ---- per digit: exclusive prefix over the warps, segment totals, exclusive scan over the digits
    uint32_t total = 0;
    if (tid < RS_DIGITS) {
#pragma unroll
      for (int w = 0; w < WARPS; w++) {
        const uint32_t c = S.whist[w][tid];
        S.whist[w][tid] = total;
        total += c;
      }
    }
    {
      uint32_t incl = total;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(FULLM, incl, o);
        if (lane >= o) incl += t;
      }
      if (tid < RS_DIGITS && lane == 31) S.warp_tot[warp] = incl;
      __syncthreads();
      if (tid < RS_DIGITS) {
        uint32_t woff = 0;
        for (int w = 0; w < warp; w++) woff += S.warp_tot[w];
        S.texcl[tid] = woff + incl - total;
      }
      __syncthreads();
    }
    // ---- every word of the segment is in registers: permute in place
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
      const int idx = warp * (32 * ITEMS) + i * 32 + lane;
      if (idx < n) {
        const uint32_t d = (uint32_t)(word[i] >> shift) & 255u;
        S.w[S.texcl[d] + S.whist[warp][d] + pos[i]] = word[i];
      }
    }
    __syncthreads();
  }
}

// LARGE = true : segments with len >= large_len (tickets [2], the first counters[3] entries of the order);
//                longer than CAP: sorted chunks + merge passes through global memory
// LARGE = false: the others, longest first, until the first empty tile
template <int THREADS, int ITEMS, bool LARGE>
__global__ void __launch_bounds__(THREADS, LARGE ? 1 : 2)
tile_sort_kernel(const uint4* __restrict__ order, int tiles, uint32_t* __restrict__ counters,
                 uint64_t* __restrict__ pairs, uint64_t* __restrict__ out_keys, uint32_t* __restrict__ out_vals) {
  using Smem = RadixSmem<THREADS, ITEMS>;
  constexpr int CAP = Smem::CAP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x;
  pdl_wait();
  pdl_trigger();
  const uint32_t n_large = counters[3];
  while (true) {
    __syncthreads();  // everyone is done with S of the previous item
    if (tid == 0) S.item = atomicAdd(&counters[LARGE ? 2 : 4], 1u) + (LARGE ? 0u : n_large);
    __syncthreads();
    const uint32_t item = S.item;
    if (LARGE ? item >= n_large : item >= (uint32_t)tiles) break;
    const uint4 d = order[item];
    const uint32_t tile = d.x, start = d.y;
    const int len = (int)(d.z - d.y);
    if (len == 0) break;  // the order is by decreasing length: nothing but empty tiles from here on
    const uint64_t hi = (uint64_t)tile << 32;
    if (len <= CAP) {
      for (int i = tid; i < len; i += THREADS) S.w[i] = pairs[start + i];
      __syncthreads();
      if (len <= BITONIC_MAX) {
        if (len > 1) bitonic_sort_smem<THREADS>(S.w, len);
      } else {
        radix_passes_smem<THREADS, ITEMS>(S, len, 4, 8);  // the depth bytes
        // Equal depths must come out in ascending Gaussian id, and the scatter order was arbitrary.  Exact float ties
        // are common in long lists (birthday paradox over 2^23 values per binade) but their runs are short: the thread
        // at the head of a run sorts it by insertion.  A run longer than TIE_RUN_MAX (coplanar Gaussians) sends the
        // whole segment through the id bytes instead.
        bool long_run = false;
        for (int i = tid; i + 1 < len; i += THREADS) {
          const uint32_t dep = (uint32_t)(S.w[i] >> 32);
          if ((uint32_t)(S.w[i + 1] >> 32) != dep || (i > 0 && (uint32_t)(S.w[i - 1] >> 32) == dep)) continue;
          int e = i + 2;  // head of a run [i, e)
          while (e < len && e - i <= TIE_RUN_MAX && (uint32_t)(S.w[e] >> 32) == dep) e++;
          if (e - i > TIE_RUN_MAX) {
            long_run = true;
            continue;
          }
          for (int a = i + 1; a < e; a++) {
            const uint64_t v = S.w[a];
            int b = a - 1;
            while (b >= i && S.w[b] > v) {
              S.w[b + 1] = S.w[b];
              b--;
            }
            S.w[b + 1] = v;
          }
        }
        if (__syncthreads_or(long_run)) {
          radix_passes_smem<THREADS, ITEMS>(S, len, 0, 4);  // id bytes first ...
          radix_passes_smem<THREADS, ITEMS>(S, len, 4, 8);  // ... then the (stable) depth passes again
        }
      }
      for (int i = tid; i < len; i += THREADS) {
        const uint64_t v = S.w[i];
        out_keys[start + i] = hi | (v >> 32);
        out_vals[start + i] = (uint32_t)v;
      }
    } else if (LARGE) {
      // longer than shared memory: sorted chunks (full 8-byte radix: unique words, no tie check), then merge passes
      // ping-ponging between the segment of `pairs` and the same segment of out_keys (free until the final conversion)
      for (int c0 = 0; c0 < len; c0 += CAP) {
        const int m = min(CAP, len - c0);
        __syncthreads();
        for (int i = tid; i < m; i += THREADS) S.w[i] = pairs[start + c0 + i];
        __syncthreads();
        radix_passes_smem<THREADS, ITEMS>(S, m, 0, 8);
        for (int i = tid; i < m; i += THREADS) pairs[start + c0 + i] = S.w[i];
      }
      __syncthreads();
      uint64_t* src = pairs + start;
      uint64_t* dst = out_keys + start;
      for (int width = CAP; width < len; width <<= 1) {
        for (int a = 0; a < len; a += 2 * width) {
          const int la = min(width, len - a), lb = min(width, len - a - la);
          if (lb > 0) {
            merge_runs<THREADS>(src + a, dst + a, la, lb);
          } else {
            for (int i = tid; i < la; i += THREADS) dst[a + i] = src[a + i];
          }
        }
        __syncthreads();
        uint64_t* t = src; src = dst; dst = t;
      }
      // `src` holds the sorted words; out_keys may be that very buffer: element-wise, in place
      for (int i = tid; i < len; i += THREADS) {
        const uint64_t v = src[i];
        out_keys[start + i] = hi | (v >> 32);
        out_vals[start + i] = (uint32_t)v;
      }
    }
  }
}

constexpr int RS_REG_THREADS = 512, RS_REG_ITEMS = 12;     // regular: segments below 6144 entries
constexpr int RS_LARGE_THREADS = 1024, RS_LARGE_ITEMS = 16;  // large: up to 16384 entries in shared memory
constexpr uint32_t LARGE_LEN = RS_REG_THREADS * RS_REG_ITEMS;  // 6144 = 1.5 * 2^12: a length-bin boundary

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace

// plan + scatter + per-tile sort; the keys / values were emitted into buffer a, the tile rectangles into the grid
int launch_tile_binning(const RasterParams& rp, char* geom, char* binning, char* img, const skgs_raster_layout& lay,
                        int64_t R_cap, int64_t R_hint, cudaStream_t st) {
  auto* hdr = reinterpret_cast<skgs_raster_header*>(geom + lay.header);
  const int tiles = rp.gx * rp.gy;
  auto* keys_a = reinterpret_cast<uint64_t*>(binning + lay.keys);
  auto* vals_a = reinterpret_cast<uint32_t*>(binning + lay.vals);
  auto* keys_b = reinterpret_cast<uint64_t*>(binning + lay.tile_pairs);
  auto* grid_cells = reinterpret_cast<int*>(binning + lay.tile_grid);
  auto* cursors = reinterpret_cast<uint32_t*>(binning + lay.tile_cursors);
  auto* ranges = reinterpret_cast<uint2*>(img + lay.ranges);
  auto* order = reinterpret_cast<uint4*>(img + lay.tile_order);
  auto* counters = reinterpret_cast<uint32_t*>(img + lay.work_counters);
  {
    const size_t cells = (size_t)(rp.gx + 1) * (rp.gy + 1);
    const size_t smem = cells <= (size_t)TP_SMEM_CELLS ? cells * sizeof(int) : 0;
    ProfScope prof_("tile_plan_kernel", st);
    SKGS_CUDA(launch_pdl(tile_plan_kernel, dim3(1), dim3(TP_THREADS), smem, st, grid_cells, rp.gx, rp.gy,
                         (const skgs_raster_header*)hdr, ranges, order, counters, LARGE_LEN, tile_cell_stride()));
    SKGS_CHECK_LAUNCH("tile_plan_kernel");
  }
  {
    const int64_t hint = R_hint > 0 ? (R_hint < R_cap ? R_hint : R_cap) : R_cap;
    int grid = (int)((hint + SC_CHUNK - 1) / SC_CHUNK);
    const int cap = sm_count() * 4;
    grid = grid < 1 ? 1 : (grid > cap ? cap : grid);
    ProfScope prof_("tile_scatter_kernel", st);
    if (tiles <= SC_MAX_TILES) {
      static bool attr_set = false;
      if (!attr_set) {
        SKGS_CUDA(cudaFuncSetAttribute(tile_scatter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(SC_MAX_TILES * sizeof(uint32_t))));
        attr_set = true;
      }
      SKGS_CUDA(launch_pdl(tile_scatter_kernel<true>, dim3(grid), dim3(SC_THREADS), (size_t)tiles * sizeof(uint32_t),
                           st, (const uint64_t*)keys_a, (const uint32_t*)vals_a, (const skgs_raster_header*)hdr,
                           (uint32_t)R_cap, tiles, (const uint2*)ranges, cursors, keys_b));
    } else {
      SKGS_CUDA(launch_pdl(tile_scatter_kernel<false>, dim3(grid), dim3(SC_THREADS), 0, st, (const uint64_t*)keys_a,
                           (const uint32_t*)vals_a, (const skgs_raster_header*)hdr, (uint32_t)R_cap, tiles,
                           (const uint2*)ranges, cursors, keys_b));
    }
    SKGS_CHECK_LAUNCH("tile_scatter_kernel");
  }
  {
    auto large = tile_sort_kernel<RS_LARGE_THREADS, RS_LARGE_ITEMS, true>;
    const size_t smem = sizeof(RadixSmem<RS_LARGE_THREADS, RS_LARGE_ITEMS>);
    static bool attr_set = false;
    if (!attr_set) {
      SKGS_CUDA(cudaFuncSetAttribute(large, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set = true;
    }
    ProfScope prof_("tile_sort_large_kernel", st);
    SKGS_CUDA(launch_pdl(large, dim3(sm_count()), dim3(RS_LARGE_THREADS), smem, st, (const uint4*)order, tiles,
                         counters, keys_b, keys_a, vals_a));
    SKGS_CHECK_LAUNCH("tile_sort_large_kernel");
  }
  {
    auto regular = tile_sort_kernel<RS_REG_THREADS, RS_REG_ITEMS, false>;
    const size_t smem = sizeof(RadixSmem<RS_REG_THREADS, RS_REG_ITEMS>);
    static bool attr_set = false;
    static int per_sm = 0;
    if (!attr_set) {
      SKGS_CUDA(cudaFuncSetAttribute(regular, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, regular, RS_REG_THREADS, smem);
      if (per_sm < 1) per_sm = 1;
      attr_set = true;
    }
    int grid = sm_count() * per_sm;
    grid = grid > tiles ? tiles : grid;
    ProfScope prof_("tile_sort_kernel", st);
    SKGS_CUDA(launch_pdl(regular, dim3(grid), dim3(RS_REG_THREADS), smem, st, (const uint4*)order, tiles, counters,
                         keys_b, keys_a, vals_a));
    SKGS_CHECK_LAUNCH("tile_sort_kernel");
  }
  return SKGS_OK;
}

}  // namespace skgs
