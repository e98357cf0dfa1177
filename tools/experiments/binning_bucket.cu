// binning_bucket.cu - tile-bucketed binning: a two-level radix sort on the (tile | depth) keys that exploits what the
// rasterizer already knows.
//
//   most-significant "digit" = the whole tile id.  Its histogram is counted while preprocessing (one atomic per
//   (Gaussian, tile) pair), one small kernel prefix-sums it - which IS the tile range table the compositing kernels
//   need - and the scatter pass drops every (depth bits, id) pair into its tile's bucket;
//   least-significant part = the 32 depth bits (+ the Gaussian id as tie break), sorted per tile in shared memory.
//
// The result is bit-identical to the reference's stable cub::DeviceRadixSort on 64-bit (tile<<32 | depth) keys
// (my_ext/_C/src/nerf/gaussian_rasterizer_forward.cu:219-241): within a tile the stable sort leaves equal depths in
// emission order, which is ascending Gaussian id because every Gaussian emits each tile at most once - exactly the
// order of the composite key (depth << 32 | id).  Bytes per list entry: 8 (scatter) + 8 + 4 (sort in, id out) instead
// of 12 + 6 x 24 for the six global radix passes, and 3 launches instead of 8.  The onesweep path (raster_fwd.cu)
// is the default; this variant is selected with settings.debug bit 3 and tested against the same oracle.  Measured at
// R = 0.83 M (c2): 370 us vs 154 us for duplicate + onesweep - the per-tile atomics serialise on the dense tiles (a
// single address receives thousands of increments) and the bitonic network is too slow for 5k-entry tiles; it needs
// sub-bucketed counters and a shared-memory radix sort before it can win (profiles/r1_binning.md).
#include "common.cuh"

namespace skgs {

constexpr int TS_THREADS = 1024;

// exclusive scan of the tile histogram -> ranges (start, end), bucket cursors; overflow check against the arena
__global__ void __launch_bounds__(TS_THREADS)
tile_scan_kernel(const uint32_t* __restrict__ tile_count, int tiles, uint32_t R_cap, uint2* __restrict__ ranges,
                 uint32_t* __restrict__ cursor, skgs_raster_header* __restrict__ hdr) {
  __shared__ uint32_t s_warp[TS_THREADS / 32];
  __shared__ uint32_t s_carry;
  __shared__ uint32_t s_total;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // pass 1: total
  uint32_t sum = 0;
  for (int t = tid; t < tiles; t += TS_THREADS) sum += tile_count[t];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) s_warp[warp] = sum;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  if (tid == 0) {
    uint32_t tot = 0;
    for (int w = 0; w < TS_THREADS / 32; w++) tot += s_warp[w];
    s_total = tot;
    if (tot > R_cap) hdr->overflow = 1;  // arena too small: every range stays empty, the caller re-runs
  }
  __syncthreads();
  const bool ok = s_total <= R_cap;
  // pass 2: scan in chunks of TS_THREADS
  for (int base = 0; base < tiles; base += TS_THREADS) {
    const int t = base + tid;
    const uint32_t c = (t < tiles) ? tile_count[t] : 0u;
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    __syncthreads();  // s_warp / s_carry of the previous chunk fully consumed
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < warp; w++) woff += s_warp[w];
    const uint32_t excl = s_carry + woff + incl - c;
    if (t < tiles) {
      ranges[t] = (ok && c > 0) ? make_uint2(excl, excl + c) : make_uint2(0u, 0u);  // empty tiles are (0,0) like the reference
      cursor[t] = excl;
    }
    __syncthreads();
    if (tid == TS_THREADS - 1) s_carry = excl + c;
  }
}

constexpr int BS_THREADS = 256;

// every (Gaussian, tile) pair -> (depth bits << 32 | id) into the tile's bucket
__global__ void __launch_bounds__(BS_THREADS)
bucket_scatter_kernel(int P, int gx, int gy, const int32_t* __restrict__ radii, const float2* __restrict__ means2D,
                      const float* __restrict__ depths, uint32_t* __restrict__ cursor, uint64_t* __restrict__ bucket,
                      const skgs_raster_header* __restrict__ hdr) {
  if (hdr->overflow) return;
  const int tid = threadIdx.x, lane = tid & 31;
  const int i = blockIdx.x * BS_THREADS + tid;
  int x0 = 0, y0 = 0, x1 = 0, y1 = 0, cnt = 0;
  uint32_t dbits = 0;
  if (i < P) {
    const int rad = radii[i];
    if (rad > 0) {
      const float2 p = means2D[i];
      get_rect(p.x, p.y, rad, gx, gy, x0, y0, x1, y1);
      cnt = (x1 - x0) * (y1 - y0);
      dbits = __float_as_uint(depths[i]);
    }
  }
  const uint64_t ent = ((uint64_t)dbits << 32) | (uint32_t)i;
  const int w = x1 - x0;
  if (cnt > 0 && cnt <= 32)
    for (int k = 0; k < cnt; k++) bucket[atomicAdd(&cursor[(y0 + k / w) * gx + x0 + k % w], 1u)] = ent;
  uint32_t big = __ballot_sync(0xffffffffu, cnt > 32);
  while (big) {
    const int src = __ffs(big) - 1;
    big &= big - 1;
    const int bx0 = __shfl_sync(0xffffffffu, x0, src), by0 = __shfl_sync(0xffffffffu, y0, src);
    const int bw = __shfl_sync(0xffffffffu, w, src), bc = __shfl_sync(0xffffffffu, cnt, src);
    const uint64_t be = __shfl_sync(0xffffffffu, ent, src);
    for (int k = lane; k < bc; k += 32) bucket[atomicAdd(&cursor[(by0 + k / bw) * gx + bx0 + k % bw], 1u)] = be;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// per-tile sort.  "Normalised" bitonic network (every compare-exchange puts the smaller key at the lower index), so an
// arbitrary length n works with virtual +inf padding: pairs whose upper index is >= n are simply skipped.
// ------------------------------------------------------------------------------------------------------------------
constexpr int SORT_THREADS = 256;
constexpr int SORT_CHUNK = 8192;  // keys per shared-memory chunk (64 KB)

__device__ __forceinline__ void cmpex(uint64_t* a, int lo, int hi) {
  const uint64_t x = a[lo], y = a[hi];
  if (x > y) {
    a[lo] = y;
    a[hi] = x;
  }
}

// all stages with block size kmin..kmax on a[0..m) (shared or global), m <= capacity of `a`
__device__ void bitonic_range(uint64_t* a, int m, int kmin, int kmax) {
  for (int k = kmin; k <= kmax; k <<= 1) {
    const int half = k >> 1;
    for (int i = threadIdx.x; i < ((m + k - 1) / k) * half; i += SORT_THREADS) {  // mirror step
      const int blk = i / half, off = i % half;
      const int lo = blk * k + off, hi = blk * k + k - 1 - off;
      if (hi < m) cmpex(a, lo, hi);
    }
    __syncthreads();
    for (int j = half >> 1; j >= 1; j >>= 1) {
      for (int i = threadIdx.x; i < (m + 1) / 2 + j; i += SORT_THREADS) {
        const int lo = (i / j) * 2 * j + (i % j), hi = lo + j;
        if (hi < m) cmpex(a, lo, hi);
      }
      __syncthreads();
    }
  }
}

// half-cleaner strides jmax..1 only (used after a global merge step), on a[0..m)
__device__ void bitonic_clean(uint64_t* a, int m, int jmax) {
  for (int j = jmax; j >= 1; j >>= 1) {
    for (int i = threadIdx.x; i < (m + 1) / 2 + j; i += SORT_THREADS) {
      const int lo = (i / j) * 2 * j + (i % j), hi = lo + j;
      if (hi < m) cmpex(a, lo, hi);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(SORT_THREADS)
tile_sort_kernel(int tiles, const uint32_t* __restrict__ order, const uint2* __restrict__ ranges,
                 uint64_t* __restrict__ bucket, uint32_t* __restrict__ point_list, uint64_t* __restrict__ keys_out,
                 const skgs_raster_header* __restrict__ hdr) {
  extern __shared__ __align__(16) unsigned char sort_smem[];
  uint64_t* sm = reinterpret_cast<uint64_t*>(sort_smem);
  if (hdr->overflow) return;
  const int tile = (int)order[blockIdx.x];
  const uint2 r = ranges[tile];
  const int n = (int)(r.y - r.x);
  if (n == 0) return;
  uint64_t* g = bucket + r.x;
  int npow2 = 1;
  while (npow2 < n) npow2 <<= 1;
  if (n <= SORT_CHUNK) {
    for (int i = threadIdx.x; i < n; i += SORT_THREADS) sm[i] = g[i];
    __syncthreads();
    if (n > 1) bitonic_range(sm, n, 2, npow2);
    for (int i = threadIdx.x; i < n; i += SORT_THREADS) {
      const uint64_t e = sm[i];
      point_list[r.x + i] = (uint32_t)e;
      if (keys_out) keys_out[r.x + i] = ((uint64_t)tile << 32) | (e >> 32);
    }
    return;
  }
  // ---- out-of-core: sort every chunk in shared memory, then merge with global steps for the large strides
  for (int base = 0; base < n; base += SORT_CHUNK) {
    const int m = min(SORT_CHUNK, n - base);
    for (int i = threadIdx.x; i < m; i += SORT_THREADS) sm[i] = g[base + i];
    __syncthreads();
    bitonic_range(sm, m, 2, SORT_CHUNK);
    for (int i = threadIdx.x; i < m; i += SORT_THREADS) g[base + i] = sm[i];
    __syncthreads();
  }
  for (int k = 2 * SORT_CHUNK; k <= npow2; k <<= 1) {
    const int half = k >> 1;
    for (int i = threadIdx.x; i < ((n + k - 1) / k) * half; i += SORT_THREADS) {  // mirror step in global memory
      const int blk = i / half, off = i % half;
      const int lo = blk * k + off, hi = blk * k + k - 1 - off;
      if (hi < n) cmpex(g, lo, hi);
    }
    __syncthreads();
    for (int j = half >> 1; j >= SORT_CHUNK; j >>= 1) {  // strides that span chunks
      for (int i = threadIdx.x; i < (n + 1) / 2 + j; i += SORT_THREADS) {
        const int lo = (i / j) * 2 * j + (i % j), hi = lo + j;
        if (hi < n) cmpex(g, lo, hi);
      }
      __syncthreads();
    }
    for (int base = 0; base < n; base += SORT_CHUNK) {  // remaining strides stay inside a chunk
      const int m = min(SORT_CHUNK, n - base);
      for (int i = threadIdx.x; i < m; i += SORT_THREADS) sm[i] = g[base + i];
      __syncthreads();
      bitonic_clean(sm, m, SORT_CHUNK >> 1);
      for (int i = threadIdx.x; i < m; i += SORT_THREADS) g[base + i] = sm[i];
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < n; i += SORT_THREADS) {
    const uint64_t e = g[i];
    point_list[r.x + i] = (uint32_t)e;
    if (keys_out) keys_out[r.x + i] = ((uint64_t)tile << 32) | (e >> 32);
  }
}

int launch_binning_bucket(const RasterParams& rp, char* geom, char* binning, char* img, const skgs_raster_layout& lay,
                          const skgs_raster_layout& phys, const int32_t* radii, int64_t R_cap, bool write_keys,
                          uint32_t* num_rendered_host, cudaStream_t st) {
  auto* hdr = reinterpret_cast<skgs_raster_header*>(geom + lay.header);
  const int tiles = rp.gx * rp.gy;
  uint2* ranges = reinterpret_cast<uint2*>(img + lay.ranges);
  uint32_t* tile_count = reinterpret_cast<uint32_t*>(geom + lay.tile_count);
  uint32_t* cursor = reinterpret_cast<uint32_t*>(geom + lay.tile_cursor);
  if (rp.P == 0 || R_cap <= 0) {
    SKGS_CUDA(cudaMemsetAsync(ranges, 0, (size_t)tiles * sizeof(uint2), st));
    int rc = launch_tile_order(rp, img, lay, st);
    return rc;
  }
  {
    ProfScope prof_("tile_scan_kernel", st);
    tile_scan_kernel<<<1, TS_THREADS, 0, st>>>(tile_count, tiles, (uint32_t)R_cap, ranges, cursor, hdr);
    SKGS_CHECK_LAUNCH("tile_scan_kernel");
  }
  if (num_rendered_host)
    SKGS_CUDA(cudaMemcpyAsync(num_rendered_host, hdr, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  // the bucket lives in the physical "B" key buffer, the final lists where the layout says (see api.cu)
  uint64_t* bucket = reinterpret_cast<uint64_t*>(binning + (lay.keys_sorted == phys.keys_unsorted ? phys.keys_sorted
                                                                                                   : phys.keys_unsorted));
  {
    ProfScope prof_("bucket_scatter_kernel", st);
    bucket_scatter_kernel<<<(rp.P + BS_THREADS - 1) / BS_THREADS, BS_THREADS, 0, st>>>(
        rp.P, rp.gx, rp.gy, radii, reinterpret_cast<const float2*>(geom + lay.means2D),
        reinterpret_cast<const float*>(geom + lay.depths), cursor, bucket, hdr);
    SKGS_CHECK_LAUNCH("bucket_scatter_kernel");
  }
  int rc = launch_tile_order(rp, img, lay, st);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    SKGS_CUDA(cudaFuncSetAttribute(tile_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   SORT_CHUNK * (int)sizeof(uint64_t)));
    attr_set = true;
  }
  {
    ProfScope prof_("tile_sort_kernel", st);
    tile_sort_kernel<<<tiles, SORT_THREADS, SORT_CHUNK * sizeof(uint64_t), st>>>(
        tiles, reinterpret_cast<const uint32_t*>(img + lay.tile_order), ranges, bucket,
        reinterpret_cast<uint32_t*>(binning + lay.point_list),
        write_keys ? reinterpret_cast<uint64_t*>(binning + lay.keys_sorted) : nullptr, hdr);
    SKGS_CHECK_LAUNCH("tile_sort_kernel");
  }
  return SKGS_OK;
}

}  // namespace skgs
