// onesweep_radix.cu - NOT part of the product build.  The device-wide radix sort (8-bit onesweep passes with a
// warp-parallel decoupled look-back, device-side skipping of constant digits, tile ranges fused into the last pass) that
// binned the keys until round 2: 5 executed passes of ~21 us at R = 0.83 M keys (c2).  Replaced by the tile-segmented
// sort of sk_gs_b200/csrc/tile_sort.cu (one scatter pass + per-tile sorts in shared memory; A/B in
// profiles/r2_binning_ab.txt).  Kept for the record: the ranking loop lives on in tile_sort.cu (radix_passes_smem).
// It was compiled inside raster_fwd.cu (helpers: ld_volatile_*, st_volatile_u32, FULL, skgs_raster_header).
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

