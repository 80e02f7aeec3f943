// sort.cu -- hand-written onesweep LSD radix sort of (u64 key, u32 value) pairs + tile ranges.
// Replaces cub::DeviceRadixSort::SortPairs and identifyTileRanges.  No CUB.
//
//   digit_histogram_kernel : one read of the keys, all passes' 256-bin histograms at once
//                            (per-warp private shared-memory histograms, warp-aggregated by digit match)
//   digit_scan_kernel      : exclusive scan of each histogram -> global digit bases
//   onesweep_pass_kernel   : per pass, one read + one write of the pairs: stable in-tile ranking,
//                            single-pass chained scan (decoupled look-back) of the per-tile digit
//                            counts, reorder through shared memory, coalesced scatter
//   tile_ranges_kernel     : (start,end) of every (view,tile) bin in the sorted list
//
// All kernels read the pair count from device memory (B200sStatus.num_pairs) so that no host
// synchronisation is needed between duplication and sort; grids are sized for the capacity and
// surplus blocks exit.  HBM-bound: 8 B/key (histogram) + 24 B/pair/pass.
#include "kernels.cuh"

namespace b200s {

constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 12;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;  // 3072 pairs per tile: 64 registers and 44 KB -> four CTAs per SM
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int RADIX = 256;
// look-back words are 64-bit (2 flag bits + count) so that one call can sort up to 2^32 - 1 pairs
constexpr uint64_t LB_FLAG_AGG = 1ull << 62, LB_FLAG_PREFIX = 2ull << 62, LB_VALUE_MASK = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long resolve_count(const CountRef& c) {
  if (c.overflow_dev && *c.overflow_dev) return 0ull;
  return c.n_dev ? *c.n_dev : c.n_host;
}

// lanes of the warp holding the same 8-bit digit (all 32 lanes must call).
// MODE 0: eight VOTE.BALLOTs + logic (ALU pipe); MODE 1: one MATCH.ANY (ADU pipe, ~20 slots each).
// The two pipes are independent, so the ranking loop alternates them (tuning knob).
template <int MODE>
__device__ __forceinline__ uint32_t match_digit(uint32_t d) {
  if (MODE == 0) {
    uint32_t peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 8; b++) {
      const bool bit = (d >> b) & 1u;
      const uint32_t m = __ballot_sync(0xffffffffu, bit);
      peers &= bit ? m : ~m;
    }
    return peers;
  } else {
    return __match_any_sync(0xffffffffu, d);
  }
}

// run-time tuning knobs (b200s_debug_set): [0] histogram variant, [1] ranking variant
std::atomic<int> g_sort_knobs[4] = {{2}, {2}, {0}, {0}};

// ---------------------------------------------------------------------------------------------
// HMODE 0/1: warp-aggregated by digit match (ballots / MATCH.ANY), plain read-modify-write by the group
// leader on a per-warp private histogram; HMODE 2: one shared-memory atomic per lane.
template <int HMODE>
__global__ void __launch_bounds__(SORT_THREADS) digit_histogram_kernel(const uint64_t* __restrict__ keys, CountRef cnt, uint32_t* __restrict__ hist,
                                                                       int passes) {
  extern __shared__ uint32_t s_hist[];  // [SORT_WARPS][passes][256]
  const unsigned long long n = resolve_count(cnt);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < SORT_WARPS * passes * RADIX; i += SORT_THREADS) s_hist[i] = 0;
  __syncthreads();
  uint32_t* my = s_hist + warp * passes * RADIX;
  const unsigned long long stride = (unsigned long long)gridDim.x * SORT_THREADS;
  // every warp iterates the same number of times so that the warp-wide match is always convergent
  for (unsigned long long base = (unsigned long long)blockIdx.x * SORT_THREADS + warp * 32; base < n; base += stride) {
    const unsigned long long idx = base + lane;
    const bool valid = idx < n;
    const uint64_t k = valid ? __ldg(keys + idx) : 0ull;
    const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
    for (int p = 0; p < passes; p++) {
      const uint32_t d = (uint32_t)(k >> (8 * p)) & 255u;
      const uint32_t d0 = __shfl_sync(0xffffffffu, d, 0);
      if (__all_sync(0xffffffffu, d == d0)) {
        if (lane == 0) my[p * RADIX + d0] += __popc(vmask);
      } else if (HMODE == 2) {
        if (valid) atomicAdd(&my[p * RADIX + d], 1u);
      } else {
        const uint32_t peers = match_digit<HMODE>(d) & vmask;
        if (valid && lane == (__ffs(peers) - 1)) my[p * RADIX + d] += __popc(peers);
      }
      __syncwarp();
    }
  }
  __syncthreads();
  for (int p = 0; p < passes; p++) {
    uint32_t s = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++) s += s_hist[(w * passes + p) * RADIX + tid];
    if (s) atomicAdd(&hist[p * RADIX + tid], s);
  }
}

// one block; exclusive scan of each pass's histogram in place
__global__ void __launch_bounds__(RADIX) digit_scan_kernel(uint32_t* hist, int passes) {
  __shared__ uint32_t s_w[RADIX / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int p = 0; p < passes; p++) {
    const uint32_t v = hist[p * RADIX + tid];
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    uint32_t off = 0;
    for (int w = 0; w < warp; w++) off += s_w[w];
    hist[p * RADIX + tid] = off + incl - v;
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
struct PassArgs {
  const uint64_t* keys_in; const uint32_t* vals_in;
  uint64_t* keys_out; uint32_t* vals_out;
  const uint32_t* gbase;   // [256] exclusive digit bases of this pass
  uint64_t* lb_cur;        // [tiles][256] look-back words of this pass (zeroed)
  uint64_t* lb_next;       // same for the next pass: this pass zeroes the rows it owns
  uint32_t* tile_counter;  // dynamic tile id
  CountRef cnt;
  int shift;
};

// 16-byte async global->shared copy; bytes past src_bytes are zero-filled and not read
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src, int src_bytes) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gmem_src, int src_bytes) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

constexpr int LOOKBACK_BATCH = 8;
// dynamic shared memory of the pass kernel
constexpr size_t SWEEP_SMEM = SORT_TILE * 8 + SORT_TILE * 4 + SORT_WARPS * RADIX * 4;

template <int RMODE>
__global__ void __launch_bounds__(SORT_THREADS, 4) onesweep_pass_kernel(const PassArgs a) {
  extern __shared__ __align__(16) unsigned char sweep_smem[];
  uint64_t* s_keys = reinterpret_cast<uint64_t*>(sweep_smem);                              // [SORT_TILE] keys in sorted order
  uint32_t* s_vin = reinterpret_cast<uint32_t*>(sweep_smem + SORT_TILE * 8);               // values as they arrive, then (same
  uint32_t* s_vout = s_vin;                                                                // storage) in sorted order
  uint32_t(*s_wc)[RADIX] = reinterpret_cast<uint32_t(*)[RADIX]>(s_vin + SORT_TILE);        // [8][256]
  __shared__ uint32_t s_goff[RADIX];
  __shared__ uint32_t s_dstart[RADIX];
  __shared__ uint32_t s_scan[SORT_WARPS];
  __shared__ uint32_t s_tile;

  const unsigned long long n = resolve_count(a.cnt);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(a.tile_counter, 1u);
#pragma unroll
  for (int w = 0; w < SORT_WARPS; w++) s_wc[w][tid] = 0;
  __syncthreads();
  const uint32_t tile = s_tile;
  const unsigned long long base = (unsigned long long)tile * SORT_TILE;
  if (base >= n) return;
  const uint32_t nvalid = (uint32_t)min((unsigned long long)SORT_TILE, n - base);
  a.lb_next[(size_t)tile * RADIX + tid] = 0;

  // ---- the tile's values start travelling to shared memory now; they are needed only at the very end
  {
    const uint32_t* vsrc = a.vals_in + base;
    if ((reinterpret_cast<uintptr_t>(vsrc) & 15) == 0) {
#pragma unroll
      for (int c = tid; c < SORT_TILE / 4; c += SORT_THREADS) {
        const int rem = (int)nvalid - c * 4;
        const int bytes = rem >= 4 ? 16 : (rem > 0 ? rem * 4 : 0);
        cp_async_16(s_vin + c * 4, vsrc + (bytes ? c * 4 : 0), bytes);
      }
    } else {
      for (int e = tid; e < SORT_TILE; e += SORT_THREADS) cp_async_4(s_vin + e, vsrc + (e < (int)nvalid ? e : 0), e < (int)nvalid ? 4 : 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

  // ---- load keys, warp-striped: item i of lane l of warp w is element w*512 + i*32 + l of the tile
  uint64_t key[SORT_ITEMS];
  const uint32_t wbase = warp * (32 * SORT_ITEMS) + lane;
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++) {
    const uint32_t e = wbase + i * 32;
    key[i] = e < nvalid ? __ldg(a.keys_in + base + e) : ~0ull;
  }
  // ---- stable rank inside the warp, item by item
  uint32_t rank[SORT_ITEMS];
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t* wc = s_wc[warp];
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++) {
    const uint32_t d = (uint32_t)(key[i] >> a.shift) & 255u;
    uint32_t peers;
    peers = RMODE == 2 ? ((i & 1) ? match_digit<1>(d) : match_digit<0>(d)) : match_digit<(RMODE == 1 ? 1 : 0)>(d);
    // the group's first lane bumps the warp's counter with one shared atomic and hands the old value to its group:
    // no read / sync / write / sync round trip per item, and the atomics of successive items pipeline (same-address
    // atomics of one warp execute in program order, so the ranking stays stable)
    const int leader = __ffs(peers) - 1;
    uint32_t r = 0;
    if (lane == leader) r = atomicAdd(&wc[d], (uint32_t)__popc(peers));
    r = __shfl_sync(0xffffffffu, r, leader);
    rank[i] = r + __popc(peers & lt);
  }
  __syncthreads();
  // ---- thread = digit: exclusive scan over warps, tile count, chained scan across tiles
  uint32_t run = 0;
#pragma unroll
  for (int w = 0; w < SORT_WARPS; w++) {
    const uint32_t t = s_wc[w][tid];
    s_wc[w][tid] = run;
    run += t;
  }
  uint32_t prev = 0;
  {
    uint64_t* row = a.lb_cur + (size_t)tile * RADIX + tid;
    if (tile == 0) {
      st_volatile_u64(row, LB_FLAG_PREFIX | run);
    } else {
      st_volatile_u64(row, LB_FLAG_AGG | run);
      // decoupled look-back, LOOKBACK_BATCH predecessors per round trip
      int t = (int)tile - 1;
      bool found = false;
      while (!found) {
        uint64_t w[LOOKBACK_BATCH];
#pragma unroll
        for (int k = 0; k < LOOKBACK_BATCH; k++)
          w[k] = (t - k >= 0) ? ld_volatile_u64(a.lb_cur + (size_t)(t - k) * RADIX + tid) : LB_FLAG_PREFIX;
        int adv = 0;
#pragma unroll
        for (int k = 0; k < LOOKBACK_BATCH; k++) {
          const uint32_t f = (uint32_t)(w[k] >> 62);
          if (!found && adv == k && f != 0) {
            prev += (uint32_t)(w[k] & LB_VALUE_MASK);  // positions are taken modulo 2^32 (< 2^32 pairs per call)
            adv = k + 1;
            if (f == 2) found = true;
          }
        }
        t -= adv;
      }
      st_volatile_u64(row, LB_FLAG_PREFIX | (uint64_t)(uint32_t)(prev + run));
    }
  }
  // ---- exclusive scan over digits of the tile counts -> start of each digit inside the tile
  uint32_t incl = run;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
  if (lane == 31) s_scan[warp] = incl;
  __syncthreads();
  uint32_t woff = 0;
#pragma unroll
  for (int w = 0; w < SORT_WARPS; w++) if (w < warp) woff += s_scan[w];
  const uint32_t dstart = woff + incl - run;
  s_dstart[tid] = dstart;
  s_goff[tid] = a.gbase[tid] + prev - dstart;
  __syncthreads();
  // ---- reorder keys through shared memory (rank[] becomes the in-tile sorted position)
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++) {
    const uint32_t d = (uint32_t)(key[i] >> a.shift) & 255u;
    rank[i] += s_dstart[d] + wc[d];
    s_keys[rank[i]] = key[i];
  }
  // values have landed by now; wait for this thread's copies, the barrier below makes all of them visible
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // the sorted values take the place of the arrived ones: read this thread's, barrier, write by rank
  uint32_t val[SORT_ITEMS];
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++) val[i] = s_vin[wbase + i * 32];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++) {
    const uint32_t e = wbase + i * 32;
    if (e < nvalid) s_vout[rank[i]] = val[i];
  }
  __syncthreads();
  // one walk over the sorted tile: the key is read once and gives the destination of both words
#pragma unroll
  for (int j = 0; j < SORT_ITEMS; j++) {
    const uint32_t e = j * SORT_THREADS + tid;
    if (e < nvalid) {
      const uint64_t k = s_keys[e];
      const uint32_t dst = s_goff[(uint32_t)(k >> a.shift) & 255u] + e;
      a.keys_out[dst] = k;
      a.vals_out[dst] = s_vout[e];
    }
  }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tile_ranges_kernel(const uint64_t* __restrict__ keys, CountRef cnt, uint2* __restrict__ ranges) {
  const unsigned long long n = resolve_count(cnt);
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += stride) {
    const uint32_t cur = (uint32_t)(__ldg(keys + idx) >> 32);
    if (idx == 0) ranges[cur].x = 0;
    else {
      const uint32_t prv = (uint32_t)(__ldg(keys + idx - 1) >> 32);
      if (cur != prv) { ranges[prv].y = (uint32_t)idx; ranges[cur].x = (uint32_t)idx; }
    }
    if (idx == n - 1) ranges[cur].y = (uint32_t)n;
  }
}

// ---------------------------------------------------------------------------------------------
// host side
size_t sort_tmp_bytes(long long n_cap) {
  const size_t tiles = (size_t)((n_cap + SORT_TILE - 1) / SORT_TILE) + 1;
  return 8 * RADIX * sizeof(uint32_t) + 2 * tiles * RADIX * sizeof(uint64_t) + CNT_WORDS * sizeof(uint32_t);
}
int sort_tiles_for(long long n_cap) { return (int)((n_cap + SORT_TILE - 1) / SORT_TILE) + 1; }

// Sorts on the low `passes*8` bits.  Input in (keys_in, vals_in) = A if passes is even, else B;
// the output always ends in the A buffers.  hist [8*256], lookback [2*tiles*256], counters [CNT_WORDS].
cudaError_t launch_sort(uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, int passes, long long n_cap,
                        CountRef cnt, uint32_t* hist, uint64_t* lookback, uint32_t* counters, int sm_count, cudaStream_t stream,
                        bool hist_ready) {
  if (passes <= 0 || n_cap <= 0) return cudaSuccess;
  cudaError_t e;
  const int tiles = sort_tiles_for(n_cap);
  if (!hist_ready && (e = cudaMemsetAsync(hist, 0, 8 * RADIX * sizeof(uint32_t), stream)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(lookback, 0, (size_t)tiles * RADIX * sizeof(uint64_t), stream)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(counters + CNT_SORT_TILE0, 0, 8 * sizeof(uint32_t), stream)) != cudaSuccess) return e;
  const bool start_in_a = (passes % 2) == 0;
  uint64_t* kin = start_in_a ? keys_a : keys_b; uint32_t* vin = start_in_a ? vals_a : vals_b;
  uint64_t* kout = start_in_a ? keys_b : keys_a; uint32_t* vout = start_in_a ? vals_b : vals_a;
  {
    const size_t smem = (size_t)SORT_WARPS * passes * RADIX * sizeof(uint32_t);
    static std::atomic<unsigned long long> hist_configured{0};
    if (!hist_ready && smem > 48 * 1024) {  // per (function, device); only the stand-alone sort runs this kernel
      if ((e = cudaFuncSetAttribute(digit_histogram_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
      if ((e = cudaFuncSetAttribute(digit_histogram_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
      if ((e = cudaFuncSetAttribute(digit_histogram_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
      (void)hist_configured;
    }
    long long blocks = (n_cap + SORT_THREADS * 16 - 1) / (SORT_THREADS * 16);
    const long long max_blocks = (long long)sm_count * 4;
    if (blocks > max_blocks) blocks = max_blocks;
    if (blocks < 1) blocks = 1;
    stage_mark(B200S_STAGE_SORT_HIST, stream);
    if (!hist_ready) {  // forward computes the histograms while it emits the keys (emit_kernel); the stand-alone sort reads them here
      if (g_sort_knobs[0].load(std::memory_order_relaxed) == 0) digit_histogram_kernel<0><<<(int)blocks, SORT_THREADS, smem, stream>>>(kin, cnt, hist, passes);
      else if (g_sort_knobs[0].load(std::memory_order_relaxed) == 1) digit_histogram_kernel<1><<<(int)blocks, SORT_THREADS, smem, stream>>>(kin, cnt, hist, passes);
      else digit_histogram_kernel<2><<<(int)blocks, SORT_THREADS, smem, stream>>>(kin, cnt, hist, passes);
      count_launches(1);
    }
    digit_scan_kernel<<<1, RADIX, 0, stream>>>(hist, passes);
    count_launches(1);
  }
  stage_mark(B200S_STAGE_SORT_PASSES, stream);
  {
    static std::atomic<unsigned long long> sweep_configured{0};
    if (first_use_on_device(sweep_configured)) {
      if ((e = cudaFuncSetAttribute(onesweep_pass_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SWEEP_SMEM)) != cudaSuccess) return e;
      if ((e = cudaFuncSetAttribute(onesweep_pass_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SWEEP_SMEM)) != cudaSuccess) return e;
      if ((e = cudaFuncSetAttribute(onesweep_pass_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SWEEP_SMEM)) != cudaSuccess) return e;
    }
  }
  for (int p = 0; p < passes; p++) {
    PassArgs a;
    a.keys_in = kin; a.vals_in = vin; a.keys_out = kout; a.vals_out = vout;
    a.gbase = hist + p * RADIX;
    a.lb_cur = lookback + (size_t)(p & 1) * tiles * RADIX;
    a.lb_next = lookback + (size_t)((p + 1) & 1) * tiles * RADIX;
    a.tile_counter = counters + CNT_SORT_TILE0 + p;
    a.cnt = cnt; a.shift = 8 * p;
    const int nblk = tiles - 1 > 0 ? tiles - 1 : 1;
    if (g_sort_knobs[1].load(std::memory_order_relaxed) == 0) onesweep_pass_kernel<0><<<nblk, SORT_THREADS, SWEEP_SMEM, stream>>>(a);
    else if (g_sort_knobs[1].load(std::memory_order_relaxed) == 1) onesweep_pass_kernel<1><<<nblk, SORT_THREADS, SWEEP_SMEM, stream>>>(a);
    else onesweep_pass_kernel<2><<<nblk, SORT_THREADS, SWEEP_SMEM, stream>>>(a);
    count_launches(1);
    uint64_t* tk = kin; kin = kout; kout = tk;
    uint32_t* tv = vin; vin = vout; vout = tv;
  }
  return cudaGetLastError();
}

cudaError_t launch_tile_ranges(const uint64_t* keys, CountRef cnt, uint2* ranges, int bins, long long n_cap, int sm_count, cudaStream_t stream) {
  cudaError_t e;
  if ((e = cudaMemsetAsync(ranges, 0, (size_t)bins * sizeof(uint2), stream)) != cudaSuccess) return e;
  long long blocks = (n_cap + 256 * 8 - 1) / (256 * 8);
  const long long max_blocks = (long long)sm_count * 8;
  if (blocks > max_blocks) blocks = max_blocks;
  if (blocks < 1) blocks = 1;
  stage_mark(B200S_STAGE_RANGES, stream);
  tile_ranges_kernel<<<(int)blocks, 256, 0, stream>>>(keys, cnt, ranges);
  count_launches(1);
  return cudaGetLastError();
}

}  // namespace b200s
