// binsort.cu -- BINNED sort mode: every (view, tile) bin's entries are put into depth order by ONE CTA, in shared
// memory.  Together with the counting / scattering kernels of preprocess.cu this replaces
// cub::DeviceRadixSort::SortPairs + identifyTileRanges behind _C.rasterize_gaussians: instead of
// six global passes over 64-bit (view | tile | depth) keys (24 B of HBM traffic per pair per pass), the pairs are
// scattered straight into their bin (8 B written), and each bin is read (8 B, twice more out of L2) and its sorted
// Gaussian indices written once (4 B).
//
// Order inside a bin = the order the stable global sort produces: ascending depth bits, ties in ascending Gaussian
// index (the global sort is stable and pairs are emitted in index order).  The scatter claims positions with atomics,
// so the arrival order is arbitrary; the segment sort therefore orders by the pair (depth bits, index).
//
// bucket_sort_kernel (bins that fit shared memory: three size classes at 4 / 2 / 1 CTAs per SM):
//   1. read the bin: min / max of the depth bits -> b = significant bits of (depth - min);
//   2. ONE counting sort on the top ~log2(2n) (8..13) bits of (depth - min): read the bin again (L2), histogram with
//      fire-and-forget shared atomics on 16-bit counters, one block-wide scan, read the bin a third time and place
//      every entry at a position claimed from its bucket (arbitrary order inside a bucket);
//   3. with about as many buckets as entries, a bucket holds a handful of entries (a tile's depths are spread over
//      thousands of 2^low-ulp intervals), so every entry finds its final place by counting the smaller
//      (depth, index) pairs of its own bucket, and writes its index there.
//   That is ~2 warp instructions per entry; an LSD radix sort of the same keys needs ~10 (four passes of two sweeps).
//   A bin where more than RUN_CAP entries crowd into one bucket (degenerate scenes: thousands of Gaussians within a few
//   ulp of depth) is handed to the robust path below instead.
// lsd_sort_kernel (bins longer than the largest shared-memory class, and crowded bins): stable LSD counting passes over
//   8-bit digits of all significant depth bits on ping-pong buffers in global memory (L2-resident for the sizes that
//   occur) -- each warp ranks a contiguous chunk with MATCH.ANY groups + one shared atomic per group, one block-wide
//   scan of the [digit][warp] counters gives the destinations -- then exact-depth ties are ordered by index (rank
//   inside the run; index passes first when a run is long).  No limit on the bin length.
// The class lists are built on the device by bin_scan_kernel; CTAs fetch bins from them dynamically.
#include "kernels.cuh"

namespace b200s {

constexpr int BS_DIGITS = 256;
constexpr int RUN_CAP = 64;  // longest run of equal sorted bits the rank fix-up takes

struct BinSortArgs {
  uint2* entries;            // [R] (depth bits, Gaussian index), grouped by bin, arbitrary order inside a bin
  uint2* entries_tmp;        // [R] XL ping-pong
  uint32_t* rank_tmp;        // [R] XL ranks
  const uint2* ranges;       // [bins] (start, end)
  uint32_t* vals_out;        // [R] sorted Gaussian indices
  const uint32_t* list;      // bin ids of this size class
  const uint32_t* count;     // how many
  uint32_t* next;            // dynamic fetch counter
  const uint32_t* overflow;
  uint32_t* retry_list;      // bucket path: bins too crowded for it are appended to the LSD class's list ...
  uint32_t* retry_count;     // ... by bumping its count
};

// ---- the (key, index) ping-pong and the ranks of the LSD path live in global memory --------------------------------
struct GmemStore {
  uint2 *A, *B;
  uint32_t* rk;
  __device__ __forceinline__ uint32_t key(uint32_t i) const { return A[i].x; }
  __device__ __forceinline__ uint32_t val(uint32_t i) const { return A[i].y; }
  __device__ __forceinline__ uint2 get(uint32_t i) const { return A[i]; }
  __device__ __forceinline__ void put(uint32_t i, uint2 e) { B[i] = e; }
  __device__ __forceinline__ void set_rank(uint32_t i, uint32_t r) { rk[i] = r; }
  __device__ __forceinline__ uint32_t rank(uint32_t i) const { return rk[i]; }
  __device__ __forceinline__ void swap() { uint2* t = A; A = B; B = t; }
};

// Block-wide exclusive scan of the [digit][warp] counters (digit-major, rows padded to W + 1 words so that a warp's
// atomics on 32 different digits fall into 32 different banks): afterwards hist[d][w] = number of elements with a smaller
// digit, or the same digit in an earlier warp chunk = the destination of warp w's first element with digit d.
template <int T>
__device__ __forceinline__ void scan_counters(uint32_t* hist, uint32_t* s_wtot) {
  constexpr int W = T / 32, STR = W + 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t* p = hist + ((8 * tid) / W) * STR + (8 * tid) % W;  // eight consecutive warps of one digit
  uint32_t c[8], s = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) { c[j] = p[j]; s += c[j]; }
  uint32_t incl = s;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
  if (lane == 31) s_wtot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const uint32_t t = lane < W ? s_wtot[lane] : 0u;
    uint32_t x = t;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
    s_wtot[lane] = x - t;
  }
  __syncthreads();
  uint32_t run = s_wtot[warp] + incl - s;
#pragma unroll
  for (int j = 0; j < 8; j++) { p[j] = run; run += c[j]; }
}

// One stable counting pass over the 8-bit digit `shift` of (key - sub) [ON_VAL: of (index - sub)].
// Both sweeps go over a warp's chunk UNROLL x 32 elements at a time with the loads, matches, atomics and shuffles of the
// UNROLL groups issued back to back: the chain load -> MATCH -> shared atomic -> shuffle -> store is ~10^2 cycles long
// and a CTA has few warps, so the instruction-level parallelism of independent groups is what keeps the SM busy
// (same-address shared atomics of one warp execute in program order, so the ranks stay stable).
constexpr int BS_UNROLL = 4;

template <int T, bool ON_VAL, class Store>
__device__ __forceinline__ void radix_pass(Store& st, uint32_t n, uint32_t sub, int shift, uint32_t* hist, uint32_t* s_wtot) {
  constexpr int W = T / 32, STR = W + 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  for (int i = tid; i < BS_DIGITS * STR; i += T) hist[i] = 0;
  __syncthreads();
  const uint32_t chunk = ((n + W - 1) / W + 31u) & ~31u;
  const uint32_t lo = min(n, warp * chunk), hi = min(n, lo + chunk);
  uint32_t* col = hist + warp;
  for (uint32_t base = lo; base < hi; base += 32 * BS_UNROLL) {
    uint32_t d[BS_UNROLL], peers[BS_UNROLL], r[BS_UNROLL];
#pragma unroll
    for (int u = 0; u < BS_UNROLL; u++) {
      const uint32_t i = base + 32 * u + lane;
      const uint32_t x = i < hi ? (ON_VAL ? st.val(i) : st.key(i)) : 0u;
      d[u] = i < hi ? ((x - sub) >> shift) & 255u : 256u;
    }
#pragma unroll
    for (int u = 0; u < BS_UNROLL; u++) peers[u] = __match_any_sync(0xffffffffu, d[u]);
#pragma unroll
    for (int u = 0; u < BS_UNROLL; u++) {
      r[u] = 0;
      if (d[u] < 256u && lane == __ffs(peers[u]) - 1) r[u] = atomicAdd(col + d[u] * STR, (uint32_t)__popc(peers[u]));
    }
#pragma unroll
    for (int u = 0; u < BS_UNROLL; u++) {
      const uint32_t i = base + 32 * u + lane;
      const uint32_t rr = __shfl_sync(0xffffffffu, r[u], __ffs(peers[u]) - 1);
      if (i < hi) st.set_rank(i, rr + __popc(peers[u] & lt));
    }
  }
  __syncthreads();
  scan_counters<T>(hist, s_wtot);
  __syncthreads();
  for (uint32_t base = lo; base < hi; base += 32 * BS_UNROLL) {
    uint2 e[BS_UNROLL];
    uint32_t dst[BS_UNROLL];
#pragma unroll
    for (int u = 0; u < BS_UNROLL; u++) {
      const uint32_t i = base + 32 * u + lane;
      if (i < hi) { e[u] = st.get(i); dst[u] = st.rank(i); }
    }
#pragma unroll
    for (int u = 0; u < BS_UNROLL; u++) {
      const uint32_t i = base + 32 * u + lane;
      if (i < hi) dst[u] += col[((((ON_VAL ? e[u].y : e[u].x) - sub) >> shift) & 255u) * STR];
    }
#pragma unroll
    for (int u = 0; u < BS_UNROLL; u++) {
      const uint32_t i = base + 32 * u + lane;
      if (i < hi) st.put(dst[u], e[u]);
    }
  }
  __syncthreads();
  st.swap();
}

// (key, index) order
__device__ __forceinline__ bool pair_less(uint2 a, uint2 b) { return a.x < b.x || (a.x == b.x && a.y < b.y); }

template <int T, class Store>
__device__ __forceinline__ void sort_segment(Store& st, uint32_t n, uint32_t kmin, uint32_t kmax, uint32_t vmin, uint32_t vmax,
                                             uint32_t* hist, uint32_t* s_wtot, uint32_t* s_flag, uint32_t* __restrict__ out) {
  const int tid = threadIdx.x;
  const int kbits = 32 - __clz(kmax - kmin);  // __clz(0) = 32
  for (int shift = 0; shift < kbits; shift += 8) radix_pass<T, false>(st, n, kmin, shift, hist, s_wtot);
  // a run of more than RUN_CAP exactly equal depths?
  if (tid == 0) *s_flag = 0;
  __syncthreads();
  {
    bool lng = false;
    for (uint32_t i = tid; i + RUN_CAP < n; i += T) lng |= st.key(i) == st.key(i + RUN_CAP);
    if (lng) *s_flag = 1;
  }
  __syncthreads();
  const bool long_run = *s_flag != 0;
  if (long_run) {  // degenerate bin: the complete LSD order -- index passes, then the depth passes again
    const int vbits = 32 - __clz(vmax - vmin);
    for (int shift = 0; shift < vbits; shift += 8) radix_pass<T, true>(st, n, vmin, shift, hist, s_wtot);
    for (int shift = 0; shift < kbits; shift += 8) radix_pass<T, false>(st, n, kmin, shift, hist, s_wtot);
  }
  for (uint32_t i = tid; i < n; i += T) {
    const uint2 e = st.get(i);
    uint32_t pos = i;
    if (!long_run) {
      const bool eq_prev = i > 0 && st.key(i - 1) == e.x, eq_next = i + 1 < n && st.key(i + 1) == e.x;
      if (eq_prev || eq_next) {  // member of an equal-depth run: its place is the number of smaller indices in the run
        uint32_t s = i, t = i + 1, smaller = 0;
        while (s > 0 && st.key(s - 1) == e.x) { s--; smaller += st.val(s) < e.y; }
        while (t < n && st.key(t) == e.x) { smaller += st.val(t) < e.y; t++; }
        pos = s + smaller;
      }
    }
    out[pos] = e.y;
  }
}

__device__ __forceinline__ void block_minmax(uint32_t kmn, uint32_t kmx, uint32_t vmn, uint32_t vmx, uint32_t* s_mm) {
  kmn = __reduce_min_sync(0xffffffffu, kmn); kmx = __reduce_max_sync(0xffffffffu, kmx);
  vmn = __reduce_min_sync(0xffffffffu, vmn); vmx = __reduce_max_sync(0xffffffffu, vmx);
  if ((threadIdx.x & 31) == 0) { atomicMin(&s_mm[0], kmn); atomicMax(&s_mm[1], kmx); atomicMin(&s_mm[2], vmn); atomicMax(&s_mm[3], vmx); }
}

// ---- LSD path: any bin length, ping-pong in global memory -------------------------------------------------------------
template <int T>
__global__ void __launch_bounds__(T, 1) lsd_sort_kernel(const BinSortArgs a) {
  constexpr int W = T / 32, STR = W + 1;
  __shared__ uint32_t hist[BS_DIGITS * STR];
  __shared__ uint32_t s_wtot[32];
  __shared__ uint32_t s_mm[4];
  __shared__ uint32_t s_flag, s_fetch;
  if (*a.overflow) return;
  const int tid = threadIdx.x;
  const uint32_t nbins = *a.count;
  for (;;) {
    __syncthreads();  // the previous bin's readers of the shared state are done
    if (tid == 0) { s_fetch = atomicAdd(a.next, 1u); s_mm[0] = 0xffffffffu; s_mm[1] = 0u; s_mm[2] = 0xffffffffu; s_mm[3] = 0u; }
    __syncthreads();
    if (s_fetch >= nbins) return;
    const uint2 range = a.ranges[a.list[s_fetch]];
    const uint32_t n = range.y - range.x;
    uint32_t kmn = 0xffffffffu, kmx = 0u, vmn = 0xffffffffu, vmx = 0u;
    GmemStore st{a.entries + range.x, a.entries_tmp + range.x, a.rank_tmp + range.x};
    for (uint32_t i = tid; i < n; i += T) {
      const uint2 e = st.A[i];
      kmn = min(kmn, e.x); kmx = max(kmx, e.x); vmn = min(vmn, e.y); vmx = max(vmx, e.y);
    }
    block_minmax(kmn, kmx, vmn, vmx, s_mm);
    __syncthreads();
    sort_segment<T>(st, n, s_mm[0], s_mm[1], s_mm[2], s_mm[3], hist, s_wtot, &s_flag, a.vals_out + range.x);
  }
}

// ---- bucket path ---------------------------------------------------------------------------------------------------------
constexpr int BK_MAX_BITS = 13;                       // at most 8192 buckets: 16-bit counters, two per word
constexpr int BK_WORDS = (1 << BK_MAX_BITS) / 2;

// exclusive scan of nb 16-bit counters packed two per word (nb a power of two >= 256); positions stay below 65536
template <int T>
__device__ __forceinline__ void scan_u16(uint32_t* h, int nb, uint32_t* s_wtot) {
  constexpr int WPT = BK_WORDS / T > 0 ? BK_WORDS / T : 1;   // words per thread at the largest histogram
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int words = nb >> 1;
  const int wpt = words / T > 0 ? words / T : 1;
  const int w0 = tid * wpt;
  uint32_t c[WPT], sum = 0;
#pragma unroll
  for (int j = 0; j < WPT; j++) {
    c[j] = (j < wpt && w0 + j < words) ? h[w0 + j] : 0u;
    sum += (c[j] & 0xffffu) + (c[j] >> 16);
  }
  uint32_t incl = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
  if (lane == 31) s_wtot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const uint32_t t = lane < T / 32 ? s_wtot[lane] : 0u;
    uint32_t x = t;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
    s_wtot[lane] = x - t;
  }
  __syncthreads();
  uint32_t run = s_wtot[warp] + incl - sum;
#pragma unroll
  for (int j = 0; j < WPT; j++) {
    if (j < wpt && w0 + j < words) {
      const uint32_t lo = c[j] & 0xffffu, hi = c[j] >> 16;
      h[w0 + j] = run | ((run + lo) << 16);
      run += lo + hi;
    }
  }
}

template <int T, int CAP, int MIN_CTAS>
__global__ void __launch_bounds__(T, MIN_CTAS) bucket_sort_kernel(const BinSortArgs a) {
  extern __shared__ __align__(16) uint32_t bs_smem[];
  __shared__ uint32_t s_wtot[32];
  __shared__ uint32_t s_mm[4];
  __shared__ uint32_t s_flag, s_fetch;
  if (*a.overflow) return;
  uint32_t* hist = bs_smem;            // [BK_WORDS] two 16-bit bucket counters per word
  uint32_t* kB = hist + BK_WORDS;      // [CAP] depth bits in bucket order
  uint32_t* vB = kB + CAP;             // [CAP] Gaussian indices
  const int tid = threadIdx.x;
  const uint32_t nbins = *a.count;
  for (;;) {
    __syncthreads();  // the previous bin's readers of the shared state are done
    if (tid == 0) { s_fetch = atomicAdd(a.next, 1u); s_mm[0] = 0xffffffffu; s_mm[1] = 0u; s_mm[2] = 0xffffffffu; s_mm[3] = 0u; s_flag = 0; }
    __syncthreads();
    if (s_fetch >= nbins) return;
    const uint32_t bin = a.list[s_fetch];
    const uint2 range = a.ranges[bin];
    const uint32_t n = range.y - range.x;
    const uint2* __restrict__ src = a.entries + range.x;
    // bucket count ~ 2n .. 4n (256 .. 8192)
    int bbits = 33 - __clz(n);
    bbits = bbits < 8 ? 8 : (bbits > BK_MAX_BITS ? BK_MAX_BITS : bbits);
    const int nb = 1 << bbits;
    for (int i = tid; i < (nb >> 1); i += T) hist[i] = 0;
    uint32_t kmn = 0xffffffffu, kmx = 0u;
    for (uint32_t i = tid; i < n; i += T) { const uint32_t k = src[i].x; kmn = min(kmn, k); kmx = max(kmx, k); }
    kmn = __reduce_min_sync(0xffffffffu, kmn); kmx = __reduce_max_sync(0xffffffffu, kmx);
    if ((tid & 31) == 0) { atomicMin(&s_mm[0], kmn); atomicMax(&s_mm[1], kmx); }
    __syncthreads();
    const uint32_t kmin = s_mm[0];
    // bucket(k) = floor((k - min) * nb / (range + 1)): monotone, and ALL nb buckets span [min, max] (a shift by whole bits
    // would leave up to half of them unused); evaluated as the high word of a 32 x 32 bit product
    const uint32_t krange = s_mm[1] - kmin;
    const uint32_t mul = (uint32_t)min(0xffffffffull, ((unsigned long long)nb << 32) / ((unsigned long long)krange + 1ull));
    const bool exact = krange < (uint32_t)nb;  // fewer distinct depths than buckets: one depth value per bucket
#define BUCKET(k) (exact ? ((k) - kmin) : __umulhi((k) - kmin, mul))
    for (uint32_t i = tid; i < n; i += T) {
      const uint32_t b = BUCKET(src[i].x);
      atomicAdd(&hist[b >> 1], 1u << ((b & 1u) * 16));
    }
    __syncthreads();
    scan_u16<T>(hist, nb, s_wtot);
    __syncthreads();
    for (uint32_t i = tid; i < n; i += T) {
      const uint2 e = src[i];
      const uint32_t b = BUCKET(e.x), sh = (b & 1u) * 16;
      const uint32_t pos = (atomicAdd(&hist[b >> 1], 1u << sh) >> sh) & 0xffffu;
      kB[pos] = e.x; vB[pos] = e.y;
    }
    __syncthreads();
    {  // more than RUN_CAP entries in one bucket?
      bool crowded = false;
      for (uint32_t i = tid; i + RUN_CAP < n; i += T) crowded |= BUCKET(kB[i]) == BUCKET(kB[i + RUN_CAP]);
      if (crowded) s_flag = 1;
    }
    __syncthreads();
    if (s_flag) {  // hand the bin to the LSD kernel, which runs after this one
      if (tid == 0) a.retry_list[atomicAdd(a.retry_count, 1u)] = bin;
      continue;
    }
    uint32_t* __restrict__ out = a.vals_out + range.x;
    for (uint32_t i = tid; i < n; i += T) {
      const uint32_t k = kB[i], v = vB[i];
      const uint32_t top = BUCKET(k);
      uint32_t pos = i;
      const bool eq_prev = i > 0 && BUCKET(kB[i - 1]) == top, eq_next = i + 1 < n && BUCKET(kB[i + 1]) == top;
      if (eq_prev || eq_next) {  // its place inside the bucket = the number of smaller (depth, index) pairs there
        uint32_t s = i, t = i + 1, smaller = 0;
        while (s > 0) { const uint32_t ok = kB[s - 1]; if (BUCKET(ok) != top) break; s--; smaller += ok < k || (ok == k && vB[s] < v); }
        while (t < n) { const uint32_t ok = kB[t]; if (BUCKET(ok) != top) break; smaller += ok < k || (ok == k && vB[t] < v); t++; }
        pos = s + smaller;
      }
      out[pos] = v;
    }
#undef BUCKET
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------
template <int T, int CAP, int MIN_CTAS>
static cudaError_t launch_bucket_class(BinSortArgs a, int cls, const BinSortWork& w, int bins, int sm_count, cudaStream_t stream) {
  constexpr size_t smem = (size_t)BK_WORDS * 4 + (size_t)CAP * 8;
  static std::atomic<unsigned long long> configured{0};  // bit per device: the attribute is per (function, device)
  cudaError_t e;
  if (first_use_on_device(configured)) {
    if ((e = cudaFuncSetAttribute(bucket_sort_kernel<T, CAP, MIN_CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(bucket_sort_kernel<T, CAP, MIN_CTAS>, cudaFuncAttributePreferredSharedMemoryCarveout, 100)) != cudaSuccess) return e;
  }
  a.list = w.class_list + (size_t)cls * bins;
  a.count = w.class_count + cls;
  a.next = w.class_next + cls;
  const int grid = bins < sm_count * MIN_CTAS ? bins : sm_count * MIN_CTAS;
  bucket_sort_kernel<T, CAP, MIN_CTAS><<<grid, T, smem, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_bin_sort(uint2* entries, uint2* entries_tmp, uint32_t* rank_tmp, const uint2* ranges, uint32_t* vals_out,
                            const BinSortWork& w, int bins, const uint32_t* overflow, int sm_count, cudaStream_t stream) {
  if (bins <= 0) return cudaSuccess;
  BinSortArgs a;
  a.entries = entries; a.entries_tmp = entries_tmp; a.rank_tmp = rank_tmp; a.ranges = ranges; a.vals_out = vals_out;
  a.list = nullptr; a.count = nullptr; a.next = nullptr; a.overflow = overflow;
  // crowded bins are appended to the list of the LSD class (3), which runs last
  a.retry_list = w.class_list + (size_t)3 * bins;
  a.retry_count = w.class_count + 3;
  stage_mark(B200S_STAGE_BIN_SORT, stream);
  cudaError_t e;
  // longest first: the long bins of a skewed scene start while every SM is still free
  if ((e = launch_bucket_class<1024, BIN_CAP_L, 1>(a, 2, w, bins, sm_count, stream)) != cudaSuccess) return e;
  if ((e = launch_bucket_class<512, BIN_CAP_S, 2>(a, 1, w, bins, sm_count, stream)) != cudaSuccess) return e;
  if ((e = launch_bucket_class<256, BIN_CAP_XS, 4>(a, 0, w, bins, sm_count, stream)) != cudaSuccess) return e;
  a.list = w.class_list + (size_t)3 * bins;
  a.count = w.class_count + 3;
  a.next = w.class_next + 3;
  lsd_sort_kernel<1024><<<bins < sm_count ? bins : sm_count, 1024, 0, stream>>>(a);
  count_launches(4);
  return cudaGetLastError();
}

}  // namespace b200s
