// collective.cu -- in-place sum of a symmetric buffer over the ranks of an NVSwitch domain (NVLS), used for the
// per-Gaussian gradients of view-sharded training.
//
// Two-shot: rank r owns the r-th slice of the buffer.  For every 16-byte vector of its slice it issues ONE
// multimem.ld_reduce (the switch fetches the vector from every rank's replica and returns the sum) and ONE multimem.st
// (the switch writes the sum into every rank's replica).  Per GPU that is 1/N of the buffer pulled and 1/N pushed, the
// other (N-1)/N arrive as the other ranks' multicast stores: each byte crosses each NVLink once per direction, which is
// what a reduce-scatter + all-gather moves, without staging buffers or a ring.  The caller brackets the launch with
// cross-rank barriers on the stream (every rank's data is complete before, every rank's stores have landed after).
#include "kernels.cuh"

namespace b200s {

__global__ void __launch_bounds__(512) nvls_allreduce_kernel(float* __restrict__ mc, unsigned long long lo, unsigned long long hi) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = lo + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) {
    float* p = mc + 4 * i;
    float x, y, z, w;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "l"(p) : "memory");
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
  }
}

struct SegArgs {
  unsigned long long off[16];   // in 16-byte vectors
  unsigned long long end[16];   // running end of every segment in the concatenated vector index space
  int nseg;
};

// The calling rank's share of a reduce-scatter: pull the cross-rank sum of every vector of its segments out of the
// switch and keep it in its own replica.  A multimem.ld_reduce is a round trip through the NVSwitch (microseconds): every
// thread keeps PULL_UNROLL of them in flight, so that a few hundred resident warps -- what the projection backward this
// kernel runs next to leaves free -- cover the latency (one load in flight per thread measured 300 GB/s at N = 2).
constexpr int PULL_UNROLL = 8, PULL_THREADS = 256;
__global__ void __launch_bounds__(PULL_THREADS) nvls_reduce_segments_kernel(const float* __restrict__ mc, float* __restrict__ local, const SegArgs a) {
  const unsigned long long total = a.end[a.nseg - 1];
  const unsigned long long stride = (unsigned long long)gridDim.x * PULL_THREADS * PULL_UNROLL;
  for (unsigned long long i0 = (unsigned long long)blockIdx.x * PULL_THREADS * PULL_UNROLL + threadIdx.x; i0 < total; i0 += stride) {
    unsigned long long v[PULL_UNROLL];
    float4 r[PULL_UNROLL];
#pragma unroll
    for (int u = 0; u < PULL_UNROLL; u++) {
      const unsigned long long i = i0 + (unsigned long long)u * PULL_THREADS;  // consecutive threads, consecutive vectors
      int sgi = 0;
      while (sgi < a.nseg - 1 && i >= a.end[sgi]) sgi++;
      v[u] = i < total ? a.off[sgi] + (i - (sgi ? a.end[sgi - 1] : 0ull)) : ~0ull;
    }
#pragma unroll
    for (int u = 0; u < PULL_UNROLL; u++)
      if (v[u] != ~0ull)
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(r[u].x), "=f"(r[u].y), "=f"(r[u].z), "=f"(r[u].w) : "l"(mc + 4 * v[u]) : "memory");
#pragma unroll
    for (int u = 0; u < PULL_UNROLL; u++)
      if (v[u] != ~0ull) *reinterpret_cast<float4*>(local + 4 * v[u]) = r[u];
  }
}

// Same share of the reduce-scatter WITHOUT the switch's reduction: the rank reads the vectors of its segments from every
// peer's replica with ordinary loads over NVLink (peer pointers of the symmetric allocation) and adds them to its own.
// Why: a multimem.ld_reduce makes the switch read EVERY replica, the caller's own included, so each GPU's whole buffer
// leaves through its NVLink once per step whatever the number of ranks, in 16-byte requests (~600 GB/s measured, 0.8 ms per
// step at 2.9 M Gaussians).  Peer loads move (N-1)/N of the buffer per GPU -- half at N = 2 -- as coalesced 512-byte warp
// requests.  Sum order: own replica, then ranks rank+1, rank+2, ... (mod N) -- fixed, so the result is deterministic.
constexpr int P2P_UNROLL = 2, P2P_MAX_WORLD = 16;
struct PeerArgs {
  const float* peer[P2P_MAX_WORLD];  // peer[0] = own replica, then the others in ring order
  int world;
};
// SYS: ld.relaxed.sys (system-scope coherent); otherwise ld.global.cg (L2 only: a peer's memory is never in this GPU's L2, and
// the L1 is not consulted) -- the replicas are complete and ordered by the cross-rank barrier before the kernel starts
template <bool SYS>
__device__ __forceinline__ float4 ld_peer_v4(const float* p) {
  float4 r;
  if (SYS) asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  else asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}
template <int WORLD, bool SYS>  // WORLD 0 = run-time world size
__global__ void __launch_bounds__(PULL_THREADS) p2p_reduce_segments_kernel(const PeerArgs pa, float* __restrict__ local, const SegArgs a) {
  const int world = WORLD ? WORLD : pa.world;
  const unsigned long long total = a.end[a.nseg - 1];
  const unsigned long long stride = (unsigned long long)gridDim.x * PULL_THREADS * P2P_UNROLL;
  for (unsigned long long i0 = (unsigned long long)blockIdx.x * PULL_THREADS * P2P_UNROLL + threadIdx.x; i0 < total; i0 += stride) {
    unsigned long long v[P2P_UNROLL];
    float4 acc[P2P_UNROLL];
#pragma unroll
    for (int u = 0; u < P2P_UNROLL; u++) {
      const unsigned long long i = i0 + (unsigned long long)u * PULL_THREADS;
      int sgi = 0;
      while (sgi < a.nseg - 1 && i >= a.end[sgi]) sgi++;
      v[u] = i < total ? a.off[sgi] + (i - (sgi ? a.end[sgi - 1] : 0ull)) : ~0ull;
    }
    if (WORLD) {  // all loads of the thread in flight at once
      float4 r[P2P_UNROLL][WORLD ? WORLD : 1];
#pragma unroll
      for (int u = 0; u < P2P_UNROLL; u++)
#pragma unroll
        for (int k = 0; k < (WORLD ? WORLD : 1); k++)
          r[u][k] = v[u] != ~0ull ? ld_peer_v4<SYS>(pa.peer[k] + 4 * v[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < P2P_UNROLL; u++) {
        acc[u] = r[u][0];
#pragma unroll
        for (int k = 1; k < (WORLD ? WORLD : 1); k++) { acc[u].x += r[u][k].x; acc[u].y += r[u][k].y; acc[u].z += r[u][k].z; acc[u].w += r[u][k].w; }
      }
    } else {
#pragma unroll
      for (int u = 0; u < P2P_UNROLL; u++) {
        acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (v[u] == ~0ull) continue;
        acc[u] = ld_peer_v4<SYS>(pa.peer[0] + 4 * v[u]);
        for (int k = 1; k < world; k++) { const float4 t = ld_peer_v4<SYS>(pa.peer[k] + 4 * v[u]); acc[u].x += t.x; acc[u].y += t.y; acc[u].z += t.z; acc[u].w += t.w; }
      }
    }
#pragma unroll
    for (int u = 0; u < P2P_UNROLL; u++)
      if (v[u] != ~0ull) *reinterpret_cast<float4*>(local + 4 * v[u]) = acc[u];
  }
}

static int fill_segments(SegArgs& a, const unsigned long long* off, const unsigned long long* cnt, int nseg, unsigned long long& run) {
  run = 0;
  a.nseg = 0;
  for (int i = 0; i < nseg; i++) {
    if (cnt[i] == 0) continue;
    a.off[a.nseg] = off[i] / 4;
    run += cnt[i] / 4;
    a.end[a.nseg++] = run;
  }
  return a.nseg;
}

cudaError_t launch_p2p_reduce_segments(const void* const* peers, int world, int rank, float* local, const unsigned long long* off,
                                       const unsigned long long* cnt, int nseg, int sm_count, cudaStream_t stream) {
  if (world < 2 || world > P2P_MAX_WORLD || rank < 0 || rank >= world) return cudaErrorInvalidValue;
  SegArgs a;
  unsigned long long run;
  if (fill_segments(a, off, cnt, nseg, run) == 0) return cudaSuccess;
  PeerArgs pa;
  pa.world = world;
  for (int k = 0; k < world; k++) pa.peer[k] = static_cast<const float*>(peers[(rank + k) % world]);
  const long long per_block = (long long)PULL_THREADS * P2P_UNROLL;
  long long blocks = (long long)((run + per_block - 1) / per_block);
  const int knob = g_sort_knobs[2].load(std::memory_order_relaxed);
  // measured at N = 2 next to the projection backward: 74 blocks 5.77 ms / step, 148: 5.40-5.45, 296: 5.37-5.39
  const long long cap = knob > 0 ? knob : 2ll * sm_count;
  if (blocks > cap) blocks = cap;
  count_launches(1);
  const bool sys = g_sort_knobs[3].load(std::memory_order_relaxed) == 1;
#define P2P_LAUNCH(W_)                                                                                        \
  do {                                                                                                        \
    if (sys) p2p_reduce_segments_kernel<W_, true><<<(int)blocks, PULL_THREADS, 0, stream>>>(pa, local, a);    \
    else p2p_reduce_segments_kernel<W_, false><<<(int)blocks, PULL_THREADS, 0, stream>>>(pa, local, a);       \
  } while (0)
  switch (world) {
    case 2: P2P_LAUNCH(2); break;
    case 4: P2P_LAUNCH(4); break;
    case 8: P2P_LAUNCH(8); break;
    default: P2P_LAUNCH(0); break;
  }
#undef P2P_LAUNCH
  return cudaGetLastError();
}

cudaError_t launch_nvls_reduce_segments(const float* multicast, float* local, const unsigned long long* off, const unsigned long long* cnt, int nseg,
                                        int sm_count, cudaStream_t stream) {
  SegArgs a;
  unsigned long long run = 0;
  a.nseg = 0;
  for (int i = 0; i < nseg; i++) {
    if (cnt[i] == 0) continue;
    a.off[a.nseg] = off[i] / 4;
    run += cnt[i] / 4;
    a.end[a.nseg++] = run;
  }
  if (a.nseg == 0) return cudaSuccess;
  const long long per_block = (long long)PULL_THREADS * PULL_UNROLL;
  long long blocks = (long long)((run + per_block - 1) / per_block);
  // a side-stream kernel that is bound by the NVLink, not by the SMs: enough blocks to keep the link's bandwidth-delay
  // product in flight (a few MB of 16-byte requests), no more -- every resident block takes registers and issue slots from
  // the projection backward it runs next to
  const int knob = g_sort_knobs[2].load(std::memory_order_relaxed);
  const long long cap = knob > 0 ? knob : (long long)sm_count / 2;
  if (blocks > cap) blocks = cap;
  count_launches(1);
  nvls_reduce_segments_kernel<<<(int)blocks, PULL_THREADS, 0, stream>>>(multicast, local, a);
  return cudaGetLastError();
}

cudaError_t launch_nvls_allreduce(float* multicast, unsigned long long n_floats, int rank, int world, int sm_count, cudaStream_t stream) {
  const unsigned long long n_vec = n_floats / 4;
  const unsigned long long per = (n_vec + world - 1) / world;
  const unsigned long long lo = per * rank, hi = lo + per < n_vec ? lo + per : n_vec;
  if (lo >= hi) return cudaSuccess;
  long long blocks = (long long)((hi - lo + 511) / 512);
  const long long cap = (long long)sm_count * 4;
  if (blocks > cap) blocks = cap;
  count_launches(1);
  nvls_allreduce_kernel<<<(int)blocks, 512, 0, stream>>>(multicast, lo, hi);
  return cudaGetLastError();
}

}  // namespace b200s
