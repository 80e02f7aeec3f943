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
// switch and keep it in its own replica.
__global__ void __launch_bounds__(512) nvls_reduce_segments_kernel(const float* __restrict__ mc, float* __restrict__ local, const SegArgs a) {
  const unsigned long long total = a.end[a.nseg - 1];
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int sgi = 0;
    while (i >= a.end[sgi]) sgi++;
    const unsigned long long v = a.off[sgi] + (i - (sgi ? a.end[sgi - 1] : 0ull));
    float x, y, z, w;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "l"(mc + 4 * v) : "memory");
    *reinterpret_cast<float4*>(local + 4 * v) = make_float4(x, y, z, w);
  }
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
  long long blocks = (long long)((run + 511) / 512);
  const long long cap = (long long)sm_count * 2;  // a side-stream kernel: leave SMs to the projection backward it overlaps
  if (blocks > cap) blocks = cap;
  count_launches(1);
  nvls_reduce_segments_kernel<<<(int)blocks, 512, 0, stream>>>(multicast, local, a);
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
