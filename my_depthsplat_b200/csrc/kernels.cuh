// kernels.cuh -- argument blocks and host launchers shared between api.cu and the kernel files.
#pragma once
#include <atomic>

#include "common.cuh"

namespace b200s {

// where the pair count lives: a device word (B200sStatus.num_pairs, no host sync) or a host value
struct CountRef {
  const unsigned long long* n_dev;
  const uint32_t* overflow_dev;
  unsigned long long n_host;
};

struct CompArgs {
  int N, H, W, grid_x, tile_bits;
  const Rec* rec;
  const uint32_t* vals;
  const uint2* ranges;
  const float* bg;         // [VV,3]
  const uint32_t* overflow;
  // forward outputs / backward inputs
  float* color;            // [VV,3,H,W]
  float* depth;            // [VV,H,W] or NULL
  float* final_T;          // [VV,H,W]
  uint32_t* n_contrib;     // [VV,H,W]
  B200sStatus* status;
  // loss-side fusion (forward epilogue)
  const float* mse_target; // [VV,3,H,W] or NULL
  float* mse_grad;         // [VV,3,H,W]
  float* mse_partials;     // [VV,tiles,2]
  float mse_scale;
  int mse_l1;
  // backward
  const float* dL_dcolor;  // [VV,3,H,W]
  const float* dL_ddepth;  // [VV,H,W] or NULL
  const float* dpix_scale; // device scalar on dL_dcolor, or NULL
  float* grad_rec;         // [VV,N,12]
};

// profiling hooks implemented in api.cu (no-ops unless b200s_profile_enable(1))
void stage_mark(int stage, cudaStream_t stream);
void count_launches(int n);
cudaError_t launch_nvls_allreduce(float* multicast, unsigned long long n_floats, int rank, int world, int sm_count, cudaStream_t stream);
cudaError_t launch_nvls_reduce_segments(const float* multicast, float* local, const unsigned long long* off, const unsigned long long* cnt, int nseg,
                                        int sm_count, cudaStream_t stream);
cudaError_t launch_p2p_reduce_segments(const void* const* peers, int world, int rank, float* local, const unsigned long long* off,
                                       const unsigned long long* cnt, int nseg, int sm_count, cudaStream_t stream);
extern std::atomic<int> g_sort_knobs[4];  // A/B switches of b200s_debug_set (debug only; never change results)
int device_sm_count();                                       // of the current device, cached per device id
bool first_use_on_device(std::atomic<unsigned long long>& seen);  // true once per (call site's mask, current device)

cudaError_t launch_preprocess_bin(const B200sScene&, const B200sViews&, const B200sPlan&, char* saved, char* scratch, const B200sOut*,
                                  cudaStream_t);
size_t sort_tmp_bytes(long long n_cap);
int sort_tiles_for(long long n_cap);
cudaError_t launch_sort(uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, int passes, long long n_cap, CountRef cnt,
                        uint32_t* hist, uint64_t* lookback, uint32_t* counters, int sm_count, cudaStream_t, bool hist_ready);
cudaError_t launch_tile_ranges(const uint64_t* keys, CountRef cnt, uint2* ranges, int bins, long long n_cap, int sm_count, cudaStream_t);
// BINNED sort mode (preprocess.cu: count / scan / scatter; binsort.cu: per-bin segment sort)
// entries a bin of the class holds in shared memory (8 B each next to 16 KB of bucket counters; 4 / 2 / 1 CTAs per SM)
constexpr int BIN_CAP_XS = 4960, BIN_CAP_S = 12224, BIN_CAP_L = 26976;
constexpr int BIN_CLASSES = 4;
constexpr unsigned long long BIN_GLOBAL_MIN_PAIRS = 1ull << 26;  // below this the LSD path is fast enough whatever the bin lengths
struct BinSortWork {
  uint32_t* class_list;   // [BIN_CLASSES][bins] bin ids per size class
  uint32_t* class_count;  // [BIN_CLASSES]
  uint32_t* class_next;   // [BIN_CLASSES] dynamic fetch counters
};
cudaError_t launch_bin_sort(uint2* entries, uint2* entries_tmp, uint32_t* rank_tmp, const uint2* ranges, uint32_t* vals_out,
                            const BinSortWork& w, int bins, const uint32_t* overflow, int sm_count, cudaStream_t);
// scan of per-bin counts -> ranges, cursors, size-class lists, pair total (+ overflow flag, host status word)
cudaError_t launch_bin_scan(const uint32_t* bin_count, int bins, uint2* ranges, uint32_t* cursor, const BinSortWork& w,
                            B200sStatus* status, unsigned long long pair_capacity, unsigned long long* status_host, cudaStream_t);
cudaError_t launch_composite_fwd(const CompArgs&, int tiles, int views, bool depth, bool count, cudaStream_t);  // loss fusion iff mse_target
cudaError_t launch_composite_bwd(const CompArgs&, int tiles, int views, bool depth, cudaStream_t);
cudaError_t launch_preprocess_bwd(const B200sScene&, const B200sViews&, const B200sPlan&, const char* saved, char* scratch,
                                  const B200sGradIn&, cudaStream_t);

}  // namespace b200s
