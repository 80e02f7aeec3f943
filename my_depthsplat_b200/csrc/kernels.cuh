// kernels.cuh -- argument blocks and host launchers shared between api.cu and the kernel files.
#pragma once
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
  // backward
  const float* dL_dcolor;  // [VV,3,H,W]
  const float* dL_ddepth;  // [VV,H,W] or NULL
  float* grad_rec;         // [VV,N,12]
};

// profiling hooks implemented in api.cu (no-ops unless b200s_profile_enable(1))
void stage_mark(int stage, cudaStream_t stream);
void count_launches(int n);
cudaError_t launch_nvls_allreduce(float* multicast, unsigned long long n_floats, int rank, int world, int sm_count, cudaStream_t stream);
extern int g_sort_knobs[4];

cudaError_t launch_preprocess_bin(const B200sScene&, const B200sViews&, const B200sPlan&, char* saved, char* scratch, const B200sOut*,
                                  cudaStream_t);
size_t sort_tmp_bytes(long long n_cap);
int sort_tiles_for(long long n_cap);
cudaError_t launch_sort(uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, int passes, long long n_cap, CountRef cnt,
                        uint32_t* hist, uint64_t* lookback, uint32_t* counters, int sm_count, cudaStream_t, bool hist_ready);
cudaError_t launch_tile_ranges(const uint64_t* keys, CountRef cnt, uint2* ranges, int bins, long long n_cap, int sm_count, cudaStream_t);
cudaError_t launch_composite_fwd(const CompArgs&, int tiles, int views, bool depth, bool count, cudaStream_t);
cudaError_t launch_composite_bwd(const CompArgs&, int tiles, int views, bool depth, cudaStream_t);
cudaError_t launch_preprocess_bwd(const B200sScene&, const B200sViews&, const B200sPlan&, const char* saved, char* scratch,
                                  const B200sGradIn&, cudaStream_t);

}  // namespace b200s
