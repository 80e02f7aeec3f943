/*
 * b200splat.h -- C ABI of the B200-native (sm_100a) tile-based Gaussian-splatting rasterizer.
 *
 * This is the drop-in boundary for DepthSplat's rendering hot path.  The reference has NO native
 * code of its own on this path: it binds the third-party pybind11 module
 * `diff_gaussian_rasterization._C` (requirements.txt:23) through
 *     src/model/decoder/cuda_splatting.py:98-123   GaussianRasterizationSettings + GaussianRasterizer(...)
 *     src/model/decoder/cuda_splatting.py:191-216  same call, orthographic cameras
 * once PER VIEW inside a Python loop (cuda_splatting.py:90), after replicating every Gaussian tensor
 * per view (decoder_splatting_cuda.py:53-56).  The two entry points that module exports,
 *     _C.rasterize_gaussians(bg, means3D, colors_precomp, opacities, scales, rotations, scale_modifier,
 *                            cov3D_precomp, viewmatrix, projmatrix, tanfovx, tanfovy, H, W, sh, degree,
 *                            campos, prefiltered, debug)
 *     _C.rasterize_gaussians_backward(...)
 * are replaced by b200s_forward_* / b200s_backward below, which take ALL views of ALL scenes of a
 * batch in one call and read the un-replicated [B,N,...] Gaussian tensors directly.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise;
 *   - the caller owns every buffer (inputs, outputs, the two workspaces); the library never
 *     allocates, frees or synchronises; all work is enqueued on the `stream` argument
 *     (a cudaStream_t passed as void*);
 *   - re-entrant: nothing a call computes depends on process state.  The only process-wide state is opt-in and
 *     debug-only -- the stage profiler (b200s_profile_*), the launch counter and the A/B knobs of b200s_debug_set, which
 *     select between kernel variants with identical results -- plus per-device caches of device attributes;
 *   - return value: B200S_OK, or B200S_EBADARG (nothing enqueued), or B200S_ECUDA (a launch failed;
 *     `b200s_last_cuda_error` holds the code).  Capacity overflow of the (tile,depth) pair buffers
 *     is NOT a return code: it is reported in the device status block (B200sStatus) so that the
 *     happy path needs no host synchronisation; the host reads it when it chooses to.
 */
#ifndef B200SPLAT_H
#define B200SPLAT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200S_ABI_VERSION 11

enum { B200S_OK = 0, B200S_EBADARG = 1, B200S_ECUDA = 3 };

/* Gaussian tensor layouts accepted directly (no transposition copies on the caller's side). */
enum { B200S_COV_3X3 = 0 /* [B,N,3,3] DepthSplat Gaussians.covariances */, B200S_COV_UPPER6 = 1 /* [B,N,6] extension cov3D_precomp */ };
enum { B200S_SH_CHANNEL_MAJOR = 0 /* [B,N,3,d_sh] DepthSplat Gaussians.harmonics */, B200S_SH_COEFF_MAJOR = 1 /* [B,N,d_sh,3] extension shs */ };
/* depth channel composited next to RGB: colour = f(z_cam) per render_depth_cuda (cuda_splatting.py:238-246) */
enum { B200S_DEPTH_NONE = 0, B200S_DEPTH_Z = 1 /* "depth", "relative_disparity" */, B200S_DEPTH_DISPARITY = 2, B200S_DEPTH_LOG = 3 };

/* The scene batch: B scenes of N Gaussians each, fp32, contiguous. */
typedef struct B200sScene {
  int32_t num_scenes;     /* B */
  int32_t num_gaussians;  /* N per scene */
  int32_t sh_degree;      /* 0..3 active degree; ignored when colors_precomp != NULL */
  int32_t sh_coeffs;      /* d_sh = coefficients stored per channel (>= (sh_degree+1)^2) */
  int32_t cov_layout;     /* B200S_COV_* */
  int32_t sh_layout;      /* B200S_SH_* */
  const float* means;          /* [B,N,3] */
  const float* covariances;    /* [B,N,3,3] or [B,N,6] */
  const float* harmonics;      /* [B,N,3,d_sh] or [B,N,d_sh,3]; NULL when colors_precomp is used */
  const float* colors_precomp; /* [B,N,3] or NULL (use_sh=False path, cuda_splatting.py:117) */
  const float* opacities;      /* [B,N] */
  /* Optional -- the Gaussians as the encoder head's RAW output, with the Gaussian adapter
   * (src/model/encoder/common/gaussian_adapter.py:49-102, gaussians.py:8-44, src/misc/sh_rotation.py:10-30) fused into the
   * projection.  When raw_head != NULL the five tensor pointers above are ignored and must be NULL,
   * num_gaussians must equal raw_views * raw_h * raw_w (order (view, y, x)), raw_h * raw_w must be a multiple of 256, and
   * sh_degree / sh_coeffs must be 2 / 9.  The world-space tensors are never materialised. */
  const float* raw_head;   /* [B, raw_views, 37, raw_h*raw_w] channel planes of the head: opacity logit | 2 xy-offset logits |
                              3 scales | quaternion xyzw | 27 SH (channel-major 3 x 9) */
  const float* raw_depth;  /* [B, raw_views, raw_h*raw_w] */
  const float* raw_image;  /* [B, raw_views, 3, raw_h*raw_w] context images (SH DC initialisation, gaussian_adapter.py:78-83) */
  const float* raw_camera; /* [B, raw_views, 56] per context view: camera-to-world rotation (9, row-major) | translation (3) |
                              inverse intrinsics (9) | pad | degree-2 SH rotation (25) | SH mask (9) */
  int32_t raw_views, raw_h, raw_w;
  float raw_scale_min, raw_scale_max;  /* gaussian_scale_min / gaussian_scale_max */
  float* raw_cooked_out;   /* optional (tests): [B,N,40] = mean 3 | covariance 3x3 | opacity | harmonics 3x9 as the kernel built them */
} B200sScene;

/* The views: VV = total number of (scene, target camera) pairs rendered by this call. */
typedef struct B200sViews {
  int32_t num_views;      /* VV */
  int32_t height, width;  /* H, W (all views of a call share the image size) */
  int32_t depth_mode;     /* B200S_DEPTH_* */
  const int32_t* scene_index; /* [VV] scene of each view */
  const float* viewmatrix;    /* [VV,16] world->camera, transposed storage (cuda_splatting.py:85) */
  const float* projmatrix;    /* [VV,16] full projection, transposed storage (cuda_splatting.py:86) */
  const float* campos;        /* [VV,3] */
  const float* tanfov;        /* [VV,2] (tanfovx, tanfovy) */
  const float* background;    /* [VV,3] */
  const float* scale;         /* [VV,2] (s, s*s): means*s and cov*s^2 are applied in-kernel (scale_invariant,
                                 cuda_splatting.py:63-70); NULL => 1 */
  const float* depth_affine;  /* [VV,4] row 2 of the UNnormalised world->camera matrix, z_cam = r.m + t
                                 (cuda_splatting.py:238-241); required when depth_mode != NONE */
  const float* depth_clamp;   /* [VV,2] (near, far) for B200S_DEPTH_LOG; may be NULL otherwise */
} B200sViews;

/* How the (tile, Gaussian) pairs get into per-(view, tile) depth order -- what cub::DeviceRadixSort::SortPairs +
 * identifyTileRanges do behind _C.rasterize_gaussians.  Both modes produce the same lists and ranges, bit for bit.
 *   BINNED (default): per-bin pair counts by warp-aggregated atomics -> one scan = the tile ranges -> (depth bits, index)
 *     entries scattered into their bin (one atomic claim per run of lanes with the same bin) -> ONE CTA per bin orders
 *     its segment by (depth bits, index) with LSD counting passes in shared memory over the significant bits of
 *     (depth - bin minimum).  20 bytes of HBM traffic per pair.  Bins of more than 11 008 entries run the same passes on
 *     global (L2-resident) ping-pong buffers: no limit on the bin length.
 *   GLOBAL: one stable onesweep LSD radix sort of all 64-bit (view | tile | depth) keys, 8 bits per pass, then a
 *     range-finding pass.  24 bytes per pair per pass. */
enum { B200S_SORT_BINNED = 0, B200S_SORT_GLOBAL = 1 };

/* Problem dimensions -> workspace plan. */
typedef struct B200sDims {
  int32_t num_scenes, num_gaussians, num_views, height, width;
  int32_t sort_mode;     /* B200S_SORT_* */
  int64_t pair_capacity; /* R_cap: capacity of the (key,value) pair buffers, < 2^32 - 8192 (list positions are 32-bit) */
} B200sDims;

/* Byte offsets of every region inside the two caller-owned workspaces.
 * `saved` must stay alive from forward to backward; `scratch` may be reused right after forward. */
typedef struct B200sPlan {
  int32_t abi_version;
  int32_t tile_bits, view_bits, sort_bits, sort_passes;
  int32_t grid_x, grid_y, tiles, bins; /* bins = num_views << tile_bits */
  int32_t pre_tickets;                 /* blocks of the preprocess kernel */
  int32_t sort_tiles_cap;              /* onesweep tiles at full capacity */
  int32_t final_in_a;                  /* 1 if the sorted values end in saved.vals_a (always 1) */
  int64_t pair_capacity;
  /* saved workspace */
  size_t saved_bytes;
  size_t off_status;     /* B200sStatus */
  size_t off_rec;        /* [VV,N] 64-byte projected records */
  size_t off_vals_a;     /* [R_cap] u32 : sorted Gaussian indices after forward */
  size_t off_ranges;     /* [bins] uint2 (start,end) into vals_a */
  size_t off_final_T;    /* [VV,H,W] f32 */
  size_t off_n_contrib;  /* [VV,H,W] u32 */
  /* scratch workspace */
  size_t scratch_bytes;
  size_t off_keys_a, off_keys_b; /* [R_cap] u64 each */
  size_t off_vals_b;             /* [R_cap] u32 */
  size_t off_scan_state;         /* [pre_tickets] u64 exclusive pair offset of every projection CTA */
  size_t off_ticket_totals;      /* [pre_tickets] u32 tile total of every projection CTA */
  size_t off_scan_blocks;        /* [pre_tickets / 2048 + 1] u64 decoupled look-back words of the scan */
  size_t off_bin_info;           /* [pre_tickets * 256] uint2 (depth bits, packed rect) per Gaussian-view */
  size_t off_hist;               /* [8,256] u32 digit histograms -> exclusive bases */
  size_t off_lookback;           /* [2, sort_tiles_cap, 256] u64 onesweep look-back words */
  size_t off_counters;           /* [64] u32 ticket / tile counters */
  size_t off_grad_rec;           /* backward only: [VV,N] 48-byte gradient records (may alias keys) */
  /* BINNED mode (off_keys_a then holds the [R_cap] 8-byte (depth bits, Gaussian index) entries, keys_b / vals_b the
   * ping-pong and rank buffers of the bins too long for shared memory; lookback is empty) */
  int32_t sort_mode;             /* B200S_SORT_* the plan was made for */
  int32_t bin_sort_cap;          /* longest bin the segment sort keeps in shared memory */
  size_t off_bin_count;          /* [bins] u32 pairs per (view, tile) bin */
  size_t off_bin_cursor;         /* [bins] u32 next free position of every bin during the scatter */
  size_t off_long_list;          /* [4, bins] u32 bin ids per size class of the segment sort */
} B200sPlan;

/* Device status block (first bytes of the saved workspace). */
typedef struct B200sStatus {
  uint64_t num_pairs;     /* R: total (tile,Gaussian) pairs over all views of the call */
  uint32_t overflow;      /* nonzero: nothing downstream of stage A ran.  bit 0: R > pair_capacity (re-run with a larger capacity);
                             bit 1 (BINNED mode): more than 2^26 pairs, most of them in bins too long for shared memory
                             (re-run with a B200S_SORT_GLOBAL plan) */
  uint32_t num_visible;   /* Gaussian-view pairs that survived culling (Nv summed over views) */
  uint64_t tested;        /* optional flop accounting (filled when B200sOut.count_work != 0) */
  uint64_t blended;
  uint32_t max_tile_len;  /* longest per-tile list */
  uint32_t max_bin_len;   /* BINNED mode: longest bin, known right after the scan of the bin counts */
  uint32_t reserved[4];
} B200sStatus;

typedef struct B200sOut {
  float* color;      /* [VV,3,H,W] */
  float* depth;      /* [VV,H,W] or NULL (required when depth_mode != NONE) */
  int32_t* radii;    /* [VV,N] or NULL */
  int32_t count_work;/* != 0: accumulate tested/blended/max_tile_len into the status block */
  void* status_host; /* optional: DEVICE-ACCESSIBLE pointer to 16 bytes of mapped pinned host memory (b200s_host_alloc).
                        Stage A stores {u64 num_pairs, u32 overflow flags (B200sStatus.overflow), u32 nonzero marker (BINNED: 0x80000000 |
                        max_bin_len)} there directly from the kernel, so the host can read the pair count after an
                        event wait (or lazily, much later) without occupying a copy engine. */
  /* Optional loss-side fusion: the reference computes weight * mean((color - target)^2) (or the mean
   * absolute error) over the decoder's colour in PyTorch (src/loss/loss_mse.py:33-44) and the PSNR of the clipped images
   * (src/evaluation/metrics.py:11-19), and autograd turns the loss into dL/dcolor = 2 weight (color - target) / n in further
   * passes over the images.  With mse_target set, the compositing epilogue does all of it while the pixel is in registers. */
  const float* mse_target; /* [VV,3,H,W] ground-truth images, or NULL */
  float* mse_grad;         /* [VV,3,H,W] dL/dcolor of the loss, ready to be passed as B200sGradOut.dL_dcolor */
  float* mse_partials;     /* [VV, tiles, 2] per (view, tile): sum of squared (l1: absolute) error; sum of squared error of the
                              images clipped to [0, 1] (PSNR).  Fixed-order sums: the loss is reproducible */
  float mse_scale;         /* weight / n, n = number of colour values the mean runs over */
  int32_t mse_l1;          /* != 0: mean absolute error (l1_loss=True) */
} B200sOut;

typedef struct B200sGradOut { /* upstream gradients */
  const float* dL_dcolor; /* [VV,3,H,W] */
  const float* dL_ddepth; /* [VV,H,W] or NULL */
  const float* dL_dcolor_scale; /* optional DEVICE scalar every dL_dcolor value is multiplied by on load -- the upstream gradient of
                                   a loss whose dL/dcolor the forward epilogue produced (B200sOut.mse_grad): no scaling pass */
} B200sGradOut;

typedef struct B200sGradIn { /* all fp32, OVERWRITTEN (summed over the views of each scene) */
  float* dL_dmeans;       /* [B,N,3] */
  float* dL_dcovariances; /* same layout as B200sScene.covariances; lower triangle of 3x3 gets 0 */
  float* dL_dharmonics;   /* same layout as B200sScene.harmonics, or NULL */
  float* dL_dcolors;      /* [B,N,3] when colors_precomp was used, or NULL */
  float* dL_dopacities;   /* [B,N] */
  float* dL_dmeans2D;     /* [VV,N,3] screen-space mean gradients (extension API parity), or NULL */
  int32_t multicast;      /* != 0: dL_dmeans / dL_dcovariances / dL_dharmonics|dL_dcolors / dL_dopacities are NVLS MULTICAST
                             addresses of zero-initialised symmetric buffers (one replica per GPU of the process group): the
                             kernel ADDS its result with multimem.red, so that after a barrier every GPU holds the sum over
                             all ranks -- the preprocess backward and the gradient all-reduce of view-sharded training are one
                             pass, the reduction happens in the NVSwitch. */
  int32_t stages;         /* 0 = everything; else bit 0: compositing backward (fills the gradient records), bit 1: projection
                             backward.  Lets the caller run the projection backward in Gaussian CHUNKS and start reducing each
                             chunk's gradients across GPUs while the next chunk is computed. */
  int32_t chunk_begin;    /* projection backward: first 256-Gaussian chunk of every scene to process ... */
  int32_t chunk_count;    /* ... and how many (0 = all remaining) */
  int32_t chunk_stride;   /* with chunk_repeat > 1 the launch covers the chunks chunk_begin + r * chunk_stride + k, r < chunk_repeat, */
  int32_t chunk_repeat;   /* k < chunk_count (clipped to the scene): the j-th piece of EVERY rank's Gaussian range in one launch,
                             so that all ranks can pull their share of it (reduce-scatter) while the next piece is computed */
  float* dL_draw_head;    /* raw scenes: [B, raw_views, 37, raw_h*raw_w] gradient w.r.t. the head's channel planes ... */
  float* dL_draw_depth;   /* ... and [B, raw_views, raw_h*raw_w] w.r.t. the depth (the five tensor gradients above are then unused) */
} B200sGradIn;

/* Reduce-scatter building block over an NVSwitch domain: for each of the nseg (<= 16) segments -- seg_offset[i] floats into a
 * SYMMETRIC buffer, seg_count[i] floats, both multiples of 4 -- the calling rank pulls the sum over all ranks' replicas
 * (multimem.ld_reduce on the multicast address) and stores it into ITS OWN replica (local_base).  Nothing is sent to the
 * other ranks: each rank calls this for the part of the gradients it owns.  The caller orders it after a cross-rank barrier
 * (all replicas complete). */
int b200s_nvls_reduce_segments(void* multicast_base, void* local_base, const unsigned long long* seg_offset,
                               const unsigned long long* seg_count, int nseg, void* stream);

/* The same share of the reduce-scatter by ordinary peer loads instead of the switch's reduction: peer_bases[r] (HOST array of
 * `world` device pointers, 2 <= world <= 16) is rank r's replica of the symmetric buffer as mapped into the calling process
 * (peer_bases[rank] = the caller's own); for every vector of the segments the caller reads all replicas (own first, then
 * ranks rank+1, rank+2, ... mod world: a fixed summation order), and stores the sum at the same offset of local_base.  Moves
 * (world-1)/world of the buffer per GPU over NVLink where the multimem pull moves all of it. */
int b200s_p2p_reduce_segments(const void* const* peer_bases, int world, int rank, void* local_base, const unsigned long long* seg_offset,
                              const unsigned long long* seg_count, int nseg, void* stream);

/* Pure host function: fills the plan for the given dimensions.  No CUDA calls. */
int b200s_plan(const B200sDims* dims, B200sPlan* plan);

/* Stage A of forward: preprocess (cull, EWA projection, SH->RGB, tile rects) fused with the
 * tile-count scan and the emission of 64-bit (view|tile|depth) keys.  Writes B200sStatus
 * (num_pairs, overflow, num_visible).  Replaces preprocessCUDA + InclusiveSum + duplicateWithKeys
 * behind _C.rasterize_gaussians. */
int b200s_forward_bin(const B200sScene* scene, const B200sViews* views, const B200sPlan* plan, void* saved,
                      void* scratch, const B200sOut* out, void* stream);

/* Stage B of forward: onesweep radix sort of the pairs, tile ranges, 16x16-tile compositing of
 * colour (+ depth).  A no-op if the status block says overflow.  Replaces SortPairs +
 * identifyTileRanges + renderCUDA behind _C.rasterize_gaussians. */
int b200s_forward_render(const B200sScene* scene, const B200sViews* views, const B200sPlan* plan, void* saved,
                         void* scratch, const B200sOut* out, void* stream);

/* Backward: reverse-traversal compositing backward (warp-reduced gradients, one atomic per
 * gradient component per warp) followed by the preprocess backward that sums over the views of
 * each scene.  Replaces _C.rasterize_gaussians_backward AND the autograd of the per-view
 * replication (decoder_splatting_cuda.py:53-56). */
int b200s_backward(const B200sScene* scene, const B200sViews* views, const B200sPlan* plan, const void* saved,
                   void* scratch, const B200sOut* fwd_out, const B200sGradOut* gout, const B200sGradIn* gin, void* stream);

/* Stand-alone stable LSD radix sort of (u64 key, u32 value) pairs on the low `bits` bits -- the
 * same onesweep kernels forward uses, exported for the parity tests and the sort micro-benchmark.
 * keys_a/vals_a hold the input; the sorted output ends in keys_a/vals_a.  `tmp` needs
 * b200s_sort_tmp_bytes(n) bytes. */
size_t b200s_sort_tmp_bytes(int64_t n);
int b200s_sort_pairs(uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, int64_t n, int32_t bits,
                     void* tmp, void* stream);

/* Stand-alone BINNED-mode segment sort (the kernels forward uses, exported for the parity tests and micro-benchmarks):
 * `entries` holds n 8-byte (key: low word, value: high word) pairs grouped by bin -- bin b owns the bin_counts[b] pairs
 * after those of bins 0..b-1 -- in arbitrary order inside a bin.  Writes ranges_out[bins][2] = (start, end) of every bin
 * and vals_out[n] = the values of every bin in ascending (key, value) order.  `entries` is clobbered.  `tmp` needs
 * b200s_segment_sort_tmp_bytes(n, bins) bytes. */
size_t b200s_segment_sort_tmp_bytes(int64_t n, int32_t bins);
int b200s_segment_sort(const uint32_t* bin_counts, int32_t bins, uint64_t* entries, int64_t n, uint32_t* vals_out, uint32_t* ranges_out,
                       void* tmp, void* stream);

/* Optional per-stage device timing (CUDA events recorded on the caller's stream at stage boundaries;
 * off by default, process-wide).  bench.py uses it for the per-kernel roofline numbers. */
enum { B200S_STAGE_PRE_BIN = 0, B200S_STAGE_SORT_HIST = 1, B200S_STAGE_SORT_PASSES = 2, B200S_STAGE_RANGES = 3,
       B200S_STAGE_COMP_FWD = 4, B200S_STAGE_GRAD_ZERO = 5, B200S_STAGE_COMP_BWD = 6, B200S_STAGE_PRE_BWD = 7,
       B200S_STAGE_END = 8, B200S_STAGE_BIN_SORT = 9 /* BINNED mode: the per-bin segment sort */, B200S_NUM_STAGES = 10 };
void b200s_profile_enable(int on);
/* Synchronises on the recorded events, ADDS the elapsed milliseconds of every stage recorded since the
 * last read into ms_by_stage[B200S_NUM_STAGES], clears the recording.  Returns the number of stages seen. */
int b200s_profile_read(float* ms_by_stage);
/* Number of this library's kernels launched by the process so far (memsets not counted). */
long long b200s_kernel_launches(void);

/* Tuning knobs for A/B measurements (not part of the stable contract; a knob selects between variants with identical
 * results): which 0 = digit-histogram variant of the GLOBAL sort (0 ballots, 1 MATCH.ANY, 2 shared atomics), which 1 = its
 * ranking variant (0 ballots, 1 MATCH.ANY, 2 alternating), which 2 = block cap of the reduce-scatter pull kernels (0 = the
 * default per kernel), which 3 = load flavour of the peer-load pull (0 ld.global.cg, 1 ld.relaxed.sys). */
void b200s_debug_set(int which, int value);

/* Host-memory utility (not on the data path): mapped, portable pinned memory that kernels can write
 * (cudaHostAlloc).  Under unified addressing the returned pointer is valid on host and device. */
void* b200s_host_alloc(size_t bytes);
void b200s_host_free(void* p);

/* In-place sum over the `world` ranks of an NVSwitch domain of a SYMMETRIC buffer of n_floats floats (n_floats a
 * multiple of 4, the buffer 16-byte aligned) given by its NVLS MULTICAST address: rank `rank` reduces its 1/world slice
 * inside the switch (multimem.ld_reduce) and multicasts the sums back (multimem.st).  Replaces the NCCL all-reduce of
 * the per-Gaussian gradients in view-sharded training (the reference is single-GPU: src/main.py:145-147).  The caller
 * must order it on `stream` between two cross-rank barriers: all ranks' data complete before, all stores landed after. */
int b200s_nvls_allreduce(void* multicast_ptr, unsigned long long n_floats, int rank, int world, void* stream);

int b200s_abi_version(void);
int b200s_last_cuda_error(void);       /* cudaError_t of the last failed launch on this thread */
const char* b200s_build_info(void);    /* "sm_100a nvcc <ver> ..." */

#ifdef __cplusplus
}
#endif
#endif /* B200SPLAT_H */
