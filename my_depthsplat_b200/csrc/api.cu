// api.cu -- the extern "C" boundary declared in include/b200splat.h.  Host code only: argument
// checks, the workspace plan, and kernel launches on the caller's stream.  No allocation, no
// synchronisation, no global state (one thread-local error slot).
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "kernels.cuh"


using namespace b200s;

static thread_local int g_last_cuda_error = 0;

// ---- optional stage timing ---------------------------------------------------------------------------
// Process-wide (forward runs on the caller's thread, backward on autograd's): guarded by a mutex, and
// only touched at all when profiling is on.  The launch counter is a relaxed atomic.
static std::atomic<bool> g_prof{false};
static std::mutex g_prof_mu;
static cudaEvent_t g_ev[64];
static int g_ev_stage[64];
static int g_nev = 0;
static std::atomic<long long> g_launches{0};
namespace b200s {
void stage_mark(int stage, cudaStream_t stream) {
  if (!g_prof.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lock(g_prof_mu);
  if (g_nev >= 64) return;
  if (!g_ev[g_nev] && cudaEventCreate(&g_ev[g_nev]) != cudaSuccess) return;
  cudaEventRecord(g_ev[g_nev], stream);
  g_ev_stage[g_nev++] = stage;
}
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace b200s

static int fail(cudaError_t e) {
  g_last_cuda_error = (int)e;
  return B200S_ECUDA;
}
// SM count of the CURRENT device (cached per device id: one process may drive several GPUs, from one thread or many)
namespace b200s {
int device_sm_count() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  const int slot = dev & 63;
  int n = cache[slot].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[slot].store(n, std::memory_order_relaxed);
  }
  return n;
}
bool first_use_on_device(std::atomic<unsigned long long>& seen) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return true;
  const unsigned long long bit = 1ull << (dev & 63);
  if (seen.load(std::memory_order_relaxed) & bit) return false;
  seen.fetch_or(bit, std::memory_order_relaxed);
  return true;
}
}  // namespace b200s
static int sm_count() { return b200s::device_sm_count(); }
static size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
static int bits_for(int n) { int b = 0; while ((1 << b) < n) b++; return b; }

extern "C" {

int b200s_abi_version(void) { return B200S_ABI_VERSION; }
void b200s_profile_enable(int on) {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  g_prof.store(on != 0);
  g_nev = 0;
}
int b200s_profile_read(float* ms) {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  int seen = 0;
  if (g_nev > 0) cudaEventSynchronize(g_ev[g_nev - 1]);
  for (int i = 0; i + 1 < g_nev; i++) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, g_ev[i], g_ev[i + 1]) == cudaSuccess) { ms[g_ev_stage[i]] += t; seen++; }
  }
  g_nev = 0;
  return seen;
}
void* b200s_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) return nullptr;
  memset(p, 0, bytes);
  return p;
}
void b200s_host_free(void* p) { if (p) cudaFreeHost(p); }
void b200s_debug_set(int which, int value) { if (which >= 0 && which < 4) b200s::g_sort_knobs[which].store(value, std::memory_order_relaxed); }
long long b200s_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }
int b200s_last_cuda_error(void) { return g_last_cuda_error; }
const char* b200s_build_info(void) {
#define B200S_STR2(x) #x
#define B200S_STR(x) B200S_STR2(x)
  return "b200splat abi " B200S_STR(B200S_ABI_VERSION) " sm_100a nvcc " B200S_STR(__CUDACC_VER_MAJOR__) "." B200S_STR(__CUDACC_VER_MINOR__);
}

int b200s_nvls_allreduce(void* multicast_ptr, unsigned long long n_floats, int rank, int world, void* stream) {
  if (!multicast_ptr || (n_floats & 3ull) || (reinterpret_cast<uintptr_t>(multicast_ptr) & 15) || world <= 0 || rank < 0 || rank >= world)
    return B200S_EBADARG;
  const cudaError_t e = b200s::launch_nvls_allreduce(static_cast<float*>(multicast_ptr), n_floats, rank, world, sm_count(), static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? B200S_OK : fail(e);
}

int b200s_nvls_reduce_segments(void* multicast_base, void* local_base, const unsigned long long* seg_offset, const unsigned long long* seg_count,
                               int nseg, void* stream) {
  if (!multicast_base || !local_base || !seg_offset || !seg_count || nseg <= 0 || nseg > 16) return B200S_EBADARG;
  if ((reinterpret_cast<uintptr_t>(multicast_base) | reinterpret_cast<uintptr_t>(local_base)) & 15) return B200S_EBADARG;
  for (int i = 0; i < nseg; i++)
    if ((seg_offset[i] | seg_count[i]) & 3ull) return B200S_EBADARG;
  const cudaError_t e = b200s::launch_nvls_reduce_segments(static_cast<const float*>(multicast_base), static_cast<float*>(local_base), seg_offset,
                                                           seg_count, nseg, sm_count(), static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? B200S_OK : fail(e);
}

int b200s_p2p_reduce_segments(const void* const* peer_bases, int world, int rank, void* local_base, const unsigned long long* seg_offset,
                              const unsigned long long* seg_count, int nseg, void* stream) {
  if (!peer_bases || !local_base || !seg_offset || !seg_count || nseg <= 0 || nseg > 16 || world < 2 || world > 16 || rank < 0 || rank >= world)
    return B200S_EBADARG;
  if (reinterpret_cast<uintptr_t>(local_base) & 15) return B200S_EBADARG;
  for (int k = 0; k < world; k++)
    if (!peer_bases[k] || (reinterpret_cast<uintptr_t>(peer_bases[k]) & 15)) return B200S_EBADARG;
  for (int i = 0; i < nseg; i++)
    if ((seg_offset[i] | seg_count[i]) & 3ull) return B200S_EBADARG;
  const cudaError_t e = b200s::launch_p2p_reduce_segments(peer_bases, world, rank, static_cast<float*>(local_base), seg_offset, seg_count, nseg,
                                                          sm_count(), static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? B200S_OK : fail(e);
}

int b200s_plan(const B200sDims* d, B200sPlan* p) {
  if (!d || !p) return B200S_EBADARG;
  if (d->num_scenes <= 0 || d->num_gaussians <= 0 || d->num_views <= 0 || d->height <= 0 || d->width <= 0) return B200S_EBADARG;
  if (d->pair_capacity <= 0) return B200S_EBADARG;
  memset(p, 0, sizeof(*p));
  p->abi_version = B200S_ABI_VERSION;
  p->grid_x = (d->width + TILE_X - 1) / TILE_X;
  p->grid_y = (d->height + TILE_Y - 1) / TILE_Y;
  if (p->grid_x > 255 || p->grid_y > 255) return B200S_EBADARG;  // rect packing in the record
  p->tiles = p->grid_x * p->grid_y;
  p->tile_bits = bits_for(p->tiles);
  p->view_bits = bits_for(d->num_views);
  p->bins = d->num_views << p->tile_bits;
  p->sort_bits = 32 + p->tile_bits + p->view_bits;
  p->sort_passes = (p->sort_bits + 7) / 8;
  if (p->sort_passes > 8) return B200S_EBADARG;
  if (d->sort_mode != B200S_SORT_BINNED && d->sort_mode != B200S_SORT_GLOBAL) return B200S_EBADARG;
  p->sort_mode = d->sort_mode;
  p->bin_sort_cap = BIN_CAP_L;
  const bool binned = d->sort_mode == B200S_SORT_BINNED;
  // list positions (tile ranges, sort scatter) are 32-bit
  if (d->pair_capacity >= (1ll << 32) - 8192) return B200S_EBADARG;
  const long long chunks = (d->num_gaussians + PRE_THREADS - 1) / PRE_THREADS;
  if (chunks * d->num_views > 0x7fffffffll) return B200S_EBADARG;
  p->pre_tickets = (int)(chunks * d->num_views);
  p->sort_tiles_cap = sort_tiles_for(d->pair_capacity);
  p->final_in_a = 1;
  p->pair_capacity = d->pair_capacity;
  const size_t VN = (size_t)d->num_views * d->num_gaussians;
  const size_t VP = (size_t)d->num_views * d->height * d->width;
  const size_t R = (size_t)d->pair_capacity;
  size_t o = 0;
  p->off_status = o; o = align_up(o + sizeof(B200sStatus));
  p->off_rec = o; o = align_up(o + VN * sizeof(Rec));
  p->off_vals_a = o; o = align_up(o + R * 4);
  p->off_ranges = o; o = align_up(o + (size_t)p->bins * 8);
  p->off_final_T = o; o = align_up(o + VP * 4);
  p->off_n_contrib = o; o = align_up(o + VP * 4);
  p->saved_bytes = o;
  o = 0;
  // GLOBAL: key ping-pong + value pong.  BINNED: keys_a = the (depth bits, index) entries, keys_b / vals_b = the ping-pong and
  // rank buffers of bins too long for shared memory (same sizes, so the two modes need the same scratch)
  p->off_keys_a = o; o = align_up(o + R * 8);
  p->off_keys_b = o; o = align_up(o + R * 8);
  p->off_vals_b = o; o = align_up(o + R * 4);
  p->off_scan_state = o; o = align_up(o + (size_t)p->pre_tickets * 8);
  p->off_ticket_totals = o; o = align_up(o + (size_t)p->pre_tickets * 4);
  p->off_scan_blocks = o; o = align_up(o + ((size_t)p->pre_tickets / 2048 + 2) * 8);
  p->off_bin_info = o; o = align_up(o + (size_t)p->pre_tickets * PRE_THREADS * 8);
  p->off_hist = o; o = align_up(o + 8 * 256 * 4);
  p->off_lookback = o; o = align_up(o + (binned ? 0 : 2 * (size_t)p->sort_tiles_cap * 256 * 8));
  p->off_counters = o; o = align_up(o + CNT_WORDS * 4);
  p->off_bin_count = o; o = align_up(o + (binned ? (size_t)p->bins * 4 : 0));
  p->off_bin_cursor = o; o = align_up(o + (binned ? (size_t)p->bins * 4 : 0));
  p->off_long_list = o; o = align_up(o + (binned ? (size_t)BIN_CLASSES * p->bins * 4 : 0));
  const size_t fwd_bytes = o;
  p->off_grad_rec = 0;  // backward reuses the scratch from its start (keys are dead by then)
  const size_t bwd_bytes = align_up(VN * GREC_FLOATS * 4);
  p->scratch_bytes = fwd_bytes > bwd_bytes ? fwd_bytes : bwd_bytes;
  return B200S_OK;
}

static int check_common(const B200sScene* sc, const B200sViews* vw, const B200sPlan* pl, const void* saved, const void* scratch) {
  if (!sc || !vw || !pl || !saved || !scratch) return B200S_EBADARG;
  if (pl->abi_version != B200S_ABI_VERSION) return B200S_EBADARG;
  if (sc->raw_head) {  // raw scenes: the adapter runs inside the projection
    if (sc->means || sc->covariances || sc->opacities || sc->harmonics || sc->colors_precomp) return B200S_EBADARG;
    if (!sc->raw_depth || !sc->raw_image || !sc->raw_camera) return B200S_EBADARG;
    if (sc->raw_views <= 0 || sc->raw_h <= 0 || sc->raw_w <= 0 || (sc->raw_h * sc->raw_w) % PRE_THREADS != 0) return B200S_EBADARG;
    if ((long long)sc->raw_views * sc->raw_h * sc->raw_w != sc->num_gaussians || sc->sh_degree != 2 || sc->sh_coeffs != 9) return B200S_EBADARG;
    if (((uintptr_t)sc->raw_head | (uintptr_t)sc->raw_depth | (uintptr_t)sc->raw_image) & 15) return B200S_EBADARG;
  } else {
  if (!sc->means || !sc->covariances || !sc->opacities) return B200S_EBADARG;
  if (!sc->harmonics && !sc->colors_precomp) return B200S_EBADARG;
  }
  if (!sc->colors_precomp) {
    if (sc->sh_degree < 0 || sc->sh_degree > 3) return B200S_EBADARG;
    if (sc->sh_coeffs < (sc->sh_degree + 1) * (sc->sh_degree + 1) || sc->sh_coeffs > 16) return B200S_EBADARG;
  }
  if (!vw->scene_index || !vw->viewmatrix || !vw->projmatrix || !vw->campos || !vw->tanfov || !vw->background) return B200S_EBADARG;
  if (vw->depth_mode != B200S_DEPTH_NONE && !vw->depth_affine) return B200S_EBADARG;
  if (vw->depth_mode == B200S_DEPTH_LOG && !vw->depth_clamp) return B200S_EBADARG;
  if (vw->depth_mode < 0 || vw->depth_mode > 3) return B200S_EBADARG;
  if ((vw->width + TILE_X - 1) / TILE_X != pl->grid_x || (vw->height + TILE_Y - 1) / TILE_Y != pl->grid_y) return B200S_EBADARG;
  if (((long long)((sc->num_gaussians + PRE_THREADS - 1) / PRE_THREADS)) * vw->num_views != pl->pre_tickets) return B200S_EBADARG;
  if ((vw->num_views << pl->tile_bits) != pl->bins) return B200S_EBADARG;
  return B200S_OK;
}

int b200s_forward_bin(const B200sScene* sc, const B200sViews* vw, const B200sPlan* pl, void* saved, void* scratch, const B200sOut* out,
                      void* stream) {
  const int rc = check_common(sc, vw, pl, saved, scratch);
  if (rc != B200S_OK) return rc;
  const cudaError_t e = launch_preprocess_bin(*sc, *vw, *pl, (char*)saved, (char*)scratch, out, (cudaStream_t)stream);
  return e == cudaSuccess ? B200S_OK : fail(e);
}

int b200s_forward_render(const B200sScene* sc, const B200sViews* vw, const B200sPlan* pl, void* saved_, void* scratch_, const B200sOut* out,
                         void* stream_) {
  const int rc = check_common(sc, vw, pl, saved_, scratch_);
  if (rc != B200S_OK) return rc;
  if (!out || !out->color) return B200S_EBADARG;
  if (vw->depth_mode != B200S_DEPTH_NONE && !out->depth) return B200S_EBADARG;
  char* saved = (char*)saved_; char* scratch = (char*)scratch_;
  cudaStream_t stream = (cudaStream_t)stream_;
  B200sStatus* st = reinterpret_cast<B200sStatus*>(saved + pl->off_status);
  CountRef cnt{reinterpret_cast<const unsigned long long*>(&st->num_pairs), &st->overflow, 0ull};
  uint64_t* keys_a = reinterpret_cast<uint64_t*>(scratch + pl->off_keys_a);
  uint64_t* keys_b = reinterpret_cast<uint64_t*>(scratch + pl->off_keys_b);
  uint32_t* vals_a = reinterpret_cast<uint32_t*>(saved + pl->off_vals_a);
  uint32_t* vals_b = reinterpret_cast<uint32_t*>(scratch + pl->off_vals_b);
  uint2* ranges = reinterpret_cast<uint2*>(saved + pl->off_ranges);
  cudaError_t e;
  if (pl->sort_mode == B200S_SORT_BINNED) {
    uint32_t* counters = reinterpret_cast<uint32_t*>(scratch + pl->off_counters);
    BinSortWork w{reinterpret_cast<uint32_t*>(scratch + pl->off_long_list), counters + CNT_BIN_CLASS_COUNT, counters + CNT_BIN_CLASS_NEXT};
    e = launch_bin_sort(reinterpret_cast<uint2*>(keys_a), reinterpret_cast<uint2*>(keys_b), vals_b, ranges, vals_a, w, pl->bins, &st->overflow,
                        sm_count(), stream);
    if (e != cudaSuccess) return fail(e);
  } else {
    e = launch_sort(keys_a, vals_a, keys_b, vals_b, pl->sort_passes, pl->pair_capacity, cnt, reinterpret_cast<uint32_t*>(scratch + pl->off_hist),
                    reinterpret_cast<uint64_t*>(scratch + pl->off_lookback), reinterpret_cast<uint32_t*>(scratch + pl->off_counters), sm_count(), stream,
                    /*hist_ready=*/true);
    if (e != cudaSuccess) return fail(e);
    e = launch_tile_ranges(keys_a, cnt, ranges, pl->bins, pl->pair_capacity, sm_count(), stream);
    if (e != cudaSuccess) return fail(e);
  }
  CompArgs a;
  memset(&a, 0, sizeof(a));
  a.N = sc->num_gaussians; a.H = vw->height; a.W = vw->width; a.grid_x = pl->grid_x; a.tile_bits = pl->tile_bits;
  a.rec = reinterpret_cast<const Rec*>(saved + pl->off_rec);
  a.vals = vals_a; a.ranges = ranges; a.bg = vw->background; a.overflow = &st->overflow;
  a.color = out->color; a.depth = out->depth;
  a.final_T = reinterpret_cast<float*>(saved + pl->off_final_T);
  a.n_contrib = reinterpret_cast<uint32_t*>(saved + pl->off_n_contrib);
  a.status = st;
  if (out->mse_target) {
    if (!out->mse_grad || !out->mse_partials) return B200S_EBADARG;
    a.mse_target = out->mse_target; a.mse_grad = out->mse_grad; a.mse_partials = out->mse_partials;
    a.mse_scale = out->mse_scale; a.mse_l1 = out->mse_l1;
  }
  e = launch_composite_fwd(a, pl->tiles, vw->num_views, vw->depth_mode != B200S_DEPTH_NONE, out->count_work != 0, stream);
  stage_mark(B200S_STAGE_END, stream);
  return e == cudaSuccess ? B200S_OK : fail(e);
}

int b200s_backward(const B200sScene* sc, const B200sViews* vw, const B200sPlan* pl, const void* saved_, void* scratch_, const B200sOut* fwd_out,
                   const B200sGradOut* gout, const B200sGradIn* gin, void* stream_) {
  const int rc = check_common(sc, vw, pl, saved_, scratch_);
  if (rc != B200S_OK) return rc;
  if (!gout || !gin || !gout->dL_dcolor) return B200S_EBADARG;
  if (vw->depth_mode != B200S_DEPTH_NONE && !gout->dL_ddepth) return B200S_EBADARG;
  if (sc->raw_head) {
    if (!gin->dL_draw_head || !gin->dL_draw_depth) return B200S_EBADARG;
  } else {
    if (!gin->dL_dmeans || !gin->dL_dcovariances || !gin->dL_dopacities) return B200S_EBADARG;
    if (sc->colors_precomp ? !gin->dL_dcolors : !gin->dL_dharmonics) return B200S_EBADARG;
  }
  (void)fwd_out;
  const char* saved = (const char*)saved_; char* scratch = (char*)scratch_;
  cudaStream_t stream = (cudaStream_t)stream_;
  const B200sStatus* st = reinterpret_cast<const B200sStatus*>(saved + pl->off_status);
  float* grad_rec = reinterpret_cast<float*>(scratch + pl->off_grad_rec);
  const int stages = gin->stages ? gin->stages : 3;
  cudaError_t e = cudaSuccess;
  if (stages & 1) {
  stage_mark(B200S_STAGE_GRAD_ZERO, stream);
  e = cudaMemsetAsync(grad_rec, 0, (size_t)vw->num_views * sc->num_gaussians * GREC_FLOATS * sizeof(float), stream);
  if (e != cudaSuccess) return fail(e);
  CompArgs a;
  memset(&a, 0, sizeof(a));
  a.N = sc->num_gaussians; a.H = vw->height; a.W = vw->width; a.grid_x = pl->grid_x; a.tile_bits = pl->tile_bits;
  a.rec = reinterpret_cast<const Rec*>(saved + pl->off_rec);
  a.vals = reinterpret_cast<const uint32_t*>(saved + pl->off_vals_a);
  a.ranges = reinterpret_cast<const uint2*>(saved + pl->off_ranges);
  a.bg = vw->background; a.overflow = &st->overflow;
  a.final_T = const_cast<float*>(reinterpret_cast<const float*>(saved + pl->off_final_T));
  a.n_contrib = const_cast<uint32_t*>(reinterpret_cast<const uint32_t*>(saved + pl->off_n_contrib));
  a.dL_dcolor = gout->dL_dcolor; a.dL_ddepth = gout->dL_ddepth; a.dpix_scale = gout->dL_dcolor_scale; a.grad_rec = grad_rec;
  e = launch_composite_bwd(a, pl->tiles, vw->num_views, vw->depth_mode != B200S_DEPTH_NONE, stream);
  if (e != cudaSuccess) return fail(e);
  }
  if (!(stages & 2)) { stage_mark(B200S_STAGE_END, stream); return B200S_OK; }
  e = launch_preprocess_bwd(*sc, *vw, *pl, saved, scratch, *gin, stream);
  stage_mark(B200S_STAGE_END, stream);
  return e == cudaSuccess ? B200S_OK : fail(e);
}

size_t b200s_segment_sort_tmp_bytes(int64_t n, int32_t bins) {
  if (n < 0 || bins <= 0) return 0;
  return align_up((size_t)n * 8) + align_up((size_t)n * 4) + align_up((size_t)bins * 4) + align_up((size_t)BIN_CLASSES * bins * 4) +
         align_up(sizeof(B200sStatus)) + align_up(CNT_WORDS * 4);
}

int b200s_segment_sort(const uint32_t* bin_counts, int32_t bins, uint64_t* entries, int64_t n, uint32_t* vals_out, uint32_t* ranges_out, void* tmp,
                       void* stream_) {
  if (!bin_counts || bins <= 0 || !entries || n < 0 || !vals_out || !ranges_out || !tmp || n >= (1ll << 32) - 8192) return B200S_EBADARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  char* t = (char*)tmp;
  uint2* entries_tmp = reinterpret_cast<uint2*>(t); t += align_up((size_t)n * 8);
  uint32_t* rank_tmp = reinterpret_cast<uint32_t*>(t); t += align_up((size_t)n * 4);
  uint32_t* cursor = reinterpret_cast<uint32_t*>(t); t += align_up((size_t)bins * 4);
  uint32_t* lists = reinterpret_cast<uint32_t*>(t); t += align_up((size_t)BIN_CLASSES * bins * 4);
  B200sStatus* st = reinterpret_cast<B200sStatus*>(t); t += align_up(sizeof(B200sStatus));
  uint32_t* counters = reinterpret_cast<uint32_t*>(t);
  BinSortWork w{lists, counters + CNT_BIN_CLASS_COUNT, counters + CNT_BIN_CLASS_NEXT};
  cudaError_t e = launch_bin_scan(bin_counts, bins, reinterpret_cast<uint2*>(ranges_out), cursor, w, st, (unsigned long long)n, nullptr, stream);
  if (e != cudaSuccess) return fail(e);
  e = launch_bin_sort(reinterpret_cast<uint2*>(entries), entries_tmp, rank_tmp, reinterpret_cast<const uint2*>(ranges_out), vals_out, w, bins,
                      &st->overflow, sm_count(), stream);
  return e == cudaSuccess ? B200S_OK : fail(e);
}

size_t b200s_sort_tmp_bytes(int64_t n) { return sort_tmp_bytes(n); }

int b200s_sort_pairs(uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, int64_t n, int32_t bits, void* tmp, void* stream) {
  if (!keys_a || !vals_a || !keys_b || !vals_b || !tmp || n < 0 || bits <= 0 || bits > 64 || n >= (1ll << 32) - 8192) return B200S_EBADARG;
  if (n == 0) return B200S_OK;
  const int passes = (bits + 7) / 8;
  uint32_t* hist = reinterpret_cast<uint32_t*>(tmp);
  uint64_t* lookback = reinterpret_cast<uint64_t*>(hist + 8 * 256);
  uint32_t* counters = reinterpret_cast<uint32_t*>(lookback + 2 * (size_t)sort_tiles_for(n) * 256);
  // the stand-alone entry point takes its input in A: an odd pass count needs it in B first
  cudaStream_t s = (cudaStream_t)stream;
  if (passes % 2) {
    cudaError_t e = cudaMemcpyAsync(keys_b, keys_a, (size_t)n * 8, cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) return fail(e);
    e = cudaMemcpyAsync(vals_b, vals_a, (size_t)n * 4, cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) return fail(e);
  }
  CountRef cnt{nullptr, nullptr, (unsigned long long)n};
  const cudaError_t e = launch_sort(keys_a, vals_a, keys_b, vals_b, passes, n, cnt, hist, lookback, counters, sm_count(), s, /*hist_ready=*/false);
  return e == cudaSuccess ? B200S_OK : fail(e);
}

}  // extern "C"
