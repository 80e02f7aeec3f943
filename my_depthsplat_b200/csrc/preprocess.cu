// preprocess.cu -- stage A of forward.  Replaces the extension's preprocessCUDA,
// cub::DeviceScan::InclusiveSum (+ its blocking D2H read of the total) and duplicateWithKeys
//  with three kernels and no host synchronisation:
//
//   project_kernel   one CTA per (256-Gaussian chunk, view), view-fastest so that the CTAs that read
//                    the same chunk for different views run together and share it through L2:
//                    coalesced float4 staging of means / covariances / SH / opacities in the caller's
//                    own tensor layouts -> cull, projection, EWA cov2D, conic, radius, tile rect,
//                    SH->RGB, depth colour -> 64-byte records out (coalesced through shared memory),
//                    8-byte (depth bits, packed rect) binning word per Gaussian, tile total per CTA.
//   scan_kernel      single-pass chained scan (decoupled look-back) of the per-CTA tile totals; writes
//                    the pair count and the overflow flag to the device status block AND straight into
//                    mapped pinned host memory.
//   emit_kernel      64-bit (view | tile | depth-bits) keys and Gaussian-index values at the scanned
//                    offsets: per thread for small rects, one warp per large-rect Gaussian otherwise.
//
// (An earlier version chained the scan through the projection CTAs themselves; with ~46k CTAs in flight
// order the look-back latency, not HBM, bounded the kernel -- 28 % of warp samples sat at the barrier
// behind the look-back, profiles/r1b.  Scanning 46k totals separately costs one 8-byte word per
// Gaussian-view of extra traffic and removes the wait.)
//
// HBM-bound: per (view, Gaussian) 148 B read (L2-shared across views) + 72 B written, + 8 B read and
// 12 B per pair written by the emission.
#include "adapter.cuh"
#include "kernels.cuh"

namespace b200s {

struct PreArgs {
  int N, VV, H, W, grid_x, grid_y, tile_bits, chunks, n_tickets;
  unsigned long long pair_capacity;
  int fpg;          // floats per Gaussian staged: 3 + cov + colour + 1
  int cov_floats;   // 9 or 6
  int col_floats;   // 3*d_sh or 3
  int col_stride;   // padded to odd
  Rec* rec;
  uint64_t* keys;
  uint32_t* vals;
  uint2* bin_info;                  // [VV*chunks*256] (depth bits, packed rect) in ticket order
  uint32_t* ticket_totals;          // [n_tickets]
  unsigned long long* ticket_offsets;  // [n_tickets] exclusive scan of the totals
  uint64_t* scan_blocks;            // [scan blocks] decoupled look-back words
  uint32_t* hist;                   // [8*256] digit histograms of the sort (accumulated by emit_kernel)
  int sort_passes;
  uint32_t* counters;
  B200sStatus* status;
  int32_t* radii;
  unsigned long long* status_host;  // mapped pinned host memory or NULL
  // BINNED sort mode
  uint32_t* bin_count;              // [bins] pairs per (view, tile) bin
  uint32_t* bin_cursor;             // [bins] next free list position of every bin
  uint2* entries;                   // [R_cap] (depth bits, Gaussian index), grouped by bin
};

constexpr uint64_t SCAN_FLAG_AGG = 1ull << 62, SCAN_FLAG_PREFIX = 2ull << 62, SCAN_VALUE_MASK = (1ull << 62) - 1;
constexpr int WARP_EMIT_THRESHOLD = 12;

// cooperative global -> shared copy of `count` floats; float4 path when both sides are 16B aligned
__device__ __forceinline__ void stage_in(float* dst, const float* __restrict__ src, int count) {
  const int tid = threadIdx.x;
  if ((((uintptr_t)src) & 15) == 0 && (count & 3) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = tid; i < (count >> 2); i += PRE_THREADS) d4[i] = __ldg(s4 + i);
  } else {
    for (int i = tid; i < count; i += PRE_THREADS) dst[i] = __ldg(src + i);
  }
}
// same, into rows padded from k to `stride` floats
__device__ __forceinline__ void stage_in_padded(float* dst, const float* __restrict__ src, int count, int k, int stride) {
  if (k == stride) { stage_in(dst, src, count); return; }
  for (int i = threadIdx.x; i < count; i += PRE_THREADS) { const int g = i / k; dst[g * stride + (i - g * k)] = __ldg(src + i); }
}

__device__ __forceinline__ float eval_sh_channel(int deg, const float* sh, int kstride, float x, float y, float z) {
#define S(k) sh[(k) * kstride]
  float r = SH_C0 * S(0);
  if (deg > 0) {
    r = r - SH_C1 * y * S(1) + SH_C1 * z * S(2) - SH_C1 * x * S(3);
    if (deg > 1) {
      const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
      r = r + SH_C2[0] * xy * S(4) + SH_C2[1] * yz * S(5) + SH_C2[2] * (2.0f * zz - xx - yy) * S(6) +
          SH_C2[3] * xz * S(7) + SH_C2[4] * (xx - yy) * S(8);
      if (deg > 2) {
        r = r + SH_C3[0] * y * (3.0f * xx - yy) * S(9) + SH_C3[1] * xy * z * S(10) +
            SH_C3[2] * y * (4.0f * zz - xx - yy) * S(11) + SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * S(12) +
            SH_C3[4] * x * (4.0f * zz - xx - yy) * S(13) + SH_C3[5] * z * (xx - yy) * S(14) +
            SH_C3[6] * x * (xx - 3.0f * yy) * S(15);
      }
    }
  }
#undef S
  return r;
}

// MODE 1 (specialised) = the layout DepthSplat hands over (3x3 covariances, SH [N,3,9] channel-major, degree 2): every
// stride is a compile-time constant.  MODE 0 (generic) takes them from the arguments.  MODE 2 (raw) = the encoder head's
// raw channel planes: the Gaussian adapter runs here, on the staged chunk (adapter.cuh), and the specialised layout is
// what it leaves in shared memory / registers -- the rest of the kernel is the same.
template <int MODE>
__global__ void __launch_bounds__(PRE_THREADS, 4) project_kernel(const B200sScene sc, const B200sViews vw, const PreArgs a) {
  constexpr bool SPECIALISED = MODE != 0;
  constexpr bool RAW = MODE == 2;
  extern __shared__ __align__(16) float smem[];
  __shared__ ViewParams s_vp[VIEW_GROUP];
  __shared__ uint32_t s_tot[VIEW_GROUP];
  __shared__ int s_vlo, s_vhi;
  __shared__ __align__(8) uint64_t s_bar;

  const int cov_floats = SPECIALISED ? 9 : a.cov_floats;
  const int col_floats = SPECIALISED ? 27 : a.col_floats;
  const int col_stride = SPECIALISED ? 27 : a.col_stride;
  const int cstride = SPECIALISED ? 9 : ((sc.sh_layout == B200S_SH_CHANNEL_MAJOR) ? sc.sh_coeffs : 1);
  const int kstride = SPECIALISED ? 1 : ((sc.sh_layout == B200S_SH_CHANNEL_MAJOR) ? 1 : 3);
  const int deg = SPECIALISED ? 2 : sc.sh_degree;
  const bool precomp = SPECIALISED ? false : (sc.colors_precomp != nullptr);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunk = blockIdx.x % a.chunks, scene = blockIdx.x / a.chunks;
  const int i0 = chunk * PRE_THREADS;
  const int n = min(PRE_THREADS, a.N - i0);
  const long long g0 = (long long)scene * a.N + i0;

  // ---- stage the Gaussian chunk ONCE for all the views of its scene ------------------------------
  float* s_mean = smem;                                  // [256*3]
  float* s_cov = s_mean + PRE_THREADS * 3;               // [256*cov_floats]
  float* s_op = s_cov + PRE_THREADS * cov_floats;        // [256]
  float* s_col = s_op + PRE_THREADS;                     // [256*col_stride]
  if (tid == 0) { s_vlo = 0x7fffffff; s_vhi = -1; }      // ordered before the range search by the barriers of the staging below
  float mraw[3] = {0.f, 0.f, 0.f}, craw[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, opac = 0.f;
  if (RAW) {
    // ---- stage the 41 raw planes of the chunk (37 head channels, depth, 3 image channels), run the adapter -----------
    __shared__ RawCam s_cam;
    const int hw = sc.raw_h * sc.raw_w;
    const int cv = i0 / hw, p0 = i0 - cv * hw;           // context view and first pixel of the chunk (hw % 256 == 0)
    const size_t sv = (size_t)scene * sc.raw_views + cv;
    if (tid == 0) {
      mbar_init(&s_bar, 1);
      mbar_expect_tx(&s_bar, (uint32_t)(RAW_PLANES * PRE_THREADS * 4));
      for (int ch = 0; ch < RAW_CH; ch++) tma_bulk_load(smem + ch * PRE_THREADS, sc.raw_head + (sv * RAW_CH + ch) * hw + p0, PRE_THREADS * 4, &s_bar);
      tma_bulk_load(smem + RAW_CH * PRE_THREADS, sc.raw_depth + sv * hw + p0, PRE_THREADS * 4, &s_bar);
      for (int ch = 0; ch < 3; ch++) tma_bulk_load(smem + (RAW_CH + 1 + ch) * PRE_THREADS, sc.raw_image + (sv * 3 + ch) * hw + p0, PRE_THREADS * 4, &s_bar);
    }
    if (tid < RAW_CAM_FLOATS) reinterpret_cast<float*>(&s_cam)[tid] = __ldg(sc.raw_camera + sv * RAW_CAM_FLOATS + tid);
    __syncthreads();  // the barrier is initialised (and the camera block staged) before anyone waits
    mbar_wait(&s_bar, 0);
    float h[RAW_CH], img[3];
#pragma unroll
    for (int ch = 0; ch < RAW_CH; ch++) h[ch] = smem[ch * PRE_THREADS + tid];
    const float depth = smem[RAW_CH * PRE_THREADS + tid];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) img[ch] = smem[(RAW_CH + 1 + ch) * PRE_THREADS + tid];
    __syncthreads();  // every raw value is in registers: the specialised layout may now overwrite the planes
    const int p = p0 + tid;
    Cooked g;
    cook(h, depth, s_cam, p % sc.raw_w, p / sc.raw_w, sc.raw_w, sc.raw_h, sc.raw_scale_min, sc.raw_scale_max, g);
    mraw[0] = g.mean[0]; mraw[1] = g.mean[1]; mraw[2] = g.mean[2];
#pragma unroll
    for (int k = 0; k < 6; k++) craw[k] = g.cov[k];
    opac = g.op;
#pragma unroll
    for (int ch = 0; ch < 3; ch++) cook_sh(h + 10 + 9 * ch, img[ch], s_cam, s_col + tid * 27 + 9 * ch);
    if (sc.raw_cooked_out) {  // tests: the world-space Gaussian as built here
      float* o = sc.raw_cooked_out + (g0 + tid) * 40;
      o[0] = g.mean[0]; o[1] = g.mean[1]; o[2] = g.mean[2];
      o[3] = g.cov[0]; o[4] = g.cov[1]; o[5] = g.cov[2]; o[6] = g.cov[1]; o[7] = g.cov[3]; o[8] = g.cov[4]; o[9] = g.cov[2]; o[10] = g.cov[4]; o[11] = g.cov[5];
      o[12] = g.op;
      for (int k = 0; k < 27; k++) o[13 + k] = s_col[tid * 27 + k];
    }
  } else {
  const float* col_src = precomp ? sc.colors_precomp : sc.harmonics;
  {
    const float* g_mean = sc.means + g0 * 3; const float* g_cov = sc.covariances + g0 * cov_floats;
    const float* g_op = sc.opacities + g0; const float* g_col = col_src + g0 * col_floats;
    // contiguous ranges with 16-byte aligned ends go through the TMA bulk copy engine (four copies, one barrier);
    // ragged / misaligned chunks and padded colour rows fall back to cooperative loads
    const bool tma = col_floats == col_stride && tma_ok(g_mean, n * 12) && tma_ok(g_cov, n * cov_floats * 4) && tma_ok(g_op, n * 4) &&
                     tma_ok(g_col, n * col_floats * 4);
    if (tma) {
      if (tid == 0) {
        mbar_init(&s_bar, 1);
        mbar_expect_tx(&s_bar, (uint32_t)(n * 4 * (3 + cov_floats + 1 + col_floats)));
        tma_bulk_load(s_mean, g_mean, n * 12, &s_bar);
        tma_bulk_load(s_cov, g_cov, n * cov_floats * 4, &s_bar);
        tma_bulk_load(s_op, g_op, n * 4, &s_bar);
        tma_bulk_load(s_col, g_col, n * col_floats * 4, &s_bar);
      }
      __syncthreads();  // the barrier is initialised before anyone waits on it
      mbar_wait(&s_bar, 0);
    } else {
      stage_in(s_mean, g_mean, n * 3);
      stage_in(s_cov, g_cov, n * cov_floats);
      stage_in(s_op, g_op, n);
      stage_in_padded(s_col, g_col, n * col_floats, col_floats, col_stride);
      __syncthreads();
    }
  }
  if (tid < n) {
    mraw[0] = s_mean[tid * 3]; mraw[1] = s_mean[tid * 3 + 1]; mraw[2] = s_mean[tid * 3 + 2];
    const float* cp = s_cov + tid * cov_floats;
    if (cov_floats == 6) {
#pragma unroll
      for (int k = 0; k < 6; k++) craw[k] = cp[k];
    } else {
      craw[0] = cp[0]; craw[1] = cp[1]; craw[2] = cp[2]; craw[3] = cp[4]; craw[4] = cp[5]; craw[5] = cp[8];
    }
    opac = s_op[tid];
  }
  }
  const float* cs = s_col + tid * col_stride;

  // ---- the views of the scene, VIEW_GROUP camera blocks at a time; no barrier inside a group: the warps run free ----
  scene_view_range(vw, a.VV, scene, &s_vlo, &s_vhi);
  __syncthreads();
  const int vlo = s_vlo, vhi = s_vhi;
  for (int v0 = vlo; v0 <= vhi; v0 += VIEW_GROUP) {
  const int vcount = min(VIEW_GROUP, vhi + 1 - v0);
  if (v0 != vlo) __syncthreads();  // the previous group's readers of s_vp / s_tot are done
  load_view_group(s_vp, vw, v0, vcount, a.H, a.W);
  if (tid < VIEW_GROUP) s_tot[tid] = 0u;
  __syncthreads();
  for (int vj = 0; vj < vcount; vj++) {
    const ViewParams& vp = s_vp[vj];
    if (vp.scene != scene) continue;  // block-uniform
    const int view = v0 + vj;
    const int ticket = chunk * a.VV + view;

    // ---- per-Gaussian projection -------------------------------------------------------------
    Rec out;
    out.q0 = make_float4(0.f, 0.f, -1e30f, -1e30f); out.q1 = make_float4(0.f, 0.f, 0.f, 0.f); out.q2 = out.q1; out.q3 = out.q1;
    uint32_t tiles = 0;
    int rx0 = 0, ry0 = 0, rx1 = 0, ry1 = 0;
    float depth = 0.f;
    if (tid < n) {
      const float m[3] = {__fmul_rn(mraw[0], vp.s), __fmul_rn(mraw[1], vp.s), __fmul_rn(mraw[2], vp.s)};
      const float phx = xform_row(vp.proj, 0, m[0], m[1], m[2]);
      const float phy = xform_row(vp.proj, 1, m[0], m[1], m[2]);
      const float phw = xform_row(vp.proj, 3, m[0], m[1], m[2]);
      const float pw = __frcp_rn(__fadd_rn(phw, 0.0000001f));
      const float ppx = __fmul_rn(phx, pw), ppy = __fmul_rn(phy, pw);
      const float pvz = xform_row(vp.view, 2, m[0], m[1], m[2]);
      if (pvz > NEAR_CULL) {
        float c6[6];
#pragma unroll
        for (int k = 0; k < 6; k++) c6[k] = __fmul_rn(craw[k], vp.s2);
        Cov2D q;
        compute_cov2d(m, c6, vp, q);
        const float det = __fsub_rn(__fmul_rn(q.a, q.c), __fmul_rn(q.b, q.b));
        if (det != 0.0f) {
          const float det_inv = __frcp_rn(det);
          const float mid = __fmul_rn(0.5f, __fadd_rn(q.a, q.c));
          const float disc = __fsqrt_rn(fmaxf(LAMBDA_FLOOR, __fsub_rn(__fmul_rn(mid, mid), det)));
          const float lam = fmaxf(__fadd_rn(mid, disc), __fsub_rn(mid, disc));
          const float radf = ceilf(__fmul_rn(3.f, __fsqrt_rn(lam)));
          const int radius = (int)radf;
          const float px = ndc2pix(ppx, a.W), py = ndc2pix(ppy, a.H);
          get_rect(px, py, radius, a.grid_x, a.grid_y, rx0, ry0, rx1, ry1);
          tiles = (uint32_t)((rx1 - rx0) * (ry1 - ry0));
          if (tiles > 0) {
            depth = pvz;
            float rgb[3];
            uint32_t flags = 0;
            if (precomp) {
              rgb[0] = cs[0]; rgb[1] = cs[1]; rgb[2] = cs[2];
            } else {
              float dx = m[0] - vp.campos[0], dy = m[1] - vp.campos[1], dz = m[2] - vp.campos[2];
              const float len = __fsqrt_rn(dot3c(dx, dx, dy, dy, dz, dz));
              dx = __fdiv_rn(dx, len); dy = __fdiv_rn(dy, len); dz = __fdiv_rn(dz, len);
#pragma unroll
              for (int ch = 0; ch < 3; ch++) {
                float r = eval_sh_channel(deg, cs + ch * cstride, kstride, dx, dy, dz) + 0.5f;
                if (r < 0.f) flags |= (1u << ch);
                rgb[ch] = fmaxf(r, 0.f);
              }
            }
            float zc = 0.f;
            if (vw.depth_mode != B200S_DEPTH_NONE) {
              const float z = __fadd_rn(__fmaf_rn(vp.daff[2], mraw[2], __fmaf_rn(vp.daff[0], mraw[0], __fmul_rn(vp.daff[1], mraw[1]))), vp.daff[3]);
              if (vw.depth_mode == B200S_DEPTH_Z) zc = z;
              else if (vw.depth_mode == B200S_DEPTH_DISPARITY) zc = __frcp_rn(z);
              else zc = logf(fmaxf(fminf(z, vp.dnear), vp.dfar));
            }
            // conservative half-extents of the region where alpha can reach 1/255
            float ex = -1e30f, ey = -1e30f;  // never reaches alpha >= 1/255: fails every sub-tile test
            const float o255 = 255.0f * opac;
            const float cA = __fmul_rn(q.c, det_inv), cB = __fmul_rn(-q.b, det_inv), cC = __fmul_rn(q.a, det_inv);
            if (o255 >= 1.0f) {
              // bounding box of {A dx^2 + 2 B dx dy + C dy^2 <= tau2} for the STORED conic -- the numbers the blend evaluates
              // (for elongated footprints det cancels, and the float conic is no longer the inverse of the float cov2D to
              // within the margin); a conic that is not positive definite in floats is never culled
              const float tau2 = 2.0f * logf(o255);
              const float dd = cA * cC - cB * cB;
              if (dd > 0.f && cA > 0.f && cC > 0.f) {
                ex = sqrtf(tau2 * cC / dd) * 1.01f + 0.1f;
                ey = sqrtf(tau2 * cA / dd) * 1.01f + 0.1f;
              } else {
                ex = 1e30f; ey = 1e30f;
              }
            }
            out.q0 = make_float4(px, py, ex, ey);
            out.q1 = make_float4(cA, cB, cC, opac);
            out.q2 = make_float4(rgb[0], rgb[1], rgb[2], zc);
            const uint32_t rect = (uint32_t)rx0 | ((uint32_t)ry0 << 8) | ((uint32_t)rx1 << 16) | ((uint32_t)ry1 << 24);
            out.q3 = make_float4(depth, __int_as_float(radius), __uint_as_float(rect), __uint_as_float(flags));
          }
        }
      }
      if (a.radii) a.radii[(long long)view * a.N + i0 + tid] = tiles > 0 ? __float_as_int(out.q3.y) : 0;
    }

    // ---- binning word, tile total of the (chunk, view), records out -------------------------------
    {
      const uint32_t rect = tiles ? ((uint32_t)rx0 | ((uint32_t)ry0 << 8) | ((uint32_t)rx1 << 16) | ((uint32_t)ry1 << 24)) : 0u;
      a.bin_info[(size_t)ticket * PRE_THREADS + tid] = make_uint2(__float_as_uint(depth), rect);
    }
    uint32_t sum = tiles;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    const uint32_t vis_mask = __ballot_sync(0xffffffffu, tiles > 0);
    if (lane == 0 && vis_mask) { atomicAdd(&s_tot[vj], sum); atomicAdd(&a.status->num_visible, (uint32_t)__popc(vis_mask)); }
    // the thread's own 64-byte record, four 16-byte stores (two full sectors per thread; L2 merges the halves)
    if (tid < n) {
      float4* dst = reinterpret_cast<float4*>(a.rec + (long long)view * a.N + i0 + tid);
      dst[0] = out.q0; dst[1] = out.q1; dst[2] = out.q2; dst[3] = out.q3;
    }
  }
  __syncthreads();  // the group's tile totals are complete
  if (tid < vcount && s_vp[tid].scene == scene) a.ticket_totals[chunk * a.VV + v0 + tid] = s_tot[tid];
  }
}

// -------------------------------------------------------------------------------------------------
// Single-pass chained scan of the per-ticket totals (decoupled look-back across scan blocks).
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS) scan_kernel(const PreArgs a) {
  __shared__ unsigned long long s_wsum[SCAN_THREADS / 32];
  __shared__ unsigned long long s_base;
  __shared__ uint32_t s_blk;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_blk = atomicAdd(&a.counters[CNT_PRE_TICKET], 1u);  // dynamic id: predecessors are running
  __syncthreads();
  const int blk = (int)s_blk;
  const int first = blk * SCAN_TILE + tid * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS];
  unsigned long long tsum = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) { v[i] = first + i < a.n_tickets ? a.ticket_totals[first + i] : 0u; tsum += v[i]; }
  unsigned long long incl = tsum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
  if (lane == 31) s_wsum[warp] = incl;
  __syncthreads();
  unsigned long long wexcl = 0, btotal = 0;
#pragma unroll
  for (int w = 0; w < SCAN_THREADS / 32; w++) { const unsigned long long t = s_wsum[w]; if (w < warp) wexcl += t; btotal += t; }
  if (warp == 0) {
    if (lane == 0) st_volatile_u64(&a.scan_blocks[blk], (blk == 0 ? SCAN_FLAG_PREFIX : SCAN_FLAG_AGG) | btotal);
    unsigned long long excl = 0;
    for (int idx = blk - 1; idx >= 0; idx -= 32) {
      const int t = idx - lane;
      uint64_t w = SCAN_FLAG_PREFIX;  // blocks < 0 count as an empty prefix
      if (t >= 0) { do { w = ld_volatile_u64(&a.scan_blocks[t]); } while ((w >> 62) == 0); }
      const uint32_t pmask = __ballot_sync(0xffffffffu, (w >> 62) == 2);
      const int firstp = pmask ? (__ffs(pmask) - 1) : 32;
      unsigned long long x = (lane <= firstp) ? (w & SCAN_VALUE_MASK) : 0ull;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
      excl += x;
      if (pmask) break;
    }
    if (lane == 0) {
      if (blk != 0) st_volatile_u64(&a.scan_blocks[blk], SCAN_FLAG_PREFIX | (excl + btotal));
      s_base = excl;
      if ((blk + 1) * SCAN_TILE >= a.n_tickets) {  // last block: the grand total
        const unsigned long long total = excl + btotal;
        a.status->num_pairs = total;
        a.status->overflow = total > a.pair_capacity ? 1u : 0u;
        if (a.status_host) {  // straight to the host over PCIe: no copy engine, no extra launch
          a.status_host[0] = total;
          a.status_host[1] = (total > a.pair_capacity ? 1ull : 0ull) | (1ull << 32);
          __threadfence_system();
        }
      }
    }
  }
  __syncthreads();
  unsigned long long run = s_base + wexcl + incl - tsum;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) { if (first + i < a.n_tickets) a.ticket_offsets[first + i] = run; run += v[i]; }
}

// -------------------------------------------------------------------------------------------------
// Persistent: gridDim.x CTAs loop over the tickets, so that the per-CTA digit histograms (all sort passes,
// accumulated in shared memory while the keys are being written) are flushed with few global atomics.
// A Gaussian's pairs share their depth bits, so the four depth digits cost one shared atomic per GAUSSIAN
// (adding its tile count); the (view | tile) digits cost one per pair.  This removes the sort's separate
// histogram pass over the keys (8 B per pair re-read).
__global__ void __launch_bounds__(PRE_THREADS) emit_kernel(const PreArgs a) {
  __shared__ uint32_t s_warp_tot[PRE_THREADS / 32];
  __shared__ uint32_t s_hist[8][256];
  if (a.status->overflow) return;  // the host re-runs with a larger capacity
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int p = 0; p < 8; p++) s_hist[p][tid] = 0;
  __syncthreads();
  const int hi_passes = a.sort_passes - 4;
  for (int ticket = blockIdx.x; ticket < a.n_tickets; ticket += gridDim.x) {
    const int view = ticket % a.VV, chunk = ticket / a.VV;
    const uint2 info = a.bin_info[(size_t)ticket * PRE_THREADS + tid];
    const int rx0 = info.y & 255, ry0 = (info.y >> 8) & 255, rx1 = (info.y >> 16) & 255, ry1 = info.y >> 24;
    const uint32_t tiles = (uint32_t)((rx1 - rx0) * (ry1 - ry0));
    uint32_t incl = tiles;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    __syncthreads();  // previous ticket's readers of s_warp_tot are done
    if (lane == 31) s_warp_tot[warp] = incl;
    __syncthreads();
    uint32_t warp_excl = 0;
#pragma unroll
    for (int w = 0; w < PRE_THREADS / 32; w++) if (w < warp) warp_excl += s_warp_tot[w];
    const unsigned long long off = a.ticket_offsets[ticket] + warp_excl + incl - tiles;
    const uint32_t dbits = info.x;
    const uint32_t vhi = (uint32_t)view << a.tile_bits;
    const uint32_t gidx = (uint32_t)(chunk * PRE_THREADS + tid);
    const bool big = tiles > (uint32_t)WARP_EMIT_THRESHOLD;
    if (tiles > 0) {
#pragma unroll
      for (int p = 0; p < 4; p++) atomicAdd(&s_hist[p][(dbits >> (8 * p)) & 255u], tiles);
    }
    if (tiles > 0 && !big) {
      unsigned long long o = off;
      for (int y = ry0; y < ry1; y++)
        for (int x = rx0; x < rx1; x++) {
          const uint32_t hi = vhi | (uint32_t)(y * a.grid_x + x);
          a.keys[o] = ((uint64_t)hi << 32) | dbits;
          a.vals[o] = gidx;
          o++;
          for (int hp = 0; hp < hi_passes; hp++) atomicAdd(&s_hist[4 + hp][(hi >> (8 * hp)) & 255u], 1u);
        }
    }
    uint32_t bmask = __ballot_sync(0xffffffffu, big);
    while (bmask) {
      const int src = __ffs(bmask) - 1;
      bmask &= bmask - 1;
      const int x0 = __shfl_sync(0xffffffffu, rx0, src), y0 = __shfl_sync(0xffffffffu, ry0, src);
      const int w = __shfl_sync(0xffffffffu, rx1, src) - x0;
      const uint32_t cnt = __shfl_sync(0xffffffffu, tiles, src);
      const unsigned long long o = __shfl_sync(0xffffffffu, off, src);
      const uint32_t db = __shfl_sync(0xffffffffu, dbits, src), gi = __shfl_sync(0xffffffffu, gidx, src);
      for (uint32_t k = lane; k < cnt; k += 32) {
        const int yy = y0 + (int)(k / (uint32_t)w), xx = x0 + (int)(k % (uint32_t)w);
        const uint32_t hi = vhi | (uint32_t)(yy * a.grid_x + xx);
        a.keys[o + k] = ((uint64_t)hi << 32) | db;
        a.vals[o + k] = gi;
        for (int hp = 0; hp < hi_passes; hp++) atomicAdd(&s_hist[4 + hp][(hi >> (8 * hp)) & 255u], 1u);
      }
    }
  }
  __syncthreads();
  for (int p = 0; p < a.sort_passes; p++) { const uint32_t c = s_hist[p][tid]; if (c) atomicAdd(&a.hist[p * 256 + tid], c); }
}


// -------------------------------------------------------------------------------------------------
// BINNED sort mode: the (tile, Gaussian) pairs go straight into their (view, tile) bin instead of through a global
// sort.  bin_walk_kernel<false> counts the pairs of every bin, bin_scan_kernel turns the counts into the tile ranges
// (which identifyTileRanges would find in the sorted keys) and bin_walk_kernel<true> writes every pair's
// (depth bits, Gaussian index) entry at a position claimed from its bin's cursor; binsort.cu then orders each bin.
// Both walks go over a warp's 32 Gaussians in lock step, tile by tile: pixel-aligned neighbours overlap the same tile, so
// lanes are grouped into runs of equal bins and ONE atomic per run counts / claims for all of them (C2T: 34 M pairs,
// ~7 M atomics).  A Gaussian covering more than BIG_RECT tiles is walked by the whole warp instead.
constexpr uint32_t BIG_RECT = 32;
constexpr int WALK_ILP = 4;

template <bool SCATTER>
__global__ void __launch_bounds__(PRE_THREADS, 6) bin_walk_kernel(const PreArgs a) {
  if (SCATTER && a.status->overflow) return;  // the host re-runs with a larger capacity
  const int tid = threadIdx.x, lane = tid & 31;
  const uint32_t le = 0xffffffffu >> (31 - lane);  // lanes <= mine
  for (int ticket = blockIdx.x; ticket < a.n_tickets; ticket += gridDim.x) {
    const int view = ticket % a.VV, chunk = ticket / a.VV;
    const uint2 info = a.bin_info[(size_t)ticket * PRE_THREADS + tid];
    const int rx0 = info.y & 255, ry0 = (info.y >> 8) & 255, rx1 = (info.y >> 16) & 255, ry1 = info.y >> 24;
    const uint32_t tiles = (uint32_t)((rx1 - rx0) * (ry1 - ry0));
    const uint32_t vhi = (uint32_t)view << a.tile_bits;
    const uint2 entry = make_uint2(info.x, (uint32_t)(chunk * PRE_THREADS + tid));
    const bool big = tiles > BIG_RECT;
    const uint32_t steps = __reduce_max_sync(0xffffffffu, big ? 0u : tiles);
    int x = rx0, y = ry0;
    // WALK_ILP steps per iteration, their atomics issued back to back: a claim is a round trip to L2, and the steps of a
    // warp are independent of each other
    for (uint32_t t0 = 0; t0 < steps; t0 += WALK_ILP) {
      uint32_t bin[WALK_ILP], base[WALK_ILP];
      int first[WALK_ILP], len[WALK_ILP];
      bool act[WALK_ILP], head[WALK_ILP];
#pragma unroll
      for (int u = 0; u < WALK_ILP; u++) {
        act[u] = !big && t0 + u < tiles;
        bin[u] = act[u] ? (vhi | (uint32_t)(y * a.grid_x + x)) : 0xffffffffu;
        if (act[u] && ++x == rx1) { x = rx0; y++; }
        const uint32_t prev = __shfl_up_sync(0xffffffffu, bin[u], 1);
        head[u] = lane == 0 || prev != bin[u];
        const uint32_t heads = __ballot_sync(0xffffffffu, head[u]);
        first[u] = 31 - __clz(heads & le);                        // my run starts at the last head at or below me ...
        const uint32_t above = heads & ~le;
        len[u] = (above ? __ffs(above) - 1 : 32) - first[u];      // ... and ends before the next head
      }
#pragma unroll
      for (int u = 0; u < WALK_ILP; u++) {
        base[u] = 0;
        if (act[u] && head[u]) {
          if (SCATTER) base[u] = atomicAdd(&a.bin_cursor[bin[u]], (uint32_t)len[u]);
          else atomicAdd(&a.bin_count[bin[u]], (uint32_t)len[u]);
        }
      }
      if (SCATTER) {
#pragma unroll
        for (int u = 0; u < WALK_ILP; u++) {
          base[u] = __shfl_sync(0xffffffffu, base[u], first[u]);
          if (act[u]) a.entries[base[u] + (uint32_t)(lane - first[u])] = entry;
        }
      }
    }
    uint32_t bmask = __ballot_sync(0xffffffffu, big);
    while (bmask) {
      const int src = __ffs(bmask) - 1;
      bmask &= bmask - 1;
      const int x0 = __shfl_sync(0xffffffffu, rx0, src), y0 = __shfl_sync(0xffffffffu, ry0, src);
      const int w = __shfl_sync(0xffffffffu, rx1, src) - x0;
      const uint32_t cnt = __shfl_sync(0xffffffffu, tiles, src);
      const uint2 e = make_uint2(__shfl_sync(0xffffffffu, entry.x, src), __shfl_sync(0xffffffffu, entry.y, src));
      for (uint32_t k = lane; k < cnt; k += 32) {
        const int yy = y0 + (int)(k / (uint32_t)w), xx = x0 + (int)(k % (uint32_t)w);
        const uint32_t bin = vhi | (uint32_t)(yy * a.grid_x + xx);
        if (SCATTER) a.entries[atomicAdd(&a.bin_cursor[bin], 1u)] = e;
        else atomicAdd(&a.bin_count[bin], 1u);
      }
    }
  }
}

// One CTA: exclusive scan of the bin counts -> tile ranges and scatter cursors; bins sorted into the size classes of the
// segment sort; pair total, overflow flag and longest bin into the status block and the host's status word.
constexpr int BSCAN_THREADS = 1024;
__global__ void __launch_bounds__(BSCAN_THREADS) bin_scan_kernel(const uint32_t* __restrict__ bin_count, int bins, uint2* __restrict__ ranges,
                                                                 uint32_t* __restrict__ cursor, const BinSortWork w, B200sStatus* status,
                                                                 unsigned long long pair_capacity, unsigned long long* status_host) {
  __shared__ unsigned long long s_wsum[BSCAN_THREADS / 32];
  __shared__ uint32_t s_cls[BIN_CLASSES];
  __shared__ uint32_t s_max;
  __shared__ unsigned long long s_long;  // pairs in bins that only the global-memory LSD path takes
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < BIN_CLASSES) s_cls[tid] = 0;
  if (tid == 0) { s_max = 0; s_long = 0ull; }
  const int per = (bins + BSCAN_THREADS - 1) / BSCAN_THREADS;
  const int b0 = min(bins, tid * per), b1 = min(bins, b0 + per);
  unsigned long long tsum = 0;
  uint32_t tmax = 0;
  for (int b = b0; b < b1; b++) { const uint32_t c = bin_count[b]; tsum += c; tmax = max(tmax, c); }
  unsigned long long incl = tsum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
  if (lane == 31) s_wsum[warp] = incl;
  tmax = __reduce_max_sync(0xffffffffu, tmax);
  __syncthreads();
  if (lane == 0 && tmax) atomicMax(&s_max, tmax);
  unsigned long long wexcl = 0, total = 0;
#pragma unroll 4
  for (int k = 0; k < BSCAN_THREADS / 32; k++) { const unsigned long long t = s_wsum[k]; if (k < warp) wexcl += t; total += t; }
  unsigned long long run = wexcl + incl - tsum;
  for (int b = b0; b < b1; b++) {
    const uint32_t c = bin_count[b];
    const uint32_t start = (uint32_t)run;  // list positions are 32-bit (pair_capacity < 2^32)
    ranges[b] = make_uint2(start, start + c);
    cursor[b] = start;
    run += c;
    if (c) {
      const int cls = c <= (uint32_t)BIN_CAP_XS ? 0 : (c <= (uint32_t)BIN_CAP_S ? 1 : (c <= (uint32_t)BIN_CAP_L ? 2 : 3));
      w.class_list[(size_t)cls * bins + atomicAdd(&s_cls[cls], 1u)] = (uint32_t)b;
      if (cls == 3) atomicAdd(&s_long, (unsigned long long)c);
    }
  }
  __syncthreads();
  if (tid < BIN_CLASSES) { w.class_count[tid] = s_cls[tid]; w.class_next[tid] = 0; }
  if (tid == 0) {
    // flags: bit 0 = more pairs than the buffers hold; bit 1 = most of a LARGE pair count sits in bins that are too long for
    // shared memory (stress scenes: every Gaussian covers hundreds of tiles) -- one CTA per 10^5..10^6-entry bin and one
    // atomic per pair are the wrong tools there, the host re-runs the call in GLOBAL sort mode.  Either way nothing
    // downstream runs (every later kernel returns when the word is nonzero).
    const uint32_t over = (total > pair_capacity ? 1u : 0u) | ((total > BIN_GLOBAL_MIN_PAIRS && 2ull * s_long > total) ? 2u : 0u);
    status->num_pairs = total;
    status->overflow = over;
    status->max_bin_len = s_max;
    if (status_host) {  // straight to the host over PCIe: no copy engine, no extra launch
      status_host[0] = total;
      status_host[1] = (unsigned long long)over | ((unsigned long long)(0x80000000u | s_max) << 32);
      __threadfence_system();
    }
  }
}

}  // namespace b200s

// ---- host launcher (called from api.cu) --------------------------------------------------------
namespace b200s {
cudaError_t launch_bin_scan(const uint32_t* bin_count, int bins, uint2* ranges, uint32_t* cursor, const BinSortWork& w, B200sStatus* status,
                            unsigned long long pair_capacity, unsigned long long* status_host, cudaStream_t stream) {
  bin_scan_kernel<<<1, BSCAN_THREADS, 0, stream>>>(bin_count, bins, ranges, cursor, w, status, pair_capacity, status_host);
  count_launches(1);
  return cudaGetLastError();
}

cudaError_t launch_preprocess_bin(const B200sScene& sc, const B200sViews& vw, const B200sPlan& plan, char* saved, char* scratch,
                                  const B200sOut* out, cudaStream_t stream) {
  PreArgs a;
  a.N = sc.num_gaussians; a.VV = vw.num_views; a.H = vw.height; a.W = vw.width;
  a.grid_x = plan.grid_x; a.grid_y = plan.grid_y; a.tile_bits = plan.tile_bits;
  a.chunks = (sc.num_gaussians + PRE_THREADS - 1) / PRE_THREADS;
  a.n_tickets = plan.pre_tickets;
  a.pair_capacity = (unsigned long long)plan.pair_capacity;
  a.cov_floats = sc.cov_layout == B200S_COV_UPPER6 ? 6 : 9;
  a.col_floats = sc.colors_precomp ? 3 : 3 * sc.sh_coeffs;
  a.col_stride = a.col_floats | 1;
  a.fpg = 3 + a.cov_floats + 1 + a.col_stride;
  // raw scenes: the 41 staged planes are dead (in registers) before the specialised layout's 40 KB overwrite them
  if (sc.raw_head) { a.cov_floats = 9; a.col_floats = 27; a.col_stride = 27; a.fpg = 40; }
  a.rec = reinterpret_cast<Rec*>(saved + plan.off_rec);
  // the sort ping-pongs `sort_passes` times and must end in the A buffers
  const bool start_in_a = (plan.sort_passes % 2) == 0;
  a.keys = reinterpret_cast<uint64_t*>(scratch + (start_in_a ? plan.off_keys_a : plan.off_keys_b));
  a.vals = start_in_a ? reinterpret_cast<uint32_t*>(saved + plan.off_vals_a) : reinterpret_cast<uint32_t*>(scratch + plan.off_vals_b);
  a.ticket_offsets = reinterpret_cast<unsigned long long*>(scratch + plan.off_scan_state);
  a.ticket_totals = reinterpret_cast<uint32_t*>(scratch + plan.off_ticket_totals);
  a.scan_blocks = reinterpret_cast<uint64_t*>(scratch + plan.off_scan_blocks);
  a.bin_info = reinterpret_cast<uint2*>(scratch + plan.off_bin_info);
  a.hist = reinterpret_cast<uint32_t*>(scratch + plan.off_hist);
  a.sort_passes = plan.sort_passes;
  a.counters = reinterpret_cast<uint32_t*>(scratch + plan.off_counters);
  a.status = reinterpret_cast<B200sStatus*>(saved + plan.off_status);
  a.radii = out ? out->radii : nullptr;
  a.status_host = out ? reinterpret_cast<unsigned long long*>(out->status_host) : nullptr;
  const bool binned = plan.sort_mode == B200S_SORT_BINNED;
  a.bin_count = reinterpret_cast<uint32_t*>(scratch + plan.off_bin_count);
  a.bin_cursor = reinterpret_cast<uint32_t*>(scratch + plan.off_bin_cursor);
  a.entries = reinterpret_cast<uint2*>(scratch + plan.off_keys_a);

  cudaError_t e;
  stage_mark(B200S_STAGE_PRE_BIN, stream);
  if ((e = cudaMemsetAsync(a.status, 0, sizeof(B200sStatus), stream)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(a.counters, 0, CNT_WORDS * sizeof(uint32_t), stream)) != cudaSuccess) return e;
  const int scan_blocks = (plan.pre_tickets + SCAN_TILE - 1) / SCAN_TILE;
  if ((e = cudaMemsetAsync(a.scan_blocks, 0, (size_t)scan_blocks * sizeof(uint64_t), stream)) != cudaSuccess) return e;
  if (binned) { if ((e = cudaMemsetAsync(a.bin_count, 0, (size_t)plan.bins * sizeof(uint32_t), stream)) != cudaSuccess) return e; }
  else if ((e = cudaMemsetAsync(a.hist, 0, 8 * 256 * sizeof(uint32_t), stream)) != cudaSuccess) return e;
  const bool raw = sc.raw_head != nullptr;
  // the staged chunk only (raw scenes: the 41 planes, overwritten by the specialised layout once they are in registers);
  // records go straight to global memory
  const size_t smem = (size_t)PRE_THREADS * (raw ? RAW_PLANES : a.fpg) * sizeof(float);
  const bool specialised = raw || (sc.cov_layout == B200S_COV_3X3 && !sc.colors_precomp && sc.sh_layout == B200S_SH_CHANNEL_MAJOR &&
                                   sc.sh_coeffs == 9 && sc.sh_degree == 2);
  // the attribute is per (function, device) and cheap to set: no cache that a second device or thread could get wrong
  if (smem > 48 * 1024) {
    if ((e = cudaFuncSetAttribute(raw ? (const void*)project_kernel<2> : (specialised ? (const void*)project_kernel<1> : (const void*)project_kernel<0>),
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
  }
  if (plan.pre_tickets > 0) {
    const int proj_blocks = a.chunks * sc.num_scenes;
    if (raw) project_kernel<2><<<proj_blocks, PRE_THREADS, smem, stream>>>(sc, vw, a);
    else if (specialised) project_kernel<1><<<proj_blocks, PRE_THREADS, smem, stream>>>(sc, vw, a);
    else project_kernel<0><<<proj_blocks, PRE_THREADS, smem, stream>>>(sc, vw, a);
    const int sms = device_sm_count();
    const int walk_blocks = plan.pre_tickets < sms * 8 ? plan.pre_tickets : sms * 8;
    if (binned) {
      bin_walk_kernel<false><<<walk_blocks, PRE_THREADS, 0, stream>>>(a);
      BinSortWork w;
      w.class_list = reinterpret_cast<uint32_t*>(scratch + plan.off_long_list);
      w.class_count = a.counters + CNT_BIN_CLASS_COUNT;
      w.class_next = a.counters + CNT_BIN_CLASS_NEXT;
      bin_scan_kernel<<<1, BSCAN_THREADS, 0, stream>>>(a.bin_count, plan.bins, reinterpret_cast<uint2*>(saved + plan.off_ranges), a.bin_cursor, w,
                                                      a.status, a.pair_capacity, a.status_host);
      bin_walk_kernel<true><<<walk_blocks, PRE_THREADS, 0, stream>>>(a);
      count_launches(4);
    } else {
      scan_kernel<<<scan_blocks, SCAN_THREADS, 0, stream>>>(a);
      const int emit_blocks = plan.pre_tickets < sms * 6 ? plan.pre_tickets : sms * 6;
      emit_kernel<<<emit_blocks, PRE_THREADS, 0, stream>>>(a);
      count_launches(3);
    }
  }
  return cudaGetLastError();
}
}  // namespace b200s
