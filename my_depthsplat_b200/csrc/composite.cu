// composite.cu -- 16x16-tile alpha compositing, forward and backward.
//
// One CTA of 256 threads per (view, tile) in the forward, two CTAs of 128 threads per (view, tile) in the backward
// (smaller scheduling units: 28 instead of 24 resident warps at its register count); thread = pixel; each WARP owns an
// 8x4-pixel sub-tile and
// walks the tile's depth-sorted list ON ITS OWN -- there is no block-wide staging and no
// __syncthreads in the list loop, so a warp whose pixels saturate early (or whose sub-tile few
// Gaussians touch) never waits for the other seven.  Per 32 list entries a warp
//   1. loads the 32 Gaussian indices (coalesced) and each lane gathers ONE 16-byte cull record
//      (x, y, ex, ey) -- the eight warps of the CTA read the same lines, so seven of them hit L1;
//   2. tests its entry's alpha>=1/255 extent against the warp's sub-tile (32 entries per instruction)
//      and ballots;
//   3. the hit lanes gather the remaining 32 bytes of their record and compact (q0,q1,q2) into the
//      warp's private shared-memory slots;
//   4. all lanes evaluate the compacted hits on their pixels, reading the slots by broadcast.
// Indices and cull records are software-prefetched two / one chunks ahead.  Results are identical to
// walking the whole list: a skipped entry is one whose alpha is below 1/255 on every pixel of the
// sub-tile.
//
// Forward composites RGB and the depth colour in the same pass (the reference renders twice,
// cuda_splatting.py:250-263).  Backward replays the list back to front; the per-pixel gradient
// contributions of three entries at a time are summed across the warp with a 31-shuffle
// reduce-scatter butterfly (instead of 5 shuffles per value), leaving one value per lane, which is
// added to the per-(view,Gaussian) gradient record with a single RED per lane.
//
// FP32-pipe bound: ~15 flop per (pixel, entry) test, +11 per blend.
#include "kernels.cuh"

namespace b200s {

constexpr int WARPS = TILE_PIX / 32;

struct TileGeom {
  int px, py;
  bool inside;
  float pfx, pfy, X0, X1, Y0, Y1;
};
// `warp` = which of the tile's eight 8x4 sub-tiles this warp owns
__device__ __forceinline__ TileGeom tile_geom(int tile, int grid_x, int H, int W, int warp) {
  const int lane = threadIdx.x & 31;
  const int tx = tile % grid_x, ty = tile / grid_x;
  const int x0 = tx * TILE_X + (warp & 1) * 8, y0 = ty * TILE_Y + (warp >> 1) * 4;
  TileGeom g;
  g.px = x0 + (lane & 7); g.py = y0 + (lane >> 3);
  g.inside = g.px < W && g.py < H;
  g.pfx = (float)g.px; g.pfy = (float)g.py;
  g.X0 = (float)x0; g.X1 = (float)(x0 + 7); g.Y0 = (float)y0; g.Y1 = (float)(y0 + 3);
  return g;
}
// q0 = (x, y, ex, ey)
__device__ __forceinline__ bool subtile_hit(const float4 q0, const TileGeom& g) {
  return (q0.x + q0.z >= g.X0) && (q0.x - q0.z <= g.X1) && (q0.y + q0.w >= g.Y0) && (q0.y - q0.w <= g.Y1);
}

// exp(x) for x <= 0 as ONE MUFU.EX2 on x * log2(e): <= 2 ulp of the exponential plus the rounding of the product
// (|x| <= 5.6 wherever alpha can reach 1/255, so <= 5e-7 relative in all) -- the same order as expf's own 1-2 ulp, at 2
// instructions instead of 8.  Forward and backward BOTH use it: alpha, hence every alpha >= 1/255 and T' >= 1e-4 cut, is
// the same number in the two passes by construction (the reference recomputes alpha with the same exp in both passes too).
__device__ __forceinline__ float exp_neg(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
  return r;
}

// A compacted hit as a warp keeps it in its private shared-memory slots: everything one entry needs behind ONE base
// address (the loop index is warp-uniform, so the loads are broadcasts with immediate offsets).
struct __align__(16) HitSlot {
  float4 q1;   // (A, B, C, opacity)
  float4 q2;   // (r, g, b, zc)
  float x, y;
  uint32_t pos, id;
};

// -------------------------------------------------------------------------------------------------
template <bool DEPTH, bool COUNT, bool LOSS>
__global__ void __launch_bounds__(TILE_PIX) composite_fwd_kernel(const CompArgs a) {
  __shared__ HitSlot s_slot[WARPS][32];
  __shared__ float s_loss[LOSS ? WARPS : 1][2];
  if (*a.overflow) return;
  const int tile = blockIdx.x, view = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  const uint2 range = a.ranges[((uint32_t)view << a.tile_bits) | (uint32_t)tile];
  const TileGeom g = tile_geom(tile, a.grid_x, a.H, a.W, warp);
  const Rec* __restrict__ vrec = a.rec + (size_t)view * a.N;
  const uint32_t* __restrict__ list = a.vals + range.x;
  const uint32_t len = range.y - range.x;
  HitSlot* const slots = s_slot[warp];

  float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, D = 0.f;
  float loss_sum = 0.f, psnr_sum = 0.f;
  uint32_t last = 0, nblend = 0;
  bool done = !g.inside;

  if (!__all_sync(0xffffffffu, done)) {
    // software pipeline: indices two chunks ahead, cull records one chunk ahead
    uint32_t id_cur = lane < len ? __ldg(list + lane) : 0u;
    uint32_t id_nxt = 32 + lane < len ? __ldg(list + 32 + lane) : 0u;
    float4 q0_cur = lane < len ? __ldg(&vrec[id_cur].q0) : make_float4(0.f, 0.f, -1.f, -1.f);
    for (uint32_t base = 0; base < len; base += 32) {
      const uint32_t id = id_cur;
      const float4 q0 = q0_cur;
      const bool valid = base + lane < len;
      id_cur = id_nxt;
      id_nxt = base + 64 + lane < len ? __ldg(list + base + 64 + lane) : 0u;
      q0_cur = base + 32 + lane < len ? __ldg(&vrec[id_cur].q0) : make_float4(0.f, 0.f, -1.f, -1.f);

      const bool hit = valid && subtile_hit(q0, g);
      const uint32_t mask = __ballot_sync(0xffffffffu, hit);
      if (mask == 0) continue;
      if (hit) {
        HitSlot* d = slots + __popc(mask & lt);
        const float4* r = reinterpret_cast<const float4*>(vrec + id);
        d->q1 = __ldg(r + 1); d->q2 = __ldg(r + 2);
        d->x = q0.x; d->y = q0.y; d->pos = base + lane + 1u; d->id = id;
      }
      __syncwarp();
      const int nh = __popc(mask);
      for (int k = 0; k < nh; k++) {
        // one warp-uniform branch per hit (does ANY pixel blend it?), selects below it: a lane the entry does not
        // reach adds exactly zero and keeps T / last
        const HitSlot* h = slots + k;
        const float4 h1 = h->q1;
        const float dx = __fsub_rn(h->x, g.pfx), dy = __fsub_rn(h->y, g.pfy);
        const float power = gauss_power(h1.x, h1.y, h1.z, dx, dy);
        const float alpha = fminf(ALPHA_MAX, __fmul_rn(h1.w, exp_neg(power)));
        const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
        const bool reach = !done && power <= 0.0f && alpha >= ALPHA_MIN;
        const bool live = reach && !(test_T < T_MIN);
        done = done || (reach && !live);
        if (!__any_sync(0xffffffffu, live)) continue;
        const float4 h2 = h->q2;
        const float a_eff = live ? alpha : 0.f;
        C0 = __fmaf_rn(__fmul_rn(h2.x, a_eff), T, C0);
        C1 = __fmaf_rn(__fmul_rn(h2.y, a_eff), T, C1);
        C2 = __fmaf_rn(__fmul_rn(h2.z, a_eff), T, C2);
        if (DEPTH) D = __fmaf_rn(__fmul_rn(h2.w, a_eff), T, D);
        T = live ? test_T : T;
        last = live ? h->pos : last;
        if (COUNT) nblend += live ? 1u : 0u;
      }
      __syncwarp();  // this chunk's slot reads are done before the next chunk overwrites the slots
      if (__all_sync(0xffffffffu, done)) break;
    }
  }
  if (g.inside) {
    const size_t HW = (size_t)a.H * a.W;
    const size_t pid = (size_t)g.py * a.W + g.px;
    const float* bg = a.bg + view * 3;
    float* col = a.color + (size_t)view * 3 * HW;
    const float c[3] = {__fmaf_rn(T, bg[0], C0), __fmaf_rn(T, bg[1], C1), __fmaf_rn(T, bg[2], C2)};
    col[pid] = c[0];
    col[HW + pid] = c[1];
    col[2 * HW + pid] = c[2];
    if (DEPTH) a.depth[(size_t)view * HW + pid] = D;
    a.final_T[(size_t)view * HW + pid] = T;
    a.n_contrib[(size_t)view * HW + pid] = last;
    if (LOSS) {  // the loss and its gradient while the pixel is in registers (loss_mse.py:33-44, metrics.py:11-19)
      const float* gt = a.mse_target + (size_t)view * 3 * HW + pid;
      float* gr = a.mse_grad + (size_t)view * 3 * HW + pid;
#pragma unroll
      for (int ch = 0; ch < 3; ch++) {
        const float t = __ldg(gt + ch * HW), d = c[ch] - t;
        gr[ch * HW] = a.mse_l1 ? (d > 0.f ? a.mse_scale : (d < 0.f ? -a.mse_scale : 0.f)) : 2.0f * a.mse_scale * d;
        loss_sum += a.mse_l1 ? fabsf(d) : d * d;
        const float dc = fminf(fmaxf(c[ch], 0.f), 1.f) - fminf(fmaxf(t, 0.f), 1.f);
        psnr_sum += dc * dc;
      }
    }
  }
  if (LOSS) {  // fixed-order sums: lanes, then the eight warps, one (view, tile) slot each -- no atomics
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { loss_sum += __shfl_xor_sync(0xffffffffu, loss_sum, d); psnr_sum += __shfl_xor_sync(0xffffffffu, psnr_sum, d); }
    if (lane == 0) { s_loss[warp][0] = loss_sum; s_loss[warp][1] = psnr_sum; }
    __syncthreads();
    if (threadIdx.x < 2) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < WARPS; w++) t += s_loss[w][threadIdx.x];
      a.mse_partials[((size_t)view * gridDim.x + tile) * 2 + threadIdx.x] = t;
    }
  }
  if (COUNT) {
    unsigned long long t = g.inside ? last : 0, b = nblend;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { t += __shfl_xor_sync(0xffffffffu, t, d); b += __shfl_xor_sync(0xffffffffu, b, d); }
    if (lane == 0) {
      atomicAdd((unsigned long long*)&a.status->tested, t);
      atomicAdd((unsigned long long*)&a.status->blended, b);
      if (threadIdx.x == 0) atomicMax(&a.status->max_tile_len, len);
    }
  }
}

// -------------------------------------------------------------------------------------------------
// Backward.  Per pixel and entry the walk computes only what depends on the pixel; what depends on the Gaussian is
// accumulated by a lane that OWNS the Gaussian:
//
//   pixel side (lane = pixel, as in the forward): entries are compacted into a warp-private ring of hit slots; for a
//     batch of up to 16 hits, in list order back to front, every lane replays alpha and T, carries the scalar image of
//     the behind-colour recurrence, and writes just TWO numbers per (hit, pixel) into a 16 x 32 shared-memory matrix:
//         q = G * dL/dalpha   (drives dL/dmean2D, dL/dconic, dL/dopacity)      w = alpha * T   (drives dL/dcolour)
//   hit side (lane = hit h = lane & 15, pixel half = lane >> 4): reads its row of the matrix (16 pixels per lane,
//     conflict-free: rows are padded to 33 words), rebuilds (dx, dy) of every pixel from its own centre, and sums the
//     six moments of q and the four colour sums in REGISTERS; the two halves are combined with one shuffle per value and
//     the hit's ten sums go out as ten REDs (five per half).
//
// This replaces the reduce-scatter butterfly of the previous version (three hits' 30 partial sums per pixel reduced
// across the warp with 31 shuffles + 62 selects + 31 adds): the transposition now moves 2 values per (hit, pixel) instead
// of reducing 10, and the moment arithmetic runs once per (hit, pixel) on the hit side instead of in every pixel lane
// before the reduction.  Per processed hit: ~45 (pixel side) + ~19 (hit side) warp instructions against ~105.
//
// Per-pixel state of the reverse walk.  The reference recurrence keeps, per channel, the colour accumulated BEHIND the
// current entry (accum = last_alpha * last_colour + (1 - last_alpha) * accum) and contracts it with dL/dpixel; the
// contraction commutes with the recurrence, so only its scalar image is carried, advanced right after each entry:
//   E = sum_ch accum[ch] * dpix[ch]   ->   E' = alpha * D + (1 - alpha) * E,   D = sum_ch colour[ch] * dpix[ch]
template <bool DEPTH>
struct PixState {
  float T, tb;               // tb = final_T * sum_ch bg[ch] * dpix[ch]
  float E;
  float dpix[DEPTH ? 4 : 3];
  uint32_t last_contributor;
};

__device__ __forceinline__ float rcp_approx(float x) {  // one MUFU.RCP; the argument is 1 - alpha in [0.01, 1]
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

constexpr int BWD_BATCH = 16;   // hits per batch: the hit side maps lane -> (hit = lane & 15, pixel half = lane >> 4)
constexpr int BWD_SLOTS = 48;   // hit slots per warp: up to 15 left over + the 32 of a new chunk
constexpr int BWD_ROW = 33;     // padded row of the (hit, pixel) matrices

// One list entry on the warp's 32 pixels: returns (warp-uniform) whether ANY pixel of the warp received a contribution --
// a quarter of the entries that pass the bounding-box cull touch no pixel that is still "alive" at that list position;
// they are dropped before anything is stored.  All 32 lanes must call.  Branch-free below the vote: a lane the entry does
// not reach runs the same arithmetic with alpha = 0 (T and the recurrence state come out unchanged, q = w = 0).
// h1 = the slot's (A, B, C, opacity) quarter, hxy = its (x, y, pos, id) quarter: the caller requests them ONE HIT AHEAD (the
// loads of hit k + 1 are in flight while hit k's dependent chain -- power, exponential, vote -- runs: 1.84 -> 1.80 ms);
// Q2AHEAD: the colour quarter comes a hit ahead as well (needs the 80 registers of six CTAs per SM); otherwise it is
// requested here, with the vote still to come
template <bool DEPTH, bool Q2AHEAD = false>
__device__ __forceinline__ bool bwd_entry(const HitSlot* __restrict__ h, const float4 h1, const float4 hxy, PixState<DEPTH>& s, const TileGeom& g,
                                          float& q_out, float& w_out, const float4 pre_q2 = make_float4(0.f, 0.f, 0.f, 0.f)) {
  const float4 h2 = Q2AHEAD ? pre_q2 : h->q2;
  const float dx = __fsub_rn(hxy.x, g.pfx), dy = __fsub_rn(hxy.y, g.pfy);
  const float power = gauss_power(h1.x, h1.y, h1.z, dx, dy);
  const float G = exp_neg(power);  // the forward's alpha, bit for bit: the cuts fall where the forward put them
  const float araw = __fmul_rn(h1.w, G);
  const float alpha = fminf(ALPHA_MAX, araw);
  const bool live = __float_as_uint(hxy.z) < s.last_contributor && power <= 0.0f && alpha >= ALPHA_MIN;
  if (!__any_sync(0xffffffffu, live)) return false;
  const float a_eff = live ? alpha : 0.f;
  const float inv = rcp_approx(1.f - a_eff);  // exactly 1 for a_eff = 0
  s.T = s.T * inv;
  w_out = a_eff * s.T;
  float D = h2.x * s.dpix[0];
  D = fmaf(h2.y, s.dpix[1], D);
  D = fmaf(h2.z, s.dpix[2], D);
  if (DEPTH) D = fmaf(h2.w, s.dpix[3], D);
  // E = the colour accumulated BEHIND this entry, contracted with dL/dpixel: after the entry it becomes alpha D + (1 - alpha) E,
  // which leaves it untouched for alpha = 0 -- no selects, and the difference D - E serves both lines
  const float dE = D - s.E;
  const float dL_dalpha = dE * s.T - s.tb * inv;
  s.E = fmaf(a_eff, dE, s.E);
  // q = G * dL/dalpha; the per-Gaussian factors (opacity, conic, half extent of the image, -1/2) are applied once per
  // (view, Gaussian) by the projection backward:
  //   dL/dmean2D = opacity * half * (-A*S_x - B*S_y, -C*S_y - B*S_x),  dL/dconic = -opacity/2 * (S_xx, S_xy, S_yy),
  //   dL/dopacity = S_1,   S_f = sum over pixels of q * f(dx, dy)
  q_out = live ? G * dL_dalpha : 0.f;
  return true;
}

// CTA_WARPS = 8: one CTA per tile; 4: two CTAs of four warps per tile (smaller scheduling units: MIN_CTAS of them fit
// where the register file holds fewer whole tiles)
template <bool DEPTH, int CTA_WARPS, int MIN_CTAS, bool Q2AHEAD = false>
__global__ void __launch_bounds__(CTA_WARPS * 32, MIN_CTAS) composite_bwd_kernel(const CompArgs a) {
  __shared__ HitSlot s_slot[CTA_WARPS][BWD_SLOTS];
  __shared__ float s_q[CTA_WARPS][BWD_BATCH * BWD_ROW];
  __shared__ float s_w[CTA_WARPS][BWD_BATCH * BWD_ROW];
  __shared__ float4 s_dpix[CTA_WARPS][32];
  if (*a.overflow) return;
  constexpr int PER_TILE = 8 / CTA_WARPS;  // CTAs per tile
  const int tile = blockIdx.x / PER_TILE, view = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wsub = (int)(blockIdx.x % PER_TILE) * CTA_WARPS + warp;
  const uint32_t lt = (1u << lane) - 1u;
  const uint2 range = a.ranges[((uint32_t)view << a.tile_bits) | (uint32_t)tile];
  if (range.y == range.x) return;
  const TileGeom g = tile_geom(tile, a.grid_x, a.H, a.W, wsub);
  const Rec* __restrict__ vrec = a.rec + (size_t)view * a.N;
  const uint32_t* __restrict__ list = a.vals + range.x;
  const size_t HW = (size_t)a.H * a.W;
  const size_t pid = (size_t)g.py * a.W + g.px;
  float* __restrict__ grec = a.grad_rec + (size_t)view * a.N * GREC_FLOATS;

  PixState<DEPTH> s;
  s.T = g.inside ? a.final_T[(size_t)view * HW + pid] : 0.f;
  s.last_contributor = g.inside ? a.n_contrib[(size_t)view * HW + pid] : 0u;
  s.E = 0.f;
  float bg_dot = 0.f;
#pragma unroll
  for (int ch = 0; ch < (DEPTH ? 4 : 3); ch++) s.dpix[ch] = 0.f;
  if (g.inside) {
    const float cs = a.dpix_scale ? __ldg(a.dpix_scale) : 1.0f;
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
      s.dpix[ch] = a.dL_dcolor[((size_t)view * 3 + ch) * HW + pid] * cs;
      bg_dot += a.bg[view * 3 + ch] * s.dpix[ch];
    }
    if (DEPTH) s.dpix[3] = a.dL_ddepth[(size_t)view * HW + pid];
  }
  s.tb = s.T * bg_dot;
  // entries at list positions >= the warp's max(last_contributor) are needed by none of its pixels
  uint32_t wmax = s.last_contributor;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, d));
  if (wmax == 0) return;

  HitSlot* const slots = s_slot[warp];
  float* const mq = s_q[warp];
  float* const mw = s_w[warp];
  s_dpix[warp][lane] = make_float4(s.dpix[0], s.dpix[1], s.dpix[2], DEPTH ? s.dpix[3] : 0.f);
  // hit-side role of this lane: hit hh of the batch, pixels half * 16 .. half * 16 + 15 of the sub-tile
  const int hh = lane & (BWD_BATCH - 1), half = lane >> 4;
  const float hx0 = g.X0, hy0 = g.Y0 + (float)(2 * half);  // pixel j of the half sits at (X0 + (j & 7), Y0 + 2 * half + (j >> 3))
  const float* const mq_row = mq + hh * BWD_ROW + 16 * half;
  const float* const mw_row = mw + hh * BWD_ROW + 16 * half;
  const float4* const dp_half = s_dpix[warp] + 16 * half;
  __syncwarp();

  int cnt = 0;  // hits waiting in slots[0 .. cnt) (warp-uniform)

  // one batch of nb <= 16 hits: slots[first .. first + nb)
  auto run_batch = [&](const HitSlot* batch, const int nb) {
    uint32_t nonempty = 0;
    float4 nq1 = batch->q1, nxy = *reinterpret_cast<const float4*>(&batch->x), nq2 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (Q2AHEAD) nq2 = batch->q2;
#pragma unroll 4
    for (int k = 0; k < nb; k++) {
      const float4 cq1 = nq1, cxy = nxy, cq2 = nq2;
      const HitSlot* nx = batch + (k + 1 < nb ? k + 1 : k);
      nq1 = nx->q1; nxy = *reinterpret_cast<const float4*>(&nx->x);
      if (Q2AHEAD) nq2 = nx->q2;
      float q, w;
      if (!bwd_entry<DEPTH, Q2AHEAD>(batch + k, cq1, cxy, s, g, q, w, cq2)) continue;
      mq[k * BWD_ROW + lane] = q;
      mw[k * BWD_ROW + lane] = w;
      nonempty |= 1u << k;
    }
    if (nonempty) {
      __syncwarp();
      float S[10];
#pragma unroll
      for (int i = 0; i < 10; i++) S[i] = 0.f;
      const bool mine = (nonempty >> hh) & 1u;
      uint32_t gid = 0;
      if (mine) {
        const HitSlot* h = batch + hh;
        const float bx = h->x, by = h->y;
        gid = h->id;
#pragma unroll
        for (int j = 0; j < 16; j++) {
          const float q = mq_row[j], w = mw_row[j];
          const float4 dp = dp_half[j];
          const float dx = __fsub_rn(bx, hx0 + (float)(j & 7)), dy = __fsub_rn(by, hy0 + (float)(j >> 3));
          const float qx = q * dx, qy = q * dy;
          S[0] += qx; S[1] += qy;
          S[2] = fmaf(qx, dx, S[2]); S[3] = fmaf(qx, dy, S[3]); S[4] = fmaf(qy, dy, S[4]);
          S[5] += q;
          S[6] = fmaf(w, dp.x, S[6]); S[7] = fmaf(w, dp.y, S[7]); S[8] = fmaf(w, dp.z, S[8]);
          if (DEPTH) S[9] = fmaf(w, dp.w, S[9]);
        }
      }
#pragma unroll
      for (int i = 0; i < 10; i++) S[i] += __shfl_xor_sync(0xffffffffu, S[i], 16);
      if (mine) {
        float* dst = grec + (size_t)gid * GREC_FLOATS + 5 * half;
#pragma unroll
        for (int i = 0; i < 5; i++) {
          const float v = half ? S[5 + i] : S[i];
          if (v != 0.f) atomicAdd(dst + i, v);
        }
      }
      __syncwarp();  // the matrix rows are read before the next batch overwrites them
    }
  };

  // walk positions wmax-1 ... 0; lane l of a chunk starting at `top` holds position top-1-l, so that the
  // compacted slots ascend as the list is walked back to front
  int top = (int)wmax;
  uint32_t id_cur = top - 1 - lane >= 0 ? __ldg(list + (top - 1 - lane)) : 0u;
  uint32_t id_nxt = top - 33 - lane >= 0 ? __ldg(list + (top - 33 - lane)) : 0u;
  float4 q0_cur = top - 1 - lane >= 0 ? __ldg(&vrec[id_cur].q0) : make_float4(0.f, 0.f, -1.f, -1.f);
  for (; top > 0; top -= 32) {
    const uint32_t id = id_cur;
    const float4 q0 = q0_cur;
    const int pos = top - 1 - lane;
    id_cur = id_nxt;
    id_nxt = top - 65 - lane >= 0 ? __ldg(list + (top - 65 - lane)) : 0u;
    q0_cur = top - 33 - lane >= 0 ? __ldg(&vrec[id_cur].q0) : make_float4(0.f, 0.f, -1.f, -1.f);

    const bool hit = pos >= 0 && subtile_hit(q0, g);
    const uint32_t mask = __ballot_sync(0xffffffffu, hit);
    if (mask == 0) continue;
    if (hit) {
      HitSlot* d = slots + cnt + __popc(mask & lt);
      const float4* r = reinterpret_cast<const float4*>(vrec + id);
      d->q1 = __ldg(r + 1); d->q2 = __ldg(r + 2);
      d->x = q0.x; d->y = q0.y; d->pos = (uint32_t)pos; d->id = id;
    }
    cnt += __popc(mask);
    __syncwarp();
    if (cnt >= BWD_BATCH) {
      int first = 0;
      for (; cnt - first >= BWD_BATCH; first += BWD_BATCH) run_batch(slots + first, BWD_BATCH);
      // the (< 16) hits left over move to the front of the buffer
      const int left = cnt - first;
      float4 t0, t1, t2;
      if (lane < left) {
        const float4* src = reinterpret_cast<const float4*>(slots + first + lane);
        t0 = src[0]; t1 = src[1]; t2 = src[2];
      }
      __syncwarp();
      if (lane < left) {
        float4* dst = reinterpret_cast<float4*>(slots + lane);
        dst[0] = t0; dst[1] = t1; dst[2] = t2;
      }
      cnt = left;
      __syncwarp();
    }
  }
  if (cnt > 0) run_batch(slots, cnt);
}

// -------------------------------------------------------------------------------------------------
cudaError_t launch_composite_fwd(const CompArgs& a, int tiles, int views, bool depth, bool count, cudaStream_t stream) {
  if (tiles <= 0 || views <= 0) return cudaSuccess;
  dim3 grid(tiles, views);
  stage_mark(B200S_STAGE_COMP_FWD, stream);
  count_launches(1);
  const bool loss = a.mse_target != nullptr;
#define FWD(D_, C_, L_) composite_fwd_kernel<D_, C_, L_><<<grid, TILE_PIX, 0, stream>>>(a)
  if (depth) { if (count) { if (loss) FWD(true, true, true); else FWD(true, true, false); } else { if (loss) FWD(true, false, true); else FWD(true, false, false); } }
  else { if (count) { if (loss) FWD(false, true, true); else FWD(false, true, false); } else { if (loss) FWD(false, false, true); else FWD(false, false, false); } }
#undef FWD
  return cudaGetLastError();
}
cudaError_t launch_composite_bwd(const CompArgs& a, int tiles, int views, bool depth, cudaStream_t stream) {
  if (tiles <= 0 || views <= 0) return cudaSuccess;
  dim3 grid(tiles, views);
  stage_mark(B200S_STAGE_COMP_BWD, stream);
  count_launches(1);
  dim3 g2(tiles * 2, views);
  // two CTAs of four warps per tile, six per SM: 80 registers hold the slot quarters of the NEXT hit (conic, position AND
  // colour) next to the current one's -- 1.77 ms; seven per SM at 72 registers (colour quarter requested per hit, it spills
  // otherwise): 1.80 ms; without any load ahead: 1.84 ms
  if (depth) composite_bwd_kernel<true, 4, 6, true><<<g2, 128, 0, stream>>>(a);
  else composite_bwd_kernel<false, 4, 6, true><<<g2, 128, 0, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace b200s
