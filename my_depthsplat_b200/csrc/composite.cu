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

// A compacted hit as a warp keeps it in its private shared-memory slots: everything one entry needs behind ONE base
// address (the loop index is warp-uniform, so the loads are broadcasts with immediate offsets).
struct __align__(16) HitSlot {
  float4 q1;   // (A, B, C, opacity)
  float4 q2;   // (r, g, b, zc)
  float x, y;
  uint32_t pos, id;
};

// -------------------------------------------------------------------------------------------------
template <bool DEPTH, bool COUNT>
__global__ void __launch_bounds__(TILE_PIX) composite_fwd_kernel(const CompArgs a) {
  __shared__ HitSlot s_slot[WARPS][32];
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
        const float alpha = fminf(ALPHA_MAX, __fmul_rn(h1.w, expf(power)));
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
    col[pid] = __fmaf_rn(T, bg[0], C0);
    col[HW + pid] = __fmaf_rn(T, bg[1], C1);
    col[2 * HW + pid] = __fmaf_rn(T, bg[2], C2);
    if (DEPTH) a.depth[(size_t)view * HW + pid] = D;
    a.final_T[(size_t)view * HW + pid] = T;
    a.n_contrib[(size_t)view * HW + pid] = last;
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
// reduce-scatter butterfly: on entry every lane holds 32 partials v[0..31]; on exit lane l holds in
// v[0] the sum over the warp of partial l.
template <int STRIDE>
__device__ __forceinline__ void butterfly_step(float* v, const uint32_t lane) {
  const bool upper = (lane & STRIDE) != 0;
#pragma unroll
  for (int i = 0; i < STRIDE; i++) {
    const float send = upper ? v[i] : v[i + STRIDE];
    const float keep = upper ? v[i + STRIDE] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, STRIDE);
  }
}

// Per-pixel state of the reverse walk.  The reference recurrence keeps, per channel, the colour accumulated BEHIND the
// current entry (accum = last_alpha * last_colour + (1 - last_alpha) * accum) and contracts it with dL/dpixel; the
// contraction commutes with the recurrence, so only its scalar image is carried:
//   E = sum_ch accum[ch] * dpix[ch]   ->   E' = last_alpha * D_last + (1 - last_alpha) * E,   D = sum_ch colour[ch] * dpix[ch]
template <bool DEPTH>
struct PixState {
  float T, tb, last_alpha;   // tb = final_T * sum_ch bg[ch] * dpix[ch]
  float E, D_last;
  float dpix[DEPTH ? 4 : 3];
  uint32_t last_contributor;
};

__device__ __forceinline__ float rcp_approx(float x) {  // one MUFU.RCP; the argument is 1 - alpha in [0.01, 1]
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// One list entry on the warp's 32 pixels; writes this pixel's 10 partial gradients to out[0..9].
// Returns (warp-uniform) whether ANY pixel of the warp received a contribution: a quarter of the entries
// that pass the bounding-box cull touch no pixel that is still "alive" at that list position, and they are
// dropped before the cross-lane reduction.  All 32 lanes must call.  Branch-free below the vote: a lane the entry
// does not reach runs the same arithmetic with alpha = 0 (T, the weights and q come out unchanged / zero) and keeps
// its recurrence state through selects.
template <bool DEPTH>
__device__ __forceinline__ bool bwd_entry(const HitSlot* __restrict__ h, PixState<DEPTH>& s, const TileGeom& g, float* out) {
  const float4 h1 = h->q1;
  const float dx = __fsub_rn(h->x, g.pfx), dy = __fsub_rn(h->y, g.pfy);
  const float power = gauss_power(h1.x, h1.y, h1.z, dx, dy);
  // ex2.approx (2 instructions, <= 2e-6 relative for the powers that can be live) instead of expf's 8: the backward
  // is held to 1e-4 of the gradient scale, not to the forward's 1e-5 on the image
  float G = __expf(power);
  float araw = __fmul_rn(h1.w, G);
  // the alpha >= 1/255 cut must fall exactly where the forward put it (same expf, same rounding): the approximate
  // exponential is off by up to ~1e-6 relative, so within 1e-5 of the threshold the decision is retaken with expf
  // (a handful of (pixel, entry) pairs per image take this branch)
  if (fabsf(__fmaf_rn(araw, 255.0f, -1.0f)) < 1e-5f) { G = expf(power); araw = __fmul_rn(h1.w, G); }
  const float alpha = fminf(ALPHA_MAX, araw);
  const bool live = h->pos < s.last_contributor && power <= 0.0f && alpha >= ALPHA_MIN;
  if (!__any_sync(0xffffffffu, live)) return false;
  const float4 h2 = h->q2;
  const float a_eff = live ? alpha : 0.f;
  const float inv = rcp_approx(1.f - a_eff);  // exactly 1 for a_eff = 0
  s.T = s.T * inv;
  const float w = a_eff * s.T;
  const float col[4] = {h2.x, h2.y, h2.z, h2.w};
  float D = 0.f;
#pragma unroll
  for (int ch = 0; ch < (DEPTH ? 4 : 3); ch++) {
    D += col[ch] * s.dpix[ch];
    out[6 + ch] = w * s.dpix[ch];
  }
  if (!DEPTH) out[9] = 0.f;
  const float E = s.last_alpha * s.D_last + (1.f - s.last_alpha) * s.E;
  s.E = live ? E : s.E;
  s.D_last = live ? D : s.D_last;
  s.last_alpha = live ? alpha : s.last_alpha;
  const float dL_dalpha = (D - E) * s.T - s.tb * inv;
  // moments of q = G * dL/dalpha over the pixels; the per-Gaussian factors (opacity, conic, half extent of the
  // image, -1/2) are applied once per (view, Gaussian) by the projection backward instead of once per pixel:
  //   dL/dmean2D = opacity * half * (-A*S_x - B*S_y, -C*S_y - B*S_x),  dL/dconic = -opacity/2 * (S_xx, S_xy, S_yy),
  //   dL/dopacity = S_1
  const float q = live ? G * dL_dalpha : 0.f;
  const float qx = q * dx, qy = q * dy;
  out[0] = qx;
  out[1] = qy;
  out[2] = qx * dx;
  out[3] = qx * dy;
  out[4] = qy * dy;
  out[5] = q;
  return true;
}

// CTA_WARPS = 8: one CTA per tile; 4: two CTAs of four warps per tile (smaller scheduling units: MIN_CTAS of them fit
// where the register file holds fewer whole tiles)
template <bool DEPTH, int CTA_WARPS, int MIN_CTAS>
__global__ void __launch_bounds__(CTA_WARPS * 32, MIN_CTAS) composite_bwd_kernel(const CompArgs a) {
  __shared__ HitSlot s_slot[CTA_WARPS][32];
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
  // after the reduction lane l holds value l of the batch: component c of entry e = l / 10
  const int red_e = lane / 10, red_c = lane - 10 * red_e;
  float* __restrict__ grec_lane = a.grad_rec + (size_t)view * a.N * GREC_FLOATS + red_c;

  PixState<DEPTH> s;
  s.T = g.inside ? a.final_T[(size_t)view * HW + pid] : 0.f;
  s.last_contributor = g.inside ? a.n_contrib[(size_t)view * HW + pid] : 0u;
  s.last_alpha = 0.f; s.E = 0.f; s.D_last = 0.f;
  float bg_dot = 0.f;
#pragma unroll
  for (int ch = 0; ch < (DEPTH ? 4 : 3); ch++) s.dpix[ch] = 0.f;
  if (g.inside) {
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
      s.dpix[ch] = a.dL_dcolor[((size_t)view * 3 + ch) * HW + pid];
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

  // walk positions wmax-1 ... 0; lane l of a chunk starting at `top` holds position top-1-l, so that the
  // compacted slots ascend as the list is walked back to front
  HitSlot* const slots = s_slot[warp];
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
      HitSlot* d = slots + __popc(mask & lt);
      const float4* r = reinterpret_cast<const float4*>(vrec + id);
      d->q1 = __ldg(r + 1); d->q2 = __ldg(r + 2);
      d->x = q0.x; d->y = q0.y; d->pos = (uint32_t)pos; d->id = id;
    }
    __syncwarp();
    const int nh = __popc(mask);
    // batches of three NON-EMPTY entries: walk the compacted slots in order, keep an entry only if some pixel
    // of the warp received a contribution from it (k and nh are warp-uniform: they live in uniform registers)
    for (int k = 0; k < nh;) {
      float v[32];
      uint32_t id0 = 0xffffffffu, id1 = 0xffffffffu, id2 = 0xffffffffu;
      bool f = false;
      while (k < nh && !f) { f = bwd_entry<DEPTH>(slots + k, s, g, v); if (f) id0 = slots[k].id; k++; }
      if (!f) break;
      f = false;
      while (k < nh && !f) { f = bwd_entry<DEPTH>(slots + k, s, g, v + 10); if (f) id1 = slots[k].id; k++; }
      if (!f) {
#pragma unroll
        for (int i = 10; i < 20; i++) v[i] = 0.f;
      }
      f = false;
      while (k < nh && !f) { f = bwd_entry<DEPTH>(slots + k, s, g, v + 20); if (f) id2 = slots[k].id; k++; }
      if (!f) {
#pragma unroll
        for (int i = 20; i < 30; i++) v[i] = 0.f;
      }
      v[30] = 0.f; v[31] = 0.f;
      butterfly_step<16>(v, lane); butterfly_step<8>(v, lane); butterfly_step<4>(v, lane);
      butterfly_step<2>(v, lane); butterfly_step<1>(v, lane);
      uint32_t gid = red_e == 0 ? id0 : id1;
      gid = red_e == 2 ? id2 : gid;
      gid = red_e > 2 ? 0xffffffffu : gid;
      if (gid != 0xffffffffu && v[0] != 0.f) atomicAdd(grec_lane + (size_t)gid * GREC_FLOATS, v[0]);
    }
    __syncwarp();  // slot reads of this chunk are done before the next chunk overwrites them
  }
}

// -------------------------------------------------------------------------------------------------
cudaError_t launch_composite_fwd(const CompArgs& a, int tiles, int views, bool depth, bool count, cudaStream_t stream) {
  if (tiles <= 0 || views <= 0) return cudaSuccess;
  dim3 grid(tiles, views);
  stage_mark(B200S_STAGE_COMP_FWD, stream);
  count_launches(1);
  if (depth) { if (count) composite_fwd_kernel<true, true><<<grid, TILE_PIX, 0, stream>>>(a); else composite_fwd_kernel<true, false><<<grid, TILE_PIX, 0, stream>>>(a); }
  else { if (count) composite_fwd_kernel<false, true><<<grid, TILE_PIX, 0, stream>>>(a); else composite_fwd_kernel<false, false><<<grid, TILE_PIX, 0, stream>>>(a); }
  return cudaGetLastError();
}
cudaError_t launch_composite_bwd(const CompArgs& a, int tiles, int views, bool depth, cudaStream_t stream) {
  if (tiles <= 0 || views <= 0) return cudaSuccess;
  dim3 grid(tiles, views);
  stage_mark(B200S_STAGE_COMP_BWD, stream);
  count_launches(1);
  // Two CTAs of four warps per tile, seven per SM (72 registers, 28 warps): 2.39 ms against 2.43 ms for one eight-warp
  // CTA per tile at three per SM (76 registers, 24 warps); four two-warp CTAs per tile measure the same as two four-warp
  // ones; capping the eight-warp kernel at 64 registers for four per SM spills and is slower (2.71 ms).
  // b200s_debug_set(2, 2) selects the eight-warp shape for A/B runs.
  if (g_sort_knobs[2].load(std::memory_order_relaxed) == 2) {
    if (depth) composite_bwd_kernel<true, 8, 3><<<grid, TILE_PIX, 0, stream>>>(a);
    else composite_bwd_kernel<false, 8, 3><<<grid, TILE_PIX, 0, stream>>>(a);
  } else {
    dim3 g2(tiles * 2, views);
    if (depth) composite_bwd_kernel<true, 4, 7><<<g2, 128, 0, stream>>>(a);
    else composite_bwd_kernel<false, 4, 7><<<g2, 128, 0, stream>>>(a);
  }
  return cudaGetLastError();
}

}  // namespace b200s
