// composite.cu -- 16x16-tile alpha compositing, forward and backward.
//
// One CTA of 256 threads per (view, tile); thread = pixel.  Each warp owns an 8x4-pixel sub-tile.
// The tile's depth-sorted list is staged 256 entries at a time into shared memory (48 B projected
// record per entry, gathered by Gaussian index).  Before a warp evaluates an entry on its 32 pixels,
// ONE lane tests that entry's alpha>=1/255 extent against the warp's sub-tile (32 entries tested per
// instruction, warp ballot), so a warp only walks the entries that can touch it; the ballot also
// gives warp-level early termination.  Results are identical to walking the whole list: a skipped
// entry is one whose alpha is below 1/255 on every pixel of the sub-tile.
//
// Forward composites RGB and the depth colour in the same pass (the reference renders twice,
// cuda_splatting.py:250-263).  Backward replays the list back to front; per-pixel gradient
// contributions of three entries at a time are summed across the warp with a 31-shuffle
// reduce-scatter butterfly (instead of 5 shuffles per value), leaving one value per lane, which is
// added to the per-(view,Gaussian) gradient record with a single RED per lane.
//
// FP32-pipe bound: ~15 flop per (pixel, entry) test, +11 per blend.
#include "kernels.cuh"

namespace b200s {


struct TileGeom {
  int px, py;
  bool inside;
  float pfx, pfy, X0, X1, Y0, Y1;
};
__device__ __forceinline__ TileGeom tile_geom(int tile, int grid_x, int H, int W) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tx = tile % grid_x, ty = tile / grid_x;
  const int x0 = tx * TILE_X + (warp & 1) * 8, y0 = ty * TILE_Y + (warp >> 1) * 4;
  TileGeom g;
  g.px = x0 + (lane & 7); g.py = y0 + (lane >> 3);
  g.inside = g.px < W && g.py < H;
  g.pfx = (float)g.px; g.pfy = (float)g.py;
  g.X0 = (float)x0; g.X1 = (float)(x0 + 7); g.Y0 = (float)y0; g.Y1 = (float)(y0 + 3);
  return g;
}
__device__ __forceinline__ bool subtile_hit(const float4 q0, const float ex, const float ey, const TileGeom& g) {
  return (q0.x + ex >= g.X0) && (q0.x - ex <= g.X1) && (q0.y + ey >= g.Y0) && (q0.y - ey <= g.Y1);
}

// -------------------------------------------------------------------------------------------------
template <bool DEPTH, bool COUNT>
__global__ void __launch_bounds__(TILE_PIX) composite_fwd_kernel(const CompArgs a) {
  __shared__ float4 s_q0[TILE_PIX], s_q1[TILE_PIX], s_q2[TILE_PIX];
  if (*a.overflow) return;
  const int tile = blockIdx.x, view = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31;
  const uint2 range = a.ranges[((uint32_t)view << a.tile_bits) | (uint32_t)tile];
  const TileGeom g = tile_geom(tile, a.grid_x, a.H, a.W);
  const Rec* __restrict__ vrec = a.rec + (size_t)view * a.N;

  float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, D = 0.f;
  uint32_t last = 0, nblend = 0;
  bool done = !g.inside;
  bool warp_done = __all_sync(0xffffffffu, done);

  for (uint32_t base = range.x; base < range.y; base += TILE_PIX) {
    if (__syncthreads_count(done) == TILE_PIX) break;
    const int nst = min((uint32_t)TILE_PIX, range.y - base);
    if (tid < nst) {
      const uint32_t id = __ldg(a.vals + base + tid);
      const float4* r = reinterpret_cast<const float4*>(vrec + id);
      s_q0[tid] = __ldg(r); s_q1[tid] = __ldg(r + 1); s_q2[tid] = __ldg(r + 2);
    }
    __syncthreads();
    if (warp_done) continue;
    for (int c = 0; c < nst; c += 32) {
      const int j = c + lane;
      bool hit = false;
      if (j < nst) { const float4 q2 = s_q2[j]; hit = subtile_hit(s_q0[j], q2.z, q2.w, g); }
      uint32_t mask = __ballot_sync(0xffffffffu, hit);
      while (mask) {
        const int jj = c + __ffs(mask) - 1;
        mask &= mask - 1;
        if (!done) {
          const float4 q0 = s_q0[jj], q1 = s_q1[jj];
          const float dx = __fsub_rn(q0.x, g.pfx), dy = __fsub_rn(q0.y, g.pfy);
          const float power = gauss_power(q0.z, q0.w, q1.x, dx, dy);
          if (power <= 0.0f) {
            const float alpha = fminf(ALPHA_MAX, __fmul_rn(q1.y, expf(power)));
            if (alpha >= ALPHA_MIN) {
              const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
              if (test_T < T_MIN) {
                done = true;
              } else {
                const float4 q2 = s_q2[jj];
                C0 = __fmaf_rn(__fmul_rn(q1.z, alpha), T, C0);
                C1 = __fmaf_rn(__fmul_rn(q1.w, alpha), T, C1);
                C2 = __fmaf_rn(__fmul_rn(q2.x, alpha), T, C2);
                if (DEPTH) D = __fmaf_rn(__fmul_rn(q2.y, alpha), T, D);
                T = test_T;
                last = base - range.x + (uint32_t)jj + 1u;
                if (COUNT) nblend++;
              }
            }
          }
        }
      }
      if (__all_sync(0xffffffffu, done)) { warp_done = true; break; }
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
      if (tid == 0) atomicMax(&a.status->max_tile_len, range.y - range.x);
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

template <bool DEPTH>
struct PixState {
  float T, T_final, last_alpha, bg_dot;
  float accum[DEPTH ? 4 : 3], last_col[DEPTH ? 4 : 3], dpix[DEPTH ? 4 : 3];
  uint32_t last_contributor;
};

// one list entry on one pixel; writes its 10 partial gradients to out[0..9]
template <bool DEPTH>
__device__ __forceinline__ void bwd_entry(const float4 q0, const float4 q1, const float4 q2, const uint32_t pos, PixState<DEPTH>& s,
                                          const TileGeom& g, const float half_w, const float half_h, float* out) {
#pragma unroll
  for (int k = 0; k < 10; k++) out[k] = 0.f;
  if (pos >= s.last_contributor) return;
  const float dx = __fsub_rn(q0.x, g.pfx), dy = __fsub_rn(q0.y, g.pfy);
  const float power = gauss_power(q0.z, q0.w, q1.x, dx, dy);
  if (power > 0.0f) return;
  const float G = expf(power);
  const float alpha = fminf(ALPHA_MAX, __fmul_rn(q1.y, G));
  if (alpha < ALPHA_MIN) return;
  s.T = s.T / (1.f - alpha);
  const float w = alpha * s.T;
  const float col[4] = {q1.z, q1.w, q2.x, q2.y};
  float dL_dalpha = 0.f;
#pragma unroll
  for (int ch = 0; ch < (DEPTH ? 4 : 3); ch++) {
    s.accum[ch] = s.last_alpha * s.last_col[ch] + (1.f - s.last_alpha) * s.accum[ch];
    s.last_col[ch] = col[ch];
    dL_dalpha += (col[ch] - s.accum[ch]) * s.dpix[ch];
    out[6 + ch] = w * s.dpix[ch];
  }
  dL_dalpha *= s.T;
  s.last_alpha = alpha;
  dL_dalpha += (-s.T_final / (1.f - alpha)) * s.bg_dot;
  const float dL_dG = q1.y * dL_dalpha;
  const float gdx = G * dx, gdy = G * dy;
  const float dG_ddelx = -gdx * q0.z - gdy * q0.w;
  const float dG_ddely = -gdy * q1.x - gdx * q0.w;
  out[0] = dL_dG * dG_ddelx * half_w;
  out[1] = dL_dG * dG_ddely * half_h;
  out[2] = -0.5f * gdx * dx * dL_dG;
  out[3] = -0.5f * gdx * dy * dL_dG;
  out[4] = -0.5f * gdy * dy * dL_dG;
  out[5] = G * dL_dalpha;
}

template <bool DEPTH>
__global__ void __launch_bounds__(TILE_PIX) composite_bwd_kernel(const CompArgs a) {
  __shared__ float4 s_q0[TILE_PIX], s_q1[TILE_PIX], s_q2[TILE_PIX];
  __shared__ uint32_t s_id[TILE_PIX];
  __shared__ uint32_t s_max[TILE_PIX / 32];
  if (*a.overflow) return;
  const int tile = blockIdx.x, view = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint2 range = a.ranges[((uint32_t)view << a.tile_bits) | (uint32_t)tile];
  if (range.y == range.x) return;
  const TileGeom g = tile_geom(tile, a.grid_x, a.H, a.W);
  const Rec* __restrict__ vrec = a.rec + (size_t)view * a.N;
  float* __restrict__ grec = a.grad_rec + (size_t)view * a.N * GREC_FLOATS;
  const size_t HW = (size_t)a.H * a.W;
  const size_t pid = (size_t)g.py * a.W + g.px;

  PixState<DEPTH> s;
  s.T_final = g.inside ? a.final_T[(size_t)view * HW + pid] : 0.f;
  s.T = s.T_final;
  s.last_contributor = g.inside ? a.n_contrib[(size_t)view * HW + pid] : 0u;
  s.last_alpha = 0.f;
  s.bg_dot = 0.f;
#pragma unroll
  for (int ch = 0; ch < (DEPTH ? 4 : 3); ch++) { s.accum[ch] = 0.f; s.last_col[ch] = 0.f; s.dpix[ch] = 0.f; }
  if (g.inside) {
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
      s.dpix[ch] = a.dL_dcolor[((size_t)view * 3 + ch) * HW + pid];
      s.bg_dot += a.bg[view * 3 + ch] * s.dpix[ch];
    }
    if (DEPTH) s.dpix[3] = a.dL_ddepth[(size_t)view * HW + pid];
  }
  // entries at list positions >= max(last_contributor) are needed by nobody
  uint32_t wmax = s.last_contributor;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, d));
  if (lane == 0) s_max[warp] = wmax;
  __syncthreads();
  uint32_t bmax = 0;
#pragma unroll
  for (int w = 0; w < TILE_PIX / 32; w++) bmax = max(bmax, s_max[w]);
  const float half_w = 0.5f * (float)a.W, half_h = 0.5f * (float)a.H;

  for (uint32_t hi = bmax; hi > 0;) {
    const uint32_t nst = min(hi, (uint32_t)TILE_PIX);
    __syncthreads();  // previous round fully consumed
    if ((uint32_t)tid < nst) {
      // slot `tid` holds list position hi-1-tid: slots ascend as the list is walked back to front
      const uint32_t id = __ldg(a.vals + range.x + (hi - 1 - tid));
      const float4* r = reinterpret_cast<const float4*>(vrec + id);
      s_id[tid] = id; s_q0[tid] = __ldg(r); s_q1[tid] = __ldg(r + 1); s_q2[tid] = __ldg(r + 2);
    }
    __syncthreads();
    if (hi - nst < wmax) {  // warp-uniform: some pixel of this warp still needs entries of this round
      for (uint32_t c = 0; c < nst; c += 32) {
        const uint32_t j = c + lane;
        bool hit = false;
        if (j < nst && (hi - 1 - j) < wmax) { const float4 q2 = s_q2[j]; hit = subtile_hit(s_q0[j], q2.z, q2.w, g); }
        uint32_t mask = __ballot_sync(0xffffffffu, hit);
        while (mask) {
          float v[32];
          uint32_t ids[3];
          bool used[3];
#pragma unroll
          for (int e = 0; e < 3; e++) {
            used[e] = mask != 0;  // warp-uniform
            if (used[e]) {
              const uint32_t jj = c + __ffs(mask) - 1;
              mask &= mask - 1;
              ids[e] = s_id[jj];
              bwd_entry<DEPTH>(s_q0[jj], s_q1[jj], s_q2[jj], hi - 1 - jj, s, g, half_w, half_h, v + 10 * e);
            } else {
              ids[e] = 0;
#pragma unroll
              for (int k = 0; k < 10; k++) v[10 * e + k] = 0.f;
            }
          }
          v[30] = 0.f; v[31] = 0.f;
          butterfly_step<16>(v, lane); butterfly_step<8>(v, lane); butterfly_step<4>(v, lane);
          butterfly_step<2>(v, lane); butterfly_step<1>(v, lane);
          const int e = lane / 10, k = lane - 10 * e;
          const bool live = e == 0 ? used[0] : (e == 1 ? used[1] : (e == 2 ? used[2] : false));
          const uint32_t id = e == 0 ? ids[0] : (e == 1 ? ids[1] : ids[2]);
          if (live && v[0] != 0.f && (DEPTH || k != 9)) atomicAdd(grec + (size_t)id * GREC_FLOATS + k, v[0]);
        }
      }
    }
    hi -= nst;
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
  if (depth) composite_bwd_kernel<true><<<grid, TILE_PIX, 0, stream>>>(a);
  else composite_bwd_kernel<false><<<grid, TILE_PIX, 0, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace b200s
