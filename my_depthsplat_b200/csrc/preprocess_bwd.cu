// preprocess_bwd.cu -- backward of the per-Gaussian projection (SURVEY.md K8 + K9), fused with the
// sum over the views of a scene that the reference gets from the autograd of its per-view
// replication (decoder_splatting_cuda.py:53-56) and with the chain rule of the scale-invariant
// normalisation (cuda_splatting.py:63-70) and of the depth colour (cuda_splatting.py:238-246).
//
// One thread per Gaussian; the thread loops over the views of its scene, re-derives the projection
// from the staged inputs, consumes the 48-byte gradient record the compositing backward
// accumulated for (view, Gaussian), and accumulates dL/d{mean, covariance, SH, opacity} in registers.
// One coalesced write per output tensor (through shared memory), in the caller's own layouts.
// No atomics; the order of the sum over views is fixed (ascending view index).
//
// HBM-bound: per Gaussian 148 B in + 148 B out, plus (48 + 16) B per (view, Gaussian).
#include "adapter.cuh"
#include "kernels.cuh"

namespace b200s {

struct PreBwdArgs {
  int N, VV, H, W;
  int cov_floats, col_floats, col_stride;
  const Rec* rec;
  const float* grad_rec;  // [VV,N,12]
  float* dL_dmeans2D;     // [VV,N,3] or NULL
  const uint32_t* overflow;
  int chunk_begin, chunk_count;  // 256-Gaussian chunks (per scene) this launch covers: chunk_begin + r * chunk_stride + k,
  int chunk_stride, chunk_repeat, chunks_total;  //   r < chunk_repeat, k < chunk_count, clipped to the scene
};

__device__ __forceinline__ void stage_in_bwd(float* dst, const float* __restrict__ src, int count, int k, int stride) {
  if (k == stride && (((uintptr_t)src) & 15) == 0 && (count & 3) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = threadIdx.x; i < (count >> 2); i += PRE_THREADS) d4[i] = __ldg(s4 + i);
  } else {
    for (int i = threadIdx.x; i < count; i += PRE_THREADS) { const int g = i / k; dst[g * stride + (i - g * k)] = __ldg(src + i); }
  }
}
// NVLS: add to every replica the multicast address maps to (the switch performs the reduction)
__device__ __forceinline__ void multimem_red_add_v4(float* mc_addr, const float4 v) {
  asm volatile("multimem.red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void multimem_red_add(float* mc_addr, const float v) {
  asm volatile("multimem.red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(mc_addr), "f"(v) : "memory");
}

// shared -> global, coalesced; MC = the destination is a multicast address and the values are ADDED (multimem.red)
template <bool MC>
__device__ __forceinline__ void stage_out_bwd(float* __restrict__ dst, const float* src, int count, int k, int stride) {
  if (k == stride && (((uintptr_t)dst) & 15) == 0 && (count & 3) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = threadIdx.x; i < (count >> 2); i += PRE_THREADS) {
      if (MC) multimem_red_add_v4(reinterpret_cast<float*>(d4 + i), s4[i]);
      else d4[i] = s4[i];
    }
  } else {
    for (int i = threadIdx.x; i < count; i += PRE_THREADS) {
      const int g = i / k;
      const float v = src[g * stride + (i - g * k)];
      if (MC) multimem_red_add(dst + i, v);
      else dst[i] = v;
    }
  }
}

__device__ __forceinline__ void dnormvdv(const float v[3], const float dv[3], float o[3]) {
  const float sum2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  const float inv = 1.0f / sqrtf(sum2 * sum2 * sum2);
  o[0] = ((+sum2 - v[0] * v[0]) * dv[0] - v[1] * v[0] * dv[1] - v[2] * v[0] * dv[2]) * inv;
  o[1] = (-v[0] * v[1] * dv[0] + (sum2 - v[1] * v[1]) * dv[1] - v[2] * v[1] * dv[2]) * inv;
  o[2] = (-v[0] * v[2] * dv[0] - v[1] * v[2] * dv[1] + (sum2 - v[2] * v[2]) * dv[2]) * inv;
}

// NC = SH coefficients per channel that are ACTIVE ((deg+1)^2); NC == 0 -> colors_precomp path.
// RAW = the scene is the encoder head's raw output (adapter.cuh): the chunk's Gaussians are rebuilt from the raw planes as
// in the forward, and the accumulated gradients go through the adapter's chain rule before they are written.
// Register budget: with degree-2 harmonics 47 accumulators and inputs stay live across the view loop next to the prefetched
// records of the next view; at three CTAs per SM (80 registers) that spilled 300 bytes per thread.  Two CTAs per SM (up to
// 128 registers, no spills to speak of) run the C2T backward in 0.577 instead of 0.646 ms -- fewer warps, but every one of
// them keeps its loads in flight; four CTAs per SM (64 registers, 420 bytes of spills): 0.686 ms.
template <int NC, bool MC, bool RAW = false>
__global__ void __launch_bounds__(PRE_THREADS, (NC >= 9 ? 2 : 3)) preprocess_bwd_kernel(const B200sScene sc, const B200sViews vw, const B200sGradIn gin,
                                                                     const PreBwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ ViewParams s_vp[VIEW_GROUP];
  __shared__ int s_vlo, s_vhi;
  __shared__ __align__(8) uint64_t s_bar;
  if (*a.overflow) return;
  constexpr int DEG = NC == 16 ? 3 : (NC == 9 ? 2 : (NC == 4 ? 1 : 0));
  constexpr int NACC = NC > 0 ? 3 * NC : 3;
  const int tid = threadIdx.x;
  const int per_scene = a.chunk_count * a.chunk_repeat;
  const int scene = blockIdx.x / per_scene, lb = blockIdx.x % per_scene;
  const int chunk = a.chunk_begin + (lb / a.chunk_count) * a.chunk_stride + lb % a.chunk_count;
  if (chunk >= a.chunks_total) return;  // block-uniform
  const int i0 = chunk * PRE_THREADS;
  const int n = min(PRE_THREADS, a.N - i0);
  const long long g0 = (long long)scene * a.N + i0;

  float* s_mean = smem;
  float* s_cov = s_mean + PRE_THREADS * 3;
  float* s_col = s_cov + PRE_THREADS * a.cov_floats;
  if (tid == 0) { s_vlo = 0x7fffffff; s_vhi = -1; }  // ordered before the range search by the barriers of the staging below
  float mraw[3] = {0.f, 0.f, 0.f}, craw[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float op = 0.f;
  __shared__ RawCam s_cam;
  float hraw[RAW ? 10 : 1];   // the ten geometry channels (the chain rule needs them again)
  Cooked ck;
  int raw_cv = 0, raw_p0 = 0;
  if (RAW) {
    const int hw = sc.raw_h * sc.raw_w;
    raw_cv = i0 / hw; raw_p0 = i0 - raw_cv * hw;
    const size_t sv = (size_t)scene * sc.raw_views + raw_cv;
    if (tid == 0) {
      mbar_init(&s_bar, 1);
      mbar_expect_tx(&s_bar, (uint32_t)(RAW_PLANES * PRE_THREADS * 4));
      for (int ch = 0; ch < RAW_CH; ch++) tma_bulk_load(smem + ch * PRE_THREADS, sc.raw_head + (sv * RAW_CH + ch) * hw + raw_p0, PRE_THREADS * 4, &s_bar);
      tma_bulk_load(smem + RAW_CH * PRE_THREADS, sc.raw_depth + sv * hw + raw_p0, PRE_THREADS * 4, &s_bar);
      for (int ch = 0; ch < 3; ch++) tma_bulk_load(smem + (RAW_CH + 1 + ch) * PRE_THREADS, sc.raw_image + (sv * 3 + ch) * hw + raw_p0, PRE_THREADS * 4, &s_bar);
    }
    if (tid < RAW_CAM_FLOATS) reinterpret_cast<float*>(&s_cam)[tid] = __ldg(sc.raw_camera + sv * RAW_CAM_FLOATS + tid);
    __syncthreads();
    mbar_wait(&s_bar, 0);
    float h[RAW_CH], img[3];
#pragma unroll
    for (int ch = 0; ch < RAW_CH; ch++) h[ch] = smem[ch * PRE_THREADS + tid];
    const float depth = smem[RAW_CH * PRE_THREADS + tid];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) img[ch] = smem[(RAW_CH + 1 + ch) * PRE_THREADS + tid];
    __syncthreads();  // every raw value is in registers: the rotated SH may now overwrite the planes
    const int p = raw_p0 + tid;
    cook(h, depth, s_cam, p % sc.raw_w, p / sc.raw_w, sc.raw_w, sc.raw_h, sc.raw_scale_min, sc.raw_scale_max, ck);
    mraw[0] = ck.mean[0]; mraw[1] = ck.mean[1]; mraw[2] = ck.mean[2];
#pragma unroll
    for (int k = 0; k < 6; k++) craw[k] = ck.cov[k];
    op = ck.op;
#pragma unroll
    for (int k = 0; k < (RAW ? 10 : 1); k++) hraw[k] = h[k];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) cook_sh(h + 10 + 9 * ch, img[ch], s_cam, s_col + tid * 27 + 9 * ch);
  } else {
  {
    const float* g_mean = sc.means + g0 * 3; const float* g_cov = sc.covariances + g0 * a.cov_floats;
    const float* g_col = NC > 0 ? sc.harmonics + g0 * a.col_floats : nullptr;
    const bool tma = a.col_floats == a.col_stride && tma_ok(g_mean, n * 12) && tma_ok(g_cov, n * a.cov_floats * 4) &&
                     (NC == 0 || tma_ok(g_col, n * a.col_floats * 4));
    if (tma) {  // TMA bulk copies of the contiguous chunk (see preprocess.cu)
      if (tid == 0) {
        mbar_init(&s_bar, 1);
        mbar_expect_tx(&s_bar, (uint32_t)(n * 4 * (3 + a.cov_floats + (NC > 0 ? a.col_floats : 0))));
        tma_bulk_load(s_mean, g_mean, n * 12, &s_bar);
        tma_bulk_load(s_cov, g_cov, n * a.cov_floats * 4, &s_bar);
        if (NC > 0) tma_bulk_load(s_col, g_col, n * a.col_floats * 4, &s_bar);
      }
      __syncthreads();
      mbar_wait(&s_bar, 0);
    } else {
      stage_in_bwd(s_mean, g_mean, n * 3, 3, 3);
      stage_in_bwd(s_cov, g_cov, n * a.cov_floats, a.cov_floats, a.cov_floats);
      if (NC > 0) stage_in_bwd(s_col, g_col, n * a.col_floats, a.col_floats, a.col_stride);
      __syncthreads();
    }
  }
  op = tid < n ? __ldg(sc.opacities + g0 + tid) : 0.f;
  if (tid < n) {
    mraw[0] = s_mean[tid * 3]; mraw[1] = s_mean[tid * 3 + 1]; mraw[2] = s_mean[tid * 3 + 2];
    const float* cp = s_cov + tid * a.cov_floats;
    if (a.cov_floats == 6) { for (int k = 0; k < 6; k++) craw[k] = cp[k]; }
    else { craw[0] = cp[0]; craw[1] = cp[1]; craw[2] = cp[2]; craw[3] = cp[4]; craw[4] = cp[5]; craw[5] = cp[8]; }
  }
  }
  const int cstride = (sc.sh_layout == B200S_SH_CHANNEL_MAJOR) ? sc.sh_coeffs : 1;
  const int kstride = (sc.sh_layout == B200S_SH_CHANNEL_MAJOR) ? 1 : 3;
  const float* my_sh = s_col + tid * a.col_stride;

  float dmean[3] = {0.f, 0.f, 0.f}, dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, dop = 0.f;
  float dcol[NACC];
#pragma unroll
  for (int k = 0; k < NACC; k++) dcol[k] = 0.f;

  // ---- the views of the scene, VIEW_GROUP camera blocks at a time; no barrier inside a group.  The (view, Gaussian)
  // record and gradient record of the NEXT view are requested before the current view is worked on: one memory latency
  // per view is hidden behind the arithmetic instead of two exposed ones (record -> radius -> gradient record). ----
  scene_view_range(vw, a.VV, scene, &s_vlo, &s_vhi);
  __syncthreads();
  const int vlo = s_vlo, vhi = s_vhi;
  const size_t rstride = (size_t)a.N;
  const bool have = tid < n;
  float2 nq3 = make_float2(0.f, 0.f);   // (radius bits, flags bits) of the next view of the scene
  float4 ng0 = make_float4(0.f, 0.f, 0.f, 0.f), ng1 = ng0;
  float2 ng2 = make_float2(0.f, 0.f);
  auto fetch = [&](int view) {
    // invisible (view, Gaussian) pairs hold zeros in their gradient record (the buffer is cleared before the compositing
    // backward): loading it unconditionally costs no correctness and frees the load from the radius
    const size_t ri = (size_t)view * rstride + i0 + tid;
    const float4 q3 = __ldg(&a.rec[ri].q3);
    nq3 = make_float2(q3.y, q3.w);
    const float4* gp = reinterpret_cast<const float4*>(a.grad_rec + ri * GREC_FLOATS);
    ng0 = __ldg(gp); ng1 = __ldg(gp + 1);
    ng2 = __ldg(reinterpret_cast<const float2*>(gp + 2));
  };
  for (int v0 = vlo; v0 <= vhi; v0 += VIEW_GROUP) {
  const int vcount = min(VIEW_GROUP, vhi + 1 - v0);
  if (v0 != vlo) __syncthreads();  // the previous group's readers of s_vp are done
  load_view_group(s_vp, vw, v0, vcount, a.H, a.W);
  __syncthreads();
  if (have) {
    int vfirst = 0;
    while (vfirst < vcount && s_vp[vfirst].scene != scene) vfirst++;
    if (vfirst < vcount) fetch(v0 + vfirst);
  }
  for (int vj = 0; vj < vcount; vj++) {
    const ViewParams& vp = s_vp[vj];
    if (vp.scene != scene) continue;  // block-uniform
    if (!have) continue;
    const int view = v0 + vj;
    const size_t ri = (size_t)view * rstride + i0 + tid;
    const float2 q3yw = nq3;
    const float4 g0v = ng0, g1v = ng1;
    const float4 g2v = make_float4(ng2.x, ng2.y, 0.f, 0.f);
    {
      int vn = vj + 1;
      while (vn < vcount && s_vp[vn].scene != scene) vn++;
      if (vn < vcount) fetch(v0 + vn);
    }
    const int radius = __float_as_int(q3yw.x);
    if (radius <= 0) {
      if (a.dL_dmeans2D) { float* o = a.dL_dmeans2D + ri * 3; o[0] = 0.f; o[1] = 0.f; o[2] = 0.f; }
      continue;
    }
    const uint32_t flags = __float_as_uint(q3yw.y);
    dop += g1v.y;
    const float m[3] = {__fmul_rn(mraw[0], vp.s), __fmul_rn(mraw[1], vp.s), __fmul_rn(mraw[2], vp.s)};
    float c6[6];
#pragma unroll
    for (int k = 0; k < 6; k++) c6[k] = __fmul_rn(craw[k], vp.s2);

    // ---- conic -> cov2D -> cov3D and camera-space point --------------------------------------
    Cov2D q;
    compute_cov2d(m, c6, vp, q);
    const float ca = q.a, cb = q.b, cc = q.c;
    const float denom = __fsub_rn(__fmul_rn(ca, cc), __fmul_rn(cb, cb));
    // the compositing backward left the pixel moments S_x, S_y, S_xx, S_xy, S_yy of q = G * dL/dalpha (composite.cu):
    // apply the per-Gaussian factors here.  The conic is the forward's, bit for bit (same cov2D, same 1/det).
    float g2x, g2y, gcx, gcy, gcz;
    {
      const float det_inv = __frcp_rn(denom);
      const float cA = __fmul_rn(cc, det_inv), cB = __fmul_rn(-cb, det_inv), cC = __fmul_rn(ca, det_inv);
      const float sx = g0v.x, sy = g0v.y;
      g2x = op * (0.5f * (float)a.W) * (-cA * sx - cB * sy);
      g2y = op * (0.5f * (float)a.H) * (-cC * sy - cB * sx);
      const float mh = -0.5f * op;
      gcx = mh * g0v.z; gcy = mh * g0v.w; gcz = mh * g1v.x;
    }
    if (a.dL_dmeans2D) { float* o = a.dL_dmeans2D + ri * 3; o[0] = g2x; o[1] = g2y; o[2] = 0.f; }
    const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
    float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
    float dm[3];  // gradient w.r.t. the NORMALISED mean
    if (denom2inv != 0.f) {
      dL_da = denom2inv * (-cc * cc * gcx + 2 * cb * cc * gcy + (denom - ca * cc) * gcz);
      dL_dc = denom2inv * (-ca * ca * gcz + 2 * ca * cb * gcy + (denom - ca * cc) * gcx);
      dL_db = denom2inv * 2 * (cb * cc * gcx - (denom + 2 * cb * cb) * gcy + ca * cb * gcz);
      const float* T0 = q.T0; const float* T1 = q.T1;
      dcov[0] += vp.s2 * (T0[0] * T0[0] * dL_da + T0[0] * T1[0] * dL_db + T1[0] * T1[0] * dL_dc);
      dcov[3] += vp.s2 * (T0[1] * T0[1] * dL_da + T0[1] * T1[1] * dL_db + T1[1] * T1[1] * dL_dc);
      dcov[5] += vp.s2 * (T0[2] * T0[2] * dL_da + T0[2] * T1[2] * dL_db + T1[2] * T1[2] * dL_dc);
      dcov[1] += vp.s2 * (2 * T0[0] * T0[1] * dL_da + (T0[0] * T1[1] + T0[1] * T1[0]) * dL_db + 2 * T1[0] * T1[1] * dL_dc);
      dcov[2] += vp.s2 * (2 * T0[0] * T0[2] * dL_da + (T0[0] * T1[2] + T0[2] * T1[0]) * dL_db + 2 * T1[0] * T1[2] * dL_dc);
      dcov[4] += vp.s2 * (2 * T0[2] * T0[1] * dL_da + (T0[1] * T1[2] + T0[2] * T1[1]) * dL_db + 2 * T1[1] * T1[2] * dL_dc);
    }
    {
      const float V[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
      float dT0[3], dT1[3];
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const float t0v = q.T0[0] * V[j][0] + q.T0[1] * V[j][1] + q.T0[2] * V[j][2];
        const float t1v = q.T1[0] * V[j][0] + q.T1[1] * V[j][1] + q.T1[2] * V[j][2];
        dT0[j] = 2 * t0v * dL_da + t1v * dL_db;
        dT1[j] = 2 * t1v * dL_dc + t0v * dL_db;
      }
      // W[c][r] = view[4r + c]
      const float* vm = vp.view;
      const float dJ00 = vm[0] * dT0[0] + vm[4] * dT0[1] + vm[8] * dT0[2];
      const float dJ02 = vm[2] * dT0[0] + vm[6] * dT0[1] + vm[10] * dT0[2];
      const float dJ11 = vm[1] * dT1[0] + vm[5] * dT1[1] + vm[9] * dT1[2];
      const float dJ12 = vm[2] * dT1[0] + vm[6] * dT1[1] + vm[10] * dT1[2];
      const float tz = 1.f / q.t[2], tz2 = tz * tz, tz3 = tz2 * tz;
      const float hx = vp.focal_x, hy = vp.focal_y;
      const float dtx = q.xmul * -hx * tz2 * dJ02;
      const float dty = q.ymul * -hy * tz2 * dJ12;
      const float dtz = -hx * tz2 * dJ00 - hy * tz2 * dJ11 + (2 * hx * q.t[0]) * tz3 * dJ02 + (2 * hy * q.t[1]) * tz3 * dJ12;
      dm[0] = vm[0] * dtx + vm[1] * dty + vm[2] * dtz;
      dm[1] = vm[4] * dtx + vm[5] * dty + vm[6] * dtz;
      dm[2] = vm[8] * dtx + vm[9] * dty + vm[10] * dtz;
    }
    // ---- screen-space mean -> 3D mean ------------------------------------------------------------
    {
      const float* pr = vp.proj;
      const float mhw = xform_row(pr, 3, m[0], m[1], m[2]);
      const float m_w = 1.0f / (mhw + 0.0000001f);
      const float mul1 = (pr[0] * m[0] + pr[4] * m[1] + pr[8] * m[2] + pr[12]) * m_w * m_w;
      const float mul2 = (pr[1] * m[0] + pr[5] * m[1] + pr[9] * m[2] + pr[13]) * m_w * m_w;
      dm[0] += (pr[0] * m_w - pr[3] * mul1) * g2x + (pr[1] * m_w - pr[3] * mul2) * g2y;
      dm[1] += (pr[4] * m_w - pr[7] * mul1) * g2x + (pr[5] * m_w - pr[7] * mul2) * g2y;
      dm[2] += (pr[8] * m_w - pr[11] * mul1) * g2x + (pr[9] * m_w - pr[11] * mul2) * g2y;
    }
    // ---- colour ----------------------------------------------------------------------------------
    const float gcol[3] = {g1v.z, g1v.w, g2v.x};
    if (NC == 0) {
      dcol[0] += gcol[0]; dcol[1] += gcol[1]; dcol[2] += gcol[2];
    } else {
      const float dir_orig[3] = {m[0] - vp.campos[0], m[1] - vp.campos[1], m[2] - vp.campos[2]};
      const float len = sqrtf(dir_orig[0] * dir_orig[0] + dir_orig[1] * dir_orig[1] + dir_orig[2] * dir_orig[2]);
      const float x = dir_orig[0] / len, y = dir_orig[1] / len, z = dir_orig[2] / len;
      float dL_ddir[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int ch = 0; ch < 3; ch++) {
        const float g = (flags >> ch) & 1u ? 0.f : gcol[ch];
        const float* sh = my_sh + ch * cstride;
#define S(k) sh[(k) * kstride]
#define DS(k) dcol[ch * NC + (k)]
        float dx_ = 0.f, dy_ = 0.f, dz_ = 0.f;
        DS(0) += SH_C0 * g;
        if (DEG > 0) {
          DS(1) += -SH_C1 * y * g; DS(2) += SH_C1 * z * g; DS(3) += -SH_C1 * x * g;
          dx_ = -SH_C1 * S(3); dy_ = -SH_C1 * S(1); dz_ = SH_C1 * S(2);
          if (DEG > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            DS(4) += SH_C2[0] * xy * g; DS(5) += SH_C2[1] * yz * g; DS(6) += SH_C2[2] * (2.f * zz - xx - yy) * g;
            DS(7) += SH_C2[3] * xz * g; DS(8) += SH_C2[4] * (xx - yy) * g;
            dx_ += SH_C2[0] * y * S(4) + SH_C2[2] * 2.f * -x * S(6) + SH_C2[3] * z * S(7) + SH_C2[4] * 2.f * x * S(8);
            dy_ += SH_C2[0] * x * S(4) + SH_C2[1] * z * S(5) + SH_C2[2] * 2.f * -y * S(6) + SH_C2[4] * 2.f * -y * S(8);
            dz_ += SH_C2[1] * y * S(5) + SH_C2[2] * 2.f * 2.f * z * S(6) + SH_C2[3] * x * S(7);
            if (DEG > 2) {
              DS(9) += SH_C3[0] * y * (3.f * xx - yy) * g; DS(10) += SH_C3[1] * xy * z * g;
              DS(11) += SH_C3[2] * y * (4.f * zz - xx - yy) * g; DS(12) += SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy) * g;
              DS(13) += SH_C3[4] * x * (4.f * zz - xx - yy) * g; DS(14) += SH_C3[5] * z * (xx - yy) * g;
              DS(15) += SH_C3[6] * x * (xx - 3.f * yy) * g;
              dx_ += SH_C3[0] * S(9) * 3.f * 2.f * xy + SH_C3[1] * S(10) * yz + SH_C3[2] * S(11) * -2.f * xy +
                     SH_C3[3] * S(12) * -3.f * 2.f * xz + SH_C3[4] * S(13) * (-3.f * xx + 4.f * zz - yy) +
                     SH_C3[5] * S(14) * 2.f * xz + SH_C3[6] * S(15) * 3.f * (xx - yy);
              dy_ += SH_C3[0] * S(9) * 3.f * (xx - yy) + SH_C3[1] * S(10) * xz + SH_C3[2] * S(11) * (-3.f * yy + 4.f * zz - xx) +
                     SH_C3[3] * S(12) * -3.f * 2.f * yz + SH_C3[4] * S(13) * -2.f * xy + SH_C3[5] * S(14) * -2.f * yz +
                     SH_C3[6] * S(15) * -3.f * 2.f * xy;
              dz_ += SH_C3[1] * S(10) * xy + SH_C3[2] * S(11) * 4.f * 2.f * yz + SH_C3[3] * S(12) * 3.f * (2.f * zz - xx - yy) +
                     SH_C3[4] * S(13) * 4.f * 2.f * xz + SH_C3[5] * S(14) * (xx - yy);
            }
          }
        }
#undef S
#undef DS
        dL_ddir[0] += dx_ * g; dL_ddir[1] += dy_ * g; dL_ddir[2] += dz_ * g;
      }
      float dmd[3];
      dnormvdv(dir_orig, dL_ddir, dmd);
      dm[0] += dmd[0]; dm[1] += dmd[1]; dm[2] += dmd[2];
    }
    // normalised -> raw mean (scale-invariant chain rule) and the depth colour's own path
    dmean[0] += vp.s * dm[0]; dmean[1] += vp.s * dm[1]; dmean[2] += vp.s * dm[2];
    if (vw.depth_mode != B200S_DEPTH_NONE) {
      const float gz = g2v.y;
      const float z = __fadd_rn(__fmaf_rn(vp.daff[2], mraw[2], __fmaf_rn(vp.daff[0], mraw[0], __fmul_rn(vp.daff[1], mraw[1]))), vp.daff[3]);
      float dz = gz;
      if (vw.depth_mode == B200S_DEPTH_DISPARITY) dz = -gz / (z * z);
      else if (vw.depth_mode == B200S_DEPTH_LOG) {
        const float lo = fminf(z, vp.dnear);
        dz = (z < vp.dnear && lo > vp.dfar) ? gz / lo : 0.f;
      }
      dmean[0] += vp.daff[0] * dz; dmean[1] += vp.daff[1] * dz; dmean[2] += vp.daff[2] * dz;
    }
  }
  }

  if (RAW) {
    // ---- the adapter's chain rule; one value per (channel plane, pixel): consecutive threads write consecutive addresses ----
    const int hw = sc.raw_h * sc.raw_w;
    const size_t sv = (size_t)scene * sc.raw_views + raw_cv;
    float dh[10], ddepth;
    cook_backward(hraw, s_cam, ck, sc.raw_w, sc.raw_h, sc.raw_scale_min, sc.raw_scale_max, dmean, dcov, dop, dh, ddepth);
    float* dst = gin.dL_draw_head + sv * RAW_CH * hw + raw_p0 + tid;
#pragma unroll
    for (int k = 0; k < 10; k++) dst[(size_t)k * hw] = dh[k];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
      float din[9];
      cook_sh_backward(dcol + ch * NC, s_cam, din);
#pragma unroll
      for (int k = 0; k < 9; k++) dst[(size_t)(10 + 9 * ch + k) * hw] = din[k];
    }
    gin.dL_draw_depth[sv * hw + raw_p0 + tid] = ddepth;
    return;
  }
  // ---- coalesced writes through shared memory ------------------------------------------------------
  __syncthreads();
  float* s_dmean = smem;
  float* s_dcov = s_dmean + PRE_THREADS * 3;
  float* s_dcol = s_dcov + PRE_THREADS * a.cov_floats;
  s_dmean[tid * 3] = dmean[0]; s_dmean[tid * 3 + 1] = dmean[1]; s_dmean[tid * 3 + 2] = dmean[2];
  {
    float* cp = s_dcov + tid * a.cov_floats;
    if (a.cov_floats == 6) { for (int k = 0; k < 6; k++) cp[k] = dcov[k]; }
    else { cp[0] = dcov[0]; cp[1] = dcov[1]; cp[2] = dcov[2]; cp[3] = 0.f; cp[4] = dcov[3]; cp[5] = dcov[4]; cp[6] = 0.f; cp[7] = 0.f; cp[8] = dcov[5]; }
  }
  {
    float* d = s_dcol + tid * a.col_stride;
    if (NC == 0) { d[0] = dcol[0]; d[1] = dcol[1]; d[2] = dcol[2]; }
    else {
      for (int k = 0; k < a.col_floats; k++) d[k] = 0.f;  // coefficients above the active degree get zero gradient
#pragma unroll
      for (int ch = 0; ch < 3; ch++)
#pragma unroll
        for (int k = 0; k < NC; k++) d[ch * cstride + k * kstride] = dcol[ch * NC + k];
    }
  }
  __syncthreads();
  stage_out_bwd<MC>(gin.dL_dmeans + g0 * 3, s_dmean, n * 3, 3, 3);
  stage_out_bwd<MC>(gin.dL_dcovariances + g0 * a.cov_floats, s_dcov, n * a.cov_floats, a.cov_floats, a.cov_floats);
  float* col_dst = NC == 0 ? gin.dL_dcolors : gin.dL_dharmonics;
  if (col_dst) stage_out_bwd<MC>(col_dst + g0 * a.col_floats, s_dcol, n * a.col_floats, a.col_floats, a.col_stride);
  if (tid < n) {
    if (MC) multimem_red_add(gin.dL_dopacities + g0 + tid, dop);
    else gin.dL_dopacities[g0 + tid] = dop;
  }
  if (MC) __threadfence_system();  // the reductions are performed at the peers before the kernel is seen as complete
}

cudaError_t launch_preprocess_bwd(const B200sScene& sc, const B200sViews& vw, const B200sPlan& plan, const char* saved, char* scratch,
                                  const B200sGradIn& gin, cudaStream_t stream) {
  PreBwdArgs a;
  a.N = sc.num_gaussians; a.VV = vw.num_views; a.H = vw.height; a.W = vw.width;
  a.cov_floats = sc.cov_layout == B200S_COV_UPPER6 ? 6 : 9;
  a.col_floats = sc.colors_precomp ? 3 : 3 * sc.sh_coeffs;
  a.col_stride = a.col_floats | 1;
  if (sc.raw_head) { a.cov_floats = 9; a.col_floats = 27; a.col_stride = 27; }
  a.rec = reinterpret_cast<const Rec*>(saved + plan.off_rec);
  a.grad_rec = reinterpret_cast<const float*>(scratch + plan.off_grad_rec);
  a.dL_dmeans2D = gin.dL_dmeans2D;
  a.overflow = &reinterpret_cast<const B200sStatus*>(saved + plan.off_status)->overflow;
  const int chunks = (sc.num_gaussians + PRE_THREADS - 1) / PRE_THREADS;
  a.chunk_begin = gin.chunk_begin > 0 ? gin.chunk_begin : 0;
  a.chunk_count = gin.chunk_count > 0 ? gin.chunk_count : chunks - a.chunk_begin;
  a.chunk_repeat = gin.chunk_repeat > 1 ? gin.chunk_repeat : 1;
  a.chunk_stride = a.chunk_repeat > 1 ? gin.chunk_stride : 0;
  a.chunks_total = chunks;
  if (a.chunk_repeat == 1 && a.chunk_begin + a.chunk_count > chunks) a.chunk_count = chunks - a.chunk_begin;
  if (a.chunk_repeat > 1 && (a.chunk_stride < a.chunk_count || a.chunk_count <= 0)) return cudaErrorInvalidValue;
  const int blocks = a.chunk_count * a.chunk_repeat * sc.num_scenes;
  if (blocks <= 0) return cudaSuccess;
  const size_t smem = sc.raw_head ? (size_t)PRE_THREADS * RAW_PLANES * sizeof(float) : (size_t)PRE_THREADS * (3 + a.cov_floats + a.col_stride) * sizeof(float);
  const int nc = sc.colors_precomp ? 0 : (sc.sh_degree + 1) * (sc.sh_degree + 1);
  cudaError_t e = cudaSuccess;
  stage_mark(B200S_STAGE_PRE_BWD, stream);
  count_launches(1);
#define LAUNCH(NCV)                                                                                                        \
  {                                                                                                                        \
    if (gin.multicast) {                                                                                                   \
      e = cudaFuncSetAttribute(preprocess_bwd_kernel<NCV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      if (e != cudaSuccess) return e;                                                                                      \
      preprocess_bwd_kernel<NCV, true><<<blocks, PRE_THREADS, smem, stream>>>(sc, vw, gin, a);                            \
    } else {                                                                                                               \
      e = cudaFuncSetAttribute(preprocess_bwd_kernel<NCV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);\
      if (e != cudaSuccess) return e;                                                                                      \
      preprocess_bwd_kernel<NCV, false><<<blocks, PRE_THREADS, smem, stream>>>(sc, vw, gin, a);                           \
    }                                                                                                                      \
  }
  if (sc.raw_head) {  // raw scenes: degree 2, no multicast variant
    if (gin.multicast || nc != 9) return cudaErrorInvalidValue;
    e = cudaFuncSetAttribute(preprocess_bwd_kernel<9, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    preprocess_bwd_kernel<9, false, true><<<blocks, PRE_THREADS, smem, stream>>>(sc, vw, gin, a);
    return cudaGetLastError();
  }
  switch (nc) {
    case 0: LAUNCH(0); break;
    case 1: LAUNCH(1); break;
    case 4: LAUNCH(4); break;
    case 9: LAUNCH(9); break;
    case 16: LAUNCH(16); break;
    default: return cudaErrorInvalidValue;
  }
#undef LAUNCH
  return cudaGetLastError();
}

}  // namespace b200s
