// common.cuh -- shared device helpers, record layout and algorithm constants (sm_100a only).
//
// Arithmetic contract (DESIGN.md "bit-exact chain"): everything that decides sort keys and tile
// ranges (camera-space depth, pixel xy, cov2D, radius, rect) is written with explicit
// __fmaf_rn/__fmul_rn/__fadd_rn/__fsub_rn so that ptxas cannot re-contract it; the contraction
// spelled out is the one nvcc applies to the published algorithm's expression forms
// (sums of products fuse: p1+p2+p3 -> fma(p3, fma(p1, mul(p2))); differences do not fuse).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200splat.h"

#if !defined(__CUDA_ARCH__) || (__CUDA_ARCH__ >= 1000)
#else
#error "b200splat kernels are written for sm_100a only"
#endif

namespace b200s {

constexpr int TILE_X = 16, TILE_Y = 16, TILE_PIX = 256;
constexpr int PRE_THREADS = 256;  // Gaussians per preprocess ticket

// named constants of the algorithm (cf. oracle/splat_oracle.c)
constexpr float NEAR_CULL = 0.2f;
constexpr float DILATION = 0.3f;
constexpr float FOV_CLAMP = 1.3f;
constexpr float ALPHA_MAX = 0.99f;
constexpr float ALPHA_MIN = 1.0f / 255.0f;
constexpr float T_MIN = 0.0001f;
constexpr float LAMBDA_FLOOR = 0.1f;

constexpr float SH_C0 = 0.28209479177387814f;
constexpr float SH_C1 = 0.4886025119029199f;
__device__ constexpr float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                       -1.0925484305920792f, 0.5462742152960396f};
__device__ constexpr float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                       0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                                       -0.5900435899266435f};

// ---- projected record: 64 bytes per (view, Gaussian), written by preprocess ---------------------
// q0 = (x, y, ex, ey) [cull record]   q1 = (conicA, conicB, conicC, opacity)   q2 = (r, g, b, zc)
// q3 = (depth, radius:int, rect packed u32 [minx | miny<<8 | maxx<<16 | maxy<<24], flags:u32)
// Compositing reads q0..q2 (48 B); q3 is for the backward preprocess, radii output and tests.
// (ex, ey) is the conservative half-extent of the alpha >= 1/255 region, used for sub-tile culling;
// negative when the Gaussian can never reach alpha >= 1/255.
struct __align__(16) Rec {
  float4 q0, q1, q2, q3;
};
static_assert(sizeof(Rec) == 64, "record must be 64 bytes");
constexpr uint32_t REC_FLAG_CLAMP_R = 1u, REC_FLAG_CLAMP_G = 2u, REC_FLAG_CLAMP_B = 4u;

// ---- gradient record: 48 bytes per (view, Gaussian), accumulated by the compositing backward ----
// the pixel moments of q = G * dL/dalpha and the colour gradients:
// (S_x, S_y, S_xx, S_xy, S_yy, S_1, dL/dr, dL/dg, dL/db, dL/dzc, pad, pad), S_f = sum over pixels of q * f(dx, dy).
// The projection backward turns them into dL/dmean2D = opacity * half_extent * (-A S_x - B S_y, -C S_y - B S_x),
// dL/dconic = -opacity / 2 * (S_xx, S_xy, S_yy), dL/dopacity = S_1 (composite.cu, preprocess_bwd.cu).
constexpr int GREC_FLOATS = 12;

// per-view camera block staged in shared memory by the per-Gaussian kernels
struct ViewParams {
  float view[16];
  float proj[16];
  float campos[3];
  float tanfovx, tanfovy, focal_x, focal_y;
  float s, s2;       // scale-invariant factors
  float daff[4];     // depth affine row
  float dnear, dfar;
  float bg[3];
  int scene;
};

__device__ __forceinline__ void load_view_params(ViewParams& vp, const B200sViews& v, int view, int H, int W) {
  for (int i = 0; i < 16; i++) { vp.view[i] = v.viewmatrix[view * 16 + i]; vp.proj[i] = v.projmatrix[view * 16 + i]; }
  for (int i = 0; i < 3; i++) { vp.campos[i] = v.campos[view * 3 + i]; vp.bg[i] = v.background[view * 3 + i]; }
  vp.tanfovx = v.tanfov[view * 2]; vp.tanfovy = v.tanfov[view * 2 + 1];
  vp.focal_y = (float)H / (2.0f * vp.tanfovy);
  vp.focal_x = (float)W / (2.0f * vp.tanfovx);
  vp.s = v.scale ? v.scale[view * 2] : 1.0f;
  vp.s2 = v.scale ? v.scale[view * 2 + 1] : 1.0f;
  for (int i = 0; i < 4; i++) vp.daff[i] = v.depth_affine ? v.depth_affine[view * 4 + i] : 0.f;
  vp.dnear = v.depth_clamp ? v.depth_clamp[view * 2] : 0.f;
  vp.dfar = v.depth_clamp ? v.depth_clamp[view * 2 + 1] : 0.f;
  vp.scene = v.scene_index[view];
}

// The camera blocks of `count` consecutive views [view0, view0 + count) loaded by the whole CTA at once (one field per
// thread: one memory latency for the group instead of ~50 dependent-issue loads by one thread per view).  Same values,
// same expressions as load_view_params.  Callers synchronise before reading.
constexpr int VIEW_GROUP = 8;
__device__ __forceinline__ void load_view_group(ViewParams* vps, const B200sViews& v, int view0, int count, int H, int W) {
  for (int t = threadIdx.x; t < count * 64; t += blockDim.x) {
    const int f = t & 63, view = view0 + (t >> 6);
    ViewParams& p = vps[t >> 6];
    if (f < 16) p.view[f] = v.viewmatrix[view * 16 + f];
    else if (f < 32) p.proj[f - 16] = v.projmatrix[view * 16 + f - 16];
    else if (f < 35) p.campos[f - 32] = v.campos[view * 3 + f - 32];
    else if (f < 38) p.bg[f - 35] = v.background[view * 3 + f - 35];
    else if (f == 38) { const float t_ = v.tanfov[view * 2]; p.tanfovx = t_; p.focal_x = (float)W / (2.0f * t_); }
    else if (f == 39) { const float t_ = v.tanfov[view * 2 + 1]; p.tanfovy = t_; p.focal_y = (float)H / (2.0f * t_); }
    else if (f == 40) p.s = v.scale ? v.scale[view * 2] : 1.0f;
    else if (f == 41) p.s2 = v.scale ? v.scale[view * 2 + 1] : 1.0f;
    else if (f < 46) p.daff[f - 42] = v.depth_affine ? v.depth_affine[view * 4 + f - 42] : 0.f;
    else if (f == 46) p.dnear = v.depth_clamp ? v.depth_clamp[view * 2] : 0.f;
    else if (f == 47) p.dfar = v.depth_clamp ? v.depth_clamp[view * 2 + 1] : 0.f;
    else if (f == 48) p.scene = v.scene_index[view];
  }
}
// [first, last] view index of `scene` (its views need not be contiguous: views of other scenes in between are skipped by
// the callers).  Shared words lo / hi are initialised by the caller before a barrier; call, then barrier, then read.
__device__ __forceinline__ void scene_view_range(const B200sViews& v, int num_views, int scene, int* lo, int* hi) {
  for (int t = threadIdx.x; t < num_views; t += blockDim.x)
    if (v.scene_index[t] == scene) { atomicMin(lo, t); atomicMax(hi, t); }
}

// p1 + p2 + p3 with the contraction nvcc applies (second product plain, first fused, third fused)
__device__ __forceinline__ float dot3c(float a, float b, float c, float d, float e, float f) {
  return __fmaf_rn(e, f, __fmaf_rn(a, b, __fmul_rn(c, d)));
}
// column-major 4x4 (transposed storage) times (p,1): rows 0..2 / row r
__device__ __forceinline__ float xform_row(const float* m, int r, float x, float y, float z) {
  return __fadd_rn(__fmaf_rn(m[8 + r], z, __fmaf_rn(m[r], x, __fmul_rn(m[4 + r], y))), m[12 + r]);
}
__device__ __forceinline__ float ndc2pix(float v, int S) {
  return (float)(__fma_rn((double)v + 1.0, (double)S, -1.0) * 0.5);
}

struct Cov2D {
  float t[3];
  float xmul, ymul;
  float T0[3], T1[3];  // T[0][r], T[1][r] of T = W*J in column-major form
  float a, b, c;       // dilated cov2D
};

// EWA projection of a 3D covariance (upper triangle c6, already scale-normalised) -- bit-exact chain.
__device__ __forceinline__ void compute_cov2d(const float mean[3], const float c6[6], const ViewParams& vp, Cov2D& o) {
  float tx = xform_row(vp.view, 0, mean[0], mean[1], mean[2]);
  float ty = xform_row(vp.view, 1, mean[0], mean[1], mean[2]);
  const float tz = xform_row(vp.view, 2, mean[0], mean[1], mean[2]);
  const float limx = __fmul_rn(FOV_CLAMP, vp.tanfovx), limy = __fmul_rn(FOV_CLAMP, vp.tanfovy);
  const float txtz = __fdiv_rn(tx, tz), tytz = __fdiv_rn(ty, tz);
  o.xmul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
  o.ymul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
  tx = __fmul_rn(fminf(limx, fmaxf(-limx, txtz)), tz);
  ty = __fmul_rn(fminf(limy, fmaxf(-limy, tytz)), tz);
  o.t[0] = tx; o.t[1] = ty; o.t[2] = tz;
  const float tz2 = __fmul_rn(tz, tz);
  const float J00 = __fdiv_rn(vp.focal_x, tz);
  const float J02 = __fdiv_rn(-__fmul_rn(vp.focal_x, tx), tz2);
  const float J11 = __fdiv_rn(vp.focal_y, tz);
  const float J12 = __fdiv_rn(-__fmul_rn(vp.focal_y, ty), tz2);
#pragma unroll
  for (int r = 0; r < 3; r++) {
    // W[c][r] = view[4r + c]
    o.T0[r] = __fmaf_rn(vp.view[4 * r + 2], J02, __fmul_rn(vp.view[4 * r + 0], J00));
    o.T1[r] = __fmaf_rn(vp.view[4 * r + 2], J12, __fmul_rn(vp.view[4 * r + 1], J11));
  }
  const float V[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
  float A0[3], A1[3];  // A[c][0], A[c][1]
#pragma unroll
  for (int c = 0; c < 3; c++) {
    A0[c] = dot3c(o.T0[0], V[c][0], o.T0[1], V[c][1], o.T0[2], V[c][2]);
    A1[c] = dot3c(o.T1[0], V[c][0], o.T1[1], V[c][1], o.T1[2], V[c][2]);
  }
  o.a = __fadd_rn(dot3c(A0[0], o.T0[0], A0[1], o.T0[1], A0[2], o.T0[2]), DILATION);
  o.b = dot3c(A1[0], o.T0[0], A1[1], o.T0[1], A1[2], o.T0[2]);
  o.c = __fadd_rn(dot3c(A1[0], o.T1[0], A1[1], o.T1[1], A1[2], o.T1[2]), DILATION);
}

__device__ __forceinline__ int imin(int a, int b) { return a < b ? a : b; }
__device__ __forceinline__ int imax(int a, int b) { return a > b ? a : b; }

__device__ __forceinline__ void get_rect(float px, float py, int r, int gx, int gy, int& x0, int& y0, int& x1, int& y1) {
  const float rf = (float)r;
  x0 = imin(gx, imax(0, (int)__fmul_rn(__fsub_rn(px, rf), 0.0625f)));
  y0 = imin(gy, imax(0, (int)__fmul_rn(__fsub_rn(py, rf), 0.0625f)));
  x1 = imin(gx, imax(0, (int)__fmul_rn(__fsub_rn(__fadd_rn(__fadd_rn(px, rf), 16.0f), 1.0f), 0.0625f)));
  y1 = imin(gy, imax(0, (int)__fmul_rn(__fsub_rn(__fadd_rn(__fadd_rn(py, rf), 16.0f), 1.0f), 0.0625f)));
}

// Gaussian falloff exponent as nvcc contracts the published expression
__device__ __forceinline__ float gauss_power(float A, float B, float C, float dx, float dy) {
  return __fsub_rn(__fmul_rn(-0.5f, __fmaf_rn(__fmul_rn(A, dx), dx, __fmul_rn(__fmul_rn(C, dy), dy))),
                   __fmul_rn(__fmul_rn(B, dx), dy));
}

// Covariance in the extension's upper-triangle order from either layout, times s2.
__device__ __forceinline__ void load_cov6(const float* cov, int layout, long long gi, float s2, float c6[6]) {
  if (layout == B200S_COV_UPPER6) {
    const float* p = cov + gi * 6;
#pragma unroll
    for (int k = 0; k < 6; k++) c6[k] = __fmul_rn(p[k], s2);
  } else {
    const float* p = cov + gi * 9;
    c6[0] = __fmul_rn(p[0], s2); c6[1] = __fmul_rn(p[1], s2); c6[2] = __fmul_rn(p[2], s2);
    c6[3] = __fmul_rn(p[4], s2); c6[4] = __fmul_rn(p[5], s2); c6[5] = __fmul_rn(p[8], s2);
  }
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u64(uint64_t* p, uint64_t v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- TMA (bulk async copy engine) staging of contiguous global ranges into shared memory ---------
// One thread arms an mbarrier with the byte count and issues cp.async.bulk (SASS: UBLKCP); the copy
// engine moves the data while the CTA's threads do nothing, and everyone waits on the barrier's phase.
// Requirements: 16-byte aligned source, destination and size.
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
      "r"(phase)
      : "memory");
}
__device__ __forceinline__ bool tma_ok(const void* src, int bytes) { return ((((uintptr_t)src) | (uintptr_t)bytes) & 15) == 0 && bytes > 0; }

// counters block layout (u32 indices)
constexpr int CNT_PRE_TICKET = 0;
constexpr int CNT_SORT_TILE0 = 8;  // + pass (8 passes)
constexpr int CNT_BIN_CLASS_COUNT = 16;  // + class (4): bins per size class of the BINNED mode
constexpr int CNT_BIN_CLASS_NEXT = 20;   // + class (4): their fetch counters
constexpr int CNT_WORDS = 64;

}  // namespace b200s
