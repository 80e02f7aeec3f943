/*
 * upstream_style.cu -- GPU COMPARATOR, NOT PRODUCT CODE.
 *
 * The rasterizer DepthSplat calls (`diff_gaussian_rasterization`, requirements.txt:23 of the reference) is a
 * third-party CUDA extension that is neither under /root/reference nor installed on the GPU box, so the
 * "reference CUDA" column of the benchmark cannot be the real thing.  This file is a straight
 * GPU restatement of that extension's PUBLISHED DESIGN (3D Gaussian Splatting, Kerbl et al. 2023), written from
 * the stage description in SURVEY.md 7.4 / appendix B and from oracle/splat_oracle.c -- one view per call:
 *
 *   preprocess        one thread per Gaussian, scattered per-attribute arrays
 *   InclusiveSum      cub::DeviceScan, then a blocking device->host copy of the pair count
 *   duplicateWithKeys one thread per Gaussian
 *   SortPairs         cub::DeviceRadixSort over 32 + msb(tiles) key bits
 *   identifyRanges    one thread per pair
 *   render            one 16x16 CTA per tile, 256 entries staged per round through shared memory behind
 *                     block-wide barriers, features gathered from global memory per blended entry
 *   render backward   same staging back to front, one atomicAdd per (pixel, Gaussian, gradient component)
 *   preprocess bwd    cov2D backward kernel + projection/SH backward kernel, one thread per Gaussian
 *
 * bench.py times it next to the product ("gpu_baseline") so that the speed-up of the B200-native design is
 * measured against the upstream design on the same GPU; tests/test_gpu_baseline.py checks that it computes the
 * same images and gradients.  It uses CUB and none of the product's kernels; the product never loads it.
 */
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int BX = 16, BY = 16, BLOCK = BX * BY;
constexpr float NEAR_CULL = 0.2f, DILATION = 0.3f, FOV_CLAMP = 1.3f, ALPHA_MAX = 0.99f, ALPHA_MIN = 1.0f / 255.0f, T_MIN = 0.0001f;

__device__ const float SH_C0 = 0.28209479177387814f;
__device__ const float SH_C1 = 0.4886025119029199f;
__device__ const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f, -1.0925484305920792f, 0.5462742152960396f};
__device__ const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                                   -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

}  // namespace

extern "C" {

typedef struct {
  int H, W, D, M;  // image, SH degree, SH coefficients per channel
  float tanfovx, tanfovy;
  const float *view, *proj, *campos, *bg;  // DEVICE pointers, as the extension takes them; matrices in the transposed storage the reference passes
} UpsView;

typedef struct {
  // geometry (per Gaussian)
  float* depths; float2* xy; float4* conic_opacity; float* rgb; uint8_t* clamped; int* radii;
  uint32_t* tiles_touched; uint32_t* offsets; void* scan_temp; size_t scan_temp_bytes;
  // binning (per pair)
  uint64_t* keys_unsorted; uint64_t* keys; uint32_t* vals_unsorted; uint32_t* vals; void* sort_temp; size_t sort_temp_bytes;
  // image
  uint2* ranges; float* final_T; uint32_t* n_contrib;
} UpsState;

}  // extern "C"

namespace {

struct Cam {
  int H, W, D, M, gx, gy;
  float tanfovx, tanfovy, focal_x, focal_y;
  const float *view, *proj, *campos, *bg;
};

Cam make_cam(const UpsView& v) {
  Cam c;
  c.H = v.H; c.W = v.W; c.D = v.D; c.M = v.M;
  c.gx = (v.W + BX - 1) / BX; c.gy = (v.H + BY - 1) / BY;
  c.tanfovx = v.tanfovx; c.tanfovy = v.tanfovy;
  c.focal_y = (float)v.H / (2.0f * v.tanfovy); c.focal_x = (float)v.W / (2.0f * v.tanfovx);
  c.view = v.view; c.proj = v.proj; c.campos = v.campos; c.bg = v.bg;
  return c;
}

__device__ __forceinline__ float3 xform4x3(const float3 p, const float* m) {
  return make_float3(m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12], m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
                     m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14]);
}
__device__ __forceinline__ float4 xform4x4(const float3 p, const float* m) {
  return make_float4(m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12], m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
                     m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14], m[3] * p.x + m[7] * p.y + m[11] * p.z + m[15]);
}
__device__ __forceinline__ float ndc2pix(float v, int S) { return ((v + 1.0) * S - 1.0) * 0.5; }

__device__ __forceinline__ void get_rect(const float2 p, int r, const Cam& c, int2& mn, int2& mx) {
  mn.x = min(c.gx, max(0, (int)((p.x - r) / BX)));
  mn.y = min(c.gy, max(0, (int)((p.y - r) / BY)));
  mx.x = min(c.gx, max(0, (int)((p.x + r + BX - 1) / BX)));
  mx.y = min(c.gy, max(0, (int)((p.y + r + BY - 1) / BY)));
}

// camera-space point (FoV-clamped), T = (J W)^T columns and the dilated 2D covariance
struct Cov2D {
  float3 t;
  float xmul, ymul;
  float T0[3], T1[3];  // T[0][r], T[1][r]
  float a, b, c;
};

__device__ void compute_cov2d(const float3 mean, const float* c3, const Cam& cam, Cov2D& o) {
  float3 t = xform4x3(mean, cam.view);
  const float limx = FOV_CLAMP * cam.tanfovx, limy = FOV_CLAMP * cam.tanfovy;
  const float txtz = t.x / t.z, tytz = t.y / t.z;
  o.xmul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
  o.ymul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
  t.x = fminf(limx, fmaxf(-limx, txtz)) * t.z;
  t.y = fminf(limy, fmaxf(-limy, tytz)) * t.z;
  o.t = t;
  const float J00 = cam.focal_x / t.z, J02 = -(cam.focal_x * t.x) / (t.z * t.z);
  const float J11 = cam.focal_y / t.z, J12 = -(cam.focal_y * t.y) / (t.z * t.z);
  const float* vm = cam.view;
#pragma unroll
  for (int r = 0; r < 3; r++) {
    // W[c][r] = view[4r + c]
    o.T0[r] = vm[4 * r + 0] * J00 + vm[4 * r + 2] * J02;
    o.T1[r] = vm[4 * r + 1] * J11 + vm[4 * r + 2] * J12;
  }
  const float V[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
  float A0[3], A1[3];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    A0[c] = o.T0[0] * V[c][0] + o.T0[1] * V[c][1] + o.T0[2] * V[c][2];
    A1[c] = o.T1[0] * V[c][0] + o.T1[1] * V[c][1] + o.T1[2] * V[c][2];
  }
  o.a = A0[0] * o.T0[0] + A0[1] * o.T0[1] + A0[2] * o.T0[2] + DILATION;
  o.b = A1[0] * o.T0[0] + A1[1] * o.T0[1] + A1[2] * o.T0[2];
  o.c = A1[0] * o.T1[0] + A1[1] * o.T1[1] + A1[2] * o.T1[2] + DILATION;
}

__device__ void sh_to_rgb(int deg, const float3 mean, const Cam& cam, const float* sh /*[M,3]*/, float* rgb, uint8_t* clamped) {
  float3 dir = make_float3(mean.x - cam.campos[0], mean.y - cam.campos[1], mean.z - cam.campos[2]);
  const float len = sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
  const float x = dir.x / len, y = dir.y / len, z = dir.z / len;
  for (int ch = 0; ch < 3; ch++) {
#define S(k) sh[(k) * 3 + ch]
    float r = SH_C0 * S(0);
    if (deg > 0) {
      r = r - SH_C1 * y * S(1) + SH_C1 * z * S(2) - SH_C1 * x * S(3);
      if (deg > 1) {
        const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        r = r + SH_C2[0] * xy * S(4) + SH_C2[1] * yz * S(5) + SH_C2[2] * (2.0f * zz - xx - yy) * S(6) + SH_C2[3] * xz * S(7) +
            SH_C2[4] * (xx - yy) * S(8);
        if (deg > 2) {
          r = r + SH_C3[0] * y * (3.0f * xx - yy) * S(9) + SH_C3[1] * xy * z * S(10) + SH_C3[2] * y * (4.0f * zz - xx - yy) * S(11) +
              SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * S(12) + SH_C3[4] * x * (4.0f * zz - xx - yy) * S(13) +
              SH_C3[5] * z * (xx - yy) * S(14) + SH_C3[6] * x * (xx - 3.0f * yy) * S(15);
        }
      }
    }
#undef S
    r += 0.5f;
    clamped[ch] = r < 0.f;
    rgb[ch] = fmaxf(r, 0.f);
  }
}

// ---------------------------------------------------------------------------------------------------
__global__ void preprocess_kernel(int P, const Cam cam, const float* __restrict__ means3D, const float* __restrict__ shs,
                                  const float* __restrict__ opacities, const float* __restrict__ cov3D, UpsState st) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  st.radii[i] = 0;
  st.tiles_touched[i] = 0;
  const float3 p = make_float3(means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]);
  const float3 pv = xform4x3(p, cam.view);
  if (pv.z <= NEAR_CULL) return;
  const float4 ph = xform4x4(p, cam.proj);
  const float pw = 1.0f / (ph.w + 0.0000001f);
  const float ppx = ph.x * pw, ppy = ph.y * pw;
  Cov2D q;
  compute_cov2d(p, cov3D + 6 * (size_t)i, cam, q);
  const float det = q.a * q.c - q.b * q.b;
  if (det == 0.0f) return;
  const float det_inv = 1.f / det;
  const float mid = 0.5f * (q.a + q.c);
  const float disc = sqrtf(fmaxf(0.1f, mid * mid - det));
  const float my_radius = ceilf(3.f * sqrtf(fmaxf(mid + disc, mid - disc)));
  const float2 pt = make_float2(ndc2pix(ppx, cam.W), ndc2pix(ppy, cam.H));
  int2 mn, mx;
  get_rect(pt, (int)my_radius, cam, mn, mx);
  if ((mx.x - mn.x) * (mx.y - mn.y) == 0) return;
  if (shs) sh_to_rgb(cam.D, p, cam, shs + (size_t)i * cam.M * 3, st.rgb + 3 * (size_t)i, st.clamped + 3 * (size_t)i);
  st.depths[i] = pv.z;
  st.radii[i] = (int)my_radius;
  st.xy[i] = pt;
  st.conic_opacity[i] = make_float4(q.c * det_inv, -q.b * det_inv, q.a * det_inv, opacities[i]);
  st.tiles_touched[i] = (uint32_t)((mx.y - mn.y) * (mx.x - mn.x));
}

__global__ void duplicate_kernel(int P, const Cam cam, UpsState st) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P || st.radii[i] <= 0) return;
  uint32_t off = i == 0 ? 0 : st.offsets[i - 1];
  int2 mn, mx;
  get_rect(st.xy[i], st.radii[i], cam, mn, mx);
  const uint32_t dbits = __float_as_uint(st.depths[i]);
  for (int y = mn.y; y < mx.y; y++)
    for (int x = mn.x; x < mx.x; x++) {
      st.keys_unsorted[off] = ((uint64_t)(y * cam.gx + x) << 32) | dbits;
      st.vals_unsorted[off] = (uint32_t)i;
      off++;
    }
}

__global__ void ranges_kernel(int64_t R, const uint64_t* __restrict__ keys, uint2* ranges) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  const uint32_t cur = (uint32_t)(keys[i] >> 32);
  if (i == 0) ranges[cur].x = 0;
  else {
    const uint32_t prev = (uint32_t)(keys[i - 1] >> 32);
    if (cur != prev) { ranges[prev].y = (uint32_t)i; ranges[cur].x = (uint32_t)i; }
  }
  if (i == R - 1) ranges[cur].y = (uint32_t)R;
}

__global__ void __launch_bounds__(BLOCK) render_kernel(const Cam cam, const UpsState st, const float* __restrict__ feat, float* __restrict__ out) {
  __shared__ uint32_t s_id[BLOCK];
  __shared__ float2 s_xy[BLOCK];
  __shared__ float4 s_co[BLOCK];
  const int px = blockIdx.x * BX + threadIdx.x, py = blockIdx.y * BY + threadIdx.y;
  const int tr = threadIdx.y * BX + threadIdx.x;
  const bool inside = px < cam.W && py < cam.H;
  const int pid = py * cam.W + px;
  const float pfx = (float)px, pfy = (float)py;
  const uint2 range = st.ranges[blockIdx.y * cam.gx + blockIdx.x];
  const int rounds = (range.y - range.x + BLOCK - 1) / BLOCK;
  int todo = range.y - range.x;
  bool done = !inside;
  float T = 1.f, C[3] = {0.f, 0.f, 0.f};
  uint32_t contributor = 0, last = 0;
  for (int i = 0; i < rounds; i++, todo -= BLOCK) {
    if (__syncthreads_count(done) == BLOCK) break;
    const int progress = i * BLOCK + tr;
    if (range.x + progress < range.y) {
      const uint32_t id = st.vals[range.x + progress];
      s_id[tr] = id; s_xy[tr] = st.xy[id]; s_co[tr] = st.conic_opacity[id];
    }
    __syncthreads();
    for (int j = 0; !done && j < min(BLOCK, todo); j++) {
      contributor++;
      const float2 xy = s_xy[j];
      const float dx = xy.x - pfx, dy = xy.y - pfy;
      const float4 co = s_co[j];
      const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
      if (power > 0.f) continue;
      const float alpha = fminf(ALPHA_MAX, co.w * expf(power));
      if (alpha < ALPHA_MIN) continue;
      const float test_T = T * (1.f - alpha);
      if (test_T < T_MIN) { done = true; continue; }
      const float* f = feat + 3 * (size_t)s_id[j];
      for (int ch = 0; ch < 3; ch++) C[ch] += f[ch] * alpha * T;
      T = test_T;
      last = contributor;
    }
  }
  if (inside) {
    st.final_T[pid] = T;
    st.n_contrib[pid] = last;
    const size_t HW = (size_t)cam.H * cam.W;
    for (int ch = 0; ch < 3; ch++) out[ch * HW + pid] = C[ch] + T * cam.bg[ch];
  }
}

__global__ void __launch_bounds__(BLOCK) render_bwd_kernel(const Cam cam, const UpsState st, const float* __restrict__ feat,
                                                           const float* __restrict__ dL_dpix, float* dL_dmean2D, float* dL_dconic,
                                                           float* dL_dopacity, float* dL_dcolor) {
  __shared__ uint32_t s_id[BLOCK];
  __shared__ float2 s_xy[BLOCK];
  __shared__ float4 s_co[BLOCK];
  __shared__ float s_col[3 * BLOCK];
  const int px = blockIdx.x * BX + threadIdx.x, py = blockIdx.y * BY + threadIdx.y;
  const int tr = threadIdx.y * BX + threadIdx.x;
  const bool inside = px < cam.W && py < cam.H;
  const int pid = py * cam.W + px;
  const float pfx = (float)px, pfy = (float)py;
  const uint2 range = st.ranges[blockIdx.y * cam.gx + blockIdx.x];
  const int rounds = (range.y - range.x + BLOCK - 1) / BLOCK;
  int todo = range.y - range.x;
  const bool done = !inside;
  const float T_final = inside ? st.final_T[pid] : 0.f;
  float T = T_final;
  uint32_t contributor = todo;
  const uint32_t last_contributor = inside ? st.n_contrib[pid] : 0;
  float accum_rec[3] = {0.f, 0.f, 0.f}, last_color[3] = {0.f, 0.f, 0.f}, last_alpha = 0.f, dpix[3] = {0.f, 0.f, 0.f};
  const size_t HW = (size_t)cam.H * cam.W;
  if (inside)
    for (int ch = 0; ch < 3; ch++) dpix[ch] = dL_dpix[ch * HW + pid];
  float bg_dot = 0.f;
  for (int ch = 0; ch < 3; ch++) bg_dot += cam.bg[ch] * dpix[ch];
  const float ddelx_dx = 0.5f * cam.W, ddely_dy = 0.5f * cam.H;
  for (int i = 0; i < rounds; i++, todo -= BLOCK) {
    __syncthreads();
    const int progress = i * BLOCK + tr;
    if (range.x + progress < range.y) {
      const uint32_t id = st.vals[range.y - progress - 1];
      s_id[tr] = id; s_xy[tr] = st.xy[id]; s_co[tr] = st.conic_opacity[id];
      for (int ch = 0; ch < 3; ch++) s_col[ch * BLOCK + tr] = feat[3 * (size_t)id + ch];
    }
    __syncthreads();
    for (int j = 0; !done && j < min(BLOCK, todo); j++) {
      contributor--;
      if (contributor >= last_contributor) continue;
      const float2 xy = s_xy[j];
      const float dx = xy.x - pfx, dy = xy.y - pfy;
      const float4 co = s_co[j];
      const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
      if (power > 0.f) continue;
      const float G = expf(power);
      const float alpha = fminf(ALPHA_MAX, co.w * G);
      if (alpha < ALPHA_MIN) continue;
      T = T / (1.f - alpha);
      const float dchannel_dcolor = alpha * T;
      const uint32_t id = s_id[j];
      float dL_dalpha = 0.f;
      for (int ch = 0; ch < 3; ch++) {
        const float c = s_col[ch * BLOCK + j];
        accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
        last_color[ch] = c;
        dL_dalpha += (c - accum_rec[ch]) * dpix[ch];
        atomicAdd(dL_dcolor + 3 * (size_t)id + ch, dchannel_dcolor * dpix[ch]);
      }
      dL_dalpha *= T;
      last_alpha = alpha;
      dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
      const float dL_dG = co.w * dL_dalpha;
      const float gdx = G * dx, gdy = G * dy;
      const float dG_ddelx = -gdx * co.x - gdy * co.y;
      const float dG_ddely = -gdy * co.z - gdx * co.y;
      atomicAdd(dL_dmean2D + 3 * (size_t)id + 0, dL_dG * dG_ddelx * ddelx_dx);
      atomicAdd(dL_dmean2D + 3 * (size_t)id + 1, dL_dG * dG_ddely * ddely_dy);
      atomicAdd(dL_dconic + 4 * (size_t)id + 0, -0.5f * gdx * dx * dL_dG);
      atomicAdd(dL_dconic + 4 * (size_t)id + 1, -0.5f * gdx * dy * dL_dG);
      atomicAdd(dL_dconic + 4 * (size_t)id + 3, -0.5f * gdy * dy * dL_dG);
      atomicAdd(dL_dopacity + id, G * dL_dalpha);
    }
  }
}

// cov2D backward: dL_dconic -> dL_dcov3D, and the part of dL_dmean3D that flows through the Jacobian (written)
__global__ void cov2d_bwd_kernel(int P, const Cam cam, const float* __restrict__ means3D, const float* __restrict__ cov3D,
                                 const int* __restrict__ radii, const float* __restrict__ dL_dconic, float* dL_dmean3D, float* dL_dcov3D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P || !(radii[i] > 0)) return;
  const float3 mean = make_float3(means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]);
  const float* c3 = cov3D + 6 * (size_t)i;
  Cov2D q;
  compute_cov2d(mean, c3, cam, q);
  const float a = q.a, b = q.b, c = q.c;
  const float gx = dL_dconic[4 * i], gy = dL_dconic[4 * i + 1], gz = dL_dconic[4 * i + 3];
  const float denom = a * c - b * b;
  float dL_da = 0, dL_db = 0, dL_dc = 0;
  const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
  const float* T0 = q.T0;
  const float* T1 = q.T1;
  float* dcov = dL_dcov3D + 6 * (size_t)i;
  if (denom2inv != 0) {
    dL_da = denom2inv * (-c * c * gx + 2 * b * c * gy + (denom - a * c) * gz);
    dL_dc = denom2inv * (-a * a * gz + 2 * a * b * gy + (denom - a * c) * gx);
    dL_db = denom2inv * 2 * (b * c * gx - (denom + 2 * b * b) * gy + a * b * gz);
    dcov[0] = T0[0] * T0[0] * dL_da + T0[0] * T1[0] * dL_db + T1[0] * T1[0] * dL_dc;
    dcov[3] = T0[1] * T0[1] * dL_da + T0[1] * T1[1] * dL_db + T1[1] * T1[1] * dL_dc;
    dcov[5] = T0[2] * T0[2] * dL_da + T0[2] * T1[2] * dL_db + T1[2] * T1[2] * dL_dc;
    dcov[1] = 2 * T0[0] * T0[1] * dL_da + (T0[0] * T1[1] + T0[1] * T1[0]) * dL_db + 2 * T1[0] * T1[1] * dL_dc;
    dcov[2] = 2 * T0[0] * T0[2] * dL_da + (T0[0] * T1[2] + T0[2] * T1[0]) * dL_db + 2 * T1[0] * T1[2] * dL_dc;
    dcov[4] = 2 * T0[2] * T0[1] * dL_da + (T0[1] * T1[2] + T0[2] * T1[1]) * dL_db + 2 * T1[1] * T1[2] * dL_dc;
  } else {
    for (int k = 0; k < 6; k++) dcov[k] = 0;
  }
  const float V[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
  float dT0[3], dT1[3];
  for (int j = 0; j < 3; j++) {
    const float t0v = T0[0] * V[j][0] + T0[1] * V[j][1] + T0[2] * V[j][2];
    const float t1v = T1[0] * V[j][0] + T1[1] * V[j][1] + T1[2] * V[j][2];
    dT0[j] = 2 * t0v * dL_da + t1v * dL_db;
    dT1[j] = 2 * t1v * dL_dc + t0v * dL_db;
  }
  const float* vm = cam.view;  // W[c][r] = vm[4r + c]
  const float dJ00 = vm[0] * dT0[0] + vm[4] * dT0[1] + vm[8] * dT0[2];
  const float dJ02 = vm[2] * dT0[0] + vm[6] * dT0[1] + vm[10] * dT0[2];
  const float dJ11 = vm[1] * dT1[0] + vm[5] * dT1[1] + vm[9] * dT1[2];
  const float dJ12 = vm[2] * dT1[0] + vm[6] * dT1[1] + vm[10] * dT1[2];
  const float tz = 1.f / q.t.z, tz2 = tz * tz, tz3 = tz2 * tz;
  const float hx = cam.focal_x, hy = cam.focal_y;
  const float dtx = q.xmul * -hx * tz2 * dJ02;
  const float dty = q.ymul * -hy * tz2 * dJ12;
  const float dtz = -hx * tz2 * dJ00 - hy * tz2 * dJ11 + (2 * hx * q.t.x) * tz3 * dJ02 + (2 * hy * q.t.y) * tz3 * dJ12;
  dL_dmean3D[3 * i + 0] = vm[0] * dtx + vm[1] * dty + vm[2] * dtz;
  dL_dmean3D[3 * i + 1] = vm[4] * dtx + vm[5] * dty + vm[6] * dtz;
  dL_dmean3D[3 * i + 2] = vm[8] * dtx + vm[9] * dty + vm[10] * dtz;
}

__device__ __forceinline__ float3 dnormvdv(const float3 v, const float3 dv) {
  const float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
  const float inv = 1.0f / sqrtf(sum2 * sum2 * sum2);
  return make_float3(((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * inv,
                     (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * inv,
                     (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * inv);
}

// screen-space mean -> 3D mean through the projection, and the SH backward (accumulates into dL_dmean3D)
__global__ void preprocess_bwd_kernel(int P, const Cam cam, const float* __restrict__ means3D, const float* __restrict__ shs,
                                      const int* __restrict__ radii, const uint8_t* __restrict__ clamped, const float* __restrict__ dL_dmean2D,
                                      const float* __restrict__ dL_dcolor, float* dL_dmean3D, float* dL_dsh) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P || !(radii[i] > 0)) return;
  const float3 mean = make_float3(means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]);
  const float* pr = cam.proj;
  const float4 mh = xform4x4(mean, pr);
  const float m_w = 1.0f / (mh.w + 0.0000001f);
  const float mul1 = (pr[0] * mean.x + pr[4] * mean.y + pr[8] * mean.z + pr[12]) * m_w * m_w;
  const float mul2 = (pr[1] * mean.x + pr[5] * mean.y + pr[9] * mean.z + pr[13]) * m_w * m_w;
  const float g2x = dL_dmean2D[3 * i], g2y = dL_dmean2D[3 * i + 1];
  float3 dmean;
  dmean.x = (pr[0] * m_w - pr[3] * mul1) * g2x + (pr[1] * m_w - pr[3] * mul2) * g2y;
  dmean.y = (pr[4] * m_w - pr[7] * mul1) * g2x + (pr[5] * m_w - pr[7] * mul2) * g2y;
  dmean.z = (pr[8] * m_w - pr[11] * mul1) * g2x + (pr[9] * m_w - pr[11] * mul2) * g2y;
  dL_dmean3D[3 * i + 0] += dmean.x; dL_dmean3D[3 * i + 1] += dmean.y; dL_dmean3D[3 * i + 2] += dmean.z;
  if (!shs) return;
  const int M = cam.M, D = cam.D;
  const float* sh = shs + (size_t)i * M * 3;
  float* dsh = dL_dsh + (size_t)i * M * 3;
  const float3 dir_orig = make_float3(mean.x - cam.campos[0], mean.y - cam.campos[1], mean.z - cam.campos[2]);
  const float len = sqrtf(dir_orig.x * dir_orig.x + dir_orig.y * dir_orig.y + dir_orig.z * dir_orig.z);
  const float x = dir_orig.x / len, y = dir_orig.y / len, z = dir_orig.z / len;
  float3 dL_ddir = make_float3(0.f, 0.f, 0.f);
  for (int ch = 0; ch < 3; ch++) {
    const float g = dL_dcolor[3 * i + ch] * (clamped[3 * i + ch] ? 0.f : 1.f);
#define S(k) sh[(k) * 3 + ch]
#define DS(k) dsh[(k) * 3 + ch]
    float dx_ = 0, dy_ = 0, dz_ = 0;
    DS(0) = SH_C0 * g;
    if (D > 0) {
      DS(1) = -SH_C1 * y * g; DS(2) = SH_C1 * z * g; DS(3) = -SH_C1 * x * g;
      dx_ = -SH_C1 * S(3); dy_ = -SH_C1 * S(1); dz_ = SH_C1 * S(2);
      if (D > 1) {
        const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        DS(4) = SH_C2[0] * xy * g; DS(5) = SH_C2[1] * yz * g; DS(6) = SH_C2[2] * (2.f * zz - xx - yy) * g;
        DS(7) = SH_C2[3] * xz * g; DS(8) = SH_C2[4] * (xx - yy) * g;
        dx_ += SH_C2[0] * y * S(4) + SH_C2[2] * 2.f * -x * S(6) + SH_C2[3] * z * S(7) + SH_C2[4] * 2.f * x * S(8);
        dy_ += SH_C2[0] * x * S(4) + SH_C2[1] * z * S(5) + SH_C2[2] * 2.f * -y * S(6) + SH_C2[4] * 2.f * -y * S(8);
        dz_ += SH_C2[1] * y * S(5) + SH_C2[2] * 2.f * 2.f * z * S(6) + SH_C2[3] * x * S(7);
        if (D > 2) {
          DS(9) = SH_C3[0] * y * (3.f * xx - yy) * g; DS(10) = SH_C3[1] * xy * z * g;
          DS(11) = SH_C3[2] * y * (4.f * zz - xx - yy) * g; DS(12) = SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy) * g;
          DS(13) = SH_C3[4] * x * (4.f * zz - xx - yy) * g; DS(14) = SH_C3[5] * z * (xx - yy) * g;
          DS(15) = SH_C3[6] * x * (xx - 3.f * yy) * g;
          dx_ += SH_C3[0] * S(9) * 3.f * 2.f * xy + SH_C3[1] * S(10) * yz + SH_C3[2] * S(11) * -2.f * xy + SH_C3[3] * S(12) * -3.f * 2.f * xz +
                 SH_C3[4] * S(13) * (-3.f * xx + 4.f * zz - yy) + SH_C3[5] * S(14) * 2.f * xz + SH_C3[6] * S(15) * 3.f * (xx - yy);
          dy_ += SH_C3[0] * S(9) * 3.f * (xx - yy) + SH_C3[1] * S(10) * xz + SH_C3[2] * S(11) * (-3.f * yy + 4.f * zz - xx) +
                 SH_C3[3] * S(12) * -3.f * 2.f * yz + SH_C3[4] * S(13) * -2.f * xy + SH_C3[5] * S(14) * -2.f * yz + SH_C3[6] * S(15) * -3.f * 2.f * xy;
          dz_ += SH_C3[1] * S(10) * xy + SH_C3[2] * S(11) * 4.f * 2.f * yz + SH_C3[3] * S(12) * 3.f * (2.f * zz - xx - yy) +
                 SH_C3[4] * S(13) * 4.f * 2.f * xz + SH_C3[5] * S(14) * (xx - yy);
        }
      }
    }
#undef S
#undef DS
    dL_ddir.x += dx_ * g; dL_ddir.y += dy_ * g; dL_ddir.z += dz_ * g;
  }
  const float3 dm = dnormvdv(dir_orig, dL_ddir);
  dL_dmean3D[3 * i + 0] += dm.x; dL_dmean3D[3 * i + 1] += dm.y; dL_dmean3D[3 * i + 2] += dm.z;
}

// sorted key bits = 32 depth bits + the "next higher most significant bit" of the tile count found by bisection
// (256 tiles -> 9, 1920 -> 11: SURVEY.md appendix B.1), one more than the tile ids need
int key_bits(int tiles) {
  uint32_t n = (uint32_t)tiles, msb = 16, step = 16;
  while (step > 1) {
    step /= 2;
    if (n >> msb) msb += step; else msb -= step;
  }
  if (n >> msb) msb++;
  return 32 + (int)msb;
}

}  // namespace

extern "C" {

size_t ups_scan_temp_bytes(int P) {
  size_t n = 0;
  cub::DeviceScan::InclusiveSum(nullptr, n, (uint32_t*)nullptr, (uint32_t*)nullptr, P);
  return n;
}

size_t ups_sort_temp_bytes(long long R, int tiles) {
  size_t n = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, n, (uint64_t*)nullptr, (uint64_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, R, 0, key_bits(tiles));
  return n;
}

/* preprocess + scan, then the blocking device->host read of the pair count (the extension does the same before
 * it can size its binning buffers) */
int ups_preprocess(const UpsView* v, int P, const float* means3D, const float* shs, const float* opacities, const float* cov3D,
                   const UpsState* st, long long* num_rendered, cudaStream_t stream) {
  *num_rendered = 0;
  if (P <= 0) return 0;
  const Cam cam = make_cam(*v);
  preprocess_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, cam, means3D, shs, opacities, cov3D, *st);
  size_t tb = st->scan_temp_bytes;
  cub::DeviceScan::InclusiveSum(st->scan_temp, tb, st->tiles_touched, st->offsets, P, stream);
  uint32_t last = 0;
  cudaError_t e = cudaMemcpyAsync(&last, st->offsets + (P - 1), 4, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  *num_rendered = last;
  return e == cudaSuccess ? 0 : (int)e;
}

int ups_bin_render(const UpsView* v, int P, long long R, const UpsState* st, const float* colors_precomp, float* out_color, cudaStream_t stream) {
  const Cam cam = make_cam(*v);
  const int tiles = cam.gx * cam.gy;
  cudaMemsetAsync(st->ranges, 0, (size_t)tiles * sizeof(uint2), stream);
  if (R > 0) {
    duplicate_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, cam, *st);
    size_t tb = st->sort_temp_bytes;
    cub::DeviceRadixSort::SortPairs(st->sort_temp, tb, st->keys_unsorted, st->keys, st->vals_unsorted, st->vals, R, 0, key_bits(tiles), stream);
    ranges_kernel<<<(unsigned)((R + 255) / 256), 256, 0, stream>>>(R, st->keys, st->ranges);
  }
  render_kernel<<<dim3(cam.gx, cam.gy), dim3(BX, BY), 0, stream>>>(cam, *st, colors_precomp ? colors_precomp : st->rgb, out_color);
  return (int)cudaGetLastError();
}

/* gradient buffers must be zero on entry (the extension's Python side allocates them with zeros) */
int ups_backward(const UpsView* v, int P, long long R, const UpsState* st, const float* means3D, const float* shs, const float* colors_precomp,
                 const float* cov3D, const float* dL_dpix, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
                 float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, cudaStream_t stream) {
  if (P <= 0) return 0;
  const Cam cam = make_cam(*v);
  if (R > 0)
    render_bwd_kernel<<<dim3(cam.gx, cam.gy), dim3(BX, BY), 0, stream>>>(cam, *st, colors_precomp ? colors_precomp : st->rgb, dL_dpix, dL_dmean2D,
                                                                        dL_dconic, dL_dopacity, dL_dcolor);
  cov2d_bwd_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, cam, means3D, cov3D, st->radii, dL_dconic, dL_dmean3D, dL_dcov3D);
  preprocess_bwd_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, cam, means3D, shs, st->radii, st->clamped, dL_dmean2D, dL_dcolor, dL_dmean3D, dL_dsh);
  return (int)cudaGetLastError();
}

int ups_struct_sizes(int which) { return which == 0 ? (int)sizeof(UpsView) : (int)sizeof(UpsState); }

}  // extern "C"
