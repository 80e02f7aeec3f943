// adapter.cuh -- the Gaussian adapter (src/model/encoder/common/gaussian_adapter.py:49-102, gaussians.py:8-44,
// src/misc/sh_rotation.py:10-30) as device functions, so that the projection kernels can start from the encoder head's RAW
// channel planes: per pixel of a context view, 37 channels
//     0      opacity logit                          -> sigmoid                          (encoder_depthsplat.py:258)
//     1..2   xy-offset logits                       -> pixel centre + (sigmoid - 1/2) pixel        (:270-273)
//     3..5   scales                                 -> clamp(softplus(x - 4), min, max)  (gaussian_adapter.py:64-67)
//     6..9   quaternion xyzw                        -> q / (|q| + 1e-8) -> R(q)          (:72, gaussians.py:8-30)
//     10..36 SH, channel-major 3 x 9                -> * sh_mask, DC += (rgb - 1/2) / C0, rotated by D^l(R_c2w)   (:75-83, :96)
// plus the depth and the context image, and per context view a camera block (c2w rotation and translation, K^-1, the
// 5 x 5 SH rotation of degree 2 -- D^1 is the rotation itself -- and the SH mask).
//     mean = t + R_c2w (K^-1 (u, v, 1) / z) depth,      covariance = (R_c2w R(q)) S^2 (R_c2w R(q))^T.
// cook() builds one Gaussian; cook_backward() applies the chain rule to the gradients w.r.t. (mean, upper-triangular
// covariance, SH, opacity) that the projection backward accumulates.  Plain fp32 with free contraction: this arithmetic is
// reassociated relative to the reference's einsum chains anyway (agreement to ~1e-6 relative, tests/test_gpu_adapter.py).
#pragma once
#include "common.cuh"

namespace b200s {

constexpr int RAW_CH = 37;          // channels of the head
constexpr int RAW_PLANES = 41;      // + depth + 3 image channels: planes staged per 256-Gaussian chunk
constexpr int RAW_CAM_FLOATS = 56;  // R 9 | t 3 | K^-1 9 | pad 1 | D2 25 | mask 9

struct RawCam {
  float R[9], t[3], Kinv[9], pad, D2[25], mask[9];
};
static_assert(sizeof(RawCam) == RAW_CAM_FLOATS * 4, "camera block layout");

struct Cooked {
  float mean[3], cov[6], op;  // cov: upper triangle (xx, xy, xz, yy, yz, zz)
  // kept for the backward
  float dcam[2], pz, depth;   // ray direction in the camera frame is (dcam, 1); pz = third component of K^-1 (u, v, 1)
  float s[3], qh[4], qn, ts, M[9];
  float o, ox, oy;
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// h[RAW_CH] raw channels of one pixel; (px, py) its integer position in the w x h context image
__device__ __forceinline__ void cook(const float* h, float depth, const RawCam& c, int px, int py, int w, int hgt, float smin, float smax,
                                     Cooked& g) {
  g.o = sigmoidf_(h[0]); g.ox = sigmoidf_(h[1]); g.oy = sigmoidf_(h[2]);
  g.op = g.o;
  const float u = ((float)px + 0.5f) / (float)w + (g.ox - 0.5f) / (float)w;
  const float v = ((float)py + 0.5f) / (float)hgt + (g.oy - 0.5f) / (float)hgt;
  const float p0 = c.Kinv[0] * u + c.Kinv[1] * v + c.Kinv[2], p1 = c.Kinv[3] * u + c.Kinv[4] * v + c.Kinv[5];
  g.pz = c.Kinv[6] * u + c.Kinv[7] * v + c.Kinv[8];
  g.dcam[0] = p0 / g.pz; g.dcam[1] = p1 / g.pz;
  g.depth = depth;
#pragma unroll
  for (int a = 0; a < 3; a++) g.mean[a] = c.t[a] + (c.R[3 * a] * g.dcam[0] + c.R[3 * a + 1] * g.dcam[1] + c.R[3 * a + 2]) * depth;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float x = h[3 + k] - 4.0f;
    const float sp = x > 20.0f ? x : log1pf(expf(x));
    g.s[k] = fminf(fmaxf(sp, smin), smax);
  }
  g.qn = sqrtf(h[6] * h[6] + h[7] * h[7] + h[8] * h[8] + h[9] * h[9]);
  const float qi = 1.0f / (g.qn + 1e-8f);
#pragma unroll
  for (int k = 0; k < 4; k++) g.qh[k] = h[6 + k] * qi;
  const float i = g.qh[0], j = g.qh[1], k = g.qh[2], r = g.qh[3];
  g.ts = 2.0f / (i * i + j * j + k * k + r * r + 1e-8f);
  const float ts = g.ts;
  const float Rq[9] = {1.f - ts * (j * j + k * k), ts * (i * j - k * r), ts * (i * k + j * r),
                       ts * (i * j + k * r), 1.f - ts * (i * i + k * k), ts * (j * k - i * r),
                       ts * (i * k - j * r), ts * (j * k + i * r), 1.f - ts * (i * i + j * j)};
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = 0; b < 3; b++) g.M[3 * a + b] = c.R[3 * a] * Rq[b] + c.R[3 * a + 1] * Rq[3 + b] + c.R[3 * a + 2] * Rq[6 + b];
  const float s2[3] = {g.s[0] * g.s[0], g.s[1] * g.s[1], g.s[2] * g.s[2]};
  int e = 0;
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = a; b < 3; b++) g.cov[e++] = g.M[3 * a] * s2[0] * g.M[3 * b] + g.M[3 * a + 1] * s2[1] * g.M[3 * b + 1] + g.M[3 * a + 2] * s2[2] * g.M[3 * b + 2];
}

// SH of one colour channel: in = raw coefficients (9), rgb = the context image's value; out = rotated coefficients
__device__ __forceinline__ void cook_sh(const float* in, float rgb, const RawCam& c, float* out) {
  float m[9];
#pragma unroll
  for (int k = 0; k < 9; k++) m[k] = in[k] * c.mask[k];
  m[0] += (rgb - 0.5f) / SH_C0;
  out[0] = m[0];
#pragma unroll
  for (int a = 0; a < 3; a++) out[1 + a] = c.R[3 * a] * m[1] + c.R[3 * a + 1] * m[2] + c.R[3 * a + 2] * m[3];
#pragma unroll
  for (int a = 0; a < 5; a++) out[4 + a] = c.D2[5 * a] * m[4] + c.D2[5 * a + 1] * m[5] + c.D2[5 * a + 2] * m[6] + c.D2[5 * a + 3] * m[7] + c.D2[5 * a + 4] * m[8];
}
__device__ __forceinline__ void cook_sh_backward(const float* dout, const RawCam& c, float* din) {
  float m[9];
  m[0] = dout[0];
#pragma unroll
  for (int b = 0; b < 3; b++) m[1 + b] = c.R[b] * dout[1] + c.R[3 + b] * dout[2] + c.R[6 + b] * dout[3];
#pragma unroll
  for (int b = 0; b < 5; b++) m[4 + b] = c.D2[b] * dout[4] + c.D2[5 + b] * dout[5] + c.D2[10 + b] * dout[6] + c.D2[15 + b] * dout[7] + c.D2[20 + b] * dout[8];
#pragma unroll
  for (int k = 0; k < 9; k++) din[k] = m[k] * c.mask[k];
}

// dmean[3], dcov[6] (gradient w.r.t. the six UPPER-triangular covariance entries), dop -> gradients of the ten geometry
// channels dh[0..9] and of the depth.  h = the raw channels (for the soft-plus / clamp / normalisation derivatives).
__device__ __forceinline__ void cook_backward(const float* h, const RawCam& c, const Cooked& g, int w, int hgt, float smin, float smax,
                                              const float* dmean, const float* dcov, float dop, float* dh, float& ddepth) {
  dh[0] = dop * g.o * (1.0f - g.o);
  // ---- mean ----
  float dirw[3], gc[3];  // world ray direction; gradient w.r.t. the camera-frame direction (times depth)
#pragma unroll
  for (int a = 0; a < 3; a++) dirw[a] = c.R[3 * a] * g.dcam[0] + c.R[3 * a + 1] * g.dcam[1] + c.R[3 * a + 2];
  ddepth = dmean[0] * dirw[0] + dmean[1] * dirw[1] + dmean[2] * dirw[2];
#pragma unroll
  for (int b = 0; b < 3; b++) gc[b] = g.depth * (c.R[b] * dmean[0] + c.R[3 + b] * dmean[1] + c.R[6 + b] * dmean[2]);
  const float ipz = 1.0f / g.pz;
  const float du = (gc[0] * (c.Kinv[0] - g.dcam[0] * c.Kinv[6]) + gc[1] * (c.Kinv[3] - g.dcam[1] * c.Kinv[6])) * ipz;
  const float dv = (gc[0] * (c.Kinv[1] - g.dcam[0] * c.Kinv[7]) + gc[1] * (c.Kinv[4] - g.dcam[1] * c.Kinv[7])) * ipz;
  dh[1] = du * g.ox * (1.0f - g.ox) / (float)w;
  dh[2] = dv * g.oy * (1.0f - g.oy) / (float)hgt;
  // ---- covariance = M S^2 M^T ----
  const float G[9] = {2.f * dcov[0], dcov[1], dcov[2], dcov[1], 2.f * dcov[3], dcov[4], dcov[2], dcov[4], 2.f * dcov[5]};  // G + G^T
  const float s2[3] = {g.s[0] * g.s[0], g.s[1] * g.s[1], g.s[2] * g.s[2]};
  float GM[9], dM[9];
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int k = 0; k < 3; k++) GM[3 * a + k] = G[3 * a] * g.M[k] + G[3 * a + 1] * g.M[3 + k] + G[3 * a + 2] * g.M[6 + k];
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int k = 0; k < 3; k++) dM[3 * a + k] = GM[3 * a + k] * s2[k];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    // d/d(s_k^2) = (M^T G_upper M)_kk = 1/2 (M^T (G + G^T) M)_kk
    const float dsk2 = 0.5f * (g.M[k] * GM[k] + g.M[3 + k] * GM[3 + k] + g.M[6 + k] * GM[6 + k]);
    const float x = h[3 + k] - 4.0f;
    const float sp = x > 20.0f ? x : log1pf(expf(x));
    const float dsp = (sp >= smin && sp <= smax) ? 2.0f * g.s[k] * dsk2 : 0.f;
    dh[3 + k] = dsp * sigmoidf_(x);
  }
  float dR[9];  // w.r.t. R(q) = R_c2w^T dM
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = 0; b < 3; b++) dR[3 * a + b] = c.R[a] * dM[b] + c.R[3 + a] * dM[3 + b] + c.R[6 + a] * dM[6 + b];
  const float i = g.qh[0], j = g.qh[1], k = g.qh[2], r = g.qh[3], ts = g.ts;
  const float dts = -dR[0] * (j * j + k * k) + dR[1] * (i * j - k * r) + dR[2] * (i * k + j * r) + dR[3] * (i * j + k * r) - dR[4] * (i * i + k * k) +
                    dR[5] * (j * k - i * r) + dR[6] * (i * k - j * r) + dR[7] * (j * k + i * r) - dR[8] * (i * i + j * j);
  float dq[4];
  dq[0] = ts * (dR[1] * j + dR[2] * k + dR[3] * j - 2.f * dR[4] * i - dR[5] * r + dR[6] * k + dR[7] * r - 2.f * dR[8] * i);
  dq[1] = ts * (-2.f * dR[0] * j + dR[1] * i + dR[2] * r + dR[3] * i + dR[5] * k - dR[6] * r + dR[7] * k - 2.f * dR[8] * j);
  dq[2] = ts * (-2.f * dR[0] * k - dR[1] * r + dR[2] * i + dR[3] * r - 2.f * dR[4] * k + dR[5] * j + dR[6] * i + dR[7] * j);
  dq[3] = ts * (-dR[1] * k + dR[2] * j + dR[3] * k - dR[5] * i - dR[6] * j + dR[7] * i);
  const float dtsq = -dts * ts * ts;
#pragma unroll
  for (int a = 0; a < 4; a++) dq[a] += dtsq * g.qh[a];
  // q_hat = q / (n + eps)
  const float n = fmaxf(g.qn, 1e-30f), ne = g.qn + 1e-8f;
  const float dot = h[6] * dq[0] + h[7] * dq[1] + h[8] * dq[2] + h[9] * dq[3];
#pragma unroll
  for (int a = 0; a < 4; a++) dh[6 + a] = dq[a] / ne - dot * h[6 + a] / (n * ne * ne);
}

}  // namespace b200s
