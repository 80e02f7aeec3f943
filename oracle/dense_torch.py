"""Dense PyTorch statement of the same rendering equation (autograd-differentiable).

TEST INFRASTRUCTURE ONLY.  Two uses:
  1. float64 cross-check of the C oracle's forward and hand-derived backward on tiny scenes
     (tests/test_oracle.py) -- the only independent check available, since the third-party
     extension the reference binds (cuda_splatting.py:5-8) is absent (PARITY UNPINNED);
  2. the "dense PyTorch CPU implementation of the same compositing equation" that
     BASELINE.json's north_star names as the CPU baseline (bench.py reports it next to the
     tile-based C oracle).

It evaluates every (pixel, Gaussian) pair, sorted once per view by depth, and restricts a
Gaussian to the 16x16 tiles of its bounding rect exactly like the tile-based algorithm.
"""
from __future__ import annotations

import math

import torch

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435)


def eval_sh(deg: int, sh, dirs):
    """sh [P,M,3], dirs [P,3] normalised -> [P,3] (before +0.5 / clamp)."""
    x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
    r = SH_C0 * sh[:, 0]
    if deg > 0:
        r = r - SH_C1 * y * sh[:, 1] + SH_C1 * z * sh[:, 2] - SH_C1 * x * sh[:, 3]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        r = (r + SH_C2[0] * xy * sh[:, 4] + SH_C2[1] * yz * sh[:, 5] + SH_C2[2] * (2 * zz - xx - yy) * sh[:, 6]
             + SH_C2[3] * xz * sh[:, 7] + SH_C2[4] * (xx - yy) * sh[:, 8])
    if deg > 2:
        r = (r + SH_C3[0] * y * (3 * xx - yy) * sh[:, 9] + SH_C3[1] * xy * z * sh[:, 10]
             + SH_C3[2] * y * (4 * zz - xx - yy) * sh[:, 11] + SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[:, 12]
             + SH_C3[4] * x * (4 * zz - xx - yy) * sh[:, 13] + SH_C3[5] * z * (xx - yy) * sh[:, 14]
             + SH_C3[6] * x * (xx - 3 * yy) * sh[:, 15])
    return r


def project(means3D, cov3D, viewmatrix, projmatrix, tanfovx, tanfovy, H, W):
    """Per-Gaussian projection.  Matrices are [4,4] in the transposed storage the reference
    passes (row i of the tensor = column i of the math matrix)."""
    P = means3D.shape[0]
    ones = torch.ones(P, 1, dtype=means3D.dtype)
    hom = torch.cat([means3D, ones], dim=1)
    p_view = hom @ viewmatrix  # [P,4]  (x^T M^T)
    p_hom = hom @ projmatrix
    p_w = 1.0 / (p_hom[:, 3] + 1e-7)
    ndc = p_hom[:, :2] * p_w[:, None]
    px = ((ndc[:, 0] + 1.0) * W - 1.0) * 0.5
    py = ((ndc[:, 1] + 1.0) * H - 1.0) * 0.5
    tz = p_view[:, 2]
    limx, limy = 1.3 * tanfovx, 1.3 * tanfovy
    tx = torch.clamp(p_view[:, 0] / tz, -limx, limx) * tz
    ty = torch.clamp(p_view[:, 1] / tz, -limy, limy) * tz
    fx, fy = W / (2.0 * tanfovx), H / (2.0 * tanfovy)
    zero = torch.zeros_like(tz)
    J = torch.stack([torch.stack([fx / tz, zero, -fx * tx / (tz * tz)], -1),
                     torch.stack([zero, fy / tz, -fy * ty / (tz * tz)], -1)], 1)  # [P,2,3]
    R = viewmatrix[:3, :3].T  # world->camera rotation (math form)
    c = cov3D
    Sigma = torch.stack([torch.stack([c[:, 0], c[:, 1], c[:, 2]], -1), torch.stack([c[:, 1], c[:, 3], c[:, 4]], -1),
                         torch.stack([c[:, 2], c[:, 4], c[:, 5]], -1)], 1)
    JR = J @ R
    cov2 = JR @ Sigma @ JR.transpose(1, 2)
    a = cov2[:, 0, 0] + 0.3
    b = cov2[:, 0, 1]
    cc = cov2[:, 1, 1] + 0.3
    det = a * cc - b * b
    conic = torch.stack([cc / det, -b / det, a / det], -1)
    return tz, torch.stack([px, py], -1), conic, (a, b, cc, det)


def render_view(*, H, W, bg, means3D, opacities, cov3D, viewmatrix, projmatrix, campos, tanfovx, tanfovy,
                shs=None, colors_precomp=None, sh_degree=0, radii=None, pixel_window=None, block=4096):
    """Returns color [3,h,w] (h,w = window).  ``radii`` [P] int: the integer screen radii that decide
    visibility and tile rects -- pass the oracle's so that tile membership is identical, or None to
    compute them here (same formula, this dtype).  ``pixel_window`` = (y0, y1, x0, x1)."""
    dt = means3D.dtype
    depth, xy, conic, (a, b, cc, det) = project(means3D, cov3D, viewmatrix, projmatrix, tanfovx, tanfovy, H, W)
    if radii is None:
        mid = 0.5 * (a + cc)
        disc = torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
        radii = torch.ceil(3.0 * torch.sqrt(torch.maximum(mid + disc, mid - disc))).to(torch.int64)
        radii = torch.where((depth > 0.2) & (det != 0), radii, torch.zeros_like(radii))
    radii = torch.as_tensor(radii).to(torch.int64)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    xyd = xy.detach().to(torch.float32)
    rf = radii.to(torch.float32)
    rminx = torch.clamp(((xyd[:, 0] - rf) / 16).to(torch.int32), 0, gx)
    rminy = torch.clamp(((xyd[:, 1] - rf) / 16).to(torch.int32), 0, gy)
    rmaxx = torch.clamp(((((xyd[:, 0] + rf) + 16.0) - 1.0) / 16).to(torch.int32), 0, gx)
    rmaxy = torch.clamp(((((xyd[:, 1] + rf) + 16.0) - 1.0) / 16).to(torch.int32), 0, gy)
    vis = (radii > 0) & ((rmaxx - rminx) * (rmaxy - rminy) > 0)

    if shs is not None:
        d = means3D - campos[None]
        d = d / d.norm(dim=1, keepdim=True)
        col = torch.clamp(eval_sh(sh_degree, shs, d) + 0.5, min=0.0)
    else:
        col = colors_precomp

    idx = torch.nonzero(vis)[:, 0]
    # stable order: depth bits ascending, ties by Gaussian index (the sort is stable, keys are emitted in index order)
    order = torch.argsort(depth.detach().to(torch.float32)[idx], stable=True)
    idx = idx[order]
    xy_s, conic_s, op_s, col_s = xy[idx], conic[idx], opacities.reshape(-1)[idx], col[idx]
    rx0, rx1, ry0, ry1 = rminx[idx], rmaxx[idx], rminy[idx], rmaxy[idx]

    y0, y1, x0, x1 = pixel_window or (0, H, 0, W)
    ys, xs = torch.meshgrid(torch.arange(y0, y1), torch.arange(x0, x1), indexing="ij")
    pix = torch.stack([xs.reshape(-1), ys.reshape(-1)], -1)  # [Npix,2] ints
    out = []
    bg = bg.to(dt)
    for s in range(0, pix.shape[0], block):
        p = pix[s:s + block]
        pf = p.to(dt)
        tx_, ty_ = (p[:, 0] // 16)[:, None], (p[:, 1] // 16)[:, None]
        member = (tx_ >= rx0[None]) & (tx_ < rx1[None]) & (ty_ >= ry0[None]) & (ty_ < ry1[None])
        dx = xy_s[None, :, 0] - pf[:, None, 0]
        dy = xy_s[None, :, 1] - pf[:, None, 1]
        power = -0.5 * (conic_s[None, :, 0] * dx * dx + conic_s[None, :, 2] * dy * dy) - conic_s[None, :, 1] * dx * dy
        alpha = torch.clamp(op_s[None] * torch.exp(torch.clamp(power, max=0.0)), max=0.99)
        live = member & (power <= 0) & (alpha >= 1.0 / 255.0)
        alpha = torch.where(live, alpha, torch.zeros_like(alpha))
        T_incl = torch.cumprod(1.0 - alpha, dim=1)
        T_excl = torch.cat([torch.ones_like(T_incl[:, :1]), T_incl[:, :-1]], dim=1)
        stop = ((T_incl < 1e-4) & live).to(torch.int8).cummax(dim=1).values.bool()
        w = torch.where(stop, torch.zeros_like(alpha), alpha * T_excl)
        # final T = transmittance after the last applied entry
        T_final = torch.where(stop, torch.zeros_like(T_incl), T_incl).clone()
        applied_any_stop = stop.any(dim=1)
        # when stopped, T stays at T_excl of the stopping entry
        first_stop = stop.to(torch.int8).argmax(dim=1)
        T_end = torch.where(applied_any_stop, T_excl.gather(1, first_stop[:, None])[:, 0],
                            T_incl[:, -1] if T_incl.shape[1] > 0 else torch.ones(p.shape[0], dtype=dt))
        del T_final
        c = w @ col_s  # [n,3]
        out.append(c + T_end[:, None] * bg[None])
    img = torch.cat(out, 0).reshape(y1 - y0, x1 - x0, 3).permute(2, 0, 1)
    return img


def focal(tanfov: float, size: int) -> float:
    return size / (2.0 * math.tan(math.atan(tanfov)))
