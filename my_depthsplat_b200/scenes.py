"""Seeded synthetic scenes for parity tests and bench.py.

The value distributions follow what the reference's encoder hands to the decoder:
  * pixel-aligned Gaussians, one per context-view pixel, order (v, y, x)
    (src/model/encoder/encoder_depthsplat.py:329-346);
  * means = ray origin + direction * depth        (common/gaussian_adapter.py:90-91);
  * scales clamp(softplus(raw - 4), min, max)     (common/gaussian_adapter.py:64-67);
  * covariance R_c2w (R S S^T R^T) R_c2w^T        (common/gaussian_adapter.py:85-87, gaussians.py:33-44);
  * SH DC from an RGB image, higher bands masked by 0.1 * 0.25**degree (gaussian_adapter.py:41-47,126-128);
  * opacity = sigmoid(raw)                        (encoder_depthsplat.py:258);
  * normalised intrinsics, OpenCV camera-to-world extrinsics (src/dataset/dataset_re10k.py:198-219).
Everything is generated on the CPU with one torch.Generator so that the CPU oracle and the GPU path
see bit-identical inputs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
import torch.nn.functional as F

from .types import Gaussians


@dataclass
class SceneConfig:
    name: str
    seed: int
    v_ctx: int
    height: int
    width: int
    batch: int = 1
    v_tgt: int = 4
    scale_mode: str = "init"  # "init" | "trained" | "stress"
    scale_max: float = 3.0
    near: float = 0.5
    far: float = 100.0
    sh_degree: int = 2
    fx: float | None = None
    near_plane_fraction: float = 0.0  # stress: fraction of depths clustered at the near plane
    pad_to: int | None = None         # pad the Gaussian count (stress config: "6 ctx views + pad")


# The five BASELINE.json configs (SURVEY.md 8d "Config -> inputs") plus small parity-test scenes.
CONFIGS = {
    "C1": SceneConfig("C1", 1001, 2, 256, 256, batch=1, v_tgt=4, scale_mode="init", scale_max=3.0, far=100.0),
    "C2": SceneConfig("C2", 1002, 6, 512, 960, batch=1, v_tgt=10, scale_mode="trained", scale_max=0.1, far=100.0),
    "C3": SceneConfig("C3", 1003, 12, 512, 960, batch=1, v_tgt=100, scale_mode="trained", scale_max=0.1, far=200.0),
    "C4": SceneConfig("C4", 1004, 2, 256, 256, batch=8, v_tgt=4, scale_mode="init", scale_max=3.0, far=100.0),
    "C5": SceneConfig("C5", 1005, 6, 512, 960, batch=1, v_tgt=4, scale_mode="stress", scale_max=0.5, far=100.0,
                      near_plane_fraction=0.1, pad_to=3_000_000),
    # headline training-shaped workload at the larger resolution (north_star target: fwd+bwd at 512x960, ~3M)
    "C2T": SceneConfig("C2T", 1002, 6, 512, 960, batch=1, v_tgt=4, scale_mode="trained", scale_max=0.1, far=100.0),
    # small scenes for oracle-sized parity tests
    "tiny": SceneConfig("tiny", 7, 2, 32, 48, batch=1, v_tgt=2, scale_mode="init", scale_max=3.0),
    "small": SceneConfig("small", 11, 2, 64, 80, batch=2, v_tgt=3, scale_mode="init", scale_max=3.0),
    "small_trained": SceneConfig("small_trained", 13, 3, 96, 160, batch=1, v_tgt=3, scale_mode="trained", scale_max=0.1),
    "small_stress": SceneConfig("small_stress", 17, 2, 64, 96, batch=1, v_tgt=2, scale_mode="stress", scale_max=0.5,
                                near_plane_fraction=0.1),
    "ragged": SceneConfig("ragged", 19, 2, 50, 70, batch=1, v_tgt=2, scale_mode="init", scale_max=3.0),
}


@dataclass
class Scene:
    cfg: SceneConfig
    gaussians: Gaussians                 # [B,N,...]
    extrinsics: torch.Tensor             # [B,V,4,4] target c2w
    intrinsics: torch.Tensor             # [B,V,3,3] normalised
    near: torch.Tensor                   # [B,V]
    far: torch.Tensor                    # [B,V]
    image_shape: tuple[int, int]
    background: torch.Tensor             # [3]
    grad_color: torch.Tensor             # [B,V,3,H,W] upstream dL/dcolor
    grad_depth: torch.Tensor             # [B,V,H,W]   upstream dL/ddepth
    ctx_extrinsics: torch.Tensor = field(default=None)  # [B,Vc,4,4]

    def to(self, device) -> "Scene":
        g = self.gaussians
        return Scene(
            self.cfg,
            Gaussians(g.means.to(device), g.covariances.to(device), g.harmonics.to(device), g.opacities.to(device)),
            self.extrinsics.to(device), self.intrinsics.to(device), self.near.to(device), self.far.to(device),
            self.image_shape, self.background.to(device), self.grad_color.to(device), self.grad_depth.to(device),
            None if self.ctx_extrinsics is None else self.ctx_extrinsics.to(device),
        )


def _yaw(a: torch.Tensor) -> torch.Tensor:
    c, s = torch.cos(a), torch.sin(a)
    z, o = torch.zeros_like(a), torch.ones_like(a)
    return torch.stack([torch.stack([c, z, s], -1), torch.stack([z, o, z], -1), torch.stack([-s, z, c], -1)], -2)


def _quat_to_matrix(q: torch.Tensor) -> torch.Tensor:
    i, j, k, r = q.unbind(-1)
    two_s = 2 / ((q * q).sum(-1) + 1e-8)
    o = torch.stack([
        1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
        two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
        two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)], -1)
    return o.reshape(*q.shape[:-1], 3, 3)


def make_scene(cfg: SceneConfig | str, *, batch: int | None = None, v_tgt: int | None = None) -> Scene:
    if isinstance(cfg, str):
        cfg = CONFIGS[cfg]
    B = batch or cfg.batch
    V = v_tgt or cfg.v_tgt
    H, W, Vc = cfg.height, cfg.width, cfg.v_ctx
    g = torch.Generator().manual_seed(cfg.seed)
    rn = lambda *s: torch.randn(*s, generator=g)
    ru = lambda *s: torch.rand(*s, generator=g)

    if cfg.fx is not None:
        fx = cfg.fx
    else:
        fx = 0.86 if W == H else 0.55
    fy = fx * W / H
    K = torch.tensor([[fx, 0.0, 0.5], [0.0, fy, 0.5], [0.0, 0.0, 1.0]])

    means, covs, shs, opacs, ext_t, ext_c = [], [], [], [], [], []
    d_sh = (cfg.sh_degree + 1) ** 2
    for _ in range(B):
        # context cameras on a line, small yaw jitter
        yaw = rn(Vc) * math.radians(2.0)
        Rc = _yaw(yaw)
        tc = torch.zeros(Vc, 3)
        tc[:, 0] = (torch.arange(Vc, dtype=torch.float32) - (Vc - 1) / 2) * 0.15
        c2w = torch.eye(4).repeat(Vc, 1, 1)
        c2w[:, :3, :3] = Rc
        c2w[:, :3, 3] = tc
        ext_c.append(c2w)
        # target cameras: interpolate first -> last context pose
        a = torch.linspace(0.0, 1.0, V) if V > 1 else torch.tensor([0.5])
        yaw_t = yaw[0] + (yaw[-1] - yaw[0]) * a
        t2w = torch.eye(4).repeat(V, 1, 1)
        t2w[:, :3, :3] = _yaw(yaw_t)
        t2w[:, :3, 3] = tc[0][None] + (tc[-1] - tc[0])[None] * a[:, None]
        ext_t.append(t2w)

        # pixel-aligned Gaussians
        ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
        centre = torch.stack([(xs + 0.5) / W, (ys + 0.5) / H], -1)  # [H,W,2]
        off = (torch.sigmoid(rn(Vc, H, W, 2)) - 0.5) / torch.tensor([W, H], dtype=torch.float32)
        uv = centre[None] + off
        depth_grid = cfg.near * (2.0 + 18.0 * ru(Vc, 1, 9, 16))
        depth = F.interpolate(depth_grid, size=(H, W), mode="bilinear", align_corners=True)[:, 0]  # [Vc,H,W]
        if cfg.near_plane_fraction > 0:
            m = ru(Vc, H, W) < cfg.near_plane_fraction
            depth = torch.where(m, cfg.near * (1.0 + 0.3 * ru(Vc, H, W)), depth)
        Kinv = torch.linalg.inv(K)
        dirs = torch.cat([uv, torch.ones(Vc, H, W, 1)], -1) @ Kinv.T
        dirs = dirs / dirs.norm(dim=-1, keepdim=True)
        dirs_w = torch.einsum("vij,vhwj->vhwi", Rc, dirs)
        mean = tc[:, None, None, :] + dirs_w * depth[..., None]

        if cfg.scale_mode == "init":
            scales = torch.clamp(F.softplus(0.5 * rn(Vc, H, W, 3) - 4.0), 1e-10, cfg.scale_max)
        elif cfg.scale_mode == "trained":
            s = depth[..., None] / (fx * W) * (0.5 + 1.5 * ru(Vc, H, W, 3))
            scales = torch.clamp(s, 1e-10, cfg.scale_max)
        elif cfg.scale_mode == "stress":
            scales = 0.05 + 0.45 * ru(Vc, H, W, 3)
            scales = torch.clamp(scales, 1e-10, cfg.scale_max)
        else:
            raise ValueError(cfg.scale_mode)
        q = rn(Vc, H, W, 4)
        q = q / (q.norm(dim=-1, keepdim=True) + 1e-8)
        Rq = _quat_to_matrix(q)
        S = torch.diag_embed(scales)
        cov = Rq @ S @ S.transpose(-1, -2) @ Rq.transpose(-1, -2)
        Rc_b = Rc[:, None, None]
        cov = Rc_b @ cov @ Rc_b.transpose(-1, -2)

        sh = torch.zeros(Vc, H, W, 3, d_sh)
        sh[..., 0] = (ru(Vc, H, W, 3) - 0.5) / 0.28209479177387814
        for deg in range(1, cfg.sh_degree + 1):
            sh[..., deg ** 2:(deg + 1) ** 2] = rn(Vc, H, W, 3, 2 * deg + 1) * 0.1 * 0.25 ** deg
        op = torch.sigmoid(rn(Vc, H, W))

        mean, cov, sh, op = mean.reshape(-1, 3), cov.reshape(-1, 3, 3), sh.reshape(-1, 3, d_sh), op.reshape(-1)
        if cfg.pad_to is not None and cfg.pad_to > mean.shape[0]:
            n_extra = cfg.pad_to - mean.shape[0]
            idx = torch.randint(0, mean.shape[0], (n_extra,), generator=g)
            jitter = rn(n_extra, 3) * 0.01
            mean = torch.cat([mean, mean[idx] + jitter])
            cov = torch.cat([cov, cov[idx]])
            sh = torch.cat([sh, sh[idx]])
            op = torch.cat([op, op[idx]])
        means.append(mean); covs.append(cov); shs.append(sh); opacs.append(op)

    gaussians = Gaussians(torch.stack(means).contiguous(), torch.stack(covs).contiguous(),
                          torch.stack(shs).contiguous(), torch.stack(opacs).contiguous())
    extr = torch.stack(ext_t)
    intr = K[None, None].repeat(B, V, 1, 1)
    near = torch.full((B, V), cfg.near)
    far = torch.full((B, V), cfg.far)
    grad_color = rn(B, V, 3, H, W) / (3 * H * W)
    grad_depth = rn(B, V, H, W) / (H * W)
    return Scene(cfg, gaussians, extr, intr, near, far, (H, W), torch.zeros(3), grad_color, grad_depth,
                 torch.stack(ext_c))
