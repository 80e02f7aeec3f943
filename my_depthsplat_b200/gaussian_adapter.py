"""The step right before the rendering path: raw per-pixel head output -> world-space Gaussians,
src/model/encoder/common/gaussian_adapter.py:49-102 with gaussians.py:8-44 (quaternion -> covariance) and
src/misc/sh_rotation.py:10-30 (SH rotated into world space).

Two implementations of the same arithmetic:
  * ``GaussianAdapter`` -- the reference's class restated in PyTorch (same cfg, same ``forward`` signature and broadcasting),
    differentiable, any dtype; it is the checker of the fused path and what runs where the fused path does not apply.
  * ``FusedAdapterDecoder`` -- the adapter FUSED INTO THE PROJECTION: the sm_100a projection kernel reads the head's raw
    channel planes, builds mean / covariance / rotated SH / opacity of its 256 Gaussians in shared memory and goes straight
    on to project them for every view; the projection backward applies the adapter's chain rule before it writes, so the
    gradients come out w.r.t. the raw channels and the depth.  The world-space tensors [B,N,3] + [B,N,3,3] + [B,N,3,9] +
    [B,N] (160 B per Gaussian, written by the adapter, read by the projection, and again in the backward) never exist.

SH rotation without e3nn.  The reference multiplies every degree's coefficient block by e3nn's Wigner-D of the
camera-to-world rotation (``wigner_D(l, *matrix_to_angles(R))``).  e3nn is not installed here; its real spherical
harmonics are, in its own (x, y, z), with y as the polar axis:
    l = 1:  sqrt(3) (x, y, z)
    l = 2:  sqrt(15) x z,  sqrt(15) x y,  sqrt(5) (y^2 - (x^2 + z^2) / 2),  sqrt(15) y z,  sqrt(15)/2 (z^2 - x^2)
and D^l(R) is defined by Y^l(R x) = D^l(R) Y^l(x).  Hence D^1(R) = R, and with A_k the symmetric matrices of the five
quadratic forms (all of Frobenius norm^2 7.5), D^2(R)[k, j] = <R^T A_k R, A_j> / 7.5.  PARITY UNPINNED for this one
sub-step: the basis above is restated from e3nn's published source, not checked against an installed e3nn; the closed form
is pinned against an independent least-squares construction from that basis (tests/test_gaussian_adapter.py).
"""
from __future__ import annotations

from dataclasses import dataclass
from math import isqrt
from typing import Optional

import torch
import torch.nn.functional as F
from torch import Tensor, nn

SH_C0 = 0.28209479177387814


@dataclass
class Gaussians:
    """The adapter's output (gaussian_adapter.py:15-22)."""
    means: Tensor
    covariances: Tensor
    scales: Tensor
    rotations: Tensor
    harmonics: Tensor
    opacities: Tensor


@dataclass
class GaussianAdapterCfg:
    gaussian_scale_min: float
    gaussian_scale_max: float
    sh_degree: int


def quaternion_to_matrix(quaternions: Tensor, eps: float = 1e-8) -> Tensor:
    """xyzw -> [...,3,3] (gaussians.py:8-30)."""
    i, j, k, r = torch.unbind(quaternions, dim=-1)
    two_s = 2 / ((quaternions * quaternions).sum(dim=-1) + eps)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(*quaternions.shape[:-1], 3, 3)


def build_covariance(scale: Tensor, rotation_xyzw: Tensor) -> Tensor:
    """R S S^T R^T (gaussians.py:33-44)."""
    s = scale.diag_embed()
    r = quaternion_to_matrix(rotation_xyzw)
    return r @ s @ s.transpose(-1, -2) @ r.transpose(-1, -2)


def _e3nn_quadratic_forms(dtype, device) -> Tensor:
    """[5,3,3]: symmetric A_k with Y^2_k(x) = x^T A_k x in e3nn's basis (module docstring)."""
    a = torch.zeros(5, 3, 3, dtype=torch.float64)
    h = 15 ** 0.5 / 2
    a[0, 0, 2] = a[0, 2, 0] = h
    a[1, 0, 1] = a[1, 1, 0] = h
    a[2] = torch.diag(torch.tensor([-0.5, 1.0, -0.5], dtype=torch.float64)) * 5 ** 0.5
    a[3, 1, 2] = a[3, 2, 1] = h
    a[4] = torch.diag(torch.tensor([-1.0, 0.0, 1.0], dtype=torch.float64)) * h
    return a.to(dtype=dtype, device=device)


def sh_rotation_matrices(rotations: Tensor, degree: int) -> list[Tensor]:
    """[..., 3, 3] rotation -> [D^0, D^1, ..., D^degree] with D^l of shape [..., 2l+1, 2l+1] (degree <= 2)."""
    if degree > 2:
        raise NotImplementedError("closed-form SH rotation is written out for degree <= 2 (DepthSplat uses 2)")
    out = [torch.ones(*rotations.shape[:-2], 1, 1, dtype=rotations.dtype, device=rotations.device)]
    if degree >= 1:
        out.append(rotations)
    if degree >= 2:
        A = _e3nn_quadratic_forms(rotations.dtype, rotations.device)
        M = torch.einsum("...ji,kjm,...mn->...kin", rotations, A, rotations)   # R^T A_k R
        out.append(torch.einsum("...kin,jin->...kj", M, A) / 7.5)
    return out


def rotate_sh(sh_coefficients: Tensor, rotations: Tensor) -> Tensor:
    """[..., n] coefficients, [..., 3, 3] rotations (broadcast) -> [..., n] (sh_rotation.py:10-30 without e3nn)."""
    n = sh_coefficients.shape[-1]
    mats = sh_rotation_matrices(rotations, isqrt(n) - 1)
    parts = [torch.einsum("...ij,...j->...i", D, sh_coefficients[..., d * d:(d + 1) * (d + 1)]) for d, D in enumerate(mats)]
    return torch.cat(parts, dim=-1)


def RGB2SH(rgb):
    return (rgb - 0.5) / SH_C0


def get_world_rays(coordinates: Tensor, extrinsics: Tensor, intrinsics: Tensor):
    """Origins and directions (z = 1 in the camera frame) of the rays through normalised image coordinates
    (src/geometry/projection.py:91-114)."""
    ones = torch.ones_like(coordinates[..., :1])
    d = torch.einsum("...ij,...j->...i", intrinsics.inverse(), torch.cat([coordinates, ones], dim=-1))
    d = d / d[..., -1:]
    d = torch.einsum("...ij,...j->...i", extrinsics[..., :3, :3], d)
    return extrinsics[..., :3, 3].broadcast_to(d.shape), d


class GaussianAdapter(nn.Module):
    cfg: GaussianAdapterCfg

    def __init__(self, cfg: GaussianAdapterCfg):
        super().__init__()
        self.cfg = cfg
        self.register_buffer("sh_mask", torch.ones((self.d_sh,), dtype=torch.float32), persistent=False)
        for degree in range(1, self.cfg.sh_degree + 1):
            self.sh_mask[degree ** 2:(degree + 1) ** 2] = 0.1 * 0.25 ** degree

    def forward(self, extrinsics: Tensor, intrinsics: Optional[Tensor], coordinates: Tensor, depths: Optional[Tensor], opacities: Tensor,
                raw_gaussians: Tensor, image_shape: tuple[int, int], eps: float = 1e-8, point_cloud: Optional[Tensor] = None,
                input_images: Optional[Tensor] = None) -> Gaussians:
        scales, rotations, sh = raw_gaussians.split((3, 4, 3 * self.d_sh), dim=-1)
        scales = torch.clamp(F.softplus(scales - 4.0), min=self.cfg.gaussian_scale_min, max=self.cfg.gaussian_scale_max)
        assert input_images is not None
        rotations = rotations / (rotations.norm(dim=-1, keepdim=True) + eps)
        sh = sh.reshape(*sh.shape[:-1], 3, self.d_sh)
        sh = sh.broadcast_to((*opacities.shape, 3, self.d_sh)) * self.sh_mask.to(sh.dtype)
        b, v = input_images.shape[:2]
        imgs = input_images.permute(0, 1, 3, 4, 2).reshape(b, v, -1, 1, 1, 3)          # "b v c h w -> b v (h w) () () c"
        sh = torch.cat([sh[..., :1] + RGB2SH(imgs)[..., None], sh[..., 1:]], dim=-1)
        covariances = build_covariance(scales, rotations)
        c2w = extrinsics[..., :3, :3]
        covariances = c2w @ covariances @ c2w.transpose(-1, -2)
        origins, directions = get_world_rays(coordinates, extrinsics, intrinsics)
        means = origins + directions * depths[..., None]
        return Gaussians(means=means, covariances=covariances, harmonics=rotate_sh(sh, c2w[..., None, :, :]), opacities=opacities,
                         scales=scales, rotations=rotations.broadcast_to((*scales.shape[:-1], 4)))

    @property
    def d_sh(self) -> int:
        return (self.cfg.sh_degree + 1) ** 2

    @property
    def d_in(self) -> int:
        return 7 + 3 * self.d_sh


def sample_image_grid_xy(h: int, w: int, device=None, dtype=torch.float32) -> Tensor:
    """[h*w, 2] normalised (x, y) pixel centres (projection.py:117-137, the float coordinates)."""
    ys, xs = torch.meshgrid((torch.arange(h, device=device, dtype=dtype) + 0.5) / h, (torch.arange(w, device=device, dtype=dtype) + 0.5) / w, indexing="ij")
    return torch.stack([xs, ys], dim=-1).reshape(h * w, 2)


def adapt_head_output(adapter: GaussianAdapter, head: Tensor, depth: Tensor, images: Tensor, extrinsics: Tensor, intrinsics: Tensor,
                      image_shape: tuple[int, int]):
    """The reference's encoder glue around the adapter (src/model/encoder/encoder_depthsplat.py:226-346, one surface, one
    sample per pixel) restated: ``head`` [B,V,1+2+3+4+3*d_sh,H,W] raw channel planes (opacity logit, xy-offset logits,
    scales, quaternion, SH), ``depth`` [B,V,H,W], ``images`` [B,V,3,H,W] -> the decoder's Gaussians
    (means [B,N,3], covariances [B,N,3,3], harmonics [B,N,3,d_sh], opacities [B,N]) with N = V*H*W in (v, y, x) order."""
    from .types import Gaussians as DecoderGaussians
    b, v, c, h, w = head.shape
    raw = head.permute(0, 1, 3, 4, 2).reshape(b, v, h * w, c)                            # "b v c h w -> b v (h w) c"
    opacities = raw[..., :1].sigmoid().unsqueeze(-1)                                     # [b v r 1 1]
    rest = raw[..., 1:].reshape(b, v, h * w, 1, c - 1)                                   # srf = 1
    xy = sample_image_grid_xy(h, w, head.device, head.dtype).reshape(h * w, 1, 2)
    pixel_size = 1 / torch.tensor((w, h), dtype=head.dtype, device=head.device)
    xy = xy + (rest[..., :2].sigmoid() - 0.5) * pixel_size                               # [b v r 1 2]
    g = adapter.forward(extrinsics[:, :, None, None, None], intrinsics[:, :, None, None, None], xy[..., None, :],
                        depth.reshape(b, v, h * w, 1, 1), opacities, rest[..., None, 2:], (h, w), input_images=images)
    n = v * h * w
    return DecoderGaussians(g.means.reshape(b, n, 3), g.covariances.reshape(b, n, 3, 3), g.harmonics.reshape(b, n, 3, adapter.d_sh),
                            g.opacities.reshape(b, n))


class FusedAdapterDecoder(nn.Module):
    """Adapter + decoder in one call: ``forward(head, depth, context images / cameras, target cameras)`` renders straight from
    the encoder head's raw channel planes; gradients come back w.r.t. ``head`` and ``depth``.  Needs h*w of the context
    views to be a multiple of 256 (the kernels' chunk) and SH degree 2; ``adapt_head_output`` + the plain decoder is the
    general route."""

    def __init__(self, adapter: GaussianAdapter, decoder: nn.Module):
        super().__init__()
        if adapter.cfg.sh_degree != 2:
            raise NotImplementedError("the fused adapter is written for sh_degree 2")
        self.adapter, self.decoder = adapter, decoder

    def forward(self, head: Tensor, depth: Tensor, context_images: Tensor, context_extrinsics: Tensor, context_intrinsics: Tensor,
                extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor, image_shape: tuple[int, int], depth_mode=None,
                cooked_out: Optional[Tensor] = None):
        from .cuda_splatting import render_views_raw
        from .types import DecoderOutput
        cfg = self.adapter.cfg
        color, dimg = render_views_raw(extrinsics, intrinsics, near, far, image_shape, self.decoder.background_color, head, depth, context_images,
                                       context_extrinsics, context_intrinsics, cfg.gaussian_scale_min, cfg.gaussian_scale_max, self.adapter.sh_mask,
                                       depth_mode=depth_mode, cooked_out=cooked_out)
        return DecoderOutput(color, dimg)
