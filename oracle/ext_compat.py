"""A `diff_gaussian_rasterization`-shaped module backed by the CPU oracle.

TEST INFRASTRUCTURE ONLY (see splat_oracle.c).  It exists so that the reference's UNMODIFIED
src/model/decoder/cuda_splatting.py can be imported and run on CPU: the
reference imports exactly two names (cuda_splatting.py:5-8) and uses them at :98-123 / :191-216.
Installing this module as ``sys.modules["diff_gaussian_rasterization"]`` gives golden outputs of
the reference's Python glue (tests/golden/make_golden.py) and the checker for the product's
own drop-in functions.
"""
from __future__ import annotations

from typing import NamedTuple

import numpy as np
import torch
from torch import nn

from . import splat_oracle as so


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


def _np(t):
    return None if t is None else t.detach().cpu().to(torch.float32).contiguous().numpy()


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, cov3Ds_precomp, rs: GaussianRasterizationSettings):
        st = so.forward_view(
            H=rs.image_height, W=rs.image_width, bg=_np(rs.bg), means3D=_np(means3D), opacities=_np(opacities),
            cov3D=_np(cov3Ds_precomp), viewmatrix=_np(rs.viewmatrix), projmatrix=_np(rs.projmatrix),
            campos=_np(rs.campos), tanfovx=float(rs.tanfovx), tanfovy=float(rs.tanfovy),
            shs=_np(sh), colors_precomp=_np(colors_precomp), sh_degree=int(rs.sh_degree),
        )
        ctx.st = st
        ctx.has_sh = sh is not None
        ctx.dev = means3D.device
        ctx.op_shape = opacities.shape
        color = torch.from_numpy(st.color).to(means3D.device)
        radii = torch.from_numpy(st.radii).to(means3D.device)
        ctx.mark_non_differentiable(radii)
        return color, radii

    @staticmethod
    def backward(ctx, grad_color, _grad_radii):
        g = so.backward_view(ctx.st, _np(grad_color))
        t = lambda a: None if a is None else torch.from_numpy(a).to(ctx.dev)
        return (
            t(g["means3D"]), t(g["means2D"]), t(g["sh"]) if ctx.has_sh else None,
            None if ctx.has_sh else t(g["colors"]), t(g["opacity"]).reshape(ctx.op_shape), t(g["cov3D"]), None,
        )


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        if (shs is None) == (colors_precomp is None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if scales is not None or rotations is not None or cov3D_precomp is None:
            raise NotImplementedError("oracle: only cov3D_precomp is on the DepthSplat path (cuda_splatting.py:122)")
        return _RasterizeGaussians.apply(means3D, means2D, shs, colors_precomp, opacities, cov3D_precomp,
                                         self.raster_settings)


def last_state_of(color_tensor):  # helper for tests that want the stage outputs
    fn = color_tensor.grad_fn
    return getattr(fn, "st", None)
