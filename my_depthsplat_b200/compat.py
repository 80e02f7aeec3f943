"""`diff_gaussian_rasterization`-shaped API on top of the sm_100a library.

The reference imports exactly ``GaussianRasterizationSettings`` and ``GaussianRasterizer`` from that
third-party module (src/model/decoder/cuda_splatting.py:5-8) and calls them per view (:98-123).
Installing this module under that name (``install()``) lets the reference's UNMODIFIED
cuda_splatting.py run on the new kernels -- the inner oracle-swap boundary of SURVEY.md 8(b).
One view per call, tensors in the extension's layouts (shs [P,M,3], cov3D_precomp [P,6]).
"""
from __future__ import annotations

import sys
from typing import NamedTuple, Optional

import torch
from torch import Tensor, nn

from . import _lib
from .rasterizer import ViewPack, rasterize


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: Tensor
    scale_modifier: float
    viewmatrix: Tensor
    projmatrix: Tensor
    sh_degree: int
    campos: Tensor
    prefiltered: bool
    debug: bool


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def forward(self, means3D: Tensor, means2D: Optional[Tensor], opacities: Tensor, shs: Optional[Tensor] = None,
                colors_precomp: Optional[Tensor] = None, scales: Optional[Tensor] = None, rotations: Optional[Tensor] = None,
                cov3D_precomp: Optional[Tensor] = None):
        rs = self.raster_settings
        if (shs is None) == (colors_precomp is None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if cov3D_precomp is None or scales is not None or rotations is not None:
            raise NotImplementedError("only cov3D_precomp is supported: DepthSplat always passes precomputed covariances "
                                      "(cuda_splatting.py:122)")
        dev = means3D.device
        f = lambda t: torch.as_tensor(t, dtype=torch.float32, device=dev)
        pack = ViewPack(
            scene_index=torch.zeros(1, dtype=torch.int32, device=dev),
            viewmatrix=f(rs.viewmatrix).reshape(1, 4, 4).contiguous(), projmatrix=f(rs.projmatrix).reshape(1, 4, 4).contiguous(),
            campos=f(rs.campos).reshape(1, 3).contiguous(),
            tanfov=torch.tensor([[float(rs.tanfovx), float(rs.tanfovy)]], dtype=torch.float32, device=dev),
            background=f(rs.bg).reshape(1, 3).contiguous(), height=int(rs.image_height), width=int(rs.image_width),
        )
        m2d = means2D[None] if (means2D is not None and means2D.requires_grad) else None
        if shs is not None:
            color, _, radii = rasterize(means3D[None], cov3D_precomp[None], shs[None], opacities.reshape(1, -1), pack, use_sh=True,
                                        sh_degree=int(rs.sh_degree), sh_layout=_lib.SH_COEFF_MAJOR, means2d=m2d, want_radii=True)
        else:
            color, _, radii = rasterize(means3D[None], cov3D_precomp[None], colors_precomp[None], opacities.reshape(1, -1), pack,
                                        use_sh=False, means2d=m2d, want_radii=True)
        return color[0], radii[0]


def install() -> None:
    """Make ``import diff_gaussian_rasterization`` resolve to this module."""
    sys.modules["diff_gaussian_rasterization"] = sys.modules[__name__]
