"""Shared helpers for the parity tests.

The checker side (oracle/) is only ever used from here, from tests and from the golden generator.
"""
from __future__ import annotations

import importlib
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

REFERENCE_SRC = Path("/root/reference/src")


def have_reference() -> bool:
    return (REFERENCE_SRC / "model/decoder/cuda_splatting.py").exists()


def load_reference_cuda_splatting(ext_module):
    """Import the reference's UNMODIFIED src/model/decoder/cuda_splatting.py on top of ``ext_module``
    (anything exposing GaussianRasterizationSettings / GaussianRasterizer).  Recipe of SURVEY.md
    appendix A: stub parent packages so that the real __init__ chains (hydra, lightning, ...) are
    not pulled in.  Only available in the build container (/root/reference is not on the GPU box)."""
    sys.modules["diff_gaussian_rasterization"] = ext_module
    root = str(REFERENCE_SRC)
    for name, path in [("src", root), ("src.model", root + "/model"), ("src.model.decoder", root + "/model/decoder"),
                       ("src.geometry", root + "/geometry")]:
        pkg = types.ModuleType(name)
        pkg.__path__ = [path]
        sys.modules[name] = pkg
    for name in ("src.model.decoder.cuda_splatting", "src.geometry.projection"):
        sys.modules.pop(name, None)
    return importlib.import_module("src.model.decoder.cuda_splatting")


def per_view_extension_inputs(scene, b: int, v: int, scale_invariant: bool = True):
    """The arguments the reference would hand to the extension for view (b, v) -- computed with the
    product's own camera code (which mirrors cuda_splatting.py:63-86 op for op)."""
    from my_depthsplat_b200.cuda_splatting import _camera_block
    from my_depthsplat_b200.projection import get_fov

    g = scene.gaussians
    ext = scene.extrinsics[b, v][None].clone().float()
    K = scene.intrinsics[b, v][None].float()
    near, far = scene.near[b, v][None].float(), scene.far[b, v][None].float()
    means, covs = g.means[b], g.covariances[b]
    if scale_invariant:
        s = 1 / near
        ext[..., :3, 3] = ext[..., :3, 3] * s[:, None]
        covs = covs * (s[:, None, None] ** 2)
        means = means * s[:, None]
        near, far = near * s, far * s
    fov_x, fov_y = get_fov(K).unbind(-1)
    tx, ty = (0.5 * fov_x).tan(), (0.5 * fov_y).tan()
    view, full, campos, tanfov = _camera_block(ext, near, far, fov_x, fov_y, tx, ty)
    row, col = torch.triu_indices(3, 3)
    return dict(
        H=scene.image_shape[0], W=scene.image_shape[1], bg=scene.background.numpy(), means3D=means.numpy(),
        opacities=g.opacities[b].numpy(), cov3D=covs[:, row, col].contiguous().numpy(),
        viewmatrix=view[0].numpy(), projmatrix=full[0].numpy(), campos=campos[0].numpy(),
        tanfovx=float(tanfov[0, 0].item()), tanfovy=float(tanfov[0, 1].item()),
        shs=g.harmonics[b].permute(0, 2, 1).contiguous().numpy(), sh_degree=scene.cfg.sh_degree,
    )


def rel_err(a, b, floor=None):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-30) if floor is None else floor
    return np.abs(a - b).max() / scale


def frac_close(a, b, atol, rtol=0.0):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.mean(np.abs(a - b) <= atol + rtol * np.abs(b)))
