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


# ---------------------------------------------------------------------------------------------------
# The reference's decoder path restated on top of the oracle (CPU, differentiable).
# Follows decoder_splatting_cuda.py:35-91 and cuda_splatting.py:46-126, 225-264 step by step: per
# view loop, scale-invariant normalisation in torch (autograd supplies its chain rule), SH re-layout,
# upper-triangle covariance gather, second render for depth with z as colour and mean over channels.
# tests/test_reference_glue.py checks it against the reference's UNMODIFIED file in this container;
# on the GPU box (no /root/reference) it is the checker for colour, depth and gradients.
def oracle_decoder_forward(gaussians, extrinsics, intrinsics, near, far, image_shape, background, depth_mode=None,
                           scale_invariant=True, use_sh=True):
    from my_depthsplat_b200.cuda_splatting import get_projection_matrix
    from my_depthsplat_b200.projection import get_fov, homogenize_points
    from oracle import ext_compat as ext

    B, V = extrinsics.shape[:2]
    h, w = image_shape

    def render(ext_bv, K_bv, near_bv, far_bv, bg, means, covs, sh, opac, use_sh_):
        # one (scene, view): cuda_splatting.py:63-123 with batch == 1
        if scale_invariant:
            scale = 1 / near_bv
            ext_bv = ext_bv.clone()
            ext_bv[:3, 3] = ext_bv[:3, 3] * scale
            covs = covs * (scale ** 2)
            means = means * scale
            near_bv, far_bv = near_bv * scale, far_bv * scale
        n = sh.shape[-1]
        degree = int(round(n ** 0.5)) - 1
        shs = sh.permute(0, 2, 1).contiguous()  # [g, n, xyz]
        fov_x, fov_y = get_fov(K_bv[None]).unbind(dim=-1)
        tan_x, tan_y = (0.5 * fov_x).tan(), (0.5 * fov_y).tan()
        proj = get_projection_matrix(near_bv[None], far_bv[None], fov_x, fov_y).transpose(1, 2)
        view = ext_bv[None].inverse().transpose(1, 2)
        full = view @ proj
        settings = ext.GaussianRasterizationSettings(
            image_height=h, image_width=w, tanfovx=tan_x[0].item(), tanfovy=tan_y[0].item(), bg=bg, scale_modifier=1.0,
            viewmatrix=view[0], projmatrix=full[0], sh_degree=degree, campos=ext_bv[:3, 3], prefiltered=False, debug=False)
        row, col = torch.triu_indices(3, 3)
        image, radii = ext.GaussianRasterizer(settings)(
            means3D=means, means2D=torch.zeros_like(means, requires_grad=True), shs=shs if use_sh_ else None,
            colors_precomp=None if use_sh_ else shs[:, 0, :], opacities=opac[..., None], cov3D_precomp=covs[:, row, col])
        return image

    colors, depths = [], []
    for b in range(B):
        cs, ds = [], []
        for v in range(V):
            args = (extrinsics[b, v].float(), intrinsics[b, v].float(), near[b, v].float(), far[b, v].float())
            bg = background if background.dim() == 1 else background[b, v]
            cs.append(render(*args, bg, gaussians.means[b], gaussians.covariances[b], gaussians.harmonics[b],
                             gaussians.opacities[b], use_sh))
            if depth_mode is not None:
                cam = torch.einsum("ij,gj->gi", extrinsics[b, v].float().inverse(), homogenize_points(gaussians.means[b]))
                fake = cam[..., 2]
                if depth_mode == "disparity":
                    fake = 1 / fake
                elif depth_mode == "log":
                    fake = fake.minimum(near[b, v]).maximum(far[b, v]).log()
                img = render(*args, torch.zeros(3), gaussians.means[b], gaussians.covariances[b],
                             fake[:, None, None].expand(-1, 3, 1), gaussians.opacities[b], False)
                ds.append(img.mean(dim=0))
        colors.append(torch.stack(cs))
        if depth_mode is not None:
            depths.append(torch.stack(ds))
    return torch.stack(colors), (torch.stack(depths) if depth_mode is not None else None)


def leaf_gaussians(scene):
    """Fresh leaf copies (requires_grad) of a scene's Gaussians on the CPU."""
    from my_depthsplat_b200.types import Gaussians
    g = scene.gaussians
    mk = lambda t: t.detach().clone().requires_grad_()
    return Gaussians(mk(g.means), mk(g.covariances), mk(g.harmonics), mk(g.opacities))
