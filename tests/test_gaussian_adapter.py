"""my_depthsplat_b200.gaussian_adapter (CPU): the restated adapter against the reference's UNMODIFIED
src/model/encoder/common/gaussian_adapter.py + gaussians.py (imported with stub parents and a stand-in for e3nn's
rotate_sh: e3nn is not installed, SH rotation is pinned separately), and the closed-form SH rotation against an
independent least-squares construction from the stated e3nn basis."""
import importlib
import sys
import types

import numpy as np
import pytest
import torch

from helpers import REFERENCE_SRC, have_reference


def _e3nn_sh(x: torch.Tensor, l: int) -> torch.Tensor:
    """e3nn's real spherical harmonics of degree l ('component' normalisation) as restated in gaussian_adapter.py."""
    X, Y, Z = x.unbind(-1)
    if l == 0:
        return torch.ones_like(X)[..., None]
    if l == 1:
        return 3 ** 0.5 * torch.stack([X, Y, Z], -1)
    s15, s5 = 15 ** 0.5, 5 ** 0.5
    return torch.stack([s15 * X * Z, s15 * X * Y, s5 * (Y * Y - 0.5 * (X * X + Z * Z)), s15 * Y * Z, 0.5 * s15 * (Z * Z - X * X)], -1)


def _random_rotations(n, seed):
    g = torch.Generator().manual_seed(seed)
    q, _ = torch.linalg.qr(torch.randn(n, 3, 3, generator=g, dtype=torch.float64))
    return q * torch.sign(torch.linalg.det(q))[:, None, None]


def test_closed_form_sh_rotation_equals_least_squares_construction():
    from my_depthsplat_b200.gaussian_adapter import rotate_sh, sh_rotation_matrices
    R = _random_rotations(6, 0)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(200, 3, generator=g, dtype=torch.float64)
    x = x / x.norm(dim=-1, keepdim=True)
    mats = sh_rotation_matrices(R, 2)
    for r in range(R.shape[0]):
        xr = x @ R[r].T                                                   # R x for every sample direction
        for l in (0, 1, 2):
            # D with Y(R x) = D Y(x): least squares over 200 directions, independent of the closed form
            D = torch.linalg.lstsq(_e3nn_sh(x, l), _e3nn_sh(xr, l)).solution.T
            torch.testing.assert_close(mats[l][r], D, rtol=0, atol=1e-10)
            torch.testing.assert_close(D @ D.T, torch.eye(2 * l + 1, dtype=torch.float64), rtol=0, atol=1e-10)   # orthogonal
    # composition D(R1 R2) = D(R1) D(R2), identity at the identity
    m01 = sh_rotation_matrices(R[0] @ R[1], 2)
    for l in (1, 2):
        torch.testing.assert_close(m01[l], mats[l][0] @ mats[l][1], rtol=0, atol=1e-10)
    eye = sh_rotation_matrices(torch.eye(3, dtype=torch.float64), 2)
    for m in eye:
        torch.testing.assert_close(m, torch.eye(m.shape[-1], dtype=torch.float64), rtol=0, atol=1e-14)
    # rotate_sh applies the blocks per degree and broadcasts the rotation
    c = torch.randn(4, 3, 9, generator=g, dtype=torch.float64)
    out = rotate_sh(c, R[2][None, None])
    torch.testing.assert_close(out[..., 4:9], torch.einsum("ij,...j->...i", mats[2][2], c[..., 4:9]))
    torch.testing.assert_close(out[..., 0], c[..., 0])


@pytest.mark.skipif(not have_reference(), reason="/root/reference not available")
def test_adapter_equals_the_reference_file():
    from my_depthsplat_b200 import gaussian_adapter as ours
    root = str(REFERENCE_SRC)
    for name, path in [("src", root), ("src.model", root + "/model"), ("src.model.encoder", root + "/model/encoder"),
                       ("src.model.encoder.common", root + "/model/encoder/common"), ("src.geometry", root + "/geometry"), ("src.misc", root + "/misc")]:
        pkg = types.ModuleType(name)
        pkg.__path__ = [path]
        sys.modules[name] = pkg
    # e3nn is not installed: the reference's rotate_sh is replaced by ours for THIS comparison (it is pinned separately above)
    stub = types.ModuleType("src.misc.sh_rotation")
    stub.rotate_sh = ours.rotate_sh
    sys.modules["src.misc.sh_rotation"] = stub
    for name in ("src.model.encoder.common.gaussian_adapter", "src.model.encoder.common.gaussians", "src.geometry.projection"):
        sys.modules.pop(name, None)
    ref = importlib.import_module("src.model.encoder.common.gaussian_adapter")

    g = torch.Generator().manual_seed(5)
    b, v, h, w = 2, 3, 6, 8
    cfg = dict(gaussian_scale_min=1e-10, gaussian_scale_max=3.0, sh_degree=2)
    r_ad, o_ad = ref.GaussianAdapter(ref.GaussianAdapterCfg(**cfg)), ours.GaussianAdapter(ours.GaussianAdapterCfg(**cfg))
    ext = torch.eye(4).repeat(b, v, 1, 1)
    ext[..., :3, :3] = _random_rotations(b * v, 2).float().reshape(b, v, 3, 3)
    ext[..., :3, 3] = torch.randn(b, v, 3, generator=g)
    K = torch.tensor([[0.8, 0.0, 0.5], [0.0, 1.1, 0.5], [0.0, 0.0, 1.0]]).repeat(b, v, 1, 1)
    head = torch.randn(b, v, 37, h, w, generator=g)
    depth = torch.rand(b, v, h, w, generator=g) * 5 + 1
    images = torch.rand(b, v, 3, h, w, generator=g)
    got = ours.adapt_head_output(o_ad, head, depth, images, ext, K, (h, w))
    # the reference, driven the way its encoder drives it (encoder_depthsplat.py:226-312)
    from einops import rearrange
    proj = importlib.import_module("src.geometry.projection")
    raw = rearrange(head, "b v c h w -> b v (h w) c")
    opac = raw[..., :1].sigmoid().unsqueeze(-1)
    raw = raw[..., 1:]
    xy_ray, _ = proj.sample_image_grid((h, w), head.device)
    xy_ray = rearrange(xy_ray, "h w xy -> (h w) () xy")
    gs = rearrange(raw, "... (srf c) -> ... srf c", srf=1)
    xy_ray = xy_ray + (gs[..., :2].sigmoid() - 0.5) * (1 / torch.tensor((w, h), dtype=torch.float32))
    want = r_ad.forward(rearrange(ext, "b v i j -> b v () () () i j"), rearrange(K, "b v i j -> b v () () () i j"),
                        rearrange(xy_ray, "b v r srf xy -> b v r srf () xy"), rearrange(depth, "b v h w -> b v (h w) () ()"), opac,
                        rearrange(gs[..., 2:], "b v r srf c -> b v r srf () c"), (h, w), input_images=images)
    torch.testing.assert_close(got.means, rearrange(want.means, "b v r srf spp xyz -> b (v r srf spp) xyz"), rtol=1e-5, atol=1e-6)  # einsum vs matmul summation order
    torch.testing.assert_close(got.covariances, rearrange(want.covariances, "b v r srf spp i j -> b (v r srf spp) i j"), rtol=1e-5, atol=1e-6)  # einsum vs matmul summation order
    torch.testing.assert_close(got.harmonics, rearrange(want.harmonics, "b v r srf spp c d_sh -> b (v r srf spp) c d_sh"), rtol=1e-5, atol=1e-6)  # einsum vs matmul summation order
    torch.testing.assert_close(got.opacities, rearrange(want.opacities, "b v r srf spp -> b (v r srf spp)"), rtol=1e-5, atol=1e-6)  # einsum vs matmul summation order
