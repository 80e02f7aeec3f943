"""Golden fixtures: outputs of the reference's UNMODIFIED Python glue running on the oracle
(tests/golden/make_golden.py, generated in the build container).  CPU: the oracle-backed helper still
reproduces them bit for bit (guards the oracle and the helper against drift).  GPU: the product's
decoder matches them within the north-star tolerances."""
from pathlib import Path

import hashlib
import numpy as np
import pytest
import torch

from helpers import input_digest, leaf_gaussians, oracle_decoder_forward, per_view_extension_inputs
from my_depthsplat_b200.scenes import make_scene

GOLDEN = Path(__file__).resolve().parent / "golden"
FIXTURES = [("tiny", "depth"), ("ragged", None), ("tiny", "disparity")]


def _load(name, depth_mode):
    return np.load(GOLDEN / f"{name}_{depth_mode or 'color'}.npz")


@pytest.mark.parametrize("name,depth_mode", FIXTURES)
def test_oracle_reproduces_golden(name, depth_mode):
    from oracle import splat_oracle as so
    gold = _load(name, depth_mode)
    scene = make_scene(name)
    # Bit-level comparison needs bit-identical inputs.  The scene and the camera block are built with host BLAS
    # (conftest.py pins MKL to its portable path); on a host where they still come out with other last bits, the
    # oracle is held to the north-star tolerances instead and the stage digests are not comparable.
    same_inputs = input_digest(scene) == str(gold["input_digest"])
    if not same_inputs:
        import warnings
        warnings.warn(f"{name}: this host builds the fixture's inputs with different last bits; tolerance comparison")

    def same(got, ref):
        if same_inputs:
            np.testing.assert_array_equal(got, ref)
        else:
            err = np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)
            assert (err > 1e-5).mean() <= 1e-3, err.max()

    g = leaf_gaussians(scene)
    color, depth = oracle_decoder_forward(g, scene.extrinsics, scene.intrinsics, scene.near, scene.far, scene.image_shape,
                                          scene.background, depth_mode)
    same(color.detach().numpy(), gold["color"])
    loss = (color * scene.grad_color).sum()
    if depth_mode is not None:
        same(depth.detach().numpy(), gold["depth"])
        loss = loss + (depth * scene.grad_depth).sum()
    loss.backward()
    for key, t in (("d_means", g.means), ("d_covariances", g.covariances), ("d_harmonics", g.harmonics), ("d_opacities", g.opacities)):
        ref = gold[key]
        if same_inputs:
            np.testing.assert_allclose(t.grad.numpy(), ref, rtol=1e-5, atol=1e-7 * np.abs(ref).max())
        else:
            assert np.abs(t.grad.numpy() - ref).max() <= 2e-4 * np.abs(ref).max(), key
    if not same_inputs:
        return
    B, V = scene.extrinsics.shape[:2]
    digests = []
    for b in range(B):
        for v in range(V):
            st = so.forward_view(**per_view_extension_inputs(scene, b, v))
            h = hashlib.sha256()
            for a in (st.keys, st.vals, st.ranges, st.radii, st.n_contrib):
                h.update(np.ascontiguousarray(a).tobytes())
            digests.append(h.hexdigest())
    assert digests == list(gold["stage_digests"])


@pytest.mark.gpu
@pytest.mark.parametrize("name,depth_mode", FIXTURES)
def test_product_matches_golden(name, depth_mode):
    from my_depthsplat_b200.decoder_splatting_cuda import DecoderSplattingCUDACfg, get_decoder
    from my_depthsplat_b200.types import Gaussians
    gold = _load(name, depth_mode)
    scene = make_scene(name)
    g = scene.gaussians
    leaves = [t.detach().clone().cuda().requires_grad_() for t in (g.means, g.covariances, g.harmonics, g.opacities)]
    dec = get_decoder(DecoderSplattingCUDACfg(name="splatting_cuda"), type("D", (), {"background_color": [0.0, 0.0, 0.0]})()).cuda()
    out = dec.forward(Gaussians(*leaves), scene.extrinsics.cuda(), scene.intrinsics.cuda(), scene.near.cuda(), scene.far.cuda(),
                      scene.image_shape, depth_mode=depth_mode)
    err = np.abs(out.color.detach().cpu().numpy() - gold["color"])
    assert (err > 1e-5).mean() <= 1e-3, (err.max(), (err > 1e-5).mean())   # 1e-5 abs; threshold flips bounded
    loss = (out.color * scene.grad_color.cuda()).sum()
    if depth_mode is not None:
        ref = gold["depth"]
        derr = np.abs(out.depth.detach().cpu().numpy() - ref) / np.maximum(np.abs(ref), 1.0)
        assert (derr > 1e-5).mean() <= 1e-3, derr.max()
        loss = loss + (out.depth * scene.grad_depth.cuda()).sum()
    loss.backward()
    for key, t in zip(("d_means", "d_covariances", "d_harmonics", "d_opacities"), leaves):
        ref = gold[key]
        # cameras are built with CUDA torch ops here (last-bit differences from the CPU-built ones of the
        # fixture can move a Gaussian across a tile-rect boundary), hence 2e-4 instead of 1e-4
        assert np.abs(t.grad.cpu().numpy() - ref).max() <= 2e-4 * np.abs(ref).max(), key
