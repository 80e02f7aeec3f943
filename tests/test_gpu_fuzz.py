"""Seeded random scene shapes against the oracle: image sizes that are no multiple of the 16x16 tile (down to a single
partial tile), 1-3 scenes x 1-3 views, every scale regime (incl. near-plane clustering that makes tile lists long and
footprints larger than the image), coloured backgrounds, depth modes.  Colour / depth within 1e-5 (flipped threshold
decisions bounded), gradients within 1e-4 of their scale."""
import numpy as np
import pytest
import torch

from helpers import leaf_gaussians, oracle_decoder_forward
from my_depthsplat_b200.scenes import SceneConfig, make_scene

pytestmark = pytest.mark.gpu


def _cases():
    rng = np.random.default_rng(2024)
    out = []
    for i in range(10):
        h, w = int(rng.integers(5, 75)), int(rng.integers(5, 90))
        mode = ["init", "trained", "stress"][i % 3]
        out.append(dict(
            cfg=SceneConfig(f"fuzz{i}", 500 + i, int(rng.integers(1, 4)), h, w, batch=int(rng.integers(1, 4)), v_tgt=int(rng.integers(1, 4)),
                            scale_mode=mode, scale_max=float([3.0, 0.1, 0.5][i % 3]), near_plane_fraction=0.2 if mode == "stress" else 0.0,
                            far=float(rng.choice([20.0, 100.0, 400.0]))),
            depth_mode=[None, "depth", "disparity", "log"][i % 4],
            bg=torch.tensor(rng.random(3), dtype=torch.float32)))
    return out


@pytest.mark.parametrize("case", _cases(), ids=lambda c: f"{c['cfg'].name}-{c['cfg'].height}x{c['cfg'].width}-b{c['cfg'].batch}v{c['cfg'].v_tgt}-{c['cfg'].scale_mode}-{c['depth_mode']}")
def test_random_scene_against_the_oracle(case):
    from my_depthsplat_b200.cuda_splatting import render_views
    cpu = make_scene(case["cfg"])
    dm, bg = case["depth_mode"], case["bg"]
    gc = leaf_gaussians(cpu)
    ref_c, ref_d = oracle_decoder_forward(gc, cpu.extrinsics, cpu.intrinsics, cpu.near, cpu.far, cpu.image_shape, bg, dm)
    loss = (ref_c * cpu.grad_color).sum()
    if dm is not None:
        loss = loss + (ref_d * cpu.grad_depth).sum()
    loss.backward()
    sc = cpu.to("cuda")
    leaves = [t.detach().clone().requires_grad_() for t in (sc.gaussians.means, sc.gaussians.covariances, sc.gaussians.harmonics, sc.gaussians.opacities)]
    col, dep = render_views(sc.extrinsics, sc.intrinsics, sc.near, sc.far, sc.image_shape, bg.cuda(), *leaves, depth_mode=dm)
    loss = (col * sc.grad_color).sum()
    if dm is not None:
        loss = loss + (dep * sc.grad_depth).sum()
    loss.backward()
    # cameras are built with CUDA torch ops here and CPU ones for the oracle (last-bit differences in the matrices): on
    # these tiny images a single flipped rect / threshold decision is a visible fraction of the pixels
    cerr = (col.detach().cpu() - ref_c.detach()).abs()
    assert float((cerr > 1e-5).float().mean()) <= 5e-3 and float(cerr.max()) <= 5e-2, (float(cerr.max()), float((cerr > 1e-5).float().mean()))
    if dm is not None:
        derr = (dep.detach().cpu() - ref_d.detach()).abs() / ref_d.detach().abs().clamp(min=1.0)
        assert float((derr > 1e-5).float().mean()) <= 5e-3, float(derr.max())
    for got, want, nm in zip(leaves, (gc.means, gc.covariances, gc.harmonics, gc.opacities), ("means", "covariances", "harmonics", "opacities")):
        r = want.grad.numpy()
        e = np.abs(got.grad.cpu().numpy() - r)
        scale = np.abs(r).max()
        if scale > 0:
            assert np.quantile(e, 0.999) <= 1e-4 * scale and e.max() <= 5e-2 * scale, (nm, e.max() / scale)
