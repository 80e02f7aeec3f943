"""The GPU comparator (baseline/upstream_style.cu + baseline/per_view_glue.py: the upstream design restated, what
bench.py reports as "gpu_baseline") computes the same thing as the oracle and the product -- otherwise the
speed-up measured against it would mean nothing.  It is not product code and nothing in my_depthsplat_b200
imports it."""
import numpy as np
import pytest
import torch

from helpers import leaf_gaussians, oracle_decoder_forward
from my_depthsplat_b200.scenes import make_scene
from my_depthsplat_b200.types import Gaussians

pytestmark = pytest.mark.gpu


def _baseline(scene_gpu, leaves, depth_mode=None):
    from baseline import per_view_glue, upstream_ext
    if not upstream_ext.available():
        pytest.skip("comparator not built (make -C baseline); it is not part of the product")
    return per_view_glue.decoder_forward(upstream_ext, Gaussians(*leaves), scene_gpu.extrinsics, scene_gpu.intrinsics, scene_gpu.near,
                                         scene_gpu.far, scene_gpu.image_shape, scene_gpu.background, depth_mode)


@pytest.mark.parametrize("name,depth_mode", [("tiny", "depth"), ("small", None), ("ragged", None)])
def test_comparator_matches_the_oracle(name, depth_mode):
    cpu = make_scene(name)
    gc = leaf_gaussians(cpu)
    ref_c, ref_d = oracle_decoder_forward(gc, cpu.extrinsics, cpu.intrinsics, cpu.near, cpu.far, cpu.image_shape, cpu.background, depth_mode)
    loss = (ref_c * cpu.grad_color).sum()
    if depth_mode is not None:
        loss = loss + (ref_d * cpu.grad_depth).sum()
    loss.backward()
    sc = cpu.to("cuda")
    leaves = [t.detach().clone().requires_grad_() for t in (sc.gaussians.means, sc.gaussians.covariances, sc.gaussians.harmonics, sc.gaussians.opacities)]
    col, dep = _baseline(sc, leaves, depth_mode)
    loss = (col * sc.grad_color).sum()
    if depth_mode is not None:
        loss = loss + (dep * sc.grad_depth).sum()
    loss.backward()
    err = np.abs(col.detach().cpu().numpy() - ref_c.detach().numpy())
    assert (err > 1e-5).mean() <= 2e-3, (err.max(), (err > 1e-5).mean())
    if depth_mode is not None:
        derr = np.abs(dep.detach().cpu().numpy() - ref_d.detach().numpy())
        assert (derr > 1e-4 * max(1.0, float(ref_d.detach().abs().max()))).mean() <= 2e-3
    for got, want in zip(leaves, (gc.means, gc.covariances, gc.harmonics, gc.opacities)):
        r = want.grad.numpy()
        e = np.abs(got.grad.cpu().numpy() - r)
        assert np.quantile(e, 0.999) <= 1e-4 * np.abs(r).max() and e.max() <= 5e-2 * np.abs(r).max()


def test_comparator_matches_the_product_at_full_size():
    from my_depthsplat_b200.cuda_splatting import render_views
    sc = make_scene("C1").to("cuda")
    g = sc.gaussians
    l1 = [t.detach().clone().requires_grad_() for t in (g.means, g.covariances, g.harmonics, g.opacities)]
    l2 = [t.detach().clone().requires_grad_() for t in (g.means, g.covariances, g.harmonics, g.opacities)]
    c1, _ = render_views(sc.extrinsics, sc.intrinsics, sc.near, sc.far, sc.image_shape, sc.background, *l1)
    c2, _ = _baseline(sc, l2)
    # the comparator is compiled from plain expressions (nvcc picks the contractions), the product spells them out to
    # match the oracle: a handful of Gaussians land on the other side of a radius / tile-rect / alpha threshold
    err = (c1 - c2).detach().abs()
    assert float((err > 1e-5).float().mean()) <= 2e-3 and float(err.max()) <= 5e-2, (float(err.max()), float((err > 1e-5).float().mean()))
    (c1 * sc.grad_color).sum().backward()
    (c2 * sc.grad_color).sum().backward()
    for a, b in zip(l1, l2):
        scale = float(b.grad.abs().max())
        e = (a.grad - b.grad).abs().flatten()
        assert float(torch.quantile(e[:: max(1, e.numel() // 4_000_000)], 0.999)) <= 1e-4 * scale and float(e.max()) <= 5e-2 * scale
