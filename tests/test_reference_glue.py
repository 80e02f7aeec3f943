"""The reference's UNMODIFIED Python glue (src/model/decoder/cuda_splatting.py, imported from
/root/reference with stub parent packages) against (a) the test helper that restates it on top of the
oracle and (b) the product's own camera code and signatures.  Skipped where /root/reference is absent
(the GPU box); the golden fixtures carry its outputs there."""
import inspect

import numpy as np
import pytest
import torch
from einops import rearrange, repeat

from helpers import have_reference, leaf_gaussians, load_reference_cuda_splatting, oracle_decoder_forward, per_view_extension_inputs
from my_depthsplat_b200.scenes import make_scene

pytestmark = pytest.mark.skipif(not have_reference(), reason="/root/reference not available")


def _ref_render(cs, scene, g, depth_mode=None):
    B, V = scene.extrinsics.shape[:2]
    f = lambda t, p: rearrange(t, p)
    cams = (f(scene.extrinsics, "b v i j -> (b v) i j"), f(scene.intrinsics, "b v i j -> (b v) i j"),
            f(scene.near, "b v -> (b v)"), f(scene.far, "b v -> (b v)"))
    rep = lambda t, p: repeat(t, p, v=V)
    color = cs.render_cuda(*cams, scene.image_shape, repeat(scene.background, "c -> (b v) c", b=B, v=V),
                           rep(g.means, "b g xyz -> (b v) g xyz"), rep(g.covariances, "b g i j -> (b v) g i j"),
                           rep(g.harmonics, "b g c d -> (b v) g c d"), rep(g.opacities, "b g -> (b v) g"))
    color = rearrange(color, "(b v) c h w -> b v c h w", b=B)
    depth = None
    if depth_mode is not None:
        depth = cs.render_depth_cuda(*cams, scene.image_shape, rep(g.means, "b g xyz -> (b v) g xyz"),
                                     rep(g.covariances, "b g i j -> (b v) g i j"), rep(g.opacities, "b g -> (b v) g"), mode=depth_mode)
        depth = rearrange(depth, "(b v) h w -> b v h w", b=B)
    return color, depth


@pytest.mark.parametrize("name,depth_mode", [("tiny", "depth"), ("small", None), ("tiny", "disparity"), ("tiny", "log")])
def test_helper_restates_the_reference_glue(name, depth_mode):
    from oracle import ext_compat
    cs = load_reference_cuda_splatting(ext_compat)
    scene = make_scene(name)
    g1, g2 = leaf_gaussians(scene), leaf_gaussians(scene)
    c1, d1 = _ref_render(cs, scene, g1, depth_mode)
    c2, d2 = oracle_decoder_forward(g2, scene.extrinsics, scene.intrinsics, scene.near, scene.far, scene.image_shape,
                                    scene.background, depth_mode)
    assert torch.equal(c1, c2)
    loss1, loss2 = (c1 * scene.grad_color).sum(), (c2 * scene.grad_color).sum()
    if depth_mode is not None:
        assert torch.equal(d1, d2)
        loss1, loss2 = loss1 + (d1 * scene.grad_depth).sum(), loss2 + (d2 * scene.grad_depth).sum()
    loss1.backward(); loss2.backward()
    for a, b in ((g1.means, g2.means), (g1.covariances, g2.covariances), (g1.harmonics, g2.harmonics), (g1.opacities, g2.opacities)):
        torch.testing.assert_close(a.grad, b.grad, rtol=1e-5, atol=1e-7 * float(b.grad.abs().max()))


def test_product_camera_block_is_what_the_reference_passes_to_the_extension():
    """Record the arguments the reference hands to GaussianRasterizer for every view and compare them,
    bit for bit, with the per-view inputs the product derives (same torch ops, batched)."""
    import types
    rec = []

    class Settings:
        def __init__(self, **kw):
            self.__dict__.update(kw)

    class Rasterizer:
        def __init__(self, s):
            self.s = s

        def __call__(self, **kw):
            rec.append((self.s, kw))
            return torch.zeros(3, self.s.image_height, self.s.image_width), torch.zeros(kw["means3D"].shape[0], dtype=torch.int32)

    stub = types.ModuleType("diff_gaussian_rasterization")
    stub.GaussianRasterizationSettings = Settings
    stub.GaussianRasterizer = Rasterizer
    cs = load_reference_cuda_splatting(stub)
    scene = make_scene("small")
    _ref_render(cs, scene, scene.gaussians)
    B, V = scene.extrinsics.shape[:2]
    assert len(rec) == B * V
    for i, (s, kw) in enumerate(rec):
        b, v = divmod(i, V)
        inp = per_view_extension_inputs(scene, b, v)
        assert s.image_height == inp["H"] and s.image_width == inp["W"] and s.sh_degree == inp["sh_degree"]
        assert s.tanfovx == inp["tanfovx"] and s.tanfovy == inp["tanfovy"]
        np.testing.assert_array_equal(s.viewmatrix.numpy().reshape(16), inp["viewmatrix"].reshape(16))
        np.testing.assert_array_equal(s.projmatrix.numpy().reshape(16), inp["projmatrix"].reshape(16))
        np.testing.assert_array_equal(s.campos.numpy(), inp["campos"])
        np.testing.assert_array_equal(kw["means3D"].numpy(), inp["means3D"])
        np.testing.assert_array_equal(kw["cov3D_precomp"].numpy(), inp["cov3D"])
        np.testing.assert_array_equal(kw["shs"].numpy(), inp["shs"])
        np.testing.assert_array_equal(kw["opacities"].numpy()[:, 0], inp["opacities"])
        assert kw["colors_precomp"] is None and s.scale_modifier == 1.0 and s.prefiltered is False


def test_signatures_match_the_reference():
    from my_depthsplat_b200 import cuda_splatting as mine
    import types
    stub = types.ModuleType("diff_gaussian_rasterization")
    stub.GaussianRasterizationSettings = object
    stub.GaussianRasterizer = object
    cs = load_reference_cuda_splatting(stub)
    for fn in ("get_projection_matrix", "render_cuda", "render_cuda_orthographic", "render_depth_cuda"):
        ref_sig, my_sig = inspect.signature(getattr(cs, fn)), inspect.signature(getattr(mine, fn))
        assert list(ref_sig.parameters) == list(my_sig.parameters), fn
        for n, p in ref_sig.parameters.items():
            assert p.default == my_sig.parameters[n].default, (fn, n)
    # decoder: same constructor / forward parameter names as decoder_splatting_cuda.py:22-44, 69-78
    from my_depthsplat_b200.decoder_splatting_cuda import DECODERS, DecoderSplattingCUDA
    assert list(inspect.signature(DecoderSplattingCUDA.__init__).parameters) == ["self", "cfg", "dataset_cfg"]
    assert list(inspect.signature(DecoderSplattingCUDA.forward).parameters) == [
        "self", "gaussians", "extrinsics", "intrinsics", "near", "far", "image_shape", "depth_mode"]
    assert list(inspect.signature(DecoderSplattingCUDA.render_depth).parameters) == [
        "self", "gaussians", "extrinsics", "intrinsics", "near", "far", "image_shape", "mode"]
    assert list(DECODERS) == ["splatting_cuda"]
    ref_proj = cs.get_projection_matrix(torch.tensor([1.0, 0.5]), torch.tensor([200.0, 50.0]), torch.tensor([1.0, 0.7]), torch.tensor([0.8, 0.9]))
    my_proj = mine.get_projection_matrix(torch.tensor([1.0, 0.5]), torch.tensor([200.0, 50.0]), torch.tensor([1.0, 0.7]), torch.tensor([0.8, 0.9]))
    assert torch.equal(ref_proj, my_proj)
    from my_depthsplat_b200.projection import get_fov
    import importlib
    ref_projection = importlib.import_module("src.geometry.projection")
    K = make_scene("small").intrinsics.reshape(-1, 3, 3)
    assert torch.equal(ref_projection.get_fov(K), get_fov(K))


@pytest.mark.parametrize("name,depth_mode", [("tiny", "depth"), ("small", None)])
def test_comparator_glue_restates_the_reference_glue(name, depth_mode):
    """baseline/per_view_glue.py (what bench.py's gpu_baseline runs over the upstream-style CUDA comparator) executes the
    reference's glue: same images and gradients as the unmodified reference file, both over the CPU oracle."""
    from baseline import per_view_glue
    from oracle import ext_compat
    cs = load_reference_cuda_splatting(ext_compat)
    scene = make_scene(name)
    g1, g2 = leaf_gaussians(scene), leaf_gaussians(scene)
    c1, d1 = _ref_render(cs, scene, g1, depth_mode)
    c2, d2 = per_view_glue.decoder_forward(ext_compat, g2, scene.extrinsics, scene.intrinsics, scene.near, scene.far,
                                           scene.image_shape, scene.background, depth_mode)
    assert torch.equal(c1, c2)
    loss1, loss2 = (c1 * scene.grad_color).sum(), (c2 * scene.grad_color).sum()
    if depth_mode is not None:
        assert torch.equal(d1, d2)
        loss1, loss2 = loss1 + (d1 * scene.grad_depth).sum(), loss2 + (d2 * scene.grad_depth).sum()
    loss1.backward(); loss2.backward()
    for a, b in ((g1.means, g2.means), (g1.covariances, g2.covariances), (g1.harmonics, g2.harmonics), (g1.opacities, g2.opacities)):
        assert torch.equal(a.grad, b.grad)


def test_comparator_glue_restates_the_reference_orthographic_variant():
    """cuda_splatting.py:129-219 (one view, as the reference uses it): the restatement the GPU box checks the product
    against equals the unmodified reference, images and gradients, both over the CPU oracle."""
    from baseline import per_view_glue
    from oracle import ext_compat
    cs = load_reference_cuda_splatting(ext_compat)
    scene = make_scene("tiny")
    g1, g2 = leaf_gaussians(scene), leaf_gaussians(scene)
    ext = scene.extrinsics[0, :1]
    width, height, near, far = torch.tensor([3.0]), torch.tensor([2.0]), torch.tensor([0.0]), torch.tensor([20.0])
    args = lambda g: (ext, width, height, near, far, (32, 48), torch.zeros(1, 3), g.means[:1], g.covariances[:1], g.harmonics[:1], g.opacities[:1])
    a = cs.render_cuda_orthographic(*args(g1))
    b = per_view_glue.render_cuda_orthographic(ext_compat, *args(g2))
    assert float(a.abs().max()) > 0 and torch.equal(a, b)
    gen = torch.Generator().manual_seed(4)
    w = torch.randn(a.shape, generator=gen)
    (a * w).sum().backward(); (b * w).sum().backward()
    for x, y in ((g1.means, g2.means), (g1.covariances, g2.covariances), (g1.harmonics, g2.harmonics), (g1.opacities, g2.opacities)):
        assert torch.equal(x.grad, y.grad)
