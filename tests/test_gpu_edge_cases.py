"""GPU edge cases of the path: ragged / misaligned shapes, every SH degree and both tensor layouts, nothing
visible, a single Gaussian, non-zero backgrounds, the orthographic variant, very large footprints."""
import numpy as np
import pytest
import torch

from helpers import oracle_decoder_forward, per_view_extension_inputs
from my_depthsplat_b200.scenes import SceneConfig, make_scene
from my_depthsplat_b200.types import Gaussians

pytestmark = pytest.mark.gpu


def _subset(scene, n, batch=None):
    g = scene.gaussians
    b = slice(None) if batch is None else slice(0, batch)
    return Gaussians(g.means[b, :n].contiguous(), g.covariances[b, :n].contiguous(), g.harmonics[b, :n].contiguous(),
                     g.opacities[b, :n].contiguous())


def _cuda(g, grad=False):
    mk = lambda t: t.detach().clone().cuda().requires_grad_(grad)
    return Gaussians(mk(g.means), mk(g.covariances), mk(g.harmonics), mk(g.opacities))


def _render(scene, g, depth_mode=None, bg=None):
    from my_depthsplat_b200.cuda_splatting import render_views
    return render_views(scene.extrinsics.cuda(), scene.intrinsics.cuda(), scene.near.cuda(), scene.far.cuda(), scene.image_shape,
                        (scene.background if bg is None else bg).cuda(), g.means, g.covariances, g.harmonics, g.opacities, depth_mode=depth_mode)


@pytest.mark.parametrize("n", [1, 3, 255, 257, 1001])
def test_odd_gaussian_counts_and_misaligned_scenes(n):
    """N not a multiple of 4 (16-byte staging falls back to scalar loads) and B = 2 (second scene's base pointer
    misaligned) -- forward and gradients against the oracle."""
    scene = make_scene("small")  # B = 2
    gs = _subset(scene, n)
    ref_g = Gaussians(*(t.clone().requires_grad_() for t in (gs.means, gs.covariances, gs.harmonics, gs.opacities)))
    ref_c, ref_d = oracle_decoder_forward(ref_g, scene.extrinsics, scene.intrinsics, scene.near, scene.far, scene.image_shape,
                                          scene.background, "depth")
    ((ref_c * scene.grad_color).sum() + (ref_d * scene.grad_depth).sum()).backward()
    g = _cuda(gs, grad=True)
    color, depth = _render(scene, g, "depth")
    ((color * scene.grad_color.cuda()).sum() + (depth * scene.grad_depth.cuda()).sum()).backward()
    assert ((color.cpu() - ref_c).abs() > 1e-5).float().mean() <= 1e-3
    assert (((depth.cpu() - ref_d).abs() / ref_d.abs().clamp(min=1)) > 1e-5).float().mean() <= 1e-3
    for got, want in ((g.means, ref_g.means), (g.covariances, ref_g.covariances), (g.harmonics, ref_g.harmonics), (g.opacities, ref_g.opacities)):
        scale = float(want.grad.abs().max())
        if scale > 0:
            assert float((got.grad.cpu() - want.grad).abs().max()) <= 2e-4 * scale


@pytest.mark.parametrize("name,views", [("tiny", 19), ("small", 9)])
def test_more_views_than_one_camera_group(name, views, capsys):
    """The projection kernels load the camera blocks of a scene's views eight at a time (common.cuh VIEW_GROUP): 19 views of
    one scene (three groups, the last one partial) and 2 scenes x 9 views (a scene's views straddle group boundaries, the
    other scene's views in between are skipped) in ONE call, held to the strict bars of helpers.strict_parity_check: stages
    bit-exact per view, colour + depth within 1e-5 off the fragile pixels, gradients summed over all views within 1e-4 of
    their scale on every Gaussian that touches no flipped pixel."""
    from helpers import strict_parity_check
    report = strict_parity_check(make_scene(name, v_tgt=views), "depth", f"{name} x{views} views")
    with capsys.disabled():
        print("\n" + "\n".join(report))


@pytest.mark.parametrize("degree", [0, 1, 2, 3])
@pytest.mark.parametrize("layout", ["channel_major", "coeff_major"])
def test_sh_degrees_and_layouts(degree, layout):
    """d_sh = 1, 4, 9, 16 through both tensor layouts (DepthSplat's [N,3,d] and the extension's [N,d,3], the latter
    with [N,6] covariances), one view, against the oracle."""
    from my_depthsplat_b200 import _lib
    from my_depthsplat_b200.rasterizer import ViewPack, rasterize
    from oracle import splat_oracle as so
    cfg = SceneConfig(f"deg{degree}", 100 + degree, 2, 48, 64, sh_degree=degree)
    scene = make_scene(cfg)
    inp = per_view_extension_inputs(scene, 0, 1, scale_invariant=False)
    st = so.forward_view(**inp)
    t = lambda a: torch.tensor(np.ascontiguousarray(a)).cuda()
    pack = ViewPack(torch.zeros(1, dtype=torch.int32, device="cuda"), t(inp["viewmatrix"]).reshape(1, 4, 4), t(inp["projmatrix"]).reshape(1, 4, 4),
                    t(inp["campos"]).reshape(1, 3), torch.tensor([[inp["tanfovx"], inp["tanfovy"]]], device="cuda"), t(inp["bg"]).reshape(1, 3),
                    inp["H"], inp["W"])
    means = t(inp["means3D"])[None].requires_grad_(); op = t(inp["opacities"])[None].requires_grad_()
    if layout == "coeff_major":
        cov = t(inp["cov3D"])[None].requires_grad_(); sh = t(inp["shs"])[None].requires_grad_()
        lay = _lib.SH_COEFF_MAJOR
    else:
        cov = scene.gaussians.covariances[:1].clone().cuda().requires_grad_(); sh = scene.gaussians.harmonics[:1].clone().cuda().requires_grad_()
        lay = _lib.SH_CHANNEL_MAJOR
    color, _, radii = rasterize(means, cov, sh, op, pack, use_sh=True, sh_degree=degree, sh_layout=lay, want_radii=True)
    np.testing.assert_array_equal(radii[0].cpu().numpy(), st.radii)
    assert (np.abs(color[0].detach().cpu().numpy() - st.color) > 1e-5).mean() <= 1e-3
    gpix = scene.grad_color[0, 1]
    (color[0] * gpix.cuda()).sum().backward()
    ref = so.backward_view(st, gpix.numpy())
    got_sh = sh.grad[0].cpu().numpy() if layout == "coeff_major" else sh.grad[0].permute(0, 2, 1).cpu().numpy()
    got_cov = cov.grad[0].cpu().numpy() if layout == "coeff_major" else cov.grad[0].cpu().numpy()[:, [0, 0, 0, 1, 1, 2], [0, 1, 2, 1, 2, 2]]
    for got, key in ((means.grad[0].cpu().numpy(), "means3D"), (got_sh, "sh"), (got_cov, "cov3D"), (op.grad[0].cpu().numpy(), "opacity")):
        assert np.abs(got - ref[key]).max() <= 1e-4 * np.abs(ref[key]).max(), key


def test_nothing_visible_renders_the_background_and_zero_gradients():
    scene = make_scene("tiny")
    g = scene.gaussians
    behind = Gaussians(g.means * torch.tensor([1.0, 1.0, -1.0]), g.covariances, g.harmonics, g.opacities)
    gc = _cuda(behind, grad=True)
    bg = torch.tensor([0.25, 0.5, 0.75])
    color, depth = _render(scene, gc, "depth", bg=bg)
    assert torch.equal(color.cpu(), bg[None, None, :, None, None].expand_as(color)) and float(depth.abs().max()) == 0.0
    (color.sum() + depth.sum()).backward()
    assert all(float(t.grad.abs().max()) == 0.0 for t in (gc.means, gc.covariances, gc.harmonics, gc.opacities))


def test_background_and_transparent_gaussians():
    """Opacity below 1/255 can never contribute; a coloured background shows through with weight T."""
    scene = make_scene("tiny")
    g = scene.gaussians
    faint = Gaussians(g.means, g.covariances, g.harmonics, torch.full_like(g.opacities, 0.003))
    bg = torch.tensor([0.9, 0.1, 0.4])
    with torch.no_grad():
        color, _ = _render(scene, _cuda(faint), bg=bg)
        ref, _ = oracle_decoder_forward(g, scene.extrinsics, scene.intrinsics, scene.near, scene.far, scene.image_shape, bg)
        full, _ = _render(scene, _cuda(g), bg=bg)
    assert torch.equal(color.cpu(), bg[None, None, :, None, None].expand_as(color))
    assert ((full.cpu() - ref).abs() > 1e-5).float().mean() <= 1e-3


def test_orthographic_variant_matches_the_oracle():
    """render_cuda_orthographic (cuda_splatting.py:129-219): colour and gradients against the reference's orthographic
    glue restated over the CPU oracle (baseline/per_view_glue.py; tests/test_reference_glue.py pins that restatement to
    the unmodified reference file where /root/reference exists).  One view per call, as in the reference (its
    ``move_back[2, 3] = -distance_to_near`` only accepts one)."""
    from baseline import per_view_glue
    from my_depthsplat_b200 import cuda_splatting as cs
    from oracle import ext_compat
    scene = make_scene("small")  # B = 2
    g = scene.gaussians
    for b, view in ((1, 0), (1, 2)):
        ext = scene.extrinsics[:b, view]
        width, height = torch.tensor([3.0, 2.5][:b]), torch.tensor([2.0, 2.2][:b])
        near, far = torch.zeros(b), torch.full((b,), 20.0)
        bg = torch.tensor([[0.1, 0.2, 0.3], [0.0, 0.0, 0.0]])[:b]
        ref_g = Gaussians(*(t[:b].clone().requires_grad_() for t in (g.means, g.covariances, g.harmonics, g.opacities)))
        ref = per_view_glue.render_cuda_orthographic(ext_compat, ext, width, height, near, far, (32, 48), bg, ref_g.means,
                                                     ref_g.covariances, ref_g.harmonics, ref_g.opacities)
        w = torch.randn(ref.shape, generator=torch.Generator().manual_seed(6)) / ref[0].numel()
        (ref * w).sum().backward()
        mine_g = _cuda(Gaussians(g.means[:b], g.covariances[:b], g.harmonics[:b], g.opacities[:b]), grad=True)
        mine = cs.render_cuda_orthographic(ext.cuda(), width.cuda(), height.cuda(), near.cuda(), far.cuda(), (32, 48), bg.cuda(),
                                           mine_g.means, mine_g.covariances, mine_g.harmonics, mine_g.opacities)
        assert mine.shape == (b, 3, 32, 48) and float(ref.abs().max()) > 0
        (mine * w.cuda()).sum().backward()
        assert ((mine.detach().cpu() - ref.detach()).abs() > 1e-5).float().mean() <= 2e-3
        for got, want in ((mine_g.means, ref_g.means), (mine_g.covariances, ref_g.covariances), (mine_g.harmonics, ref_g.harmonics),
                          (mine_g.opacities, ref_g.opacities)):
            scale = float(want.grad.abs().max())
            e = (got.grad.cpu() - want.grad).abs()
            assert float(torch.quantile(e.flatten(), 0.999)) <= 1e-4 * scale and float(e.max()) <= 5e-2 * scale


def test_huge_footprints_split_views_and_stay_correct():
    """Gaussians that cover hundreds of tiles each (stress config, scaled down): the pair count per call is large,
    big-rect emission takes the warp-cooperative path and the first capacity guess overflows (retry protocol)."""
    from my_depthsplat_b200 import rasterizer as R
    scene = make_scene("small_stress")
    gc = _cuda(scene.gaussians)
    with torch.no_grad():
        ref, _ = oracle_decoder_forward(scene.gaussians, scene.extrinsics, scene.intrinsics, scene.near, scene.far, scene.image_shape, scene.background)
        color, _ = _render(scene, gc)
    assert ((color.cpu() - ref).abs() > 1e-5).float().mean() <= 1e-3
    assert R.last_stats.num_pairs > 10 * gc.means.shape[1]  # most Gaussians cover a large part of the 24-tile image


def test_workspace_budget_splits_the_views_and_changes_nothing():
    """rasterizer.max_workspace_bytes bounds the workspaces of one call: a call whose measured pair count needs more is split
    into smaller groups of views (and of scenes, when every scene has one view) -- same images, same gradients."""
    from my_depthsplat_b200 import rasterizer as R
    scene = make_scene("small_stress", batch=2, v_tgt=2)
    g0, g1 = _cuda(scene.gaussians), _cuda(scene.gaussians)
    for g in (g0, g1):
        for t in (g.means, g.covariances, g.harmonics, g.opacities):
            t.requires_grad_()
    ref, _ = _render(scene, g0)
    (ref * scene.grad_color.cuda()).sum().backward()
    pairs = R.last_stats.num_pairs
    old = R.max_workspace_bytes
    R._capacity_hint.clear()
    N = g0.means.shape[1]
    H, W = scene.image_shape
    # room for the records of all four views but for the pairs of about one
    R.max_workspace_bytes = 4 * (N * 72 + H * W * 8) + (pairs // 3) * 24
    remat0 = R.remat_count
    try:
        got, _ = _render(scene, g1)
        assert R.last_stats.num_pairs < pairs           # the last call rendered a subset of the views
        (got * scene.grad_color.cuda()).sum().backward()
        # the parts gave their forward->backward workspaces back after the forward and rebuilt them in the backward
        # (rasterizer.remat_depth): the bound on one call bounds the step
        assert R.remat_count - remat0 >= 2
    finally:
        R.max_workspace_bytes = old
        R._capacity_hint.clear()
    assert torch.equal(got, ref)
    for a, b in ((g1.means, g0.means), (g1.covariances, g0.covariances), (g1.harmonics, g0.harmonics), (g1.opacities, g0.opacities)):
        # the compositing backward accumulates with REDs (order varies from run to run) and the split sums the views in
        # another order: a few 1e-6 of the gradient scale apart, an order below the 1e-4 bar
        torch.testing.assert_close(a.grad, b.grad, rtol=1e-5, atol=1e-5 * float(b.grad.abs().max()))


def test_camera_block_graph_replays_the_eager_sequence_bit_for_bit():
    """The ~60 torch kernels of the camera block are captured once per (device, views, flags) in a CUDA graph
    (cuda_splatting._camera_tensors_cached): same outputs as the eager sequence for every new set of cameras, and a
    result handed out earlier is not overwritten by a later replay."""
    from my_depthsplat_b200 import cuda_splatting as cs
    outs = []
    for seed, name in ((0, "small"), (1, "small"), (2, "small")):
        sc = make_scene(name, v_tgt=3).to("cuda")
        ext = sc.extrinsics.reshape(-1, 4, 4).clone()
        ext[:, :3, 3] += 0.01 * seed
        K, near, far = sc.intrinsics.reshape(-1, 3, 3), sc.near.reshape(-1) * (1 + 0.1 * seed), sc.far.reshape(-1)
        for want_depth in (False, True):
            eager = cs._camera_tensors(ext, K, near, far, want_depth, True)
            graphed = cs._camera_tensors_cached(ext, K, near, far, want_depth, True)
            assert cs._camera_graphs[(ext.device, ext.shape[0], want_depth, True)] is not False, "capture failed"
            for a, b in zip(eager, graphed):
                assert (a is None) == (b is None)
                if a is not None:
                    assert torch.equal(a, b)
            outs.append((eager, graphed))
    for eager, graphed in outs:  # earlier results survived the later replays
        for a, b in zip(eager, graphed):
            if a is not None:
                assert torch.equal(a, b)
