"""GPU tests at BASELINE.json's FULL sizes (C1: 131 072 Gaussians 256x256 x4 views; C2T: 2 949 120 Gaussians
512x960 x4 views), where the CPU oracle is too slow to be the checker for every view.  Size-independent
properties of the path instead:
  * binning: keys non-decreasing, ranges partition the list by (view, tile), pair count = sum of tile
    rect areas, the multiset of values is preserved by the sort (checksums);
  * compositing: exact linearity in the colours and the background, final_T in [0, 1], n_contrib within the
    tile's list, depth = compositing of z with the same weights (a constant depth colour gives 1 - T);
  * backward: the gradient w.r.t. the colours is the exact adjoint of that linear map
    (<dL/dc, delta> == L(c + delta) - L(c)), gradients of culled Gaussians are exactly zero
    (finite differences are NOT used: the alpha >= 1/255 and T >= 1e-4 cuts make the rendered image
    discontinuous in opacity and position, so differences carry jump terms the analytic gradient --
    the reference's too -- ignores);
(The oracle comparison at these sizes -- every view, strict bars -- is tests/test_gpu_fullsize_parity.py.)
"""
import numpy as np
import pytest
import torch

from my_depthsplat_b200.scenes import make_scene

pytestmark = pytest.mark.gpu

CONFIGS = ["C1", "C2T"]
_cache = {}


def _scene(name):
    if name not in _cache:
        _cache.clear()
        cpu = make_scene(name)
        _cache[name] = (cpu, cpu.to("cuda"))
    return _cache[name]


def _render(sc, colors=None, use_sh=True, depth_mode=None, bg=None, **kw):
    from my_depthsplat_b200.cuda_splatting import render_views
    g = sc.gaussians
    return render_views(sc.extrinsics, sc.intrinsics, sc.near, sc.far, sc.image_shape, sc.background if bg is None else bg,
                        g.means, g.covariances, g.harmonics if colors is None else colors, g.opacities, use_sh=use_sh,
                        depth_mode=depth_mode, **kw)


@pytest.mark.parametrize("mode", ["binned", "global"])
@pytest.mark.parametrize("name", CONFIGS)
def test_binning_invariants(name, mode):
    from my_depthsplat_b200 import rasterizer as R
    _, sc = _scene(name)
    R.debug_keep = True
    old_mode, R.sort_mode = R.sort_mode, mode
    try:
        with torch.no_grad():
            color, depth, radii = _render(sc, depth_mode="depth", want_radii=True)
        d = R.debug_last
        torch.cuda.synchronize()
        plan, n_pairs, N, VV = d["plan"], d["num_pairs"], d["N"], d["VV"]
        vals = d["saved"][plan.off_vals_a: plan.off_vals_a + n_pairs * 4].view(torch.int32)
        ranges = d["saved"][plan.off_ranges: plan.off_ranges + plan.bins * 8].view(torch.int32).reshape(plan.bins, 2).long()
        rec = d["saved"][plan.off_rec: plan.off_rec + VV * N * 64].view(torch.int32).reshape(VV, N, 16)
        rect = rec[..., 14]
        area = (((rect >> 16) & 255) - (rect & 255)) * (((rect >> 24) & 255) - ((rect >> 8) & 255))
        vis = rec[..., 13] > 0
        assert int(area[vis].sum()) == n_pairs and int(area[~vis].abs().sum()) == 0
        assert torch.equal(radii.reshape(VV, N), rec[..., 13])
        # the sort preserved the multiset of values: sum and sum of squares of the Gaussian indices
        idx = torch.arange(N, device="cuda", dtype=torch.int64)[None].expand(VV, N)
        assert int((idx * area.long()).sum()) == int(vals.long().sum())
        assert int((idx * idx % 1000003 * area.long()).sum()) == int((vals.long() * vals.long() % 1000003).sum())
        # ranges partition the list by bin = view << tile_bits | tile, in bin order
        lens = ranges[:, 1] - ranges[:, 0]
        assert int(lens.sum()) == n_pairs and bool((lens >= 0).all())
        nz = lens > 0
        assert torch.equal(ranges[nz, 0][1:], ranges[nz, 1][:-1]) and int(ranges[nz, 0][0]) == 0 and int(ranges[nz, 1][-1]) == n_pairs
        bins = torch.repeat_interleave(torch.arange(plan.bins, device="cuda"), lens)       # bin of every list position
        view_of = bins >> plan.tile_bits
        tile_of = bins & ((1 << plan.tile_bits) - 1)
        r = rec[view_of, vals.long()]                                                        # the entry's projected record
        rc = r[:, 14]
        tx, ty = tile_of % plan.grid_x, tile_of // plan.grid_x
        assert bool(((rc & 255) <= tx).all()) and bool((tx < ((rc >> 16) & 255)).all())      # the tile lies in the Gaussian's rect
        assert bool((((rc >> 8) & 255) <= ty).all()) and bool((ty < ((rc >> 24) & 255)).all())
        # inside a bin: ascending (depth bits, Gaussian index) -- the order of the stable sort of pairs emitted in index order
        key = (r[:, 12].long() << 32) | vals.long()
        same_bin = bins[1:] == bins[:-1]
        assert bool((key[1:][same_bin] > key[:-1][same_bin]).all())
        if plan.sort_mode == 1:  # GLOBAL mode materialises the sorted 64-bit keys
            keys = d["scratch"][plan.off_keys_a: plan.off_keys_a + n_pairs * 8].view(torch.int64)
            assert bool((keys[1:] >= keys[:-1]).all()) and torch.equal(keys >> 32, bins) and torch.equal(keys & 0xFFFFFFFF, r[:, 12].long())
        # image state
        HW = d["H"] * d["W"]
        final_T = d["saved"][plan.off_final_T: plan.off_final_T + VV * HW * 4].view(torch.float32).reshape(VV, d["H"], d["W"])
        n_contrib = d["saved"][plan.off_n_contrib: plan.off_n_contrib + VV * HW * 4].view(torch.int32).reshape(VV, d["H"], d["W"])
        assert bool((final_T >= 0).all()) and bool((final_T <= 1).all())
        tile_of_pixel = (torch.arange(d["H"], device="cuda")[:, None] // 16) * plan.grid_x + torch.arange(d["W"], device="cuda")[None] // 16
        per_pixel_len = lens.reshape(VV, -1)[:, : 1 << plan.tile_bits][torch.arange(VV, device="cuda")[:, None, None], tile_of_pixel[None]]
        assert bool((n_contrib <= per_pixel_len).all())
    finally:
        R.debug_keep = False
        R.debug_last = None
        R.sort_mode = old_mode


@pytest.mark.parametrize("name", CONFIGS)
def test_compositing_is_linear_in_colours_and_background(name):
    _, sc = _scene(name)
    g = sc.gaussians
    B, N = g.opacities.shape
    gen = torch.Generator(device="cuda").manual_seed(1)
    c1 = torch.rand(B, N, 3, 1, device="cuda", generator=gen)
    c2 = torch.rand(B, N, 3, 1, device="cuda", generator=gen)
    bg1, bg2 = torch.tensor([0.2, 0.4, 0.6], device="cuda"), torch.tensor([0.5, 0.1, 0.3], device="cuda")
    with torch.no_grad():
        i1, _ = _render(sc, colors=c1, use_sh=False, bg=bg1)
        i2, _ = _render(sc, colors=c2, use_sh=False, bg=bg2)
        i12, _ = _render(sc, colors=2 * c1 + 0.5 * c2, use_sh=False, bg=2 * bg1 + 0.5 * bg2)
        ones, dep = _render(sc, colors=torch.ones_like(c1), use_sh=False, bg=torch.zeros(3, device="cuda"), depth_mode="depth")
    torch.testing.assert_close(i12, 2 * i1 + 0.5 * i2, atol=2e-5, rtol=1e-5)
    # a constant colour of one over a black background renders the accumulated opacity 1 - T: in [0, 1], same in all channels
    assert bool((ones >= -1e-6).all()) and bool((ones <= 1 + 1e-5).all())
    assert torch.equal(ones[:, :, 0], ones[:, :, 1]) and torch.equal(ones[:, :, 0], ones[:, :, 2])
    # depth / accumulated opacity is a convex combination of camera-space depths: within [near cull, max depth]
    covered = ones[:, :, 0] > 0.5
    zbar = dep[covered] / ones[:, :, 0][covered]
    assert float(zbar.min()) > 0.1 and float(zbar.max()) < 1e3


@pytest.mark.parametrize("name", CONFIGS)
def test_colour_gradient_is_the_exact_adjoint(name):
    _, sc = _scene(name)
    g = sc.gaussians
    B, N = g.opacities.shape
    gen = torch.Generator(device="cuda").manual_seed(2)
    colors = torch.rand(B, N, 3, 1, device="cuda", generator=gen).requires_grad_()
    opac = g.opacities.detach().clone().requires_grad_()
    means = g.means.detach().clone().requires_grad_()
    from my_depthsplat_b200.cuda_splatting import render_views

    def loss_of(m, o, c):
        img, _ = render_views(sc.extrinsics, sc.intrinsics, sc.near, sc.far, sc.image_shape, sc.background, m, g.covariances, c, o, use_sh=False)
        return (img.double() * sc.grad_color.double()).sum() * 1e3

    L0 = loss_of(means, opac, colors)
    dm, do, dc = torch.autograd.grad(L0, (means, opac, colors))
    # exact adjoint in the colours
    delta = torch.randn(colors.shape, device="cuda", generator=gen)
    with torch.no_grad():
        L1 = loss_of(means, opac, colors + delta)
    lhs, rhs = float((dc.double() * delta.double()).sum()), float(L1 - L0)
    assert abs(lhs - rhs) <= 2e-3 * max(abs(rhs), abs(lhs), 1e-6) + 1e-6, (lhs, rhs)
    # Gaussians culled in every view get exactly zero gradient
    with torch.no_grad():
        _, _, radii = _render(sc, want_radii=True)
    dead = (radii.reshape(B, -1, N) <= 0).all(dim=1)
    if bool(dead.any()):
        assert float(dm[dead].abs().max()) == 0.0 and float(do[dead].abs().max()) == 0.0
