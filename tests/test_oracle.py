"""CPU tests of the oracle itself (no GPU): the C restatement against an independent float64 dense
autograd formulation, and the internal consistency of its stage outputs.  The third-party
extension the reference binds is absent, so this is the strongest pin available for the per-view
arithmetic (PARITY UNPINNED against the real extension; see oracle/splat_oracle.c)."""
import numpy as np
import pytest
import torch

from helpers import per_view_extension_inputs
from my_depthsplat_b200.scenes import make_scene
from oracle import dense_torch as dt
from oracle import splat_oracle as so


def _dense_inputs(inp):
    T = lambda a: torch.tensor(a, dtype=torch.float64)
    kw = dict(H=inp["H"], W=inp["W"], bg=T(inp["bg"]), viewmatrix=T(inp["viewmatrix"]), projmatrix=T(inp["projmatrix"]),
              campos=T(inp["campos"]), tanfovx=inp["tanfovx"], tanfovy=inp["tanfovy"], sh_degree=inp["sh_degree"])
    leaves = dict(means3D=T(inp["means3D"]).requires_grad_(), cov3D=T(inp["cov3D"]).requires_grad_(),
                  shs=T(inp["shs"]).requires_grad_(), opacities=T(inp["opacities"]).requires_grad_())
    return kw, leaves


@pytest.mark.parametrize("name,view", [("tiny", (0, 0)), ("tiny", (0, 1)), ("small_stress", (0, 1))])
def test_oracle_matches_dense_float64(name, view):
    sc = make_scene(name)
    inp = per_view_extension_inputs(sc, *view)
    if name == "small_stress":  # keep the dense O(P*G) evaluation small: every 6th Gaussian
        keep = slice(0, None, 6)
        for k in ("means3D", "opacities", "cov3D", "shs"):
            inp[k] = np.ascontiguousarray(inp[k][keep])
    st = so.forward_view(**inp)
    assert st.num_rendered > 0 and (st.radii > 0).sum() > 0
    kw, lv = _dense_inputs(inp)
    img = dt.render_view(radii=torch.tensor(st.radii), **kw, **lv)
    d = np.abs(img.detach().numpy() - st.color)
    # threshold decisions (alpha < 1/255, T < 1e-4) are taken in fp32 by the oracle and in fp64 by the dense
    # model: allow a handful of flipped pixels, everything else agrees to fp32 round-off
    assert (d > 2e-5).mean() < 2e-3, d.max()
    assert np.median(d) < 1e-6

    g = torch.randn(img.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
    (img * g).sum().backward()
    gr = so.backward_view(st, g.numpy().astype(np.float32))
    # Where the 1.3*tanfov clamp of t.x/t.z is active, the published backward deliberately differs from
    # the true derivative: it zeroes dL/dt.x but ignores that the clamped t.x now depends on t.z.  Those
    # Gaussians are excluded from the autograd comparison of dL/dmean (the oracle follows the algorithm).
    hom = np.concatenate([inp["means3D"], np.ones((len(inp["means3D"]), 1), np.float32)], 1) @ inp["viewmatrix"].reshape(4, 4)
    clamped = (np.abs(hom[:, 0] / hom[:, 2]) > 1.3 * inp["tanfovx"]) | (np.abs(hom[:, 1] / hom[:, 2]) > 1.3 * inp["tanfovy"])
    for key, leaf in (("means3D", "means3D"), ("cov3D", "cov3D"), ("sh", "shs"), ("opacity", "opacities")):
        ref = lv[leaf].grad.numpy()
        err = np.abs(gr[key] - ref)
        if key == "means3D":
            err = err[~clamped]
        scale = np.abs(ref).max()
        # hand-derived backward == autograd of the rendering equation, up to fp32 accumulation
        assert err.max() <= 2e-4 * scale, (key, err.max(), scale)


@pytest.mark.parametrize("degree", [0, 1, 3])
def test_oracle_sh_degrees_match_dense_float64(degree):
    """The SH evaluation and its hand-derived backward (view-direction chain included) for the degrees the default
    configs do not use: 1, 4 and 16 coefficients per channel."""
    from my_depthsplat_b200.scenes import SceneConfig
    sc = make_scene(SceneConfig(f"deg{degree}", 300 + degree, 2, 24, 32, sh_degree=degree))
    inp = per_view_extension_inputs(sc, 0, 1)
    st = so.forward_view(**inp)
    assert (st.radii > 0).sum() > 0
    kw, lv = _dense_inputs(inp)
    img = dt.render_view(radii=torch.tensor(st.radii), **kw, **lv)
    d = np.abs(img.detach().numpy() - st.color)
    assert (d > 2e-5).mean() < 2e-3 and np.median(d) < 1e-6, d.max()
    g = torch.randn(img.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(5))
    (img * g).sum().backward()
    gr = so.backward_view(st, g.numpy().astype(np.float32))
    hom = np.concatenate([inp["means3D"], np.ones((len(inp["means3D"]), 1), np.float32)], 1) @ inp["viewmatrix"].reshape(4, 4)
    clamped = (np.abs(hom[:, 0] / hom[:, 2]) > 1.3 * inp["tanfovx"]) | (np.abs(hom[:, 1] / hom[:, 2]) > 1.3 * inp["tanfovy"])
    for key, leaf in (("means3D", "means3D"), ("cov3D", "cov3D"), ("sh", "shs"), ("opacity", "opacities")):
        ref = lv[leaf].grad.numpy()
        err = np.abs(gr[key] - ref)
        if key == "means3D":
            err = err[~clamped]
        assert err.max() <= 2e-4 * np.abs(ref).max(), (key, err.max(), np.abs(ref).max())


def test_oracle_precomputed_colors_and_background():
    sc = make_scene("tiny")
    inp = per_view_extension_inputs(sc, 0, 0)
    P = inp["means3D"].shape[0]
    rng = np.random.default_rng(5)
    cols = rng.random((P, 3), dtype=np.float32) * 3 - 1  # unclamped, may be negative
    inp2 = dict(inp); inp2.pop("shs"); inp2["colors_precomp"] = cols; inp2["bg"] = np.array([0.2, 0.5, 0.9], np.float32)
    st = so.forward_view(**inp2)
    kw, lv = _dense_inputs(inp)
    kw["bg"] = torch.tensor(inp2["bg"], dtype=torch.float64)
    lv.pop("shs")
    img = dt.render_view(radii=torch.tensor(st.radii), colors_precomp=torch.tensor(cols, dtype=torch.float64), **kw, **lv)
    d = np.abs(img.detach().numpy() - st.color)
    assert (d > 2e-5).mean() < 2e-3
    # linearity in the colours (the compositing weights do not depend on them)
    inp3 = dict(inp2); inp3["colors_precomp"] = 2 * cols; inp3["bg"] = 2 * inp2["bg"]
    st3 = so.forward_view(**inp3)
    np.testing.assert_allclose(st3.color, 2 * st.color, rtol=2e-6, atol=1e-6)


def test_oracle_stage_invariants():
    sc = make_scene("small")
    inp = per_view_extension_inputs(sc, 1, 2)
    st = so.forward_view(**inp)
    gx, gy = st.grid
    vis = st.radii > 0
    # scan / duplicate
    assert st.offsets[-1] == st.num_rendered == st.tiles_touched.sum()
    assert np.all(st.tiles_touched[~vis] == 0) and np.all(st.tiles_touched[vis] > 0)
    assert np.all(st.depths[vis] > 0.2)
    # unsorted keys: values ascend, depth bits belong to the Gaussian
    assert np.all(np.diff(st.vals_unsorted.astype(np.int64)) >= 0)
    np.testing.assert_array_equal((st.keys_unsorted & 0xFFFFFFFF).astype(np.uint32), st.depths[st.vals_unsorted].view(np.uint32))
    # sorted: non-decreasing keys, a permutation of the unsorted pairs, stable for equal keys
    assert np.all(np.diff(st.keys.astype(np.int64)) >= 0) or np.all(st.keys[1:] >= st.keys[:-1])
    order = np.argsort(st.keys_unsorted, kind="stable")
    np.testing.assert_array_equal(st.keys, st.keys_unsorted[order])
    np.testing.assert_array_equal(st.vals, st.vals_unsorted[order])
    # ranges partition the sorted list by tile id
    tiles = (st.keys >> 32).astype(np.int64)
    for t in range(gx * gy):
        a, b = st.ranges[t]
        if a == b:
            assert not np.any(tiles == t)
        else:
            assert np.all(tiles[a:b] == t) and (a == 0 or tiles[a - 1] != t) and (b == len(tiles) or tiles[b] != t)
    # image state
    assert np.all(st.final_T <= 1.0) and np.all(st.final_T >= 0.0)
    lens = (st.ranges[:, 1] - st.ranges[:, 0]).reshape(gy, gx)
    H, W = st.H, st.W
    per_pix_len = np.repeat(np.repeat(lens, 16, 0), 16, 1)[:H, :W]
    assert np.all(st.n_contrib <= per_pix_len)


def test_oracle_empty_and_culled():
    sc = make_scene("tiny")
    inp = per_view_extension_inputs(sc, 0, 0)
    # everything behind the camera -> background only, nothing rendered
    inp_b = dict(inp); inp_b["bg"] = np.array([0.1, 0.2, 0.3], np.float32)
    inp_b["means3D"] = inp["means3D"] * np.array([1, 1, -1], np.float32)
    st = so.forward_view(**inp_b)
    assert st.num_rendered == 0 and np.all(st.radii == 0)
    np.testing.assert_allclose(st.color, np.broadcast_to(inp_b["bg"][:, None, None], st.color.shape))
    gr = so.backward_view(st, np.ones_like(st.color))
    assert all(np.all(v == 0) for v in gr.values() if v is not None)
    # zero Gaussians
    inp_0 = dict(inp)
    for k, shp in (("means3D", (0, 3)), ("opacities", (0,)), ("cov3D", (0, 6)), ("shs", (0, 9, 3))):
        inp_0[k] = np.zeros(shp, np.float32)
    st0 = so.forward_view(**inp_0)
    assert st0.num_rendered == 0 and st0.color.shape == (3, inp["H"], inp["W"])
