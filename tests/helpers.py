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


def input_digest(scene) -> str:
    """sha256 over everything the oracle is fed for `scene` (Gaussians, upstream gradients, per-view extension
    arguments).  Scene generation and camera glue use host BLAS; a fixture records the digest of the inputs it was made
    from so that a test can tell "the oracle drifted" from "this host builds the inputs with different last bits"."""
    import hashlib
    h = hashlib.sha256()
    g = scene.gaussians
    for t in (g.means, g.covariances, g.harmonics, g.opacities, scene.grad_color, scene.grad_depth):
        h.update(t.detach().contiguous().numpy().tobytes())
    B, V = scene.extrinsics.shape[:2]
    for b in range(B):
        for v in range(V):
            for k, a in sorted(per_view_extension_inputs(scene, b, v).items()):
                h.update(np.ascontiguousarray(a).tobytes() if isinstance(a, np.ndarray) else repr(a).encode())
    return h.hexdigest()


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


# ---------------------------------------------------------------------------------------------------
# GPU-side helpers of the parity tests (imported lazily: the CPU suite never touches CUDA)
def render_cpu_cameras(scene, g, depth_mode=None, **kw):
    """The product's rasterizer fed a camera block (view / projection matrices, tanfov, scale) that was built on the CPU
    exactly as the oracle's inputs are, then moved to the GPU.  torch's CPU and CUDA ``inverse`` / ``matmul`` differ in the
    last bits; bit-exactness of keys is a statement about the kernels given IDENTICAL camera matrices (the boundary the
    extension sees, cuda_splatting.py:98-111), so every oracle comparison feeds identical ones."""
    from my_depthsplat_b200 import cuda_splatting as cs
    from my_depthsplat_b200.rasterizer import ViewPack, rasterize
    B, V = scene.extrinsics.shape[:2]
    h, w = scene.image_shape
    ext = scene.extrinsics.reshape(B * V, 4, 4).float().cpu()
    near, far = scene.near.reshape(-1).float().cpu(), scene.far.reshape(-1).float().cpu()
    daff, dclamp = cs._depth_block(ext, near, far) if depth_mode is not None else (None, None)
    scale = 1 / near
    ext = ext.clone()
    ext[..., :3, 3] = ext[..., :3, 3] * scale[:, None]
    fov_x, fov_y = cs.get_fov(scene.intrinsics.reshape(B * V, 3, 3).float().cpu()).unbind(-1)
    view, full, campos, tanfov = cs._camera_block(ext, near * scale, far * scale, fov_x, fov_y, (0.5 * fov_x).tan(), (0.5 * fov_y).tan())
    c = lambda t: None if t is None else t.contiguous().cuda()
    bg = kw.pop("background", scene.background).cpu()
    bg = bg.expand(B * V, 3) if bg.dim() == 1 else bg.reshape(B * V, 3)
    pack = ViewPack(torch.arange(B, dtype=torch.int32).repeat_interleave(V).cuda(), c(view), c(full), c(campos), c(tanfov),
                    c(bg), h, w, c(torch.stack([scale, scale ** 2], -1)), depth_mode, c(daff), c(dclamp))
    color, depth, radii = rasterize(g.means, g.covariances, g.harmonics, g.opacities, pack, want_radii=True, **kw)
    return color.reshape(B, V, 3, h, w), (None if depth is None else depth.reshape(B, V, h, w)), radii.reshape(B, V, -1)


def stage_dump():
    """Stage outputs of the last rasterizer call (rasterizer.debug_keep), as numpy.  ``keys`` (the sorted 64-bit
    (view | tile | depth) keys) exist only after a GLOBAL-mode sort; the BINNED mode never materialises them."""
    from my_depthsplat_b200 import _lib
    from my_depthsplat_b200 import rasterizer as R
    d = R.debug_last
    plan, saved, scratch = d["plan"], d["saved"], d["scratch"]
    torch.cuda.synchronize()
    VV, N, Rn = d["VV"], d["N"], d["num_pairs"]
    part = lambda buf, off, nbytes: buf[off: off + nbytes].cpu().numpy()
    rec = part(saved, plan.off_rec, VV * N * 64).view(np.float32).reshape(VV, N, 16)
    keys = None
    if getattr(plan, "sort_mode", _lib.SORT_GLOBAL) == _lib.SORT_GLOBAL:
        keys = part(scratch, plan.off_keys_a, Rn * 8).view(np.uint64)
    vals = part(saved, plan.off_vals_a, Rn * 4).view(np.uint32)
    ranges = part(saved, plan.off_ranges, plan.bins * 8).view(np.uint32).reshape(plan.bins, 2)
    HW = d["H"] * d["W"]
    final_T = part(saved, plan.off_final_T, VV * HW * 4).view(np.float32).reshape(VV, d["H"], d["W"])
    n_contrib = part(saved, plan.off_n_contrib, VV * HW * 4).view(np.uint32).reshape(VV, d["H"], d["W"])
    return dict(plan=plan, rec=rec, keys=keys, vals=vals, ranges=ranges, final_T=final_T, n_contrib=n_contrib)


def cuda_leaf_gaussians(scene):
    from my_depthsplat_b200.types import Gaussians
    g = scene.gaussians
    mk = lambda t: t.detach().clone().cuda().requires_grad_()
    return Gaussians(mk(g.means), mk(g.covariances), mk(g.harmonics), mk(g.opacities))


def check_view_stages(d, vi, st, radii_view):
    """Bit-exact comparison of one view's stage outputs (dump ``d`` of the multi-view call) with the oracle's ViewState:
    radii, depth bits, pixel xy, conic, the view's slice of the sorted list (keys when the sort materialised them, the
    Gaussian indices always) and the tile ranges.  Returns the first list position of the view."""
    plan = d["plan"]
    rec = d["rec"][vi]
    vis = st.radii > 0
    np.testing.assert_array_equal(radii_view, st.radii)
    np.testing.assert_array_equal(rec[vis, 13].view(np.int32), st.radii[vis])
    np.testing.assert_array_equal(rec[vis, 12].view(np.uint32), st.depths[vis].view(np.uint32))
    np.testing.assert_array_equal(rec[vis, 0:2].view(np.uint32), st.xy[vis].view(np.uint32))
    np.testing.assert_array_equal(rec[vis, 4:7].view(np.uint32), st.conic_opacity[vis, 0:3].view(np.uint32))
    rg = d["ranges"][(vi << plan.tile_bits): (vi << plan.tile_bits) + plan.tiles].astype(np.int64)
    nonempty = rg[:, 1] > rg[:, 0]
    np.testing.assert_array_equal(nonempty, st.ranges[:, 1] > st.ranges[:, 0])
    first = int(rg[nonempty, 0].min()) if nonempty.any() else 0
    last = int(rg[nonempty, 1].max()) if nonempty.any() else 0
    assert last - first == st.num_rendered
    np.testing.assert_array_equal(rg[nonempty] - first, st.ranges[nonempty].astype(np.int64))
    np.testing.assert_array_equal(d["vals"][first:last], st.vals)
    if d["keys"] is not None:
        k = d["keys"][first:last]
        tile_mask = np.uint64((1 << plan.tile_bits) - 1)
        assert bool(((k >> np.uint64(32 + plan.tile_bits)) == np.uint64(vi)).all())
        k_view = ((k >> np.uint64(32)) & tile_mask) << np.uint64(32) | (k & np.uint64(0xFFFFFFFF))
        np.testing.assert_array_equal(k_view, st.keys)
    return first


def gaussians_touching(st, pixels):
    """Indices of the Gaussians whose footprint (centre +- radius) covers one of ``pixels`` [(y, x), ...] of the view the
    oracle state ``st`` belongs to: the ones a flipped threshold decision at such a pixel can reach."""
    gx = st.grid[0]
    hit = []
    for (y, x) in pixels:
        t = (y // 16) * gx + (x // 16)
        ids = st.vals[st.ranges[t, 0]: st.ranges[t, 1]]
        if ids.size == 0:
            continue
        xy, r = st.xy[ids], st.radii[ids].astype(np.float32) + 1.0
        m = (np.abs(xy[:, 0] - x) <= r) & (np.abs(xy[:, 1] - y) <= r)
        hit.append(ids[m])
    return np.unique(np.concatenate(hit)) if hit else np.zeros(0, np.int64)


def strict_parity_check(scene, dm, name="scene"):
    """The product (one multi-view call, identical camera block) against the oracle on a CPU ``scene``:
      (i)   stages bit-exact per view (check_view_stages);
      (ii)  colour (+ depth, relative to max(1, |d|)) within 1e-5 on every pixel the oracle does not mark fragile; every
            pixel that differs more, or whose n_contrib differs, must be a fragile one;
      (iii) gradients within 1e-4 of each tensor's scale on every Gaussian whose footprint covers no flipped pixel.
    Returns the lines of a report (counts of fragile / flipped pixels, excluded Gaussians, worst errors)."""
    from my_depthsplat_b200 import rasterizer as R
    from oracle import splat_oracle as so
    B, V = scene.extrinsics.shape[:2]
    h, w = scene.image_shape

    # ---- the product, one call for all views --------------------------------------------------------------------
    g = cuda_leaf_gaussians(scene)
    R.debug_keep = True
    try:
        color, depth, radii = render_cpu_cameras(scene, g, depth_mode=dm)
        d = stage_dump()
    finally:
        R.debug_keep = False
        R.debug_last = None
    loss = (color * scene.grad_color.cuda()).sum()
    if dm is not None:
        loss = loss + (depth * scene.grad_depth.cuda()).sum()
    loss.backward()
    got = {k: getattr(g, k).grad.cpu().numpy() for k in ("means", "covariances", "harmonics", "opacities")}
    color_np = color.detach().cpu().numpy()
    depth_np = None if depth is None else depth.detach().cpu().numpy()
    radii_np = radii.cpu().numpy()
    del color, depth, radii, loss
    torch.cuda.empty_cache()

    # ---- the oracle: colour + depth + gradients through the reference's glue restated (per-view loop) -----------
    gc = leaf_gaussians(scene)
    ref_c, ref_d = oracle_decoder_forward(gc, scene.extrinsics, scene.intrinsics, scene.near, scene.far, scene.image_shape,
                                          scene.background, dm)
    rl = (ref_c * scene.grad_color).sum()
    if dm is not None:
        rl = rl + (ref_d * scene.grad_depth).sum()
    rl.backward()
    ref_c = ref_c.detach().numpy()
    ref_d = None if ref_d is None else ref_d.detach().numpy()

    # ---- (i) stages bit-exact, (ii) images, per view ------------------------------------------------------------
    N = scene.gaussians.means.shape[1]
    excluded = [np.zeros(N, bool) for _ in range(B)]
    n_fragile = n_flipped = n_contrib_diff = 0
    worst_solid = 0.0
    for b in range(B):
        for v in range(V):
            vi = b * V + v
            st = so.forward_view(**per_view_extension_inputs(scene, b, v))
            check_view_stages(d, vi, st, radii_np[b, v])
            np.testing.assert_array_equal(st.color, ref_c[b, v])  # the two oracle entry points are the same computation
            err = np.abs(color_np[b, v] - ref_c[b, v]).max(axis=0)
            if dm is not None:
                err = np.maximum(err, np.abs(depth_np[b, v] - ref_d[b, v]) / np.maximum(1.0, np.abs(ref_d[b, v])))
            ndiff = d["n_contrib"][vi] != st.n_contrib
            flipped = (err > 1e-5) | ndiff
            fragile = st.fragile != 0
            stray = flipped & ~fragile
            assert not stray.any(), (name, b, v, int(stray.sum()), float(err[stray].max()), np.argwhere(stray)[:5].tolist())
            worst_solid = max(worst_solid, float(err[~fragile].max()))
            n_fragile += int(fragile.sum()); n_flipped += int(flipped.sum()); n_contrib_diff += int(ndiff.sum())
            if flipped.any():
                excluded[b][gaussians_touching(st, [tuple(p) for p in np.argwhere(flipped)])] = True
            st.close()

    # ---- (iii) gradients --------------------------------------------------------------------------------------------
    report = [f"[{name}] views {B}x{V}  pixels {B * V * h * w}  fragile {n_fragile}  flipped {n_flipped} (n_contrib differs on "
              f"{n_contrib_diff})  worst error off the fragile pixels {worst_solid:.2e}  Gaussians excluded {int(sum(e.sum() for e in excluded))} of {B * N}"]
    keep = ~np.stack(excluded)
    for k in ("means", "covariances", "harmonics", "opacities"):
        r = getattr(gc, k).grad.numpy()
        e = np.abs(got[k] - r).reshape(B, N, -1).max(axis=-1)
        scale = np.abs(r).max()
        assert scale > 0, k
        report.append(f"   d{k}: max err / scale = {e[keep].max() / scale:.2e} (kept), {e.max() / scale:.2e} (all)")
        assert e[keep].max() <= 1e-4 * scale, (name, k, e[keep].max() / scale)
    return report


