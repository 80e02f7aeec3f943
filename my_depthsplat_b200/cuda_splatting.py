"""Drop-in for src/model/decoder/cuda_splatting.py of the reference.

Same public names, argument meaning and results:
    get_projection_matrix      (reference :16-43)
    render_cuda                (reference :46-126)
    render_cuda_orthographic   (reference :129-219)
    render_depth_cuda          (reference :225-264)
    DepthRenderingMode         (reference :222)

What differs is below the signatures: no per-view Python loop, no ``.item()`` syncs, no SH / covariance
re-layout copies, no second rasterization for depth -- one call into the sm_100a library for all
views (my_depthsplat_b200.rasterizer).  The scale-invariant normalisation of the Gaussians
(reference :63-70) is folded into the projection kernel; the camera matrices are built with the same
torch operations as the reference so that they are bit-identical.

``render_views`` is the multi-view entry the decoder uses: Gaussians stay ``[B,N,...]``, cameras are
``[B,V,...]``; nothing is replicated per view.
"""
from __future__ import annotations

from math import isqrt
from typing import Literal, Optional

import torch
from torch import Tensor

from . import _lib
from .projection import get_fov, homogenize_points, inverse_nosync
from .rasterizer import PairLimitExceeded, ViewPack, rasterize

DepthRenderingMode = Literal["depth", "disparity", "relative_disparity", "log"]


def get_projection_matrix(near: Tensor, far: Tensor, fov_x: Tensor, fov_y: Tensor) -> Tensor:
    """[b] each -> [b,4,4].  Maps the viewing frustum to (-1, 1) in X/Y and (0, 1) in Z, with w = z
    (Z is not flipped to (-1, 1) as OpenGL would)."""
    tan_x = (0.5 * fov_x).tan()
    tan_y = (0.5 * fov_y).tan()
    top, right = tan_y * near, tan_x * near
    bottom, left = -top, -right
    proj = torch.zeros((near.shape[0], 4, 4), dtype=torch.float32, device=near.device)
    proj[:, 0, 0] = 2 * near / (right - left)
    proj[:, 1, 1] = 2 * near / (top - bottom)
    proj[:, 0, 2] = (right + left) / (right - left)
    proj[:, 1, 2] = (top + bottom) / (top - bottom)
    proj[:, 3, 2] = 1
    proj[:, 2, 2] = far / (far - near)
    proj[:, 2, 3] = -(far * near) / (far - near)
    return proj


def _camera_block(extrinsics: Tensor, near: Tensor, far: Tensor, fov_x: Tensor, fov_y: Tensor, tan_fov_x: Tensor,
                  tan_fov_y: Tensor):
    """View / full-projection matrices in the transposed storage the rasterizer consumes, camera
    positions and tanfov -- the reference's lines :83-86 and :101-110, batched."""
    projection = get_projection_matrix(near, far, fov_x, fov_y).transpose(1, 2)
    view = inverse_nosync(extrinsics).transpose(1, 2)
    full = view @ projection
    campos = extrinsics[:, :3, 3]
    tanfov = torch.stack([tan_fov_x.expand(near.shape[0]), tan_fov_y.expand(near.shape[0])], dim=-1)
    return view.contiguous(), full.contiguous(), campos.contiguous(), tanfov.to(torch.float32).contiguous()


def _depth_block(extrinsics_raw: Tensor, near_raw: Tensor, far_raw: Tensor):
    """Row 2 of the UNnormalised world->camera matrix: z_cam = row . (mean, 1) (reference :238-241)."""
    w2c = inverse_nosync(extrinsics_raw)
    return w2c[:, 2, :].contiguous(), torch.stack([near_raw, far_raw], dim=-1).contiguous()


def _camera_tensors(ext: Tensor, K: Tensor, near_f: Tensor, far_f: Tensor, want_depth: bool, scale_invariant: bool):
    """Everything the kernels need per view, as a pure function of the cameras ([VV,4,4], [VV,3,3], [VV], [VV]): the
    reference's lines :63-70 (scale-invariant normalisation of the camera side), :79-86 (field of view, projection and
    view matrices) and :238-241 (row of the un-normalised world->camera matrix for the depth colour), same torch
    operations in the same order.  -> (view, full, campos, tanfov, scale_pack | None, depth_affine | None,
    depth_clamp | None)."""
    depth_affine = depth_clamp = None
    if want_depth:
        depth_affine, depth_clamp = _depth_block(ext, near_f, far_f)
    scale_pack = None
    if scale_invariant:
        scale = 1 / near_f
        ext = ext.clone()
        ext[..., :3, 3] = ext[..., :3, 3] * scale[:, None]
        scale_pack = torch.stack([scale, scale ** 2], dim=-1).contiguous()
        near_f = near_f * scale
        far_f = far_f * scale
    fov_x, fov_y = get_fov(K).unbind(dim=-1)
    tan_fov_x, tan_fov_y = (0.5 * fov_x).tan(), (0.5 * fov_y).tan()
    view, full, campos, tanfov = _camera_block(ext, near_f, far_f, fov_x, fov_y, tan_fov_x, tan_fov_y)
    return view, full, campos, tanfov, scale_pack, depth_affine, depth_clamp


# The camera block is ~60 tiny torch kernels: at 256x256 their launches cost as much host time as the whole
# rasterizer.  They are a fixed sequence for a given number of views, so they are captured ONCE per
# (device, views, flags) in a CUDA graph and replayed: same kernels, same arithmetic, one launch.
use_camera_graph = True
_camera_graphs: dict = {}
_scene_index_cache: dict = {}


class _CameraGraph:
    def __init__(self, ext, K, near_f, far_f, want_depth, scale_invariant):
        dev = ext.device
        self.static_in = [t.detach().clone() for t in (ext, K, near_f, far_f)]
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):  # warm-up outside the capture: cuBLAS / cuSOLVER handles and workspaces
            for _ in range(2):
                _camera_tensors(*self.static_in, want_depth, scale_invariant)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            outs = _camera_tensors(*self.static_in, want_depth, scale_invariant)
            self.shapes = [None if o is None else tuple(o.shape) for o in outs]
            self.flat = torch.cat([o.reshape(-1) for o in outs if o is not None])

    def __call__(self, ext, K, near_f, far_f):
        for dst, src in zip(self.static_in, (ext, K, near_f, far_f)):
            dst.copy_(src)
        self.graph.replay()
        flat = self.flat.clone()  # the static result is overwritten by the next replay; the views are saved for backward
        outs, o = [], 0
        for shp in self.shapes:
            if shp is None:
                outs.append(None)
                continue
            n = 1
            for d in shp:
                n *= d
            outs.append(flat[o:o + n].view(shp))
            o += n
        return tuple(outs)


def _camera_tensors_cached(ext, K, near_f, far_f, want_depth, scale_invariant):
    eligible = (use_camera_graph and ext.is_cuda and ext.device.index == torch.cuda.current_device()
                and not torch.cuda.is_current_stream_capturing()
                and not (ext.requires_grad or K.requires_grad or near_f.requires_grad or far_f.requires_grad))
    if not eligible:
        return _camera_tensors(ext, K, near_f, far_f, want_depth, scale_invariant)
    key = (ext.device, ext.shape[0], want_depth, scale_invariant)
    g = _camera_graphs.get(key)
    if g is None:
        try:
            g = _CameraGraph(ext, K, near_f, far_f, want_depth, scale_invariant)
        except Exception as e:  # capture is an optimisation: fall back to the eager sequence, once and loudly
            import warnings
            warnings.warn(f"camera-block CUDA graph capture failed ({type(e).__name__}: {e}); running the torch ops eagerly")
            g = False
        _camera_graphs[key] = g
    if g is False:
        return _camera_tensors(ext, K, near_f, far_f, want_depth, scale_invariant)
    return g(ext, K, near_f, far_f)


def _scene_index(dev, B: int, V: int) -> Tensor:
    key = (dev, B, V)
    t = _scene_index_cache.get(key)
    if t is None:
        t = torch.arange(B, device=dev, dtype=torch.int32).repeat_interleave(V)
        _scene_index_cache[key] = t
    return t


def render_views(
    extrinsics: Tensor,            # [B,V,4,4] camera-to-world, OpenCV
    intrinsics: Tensor,            # [B,V,3,3] normalised
    near: Tensor,                  # [B,V]
    far: Tensor,                   # [B,V]
    image_shape: tuple[int, int],
    background_color: Tensor,      # [3] or [B,V,3]
    gaussian_means: Tensor,        # [B,N,3]
    gaussian_covariances: Tensor,  # [B,N,3,3]
    gaussian_sh_coefficients: Tensor,  # [B,N,3,d_sh]
    gaussian_opacities: Tensor,    # [B,N]
    scale_invariant: bool = True,
    use_sh: bool = True,
    depth_mode: Optional[DepthRenderingMode] = None,
    want_radii: bool = False,
    count_work: bool = False,
    grad_reducer=None,
    mse: Optional[dict] = None,
):
    """All V views of all B scenes in one rasterizer call.  Returns (color [B,V,3,H,W],
    depth [B,V,H,W] | None) (+ radii [B,V,N] when want_radii).

    ``mse`` (loss-side fusion) = dict(target=[B,V,3,H,W], weight=float, l1=bool[, count=int]): the
    compositing epilogue also computes weight * mean((color - target)^2) (src/loss/loss_mse.py:33-44), its dL/dcolor and the
    clipped squared error of compute_psnr (src/evaluation/metrics.py:11-19); the dict comes back with "loss" (differentiable
    scalar; its backward feeds the rasterizer's backward directly) and "sse_clipped" [B,V]."""
    B, V = extrinsics.shape[:2]
    h, w = image_shape
    assert use_sh or gaussian_sh_coefficients.shape[-1] == 1
    if torch.is_grad_enabled() and any(t.requires_grad for t in (extrinsics, intrinsics, near, far)):
        # the extension the reference binds has no camera gradients either, but its depth pass differentiates z through
        # extrinsics.inverse() in PyTorch (cuda_splatting.py:238-241); here z is computed in-kernel
        import warnings
        warnings.warn("render_views: cameras / near / far that require grad get no gradient from the rasterizer", stacklevel=2)
    if mse is not None and "count" not in mse:
        mse["count"] = B * V * 3 * h * w
    try:
        return _render_views_once(extrinsics, intrinsics, near, far, image_shape, background_color, gaussian_means,
                                  gaussian_covariances, gaussian_sh_coefficients, gaussian_opacities, scale_invariant, use_sh,
                                  depth_mode, want_radii, count_work, grad_reducer, mse)
    except PairLimitExceeded:
        if V == 1 and B == 1:
            raise
    # the parts of a split call give their workspaces back after the forward and rebuild them in the backward
    # (rasterizer.remat_depth): the workspace bound of one call then bounds the whole step
    from . import rasterizer as _R
    _R.remat_depth += 1
    try:
        return _render_views_split(extrinsics, intrinsics, near, far, image_shape, background_color, gaussian_means, gaussian_covariances,
                                   gaussian_sh_coefficients, gaussian_opacities, scale_invariant, use_sh, depth_mode, want_radii, count_work,
                                   grad_reducer, mse)
    finally:
        _R.remat_depth -= 1


def _render_views_split(extrinsics, intrinsics, near, far, image_shape, background_color, gaussian_means, gaussian_covariances,
                        gaussian_sh_coefficients, gaussian_opacities, scale_invariant, use_sh, depth_mode, want_radii, count_work,
                        grad_reducer, mse):
    B, V = extrinsics.shape[:2]
    if V == 1:  # one view per scene: split the SCENES instead
        hb = B // 2
        bsl = (slice(0, hb), slice(hb, B))
        sub = [None if mse is None else dict(target=mse["target"][sl], weight=mse["weight"], l1=mse.get("l1", False), count=mse["count"]) for sl in bsl]
        parts = [render_views(extrinsics[sl], intrinsics[sl], near[sl], far[sl], image_shape, background_color if background_color.dim() == 1 else background_color[sl],
                              gaussian_means[sl], gaussian_covariances[sl], gaussian_sh_coefficients[sl], gaussian_opacities[sl], scale_invariant, use_sh,
                              depth_mode, want_radii, count_work, grad_reducer, m) for sl, m in zip(bsl, sub)]
        if mse is not None:
            mse["loss"] = sub[0]["loss"] + sub[1]["loss"]
            mse["sse_clipped"] = torch.cat([sub[0]["sse_clipped"], sub[1]["sse_clipped"]], dim=0)
        out = [torch.cat([p[0] for p in parts], dim=0), None if parts[0][1] is None else torch.cat([p[1] for p in parts], dim=0)]
        if want_radii:
            out.append(torch.cat([p[2] for p in parts], dim=0))
        return tuple(out)
    # more than 2^32 (tile, Gaussian) pairs in one call (huge Gaussians, many views): split the views and
    # concatenate -- each half is its own autograd node, gradients add up as usual
    half = V // 2
    bgs = (background_color, background_color) if background_color.dim() == 1 else (background_color[:, :half], background_color[:, half:])
    sls = (slice(0, half), slice(half, V))
    sub = [None if mse is None else dict(target=mse["target"][:, sl], weight=mse["weight"], l1=mse.get("l1", False), count=mse["count"]) for sl in sls]
    parts = [render_views(extrinsics[:, sl], intrinsics[:, sl], near[:, sl], far[:, sl], image_shape, bg, gaussian_means,
                          gaussian_covariances, gaussian_sh_coefficients, gaussian_opacities, scale_invariant, use_sh, depth_mode,
                          want_radii, count_work, grad_reducer, m) for sl, bg, m in zip(sls, bgs, sub)]
    if mse is not None:  # both halves divide by the count of the whole call: the loss is their sum
        mse["loss"] = sub[0]["loss"] + sub[1]["loss"]
        mse["sse_clipped"] = torch.cat([sub[0]["sse_clipped"], sub[1]["sse_clipped"]], dim=1)
    out = [torch.cat([p[0] for p in parts], dim=1),
           None if parts[0][1] is None else torch.cat([p[1] for p in parts], dim=1)]
    if want_radii:
        out.append(torch.cat([p[2] for p in parts], dim=1))
    return tuple(out)


def _render_views_once(extrinsics, intrinsics, near, far, image_shape, background_color, gaussian_means, gaussian_covariances,
                       gaussian_sh_coefficients, gaussian_opacities, scale_invariant, use_sh, depth_mode, want_radii, count_work,
                       grad_reducer=None, mse=None):
    B, V = extrinsics.shape[:2]
    h, w = image_shape
    dev = gaussian_means.device
    ext = extrinsics.reshape(B * V, 4, 4).to(torch.float32)
    K = intrinsics.reshape(B * V, 3, 3).to(torch.float32)
    near_f, far_f = near.reshape(B * V).to(torch.float32), far.reshape(B * V).to(torch.float32)

    view, full, campos, tanfov, scale_pack, depth_affine, depth_clamp = _camera_tensors_cached(
        ext, K, near_f, far_f, depth_mode is not None, scale_invariant)

    bg = background_color.to(device=dev, dtype=torch.float32)
    bg = bg.expand(B, V, 3).reshape(B * V, 3).contiguous() if bg.dim() == 1 else bg.reshape(B * V, 3).contiguous()
    scene_index = _scene_index(dev, B, V)

    pack = ViewPack(scene_index, view, full, campos, tanfov, bg, h, w, scale_pack, depth_mode, depth_affine, depth_clamp,
                    grad_reducer)
    if mse is not None:
        pack.mse_target = mse["target"].reshape(B * V, 3, h, w)
        pack.mse_weight, pack.mse_l1, pack.mse_count = float(mse["weight"]), bool(mse.get("l1", False)), int(mse["count"])
    if use_sh:
        degree = isqrt(gaussian_sh_coefficients.shape[-1]) - 1
        colors = gaussian_sh_coefficients
    else:
        degree = 0
        colors = gaussian_sh_coefficients[..., 0]
    color, depth, radii = rasterize(gaussian_means, gaussian_covariances, colors, gaussian_opacities, pack, use_sh=use_sh,
                                    sh_degree=degree, sh_layout=_lib.SH_CHANNEL_MAJOR, want_radii=want_radii, count_work=count_work)
    color = color.reshape(B, V, 3, h, w)
    depth = None if depth is None else depth.reshape(B, V, h, w)
    if mse is not None:
        mse["loss"], mse["sse_clipped"] = pack.mse_result[0], pack.mse_result[1].reshape(B, V)
    if want_radii:
        return color, depth, radii.reshape(B, V, -1)
    return color, depth


def render_views_raw(
    extrinsics: Tensor,            # [B,V,4,4] TARGET cameras, camera-to-world
    intrinsics: Tensor,            # [B,V,3,3] normalised
    near: Tensor,                  # [B,V]
    far: Tensor,                   # [B,V]
    image_shape: tuple[int, int],
    background_color: Tensor,      # [3] or [B,V,3]
    head: Tensor,                  # [B,Vc,37,h,w] raw channel planes of the encoder head
    depth: Tensor,                 # [B,Vc,h,w]
    context_images: Tensor,        # [B,Vc,3,h,w]
    context_extrinsics: Tensor,    # [B,Vc,4,4]
    context_intrinsics: Tensor,    # [B,Vc,3,3]
    scale_min: float,
    scale_max: float,
    sh_mask: Tensor,               # [9]
    depth_mode: Optional[DepthRenderingMode] = None,
    cooked_out: Optional[Tensor] = None,
):
    """Encoder head output -> images, with the Gaussian adapter fused into the projection (see
    gaussian_adapter.FusedAdapterDecoder).  Returns (color [B,V,3,H,W], depth [B,V,H,W] | None)."""
    from .gaussian_adapter import sh_rotation_matrices
    from .rasterizer import RawScene, rasterize_raw
    B, V = extrinsics.shape[:2]
    h, w = image_shape
    Vc, hc, wc = head.shape[1], head.shape[3], head.shape[4]
    dev = head.device
    ext = extrinsics.reshape(B * V, 4, 4).to(torch.float32)
    K = intrinsics.reshape(B * V, 3, 3).to(torch.float32)
    near_f, far_f = near.reshape(B * V).to(torch.float32), far.reshape(B * V).to(torch.float32)
    view, full, campos, tanfov, scale_pack, depth_affine, depth_clamp = _camera_tensors_cached(ext, K, near_f, far_f, depth_mode is not None, True)
    bg = background_color.to(device=dev, dtype=torch.float32)
    bg = bg.expand(B, V, 3).reshape(B * V, 3).contiguous() if bg.dim() == 1 else bg.reshape(B * V, 3).contiguous()
    pack = ViewPack(_scene_index(dev, B, V), view, full, campos, tanfov, bg, h, w, scale_pack, depth_mode, depth_affine, depth_clamp)
    # per context view: rotation | translation | K^-1 | pad | degree-2 SH rotation | SH mask
    R = context_extrinsics[..., :3, :3].to(torch.float32)
    cam = torch.cat([R.reshape(B, Vc, 9), context_extrinsics[..., :3, 3].to(torch.float32), context_intrinsics.to(torch.float32).inverse().reshape(B, Vc, 9),
                     torch.zeros(B, Vc, 1, device=dev), sh_rotation_matrices(R, 2)[2].reshape(B, Vc, 25),
                     sh_mask.to(device=dev, dtype=torch.float32).expand(B, Vc, 9)], dim=-1).contiguous()
    raw = RawScene(Vc, hc, wc, scale_min, scale_max, context_images.reshape(B, Vc, 3, hc * wc), cam, cooked_out)
    color, dimg, _ = rasterize_raw(head.reshape(B, Vc, 37, hc * wc), depth.reshape(B, Vc, hc * wc), raw, pack)
    return color.reshape(B, V, 3, h, w), (None if dimg is None else dimg.reshape(B, V, h, w))


def render_cuda(
    extrinsics: Tensor,                # [batch,4,4]
    intrinsics: Tensor,                # [batch,3,3]
    near: Tensor,                      # [batch]
    far: Tensor,                       # [batch]
    image_shape: tuple[int, int],
    background_color: Tensor,          # [batch,3]
    gaussian_means: Tensor,            # [batch,gaussian,3]
    gaussian_covariances: Tensor,      # [batch,gaussian,3,3]
    gaussian_sh_coefficients: Tensor,  # [batch,gaussian,3,d_sh]
    gaussian_opacities: Tensor,        # [batch,gaussian]
    scale_invariant: bool = True,
    use_sh: bool = True,
) -> Tensor:
    """-> [batch,3,height,width].  Every batch element is its own (scene, camera) pair, as in the
    reference; all of them are rendered by one call."""
    color, _ = render_views(
        extrinsics[:, None], intrinsics[:, None], near[:, None], far[:, None], image_shape, background_color[:, None],
        gaussian_means, gaussian_covariances, gaussian_sh_coefficients, gaussian_opacities,
        scale_invariant=scale_invariant, use_sh=use_sh,
    )
    return color[:, 0]


def render_cuda_orthographic(
    extrinsics: Tensor,                # [batch,4,4]
    width: Tensor,                     # [batch]
    height: Tensor,                    # [batch]
    near: Tensor,                      # [batch]
    far: Tensor,                       # [batch]
    image_shape: tuple[int, int],
    background_color: Tensor,          # [batch,3]
    gaussian_means: Tensor,            # [batch,gaussian,3]
    gaussian_covariances: Tensor,      # [batch,gaussian,3,3]
    gaussian_sh_coefficients: Tensor,  # [batch,gaussian,3,d_sh]
    gaussian_opacities: Tensor,        # [batch,gaussian]
    fov_degrees: float = 0.1,
    use_sh: bool = True,
    dump: dict | None = None,
) -> Tensor:
    """-> [batch,3,height,width].  "Orthographic" = a camera moved far back with a tiny field of view
    (reference :129-219); same kernels, different matrices, no scale-invariant normalisation."""
    b = extrinsics.shape[0]
    h, w = image_shape
    assert use_sh or gaussian_sh_coefficients.shape[-1] == 1
    dev = extrinsics.device

    fov_x = torch.tensor(fov_degrees, device=dev).deg2rad()
    tan_fov_x = (0.5 * fov_x).tan()
    distance_to_near = (0.5 * width) / tan_fov_x
    tan_fov_y = 0.5 * height / distance_to_near
    fov_y = (2 * tan_fov_y).atan()
    near = near + distance_to_near
    far = far + distance_to_near
    move_back = torch.eye(4, dtype=torch.float32, device=dev)
    move_back[2, 3] = -distance_to_near
    extrinsics = extrinsics @ move_back

    if dump is not None:  # escape hatch for visualisation code, as in the reference
        dump["extrinsics"] = extrinsics
        dump["fov_x"] = fov_x
        dump["fov_y"] = fov_y
        dump["near"] = near
        dump["far"] = far

    view, full, campos, tanfov = _camera_block(extrinsics.to(torch.float32), near, far, fov_x.expand(b), fov_y, tan_fov_x, tan_fov_y)
    scene_index = torch.arange(b, device=dev, dtype=torch.int32)
    bg = background_color.to(device=dev, dtype=torch.float32).reshape(b, 3).contiguous()
    pack = ViewPack(scene_index, view, full, campos, tanfov, bg, h, w)
    if use_sh:
        degree, colors = isqrt(gaussian_sh_coefficients.shape[-1]) - 1, gaussian_sh_coefficients
    else:
        degree, colors = 0, gaussian_sh_coefficients[..., 0]
    color, _, _ = rasterize(gaussian_means, gaussian_covariances, colors, gaussian_opacities, pack, use_sh=use_sh, sh_degree=degree)
    return color


def render_depth_cuda(
    extrinsics: Tensor,            # [batch,4,4]
    intrinsics: Tensor,            # [batch,3,3]
    near: Tensor,                  # [batch]
    far: Tensor,                   # [batch]
    image_shape: tuple[int, int],
    gaussian_means: Tensor,        # [batch,gaussian,3]
    gaussian_covariances: Tensor,  # [batch,gaussian,3,3]
    gaussian_opacities: Tensor,    # [batch,gaussian]
    scale_invariant: bool = True,
    mode: DepthRenderingMode = "depth",
) -> Tensor:
    """-> [batch,height,width]: alpha-composited camera-space depth (or 1/z, or the reference's
    clamped log) over a zero background."""
    b, g = gaussian_opacities.shape
    dummy = torch.zeros((b, g, 3, 1), dtype=torch.float32, device=gaussian_means.device)
    _, depth = render_views(
        extrinsics[:, None], intrinsics[:, None], near[:, None], far[:, None], image_shape,
        torch.zeros((b, 1, 3), dtype=torch.float32, device=gaussian_means.device),
        gaussian_means, gaussian_covariances, dummy, gaussian_opacities,
        scale_invariant=scale_invariant, use_sh=False, depth_mode=mode,
    )
    return depth[:, 0]


__all__ = ["DepthRenderingMode", "get_projection_matrix", "render_cuda", "render_cuda_orthographic", "render_depth_cuda",
           "render_views", "homogenize_points"]
