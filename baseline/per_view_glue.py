"""The reference's rendering glue restated on top of ANY `diff_gaussian_rasterization`-shaped module
(comparator / test infrastructure, NOT product code).

Follows, step for step, what the reference does around its rasterizer:
  * decoder_splatting_cuda.py:35-91 -- flatten (scene, view), REPLICATE the Gaussians once per view, colour render,
    optional second render for depth;
  * cuda_splatting.py:46-126 -- scale-invariant normalisation, SH re-layout, field of view from the normalised
    intrinsics, projection / view matrices in transposed storage, then a PYTHON LOOP over the flattened views
    that builds the settings (two ``.item()`` host reads per view), gathers the covariance upper triangle and
    calls the extension once per view;
  * cuda_splatting.py:225-264 -- camera-space z (or its transform) rendered as a colour, mean over channels.

``ext`` is the extension module: ``oracle.ext_compat`` (CPU; tests/test_reference_glue.py checks this file
against the reference's unmodified source in the build container) or ``baseline.upstream_ext`` (GPU; what
bench.py times as "gpu_baseline").  Unlike the product's camera code this one keeps ``Tensor.inverse()`` and its
host synchronisation, because that is what the reference executes.
"""
from __future__ import annotations

from math import isqrt

import torch

from my_depthsplat_b200.cuda_splatting import get_projection_matrix
from my_depthsplat_b200.projection import homogenize_points


def get_fov(intrinsics):
    """projection.py:233-247 of the reference: angle between the rays through opposite edge mid-points."""
    k_inv = intrinsics.inverse()

    def ray(u, v):
        p = torch.tensor([u, v, 1.0], dtype=torch.float32, device=intrinsics.device)
        r = torch.einsum("bij,j->bi", k_inv, p)
        return r / r.norm(dim=-1, keepdim=True)

    left, right, top, bottom = ray(0.0, 0.5), ray(1.0, 0.5), ray(0.5, 0.0), ray(0.5, 1.0)
    return torch.stack(((left * right).sum(dim=-1).acos(), (top * bottom).sum(dim=-1).acos()), dim=-1)


def render_cuda(ext, extrinsics, intrinsics, near, far, image_shape, background_color, gaussian_means, gaussian_covariances,
                gaussian_sh_coefficients, gaussian_opacities, scale_invariant=True, use_sh=True):
    assert use_sh or gaussian_sh_coefficients.shape[-1] == 1
    if scale_invariant:
        scale = 1 / near
        extrinsics = extrinsics.clone()
        extrinsics[..., :3, 3] = extrinsics[..., :3, 3] * scale[:, None]
        gaussian_covariances = gaussian_covariances * (scale[:, None, None, None] ** 2)
        gaussian_means = gaussian_means * scale[:, None, None]
        near = near * scale
        far = far * scale
    n = gaussian_sh_coefficients.shape[-1]
    degree = isqrt(n) - 1
    shs = gaussian_sh_coefficients.permute(0, 1, 3, 2).contiguous()  # [b, g, n, xyz]
    b = extrinsics.shape[0]
    h, w = image_shape
    fov_x, fov_y = get_fov(intrinsics).unbind(dim=-1)
    tan_fov_x = (0.5 * fov_x).tan()
    tan_fov_y = (0.5 * fov_y).tan()
    projection_matrix = get_projection_matrix(near, far, fov_x, fov_y).transpose(1, 2)
    view_matrix = extrinsics.inverse().transpose(1, 2)
    full_projection = view_matrix @ projection_matrix
    images = []
    for i in range(b):
        mean_gradients = torch.zeros_like(gaussian_means[i], requires_grad=True)
        settings = ext.GaussianRasterizationSettings(
            image_height=h, image_width=w, tanfovx=tan_fov_x[i].item(), tanfovy=tan_fov_y[i].item(), bg=background_color[i],
            scale_modifier=1.0, viewmatrix=view_matrix[i], projmatrix=full_projection[i], sh_degree=degree,
            campos=extrinsics[i, :3, 3], prefiltered=False, debug=False)
        rasterizer = ext.GaussianRasterizer(settings)
        row, col = torch.triu_indices(3, 3)
        image, _radii = rasterizer(
            means3D=gaussian_means[i], means2D=mean_gradients, shs=shs[i] if use_sh else None,
            colors_precomp=None if use_sh else shs[i, :, 0, :], opacities=gaussian_opacities[i, ..., None],
            cov3D_precomp=gaussian_covariances[i, :, row, col])
        images.append(image)
    return torch.stack(images)


def render_depth_cuda(ext, extrinsics, intrinsics, near, far, image_shape, gaussian_means, gaussian_covariances, gaussian_opacities,
                      scale_invariant=True, mode="depth"):
    camera_space = torch.einsum("bij,bgj->bgi", extrinsics.inverse(), homogenize_points(gaussian_means))
    fake_color = camera_space[..., 2]
    if mode == "disparity":
        fake_color = 1 / fake_color
    elif mode == "log":
        fake_color = fake_color.minimum(near[:, None]).maximum(far[:, None]).log()
    b = fake_color.shape[0]
    result = render_cuda(ext, extrinsics, intrinsics, near, far, image_shape,
                         torch.zeros((b, 3), dtype=fake_color.dtype, device=fake_color.device), gaussian_means, gaussian_covariances,
                         fake_color[:, :, None, None].repeat(1, 1, 3, 1), gaussian_opacities, scale_invariant=scale_invariant, use_sh=False)
    return result.mean(dim=1)


def render_cuda_orthographic(ext, extrinsics, width, height, near, far, image_shape, background_color, gaussian_means,
                             gaussian_covariances, gaussian_sh_coefficients, gaussian_opacities, fov_degrees=0.1, use_sh=True):
    """cuda_splatting.py:129-219 of the reference: a fake orthographic camera (tiny field of view, camera moved back so that
    the near plane keeps the requested width), then the same per-view loop; no scale-invariant normalisation.  The reference
    hands the whole ``tan_fov_y`` tensor to every view's settings (it is only ever called with one view); here view i gets
    element i, which is the same thing for one view."""
    b = extrinsics.shape[0]
    h, w = image_shape
    assert use_sh or gaussian_sh_coefficients.shape[-1] == 1
    n = gaussian_sh_coefficients.shape[-1]
    degree = isqrt(n) - 1
    shs = gaussian_sh_coefficients.permute(0, 1, 3, 2).contiguous()
    fov_x = torch.tensor(fov_degrees, device=extrinsics.device).deg2rad()
    tan_fov_x = (0.5 * fov_x).tan()
    distance_to_near = (0.5 * width) / tan_fov_x
    tan_fov_y = 0.5 * height / distance_to_near
    fov_y = (2 * tan_fov_y).atan()
    near = near + distance_to_near
    far = far + distance_to_near
    move_back = torch.eye(4, dtype=torch.float32, device=extrinsics.device)
    move_back[2, 3] = -distance_to_near
    extrinsics = extrinsics @ move_back
    projection_matrix = get_projection_matrix(near, far, fov_x.expand(b), fov_y).transpose(1, 2)
    view_matrix = extrinsics.inverse().transpose(1, 2)
    full_projection = view_matrix @ projection_matrix
    images = []
    for i in range(b):
        mean_gradients = torch.zeros_like(gaussian_means[i], requires_grad=True)
        settings = ext.GaussianRasterizationSettings(
            image_height=h, image_width=w, tanfovx=tan_fov_x, tanfovy=tan_fov_y[i], bg=background_color[i], scale_modifier=1.0,
            viewmatrix=view_matrix[i], projmatrix=full_projection[i], sh_degree=degree, campos=extrinsics[i, :3, 3],
            prefiltered=False, debug=False)
        row, col = torch.triu_indices(3, 3)
        image, _radii = ext.GaussianRasterizer(settings)(
            means3D=gaussian_means[i], means2D=mean_gradients, shs=shs[i] if use_sh else None,
            colors_precomp=None if use_sh else shs[i, :, 0, :], opacities=gaussian_opacities[i, ..., None],
            cov3D_precomp=gaussian_covariances[i, :, row, col])
        images.append(image)
    return torch.stack(images)


def _per_view(t, v):
    """[b, ...] -> [(b v), ...]: a materialised copy per view, like einops.repeat in the reference."""
    return t.repeat_interleave(v, dim=0)


def decoder_forward(ext, gaussians, extrinsics, intrinsics, near, far, image_shape, background, depth_mode=None):
    """-> (color [b,v,3,h,w], depth [b,v,h,w] | None)."""
    b, v = extrinsics.shape[:2]
    flat = lambda t: t.reshape(b * v, *t.shape[2:])
    bg = background.reshape(1, 3).repeat(b * v, 1) if background.dim() == 1 else flat(background)
    color = render_cuda(ext, flat(extrinsics), flat(intrinsics), flat(near), flat(far), image_shape, bg,
                        _per_view(gaussians.means, v), _per_view(gaussians.covariances, v), _per_view(gaussians.harmonics, v),
                        _per_view(gaussians.opacities, v))
    color = color.reshape(b, v, *color.shape[1:])
    depth = None
    if depth_mode is not None:
        depth = render_depth_cuda(ext, flat(extrinsics), flat(intrinsics), flat(near), flat(far), image_shape,
                                  _per_view(gaussians.means, v), _per_view(gaussians.covariances, v), _per_view(gaussians.opacities, v),
                                  mode=depth_mode)
        depth = depth.reshape(b, v, *depth.shape[1:])
    return color, depth
