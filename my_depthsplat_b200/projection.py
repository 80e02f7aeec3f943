"""The two camera helpers the hot path uses from src/geometry/projection.py of the reference
(get_fov :233-247, homogenize_points :9-13).  The sequence of torch operations is kept (inverse,
mat-vec, normalise, dot, acos) so that the field of view -- and hence tanfov, focal lengths and
tile rects -- comes out bit-identical to the reference's for the same normalised intrinsics.
"""
from __future__ import annotations

import torch
from torch import Tensor

# mid-points of the left, right, top and bottom image edges in normalised image coordinates
_EDGE_MIDPOINTS = ((0.0, 0.5), (1.0, 0.5), (0.5, 0.0), (0.5, 1.0))


def homogenize_points(points: Tensor) -> Tensor:
    """(..., d) -> (..., d+1) with a trailing one."""
    one = torch.ones_like(points[..., :1])
    return torch.cat([points, one], dim=-1)


def inverse_nosync(m: Tensor) -> Tensor:
    """``m.inverse()`` without its device->host synchronisation: ``Tensor.inverse`` reads back LAPACK's
    ``info`` to raise on singular input, which stalls the host until the GPU has drained.  ``inv_ex`` runs
    the same factorisation (bit-identical result) and leaves ``info`` on the device."""
    return torch.linalg.inv_ex(m, check_errors=False).inverse


_pixel_cache: dict = {}


def _edge_pixel(device, u: float, v: float) -> Tensor:
    """Homogeneous edge mid-point as a device tensor, created once per device: building it per call
    (``torch.tensor(..., device=cuda)``) is a pageable host->device copy that SYNCHRONISES the stream, i.e.
    stalls the host until the previous step's kernels have drained and exposes every following launch."""
    key = (device, u, v)
    t = _pixel_cache.get(key)
    if t is None:
        t = torch.tensor([u, v, 1.0], dtype=torch.float32, device=device)
        _pixel_cache[key] = t
    return t


def _unit_ray(k_inv: Tensor, u: float, v: float) -> Tensor:
    ray = torch.einsum("bij,j->bi", k_inv, _edge_pixel(k_inv.device, u, v))
    return ray / ray.norm(dim=-1, keepdim=True)


def get_fov(intrinsics: Tensor) -> Tensor:
    """Normalised intrinsics [b,3,3] -> [b,2] = (fov_x, fov_y): the angle between the unprojected
    rays through opposite edge mid-points.  The principal point does not survive this (the
    frustum built from it is symmetric), exactly as in the reference."""
    k_inv = inverse_nosync(intrinsics)
    left, right, top, bottom = (_unit_ray(k_inv, u, v) for u, v in _EDGE_MIDPOINTS)
    fov_x = (left * right).sum(dim=-1).acos()
    fov_y = (top * bottom).sum(dim=-1).acos()
    return torch.stack((fov_x, fov_y), dim=-1)
