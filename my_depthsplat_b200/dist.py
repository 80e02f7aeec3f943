"""Multi-GPU sharding of the rendering path: one process per GPU, torch.distributed (NCCL on GPUs,
gloo in the CPU tests).

The reference runs on one device only (src/main.py:145-147 hard-codes ``devices=1``), and renders the
(scene, view) pairs of a batch one after the other (cuda_splatting.py:90).  Those pairs are independent
given the Gaussians, so the path shards without any data-path collective:

  * inference / video (BASELINE config 3): the Gaussians are replicated, the target views are split
    contiguously over the ranks, each rank renders its slice; ``all_gather_views`` optionally
    reassembles the frames.
  * training with SCENES sharded (config 4, one scene per GPU): nothing to exchange on this path.
  * training with the VIEWS of one scene sharded: forward as above; the backward produces, on every
    rank, the partial gradient of ITS views w.r.t. the (replicated) Gaussian tensors.  The sum over
    views that the reference gets from the autograd of its per-view ``repeat``
    (decoder_splatting_cuda.py:53-56) therefore crosses ranks: ``sync_gaussian_grads`` is an identity
    in forward whose backward packs the four gradients into one flat buffer and sums it with ONE
    all-reduce (NCCL over NVLink 5 / NVSwitch; 37-40 floats per Gaussian).
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch
import torch.distributed as dist
from torch import Tensor

from .types import DecoderOutput, Gaussians


def shard_bounds(num_items: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced split: the first ``num_items % world_size`` ranks get one item more."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, extra = divmod(num_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_views(t: Tensor, world_size: int, rank: int, dim: int = 1) -> Tensor:
    """Slice of a ``[B, V, ...]`` camera tensor that belongs to ``rank``."""
    a, b = shard_bounds(t.shape[dim], world_size, rank)
    return t.narrow(dim, a, b - a)


class _SyncGrads(torch.autograd.Function):
    """Identity on the tensors; the backward sums the incoming gradients over the process group with a
    single all-reduce of one flat buffer."""

    @staticmethod
    def forward(ctx, group, *tensors):
        ctx.group = group
        return tuple(t.view_as(t) for t in tensors)

    @staticmethod
    def backward(ctx, *grads):
        shapes = [g.shape for g in grads]
        flat = torch.cat([g.contiguous().reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=ctx.group)
        out, o = [], 0
        for s in shapes:
            n = s.numel()
            out.append(flat[o:o + n].view(s))
            o += n
        return (None, *out)


def sync_gaussian_grads(gaussians: Gaussians, group: Optional[dist.ProcessGroup] = None) -> Gaussians:
    """Gaussians whose gradients are summed over the ranks on the way back (no-op without a process
    group of more than one rank)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return gaussians
    m, c, h, o = _SyncGrads.apply(group, gaussians.means, gaussians.covariances, gaussians.harmonics, gaussians.opacities)
    return Gaussians(m, c, h, o)


def all_gather_views(local: Tensor, num_views: int, dim: int = 1, group: Optional[dist.ProcessGroup] = None) -> Tensor:
    """Reassemble ``[B, V, ...]`` from the per-rank ``[B, v_r, ...]`` slices (uneven slices allowed)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_bounds(num_views, world, r) for r in range(world)]
    vmax = max(b - a for a, b in sizes)
    pad_shape = list(local.shape)
    pad_shape[dim] = vmax
    padded = local.new_zeros(pad_shape)
    padded.narrow(dim, 0, local.shape[dim]).copy_(local)
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded.contiguous(), group=group)
    return torch.cat([b_.narrow(dim, 0, hi - lo) for b_, (lo, hi) in zip(bufs, sizes)], dim=dim)


class ViewShardedDecoder(torch.nn.Module):
    """Wraps a decoder (``DecoderSplattingCUDA`` or anything with its ``forward`` signature): every rank
    renders its contiguous slice of the target views of every scene; per-Gaussian gradients are summed
    across ranks in the backward.  ``gather=True`` returns the full ``[B, V, ...]`` frames on every rank
    (inference); otherwise the local slice (training: the loss is computed on the local views)."""

    def __init__(self, decoder: torch.nn.Module, group: Optional[dist.ProcessGroup] = None, gather: bool = False):
        super().__init__()
        self.decoder = decoder
        self.group = group
        self.gather = gather

    def forward(self, gaussians: Gaussians, extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor,
                image_shape: tuple[int, int], depth_mode=None) -> DecoderOutput:
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        V = extrinsics.shape[1]
        sl = [shard_views(t, world, rank) for t in (extrinsics, intrinsics, near, far)]
        g = sync_gaussian_grads(gaussians, self.group) if torch.is_grad_enabled() else gaussians
        out = self.decoder.forward(g, *sl, image_shape, depth_mode=depth_mode)
        if self.gather and world > 1:
            color = all_gather_views(out.color, V, group=self.group)
            depth = None if out.depth is None else all_gather_views(out.depth, V, group=self.group)
            return DecoderOutput(color, depth)
        return out


def render_sharded(render_fn: Callable[..., Sequence[Tensor]], gaussians: Gaussians, cameras: Sequence[Tensor], *args,
                   group: Optional[dist.ProcessGroup] = None, **kw):
    """Functional form used by the tests: ``render_fn(gaussians, *camera_slices, *args, **kw)`` on this
    rank's view slice with gradient synchronisation."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    sl = [shard_views(t, world, rank) for t in cameras]
    return render_fn(sync_gaussian_grads(gaussians, group), *sl, *args, **kw)
