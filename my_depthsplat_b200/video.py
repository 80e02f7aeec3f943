"""Clip rendering: the caller-side loop of the reference around the decoder.

The reference renders a video in chunks of ``test.render_chunk_size`` target views, one ``decoder.forward``
per chunk on one GPU, and grows the result with ``torch.cat`` after every chunk
(src/model/model_wrapper.py:455-484 -- each cat re-copies all frames so far).  ``render_clip`` keeps that
contract (same decoder signature, colour of all V views in order, depth ignored unless asked for) and
changes the schedule:

  * the V target views are split contiguously over the ranks of the process group (Gaussians replicated,
    no data-path collective -- ``dist.shard_bounds``); every rank renders its slice in chunks;
  * frames are written straight into ONE preallocated result (no cat chain);
  * ``to_host=True`` streams every finished chunk to pinned host memory on a copy stream while the next
    chunk renders (a clip is consumed by the video encoder / the metrics on the host), two chunk buffers
    in flight, so device memory holds two chunks of frames instead of the clip;
  * ``gather=True`` reassembles the whole clip on every rank (device results only).

Runs under ``torch.no_grad`` (the reference's test/video path is inference).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist
from torch import Tensor

from .dist import _require_views, all_gather_views, shard_bounds
from .types import DecoderOutput, Gaussians


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


@torch.no_grad()
def render_clip(decoder, gaussians: Gaussians, extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor,
                image_shape: tuple[int, int], chunk_size: Optional[int] = 10, depth_mode=None, group=None,
                gather: bool = False, to_host: bool = False) -> DecoderOutput:
    """extrinsics [B,V,4,4], intrinsics [B,V,3,3], near/far [B,V] -> DecoderOutput(color [B,v,3,H,W], depth [B,v,H,W] | None)
    with v = this rank's views (all V when the group has one rank or ``gather``).  ``chunk_size=None`` renders the
    slice in one call (the rasterizer still splits a call that exceeds its pair limit)."""
    world, rank = _world(group)
    B, V = extrinsics.shape[:2]
    H, W = image_shape
    _require_views(V, world)  # a rank without frames would skip the gather and hang the others
    lo, hi = shard_bounds(V, world, rank)
    v = hi - lo
    dev = extrinsics.device
    step = v if not chunk_size else max(1, int(chunk_size))
    if to_host and gather:
        raise ValueError("gather=True returns device tensors; use to_host on the rank-local slice")
    if to_host and dev.type != "cuda":
        raise ValueError("to_host needs CUDA tensors")  # no CPU path
    where = dict(dtype=torch.float32, pin_memory=True) if to_host else dict(dtype=torch.float32, device=dev)
    color = torch.empty((B, v, 3, H, W), **where)
    depth = torch.empty((B, v, H, W), **where) if depth_mode is not None else None

    copy_stream = torch.cuda.Stream(dev) if to_host else None
    pending = []  # (event, chunk outputs kept alive until their copy is done)
    for a in range(0, v, step):
        b = min(a + step, v)
        sl = slice(lo + a, lo + b)
        out = decoder.forward(gaussians, extrinsics[:, sl], intrinsics[:, sl], near[:, sl], far[:, sl], image_shape, depth_mode=depth_mode)
        if not to_host:
            color[:, a:b].copy_(out.color)
            if depth is not None:
                depth[:, a:b].copy_(out.depth)
            continue
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(dev))
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done)
            color[:, a:b].copy_(out.color, non_blocking=True)
            if depth is not None:
                depth[:, a:b].copy_(out.depth, non_blocking=True)
            copied = torch.cuda.Event()
            copied.record(copy_stream)
        pending.append((copied, out))
        while len(pending) > 2:  # at most two chunks of frames live on the device
            pending.pop(0)[0].synchronize()
    for ev, _ in pending:
        ev.synchronize()
    if gather and world > 1:
        color = all_gather_views(color, V, group=group)
        depth = None if depth is None else all_gather_views(depth, V, group=group)
    return DecoderOutput(color, depth)
