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
    in forward whose backward all-reduces the four gradient tensors in place (NCCL over NVLink 5 /
    NVSwitch; 40 floats per Gaussian).  Three alternatives are implemented and measured, all opt-in:
    ``NvlsAllReducer`` (gradients land in symmetric memory, the library's own two-shot NVLS kernel sums them in place:
    slightly faster than NCCL at N=8, slower at N=2), ``FusedGradReducer`` (the backward kernel adds into NVLS
    multicast memory) and ``ChunkedAllReducer`` (projection backward in ranges, async all-reduce per range).
"""
from __future__ import annotations

import os
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


def _require_views(num_views: int, world_size: int) -> None:
    """Every rank must own at least one view: a rank without views would skip the rasterizer and with it the collectives of
    the backward, and the other ranks would wait for it forever.  The check is on values every rank knows, so all of
    them raise together."""
    if num_views < world_size:
        raise ValueError(f"{num_views} target views cannot be sharded over {world_size} ranks: every rank needs at least one "
                         "(use a smaller process group for this call, or shard scenes instead)")


def shard_views(t: Tensor, world_size: int, rank: int, dim: int = 1, interleave: bool = False) -> Tensor:
    """The views of a ``[B, V, ...]`` camera tensor that belong to ``rank``: a contiguous slice, or -- ``interleave`` -- the
    views rank, rank + world, rank + 2 world, ...  Along a camera path neighbouring views cost about the same, so
    contiguous slices give every rank a different part of the path (and a different load); interleaved, every rank's
    views span the whole path."""
    if interleave:
        return t[(slice(None),) * dim + (slice(rank, None, world_size),)]
    a, b = shard_bounds(t.shape[dim], world_size, rank)
    return t.narrow(dim, a, b - a)


class _SyncGrads(torch.autograd.Function):
    """Identity on the tensors; the backward sums the incoming gradients over the process group, in place."""

    @staticmethod
    def forward(ctx, group, *tensors):
        ctx.group = group
        return tuple(t.view_as(t) for t in tensors)

    @staticmethod
    def backward(ctx, *grads):
        # in-place all-reduce of the fresh outputs of the rasterizer's backward: no flatten / split copies of the 160 B
        # per Gaussian.  The rasterizer carves the four tensors out of one allocation, so it is ONE collective over
        # 40 floats per Gaussian; tensors from elsewhere get one call each.  (torch's _coalescing_manager would also
        # make it one NCCL group call, but it gave erratic step times on the B200 box: 9.8 - 15 ms against 9.8 ms.)
        grads = [g.contiguous() for g in grads]
        flat = _flat_span(grads)
        if flat is not None:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=ctx.group)   # one call over the rasterizer's flat gradient buffer
        else:
            for g in grads:
                dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
        return (None, *grads)


def _flat_span(tensors):
    """The 1-D tensor that covers ``tensors`` exactly when they tile one contiguous stretch of a single storage without
    gaps or overlaps (the rasterizer's backward allocates its four gradient tensors that way), else None."""
    if len(tensors) < 2 or any(t.dtype != tensors[0].dtype or not t.is_contiguous() for t in tensors):
        return None
    st = tensors[0].untyped_storage()
    if any(t.untyped_storage().data_ptr() != st.data_ptr() for t in tensors):
        return None
    spans = sorted((t.storage_offset(), t.storage_offset() + t.numel()) for t in tensors)
    if any(a[1] != b[0] for a, b in zip(spans, spans[1:])):
        return None
    return torch.empty(0, dtype=tensors[0].dtype, device=tensors[0].device).set_(st, spans[0][0], (spans[-1][1] - spans[0][0],))


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


class FusedGradReducer:
    """Sum of the per-Gaussian gradients over the ranks WITHOUT a separate collective: the preprocess-backward
    kernel adds its results straight into NVLS multicast memory (``multimem.red``: the NVSwitch applies every
    rank's contribution to every rank's replica), so the all-reduce of view-sharded training overlaps the
    kernel that produces the data instead of following it.  Symmetric buffers come from
    ``torch.distributed._symmetric_memory``; two of them alternate so that the consumer of step i may still
    be reading while step i+1 accumulates.  Falls back (``available == False``) when the group has one rank
    or the fabric offers no multicast."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None):
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        self._bufs = {}     # total floats -> [(tensor, handle)] * 2
        self._turn = 0
        self.available = self.world > 1 and torch.cuda.is_available()

    def _get(self, nfloats: int, device):
        key = (nfloats, device)
        if key not in self._bufs:
            import torch.distributed._symmetric_memory as symm_mem
            pair = []
            for _ in range(2):
                t = symm_mem.empty(nfloats, dtype=torch.float32, device=device)
                h = symm_mem.rendezvous(t, self.group)
                if not h.multicast_ptr:
                    self.available = False
                    raise RuntimeError("no NVLS multicast on this fabric")
                pair.append((t, h))
            self._bufs[key] = pair
        self._turn ^= 1
        return self._bufs[key][self._turn]

    def begin(self, shapes, device):
        """-> (local gradient tensors, multicast addresses).  Zeroes this rank's replica and waits (on the
        stream) until every rank has done so."""
        sizes = [int(torch.Size(s).numel()) for s in shapes]
        offs, o = [], 0
        for n in sizes:
            offs.append(o)
            o += (n + 3) // 4 * 4  # keep every segment 16-byte aligned for multimem.red.v4
        buf, h = self._get(o, device)
        buf.zero_()
        h.barrier(channel=0)
        self._cur = h
        locals_ = [buf[a:a + n].view(s) for a, n, s in zip(offs, sizes, shapes)]
        return locals_, [h.multicast_ptr + 4 * a for a in offs]

    def end(self):
        """All ranks' contributions have landed once this barrier has passed on the stream."""
        self._cur.barrier(channel=1)


class NvlsAllReducer:
    """Cross-rank gradient sum WITHOUT NCCL: the backward kernel writes its four gradient tensors (ordinary stores)
    into ONE buffer of torch symmetric memory, and the library's own two-shot NVLS kernel (``b200s_nvls_allreduce``,
    csrc/collective.cu: ``multimem.ld_reduce`` + ``multimem.st``, 1/N of the buffer per rank) sums it in place between
    two cross-rank barriers.  ``ROTATE`` buffers alternate: the tensors handed out by one backward stay valid until
    ``ROTATE - 1`` further backward passes have run (their consumer -- the encoder's backward, an optimiser -- reads
    them long before)."""

    ROTATE = 3
    in_place = True

    def __init__(self, group: Optional[dist.ProcessGroup] = None):
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        self._bufs = {}
        self._turn = 0
        self.available = self.world > 1 and torch.cuda.is_available()

    def begin(self, shapes, device):
        """-> local gradient tensors (views of this rank's replica of the symmetric buffer)."""
        import torch.distributed._symmetric_memory as symm_mem
        sizes = [int(torch.Size(s).numel()) for s in shapes]
        offs, o = [], 0
        for n in sizes:
            offs.append(o)
            o += (n + 3) // 4 * 4
        key = (o, device)
        if key not in self._bufs:
            ring = []
            for _ in range(self.ROTATE):
                t = symm_mem.empty(o, dtype=torch.float32, device=device)
                h = symm_mem.rendezvous(t, self.group)
                if not h.multicast_ptr:
                    self.available = False
                    raise RuntimeError("no NVLS multicast on this fabric")
                t.zero_()  # the alignment gaps between the tensors are summed too
                ring.append((t, h))
            self._bufs[key] = ring
        self._turn = (self._turn + 1) % self.ROTATE
        self._cur = self._bufs[key][self._turn]
        buf = self._cur[0]
        return [buf[a:a + n].view(s) for a, n, s in zip(offs, sizes, shapes)]

    def end(self):
        from . import _lib
        buf, h = self._cur
        stream = torch.cuda.current_stream(buf.device).cuda_stream
        h.barrier(channel=0)   # every rank's gradients are complete
        _lib.check(_lib.load().b200s_nvls_allreduce(h.multicast_ptr, buf.numel(), self.rank, self.world, stream), "b200s_nvls_allreduce")
        h.barrier(channel=1)   # every rank's multicast stores have landed


class ChunkedAllReducer:
    """Sum of the per-Gaussian gradients over the ranks, OVERLAPPED with the kernel that produces them: the
    rasterizer runs its projection backward in ``chunks`` Gaussian ranges and hands each finished range to an
    asynchronous NCCL all-reduce (NCCL's own stream), so the NVLink transfer of range k runs under the
    computation of range k+1.  The backward returns after waiting (on the stream) for all of them."""

    chunked = True

    def __init__(self, group: Optional[dist.ProcessGroup] = None, chunks: int = 2):
        self.group = group
        self.chunks = max(1, int(chunks))
        self.available = dist.is_initialized() and dist.get_world_size(group) > 1

    def reduce_async(self, tensors):
        """Async all-reduces of the (up to four) gradient ranges of a chunk -> list of work handles."""
        return [dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True) for t in tensors if t.numel() > 0]


class ViewShardedDecoder(torch.nn.Module):
    """Wraps a decoder (``DecoderSplattingCUDA`` or anything with its ``forward`` signature): every rank
    renders its contiguous slice of the target views of every scene; per-Gaussian gradients are summed
    across ranks in the backward.  ``gather=True`` returns the full ``[B, V, ...]`` frames on every rank
    (inference); otherwise the local slice (training: the loss is computed on the local views)."""

    def __init__(self, decoder: torch.nn.Module, group: Optional[dist.ProcessGroup] = None, gather: bool = False,
                 fused_reduce: bool = False, overlap_reduce: bool = False, nvls_reduce: bool = False, scatter_grads: bool = False,
                 pieces: int = 4, interleave: bool = False):
        super().__init__()
        self.decoder = decoder
        self.group = group
        self.gather = gather
        self.interleave = interleave  # rank r renders views r, r + world, ... instead of a contiguous slice (training only)
        if interleave and gather:
            raise ValueError("gather=True reassembles contiguous slices; use interleave for training (local views only)")
        # scatter_grads: REDUCE-SCATTER instead of all-reduce -- every rank gets the summed gradient of its own Gaussian range
        # (range_bounds) and zeros elsewhere, pulled out of the NVSwitch in pieces under the projection backward
        if scatter_grads and dist.is_initialized() and dist.get_world_size(group) > 1 and hasattr(decoder, "grad_reducer"):
            self.reducer = RangeScatterReducer(group, pieces)
            decoder.grad_reducer = self.reducer
            return
        # nvls_reduce: gradients land in symmetric memory and are summed in place by the library's NVLS kernel
        if nvls_reduce and dist.is_initialized() and dist.get_world_size(group) > 1 and hasattr(decoder, "grad_reducer"):
            self.reducer = NvlsAllReducer(group)
            decoder.grad_reducer = self.reducer
            return
        # overlap_reduce: chunked projection backward with one async NCCL all-reduce per chunk (ChunkedAllReducer)
        if overlap_reduce and not fused_reduce and dist.is_initialized() and dist.get_world_size(group) > 1 and hasattr(decoder, "grad_reducer"):
            self.reducer = ChunkedAllReducer(group)
            decoder.grad_reducer = self.reducer
            return
        # fused_reduce: the backward kernel reduces across ranks itself (NVLS multimem), no all-reduce afterwards
        self.reducer = None
        if fused_reduce and dist.is_initialized() and dist.get_world_size(group) > 1 and hasattr(decoder, "grad_reducer"):
            self.reducer = FusedGradReducer(group)
            decoder.grad_reducer = self.reducer

    def forward(self, gaussians: Gaussians, extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor,
                image_shape: tuple[int, int], depth_mode=None) -> DecoderOutput:
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        V = extrinsics.shape[1]
        _require_views(V, world)
        sl = [shard_views(t, world, rank, interleave=self.interleave) for t in (extrinsics, intrinsics, near, far)]
        fused = self.reducer is not None and self.reducer.available
        g = sync_gaussian_grads(gaussians, self.group) if (torch.is_grad_enabled() and not fused) else gaussians
        out = self.decoder.forward(g, *sl, image_shape, depth_mode=depth_mode)
        if self.gather and world > 1:
            color = all_gather_views(out.color, V, group=self.group)
            depth = None if out.depth is None else all_gather_views(out.depth, V, group=self.group)
            return DecoderOutput(color, depth)
        return out


# ---------------------------------------------------------------------------------------------------------------------
# Gaussians sharded by RANGE: rank r holds (and its share of the encoder produced) the Gaussians
# [range_bounds(N, world, r)) of every scene -- the Gaussian order is (context view, y, x), so a range is a set of context
# views.  Forward: the ranges are all-gathered over NVLink into the full [B, N, ...] tensors every rank renders its views
# from.  Backward: REDUCE-SCATTER -- every rank needs the summed gradient of its own range only, which halves the bytes of
# the all-reduce, and with RangeScatterReducer the projection backward runs in pieces that each cover the j-th part of
# EVERY rank's range, so that all ranks pull their share of piece j out of the switch (multimem.ld_reduce) on a side
# stream while piece j + 1 is being computed.
def range_bounds(num_gaussians: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous split in units of the kernels' 256-Gaussian chunks (the last rank's range may be short or empty)."""
    blocks = (num_gaussians + 255) // 256
    per = (blocks + world_size - 1) // world_size
    return min(num_gaussians, rank * per * 256), min(num_gaussians, (rank + 1) * per * 256)


class RangeScatterReducer:
    """Reduce-scatter of the per-Gaussian gradients by the library's own NVLS kernel, overlapped with the projection
    backward (see above).  The backward writes its gradients (ordinary stores) into one buffer of torch symmetric memory;
    ``ROTATE`` buffers alternate.  The tensors it hands out have the full shape: this rank's Gaussian range holds the sum over
    all ranks, the rest is zero -- so the sum over the ranks of what the ranks get IS the all-reduced gradient (a replicated
    encoder can backpropagate just that on every rank: its parameter gradients are summed over the ranks anyway), and
    ``_GatherRanges.backward`` returns exactly the range."""

    scatter = True
    ROTATE = 3

    def __init__(self, group: Optional[dist.ProcessGroup] = None, pieces: int = 4, pull: Optional[str] = None):
        # how a rank gets the sum of its share of a piece: "p2p" (default) = ordinary loads from every peer's replica of the
        # symmetric buffer, summed in registers (b200s_p2p_reduce_segments) -- (N-1)/N of the buffer per GPU over NVLink in
        # coalesced requests; "nvls" = multimem.ld_reduce on the multicast address (b200s_nvls_reduce_segments) -- the
        # switch reads EVERY replica, the caller's own included: the whole buffer leaves every GPU per step, in 16-byte
        # requests (N = 2: 0.84 ms of pulls per step against 0.35 ms; tools/dist_timeline.py)
        self.pull = (pull or os.environ.get("B200S_SCATTER_PULL", "p2p")).lower()
        if self.pull not in ("p2p", "nvls"):
            raise ValueError("pull must be 'p2p' or 'nvls'")
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        self.num_pieces = max(1, int(pieces))
        self._bufs = {}
        self._outs = {}
        self._turn = 0
        self._side = None
        self.available = self.world > 1 and torch.cuda.is_available() and self._probe()

    def _probe(self) -> bool:
        """Symmetric memory with an NVLS multicast mapping on this fabric?  (Collective: every rank constructs the reducer
        at the same point.)  Without it the callers fall back to NCCL."""
        try:
            import torch.distributed._symmetric_memory as symm_mem
            t = symm_mem.empty(1024, dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))
            h = symm_mem.rendezvous(t, self.group)
            return bool(h.multicast_ptr) if self.pull == "nvls" else len(h.buffer_ptrs) == self.world
        except Exception:
            return False

    def begin(self, shapes, device):
        import torch.distributed._symmetric_memory as symm_mem
        sizes = [int(torch.Size(s).numel()) for s in shapes]
        self._B, self._N = int(shapes[0][0]), int(shapes[0][1])
        self._widths = [n // (self._B * self._N) for n in sizes]
        offs, o = [], 0
        for n in sizes:
            offs.append(o)
            o += (n + 3) // 4 * 4
        self._offs = offs
        key = (o, device)
        if key not in self._bufs:
            ring = []
            for _ in range(self.ROTATE):
                t = symm_mem.empty(o, dtype=torch.float32, device=device)
                h = symm_mem.rendezvous(t, self.group)
                if self.pull == "nvls" and not h.multicast_ptr:
                    self.available = False
                    raise RuntimeError("no NVLS multicast on this fabric")
                t.zero_()
                ring.append((t, h))
            self._bufs[key] = ring
        if self._side is None:
            # high priority: the pulls and their barriers are small kernels that must get SM slots while the next piece of the
            # projection backward (tens of thousands of CTAs) is queueing for the same SMs
            self._side = torch.cuda.Stream(device, priority=-1)
        self._turn = (self._turn + 1) % self.ROTATE
        self._cur = self._bufs[key][self._turn]
        self._piece = 0
        buf = self._cur[0]
        # results: a local (non-symmetric) buffer per rotation slot, zero outside this rank's range for good -- only the
        # range is ever written, by the pulls.  Summed over the ranks, the tensors handed out are the all-reduced gradient.
        okey = (o, device, self._turn)
        if okey not in self._outs:
            self._outs[okey] = torch.zeros(o, dtype=torch.float32, device=device)
        self._out = self._outs[okey]
        self._views = lambda flat: [flat[a:a + n].view(s) for a, n, s in zip(offs, sizes, shapes)]
        # the side stream's previous pulls (of an older buffer) are long done; order it behind the current stream once
        self._side.wait_stream(torch.cuda.current_stream(device))
        return self._views(buf)

    def pieces(self):
        """(chunk_begin, chunk_count, chunk_stride, chunk_repeat) of every piece of the projection backward."""
        blocks = (self._N + 255) // 256
        per_rank = (blocks + self.world - 1) // self.world
        k = self.num_pieces
        # EVEN pieces.  Measured at N = 2 (tools/dist_timeline.py, both pull kernels): decreasing sizes (the last, un-hidden
        # pull smallest) gain nothing (5.61 against 5.60 ms with the multimem pull, 5.45 against 5.45 ms with peer loads), and
        # more than four pieces lose to the tails of the shorter kernels (8: 5.46-5.60, 16: 5.82 ms against 5.40-5.55 ms)
        per_piece = (per_rank + k - 1) // k
        out = []
        for j in range(k):
            cn = min(per_piece, per_rank - j * per_piece)
            if cn > 0:
                out.append((j * per_piece, cn, per_rank, self.world))
        self._per_rank = per_rank
        return out

    def piece_done(self, piece):
        """The piece's kernel has been enqueued on the current stream: on the side stream, wait for it, meet the other
        ranks, pull this rank's share of the piece."""
        from . import _lib
        import ctypes as C
        buf, h = self._cur
        dev = buf.device
        c0, cn, _, _ = piece
        g0 = min(self._N, (self.rank * self._per_rank + c0) * 256)
        g1 = min(self._N, (self.rank * self._per_rank + c0 + cn) * 256)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self._side):
            self._side.wait_event(ev)
            h.barrier(channel=2)  # every rank's piece is complete in its replica
            segs = []
            for b in range(self._B):
                for off, w in zip(self._offs, self._widths):
                    a0, a1 = off + (b * self._N + g0) * w, off + (b * self._N + g1) * w
                    a1 = (a1 + 3) // 4 * 4  # the ragged end of a tensor runs into its alignment padding
                    if a1 > a0:
                        segs.append((a0, a1 - a0))
            L = _lib.load()
            peers = (C.c_void_p * self.world)(*[int(p) for p in h.buffer_ptrs]) if self.pull == "p2p" else None
            for i in range(0, len(segs), 16):
                part = segs[i:i + 16]
                so = (C.c_ulonglong * len(part))(*[p[0] for p in part])
                sn = (C.c_ulonglong * len(part))(*[p[1] for p in part])
                if peers is not None:
                    _lib.check(L.b200s_p2p_reduce_segments(peers, self.world, self.rank, self._out.data_ptr(), so, sn, len(part),
                                                           self._side.cuda_stream), "b200s_p2p_reduce_segments")
                else:
                    _lib.check(L.b200s_nvls_reduce_segments(h.multicast_ptr, self._out.data_ptr(), so, sn, len(part), self._side.cuda_stream),
                               "b200s_nvls_reduce_segments")

    def end(self):
        """-> the gradient tensors: this rank's Gaussian range summed over all ranks, zero elsewhere."""
        buf, _ = self._cur
        torch.cuda.current_stream(buf.device).wait_stream(self._side)
        return self._views(self._out)


class _GatherRanges(torch.autograd.Function):
    """forward: the ranks' Gaussian ranges all-gathered into the full tensors; backward: this rank's range of the gradient
    of the full tensors, summed over the ranks (by the rasterizer's RangeScatterReducer when ``kernel_reduced``, else by a
    reduce-scatter / all-reduce here)."""

    @staticmethod
    def forward(ctx, group, num_gaussians, kernel_reduced, *local):
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        N = int(num_gaussians)
        bounds = [range_bounds(N, world, r) for r in range(world)]
        ctx.group, ctx.bounds, ctx.rank, ctx.kernel_reduced = group, bounds, rank, kernel_reduced
        lo, hi = bounds[rank]
        even = all(b - a == hi - lo for a, b in bounds)
        outs = []
        for t in local:
            B = t.shape[0]
            if t.shape[1] != hi - lo:
                raise ValueError(f"rank {rank} must pass its Gaussian range [{lo}, {hi}) of every tensor, got {tuple(t.shape)}")
            full = t.new_empty((B, N) + tuple(t.shape[2:]))
            if even and B == 1:
                dist.all_gather_into_tensor(full.view(-1), t.contiguous().view(-1), group=group)
            else:
                nmax = max(b - a for a, b in bounds)
                pad = t.new_zeros((B, nmax) + tuple(t.shape[2:]))
                pad[:, :hi - lo] = t
                parts = [torch.empty_like(pad) for _ in range(world)]
                dist.all_gather(parts, pad, group=group)
                for p_, (a, b) in zip(parts, bounds):
                    full[:, a:b] = p_[:, :b - a]
            outs.append(full)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        lo, hi = ctx.bounds[ctx.rank]
        if ctx.kernel_reduced:
            return (None, None, None, *[g[:, lo:hi] for g in grads])
        world = len(ctx.bounds)
        even = all(b - a == hi - lo for a, b in ctx.bounds)
        out = []
        for g in grads:
            g = g.contiguous()
            if even and g.shape[0] == 1 and g.is_cuda:
                mine = g.new_empty((1, hi - lo) + tuple(g.shape[2:]))
                dist.reduce_scatter_tensor(mine.view(-1), g.view(-1), op=dist.ReduceOp.SUM, group=ctx.group)
                out.append(mine)
            else:
                dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
                out.append(g[:, lo:hi])
        return (None, None, None, *out)


class RangeShardedDecoder(torch.nn.Module):
    """Every rank passes ITS range of the Gaussians (``range_bounds``) and the cameras of ALL target views; it renders its
    contiguous slice of the views from the all-gathered Gaussians and gets back the gradient of its own range, summed
    over all ranks' views.  ``num_gaussians`` is N, the Gaussians per scene over all ranks."""

    def __init__(self, decoder: torch.nn.Module, group: Optional[dist.ProcessGroup] = None, pieces: int = 4, kernel_reduce: bool = True,
                 interleave: bool = False):
        super().__init__()
        self.decoder = decoder
        self.group = group
        self.interleave = interleave
        self.reducer = None
        if kernel_reduce and dist.is_initialized() and dist.get_world_size(group) > 1 and torch.cuda.is_available() and hasattr(decoder, "grad_reducer"):
            self.reducer = RangeScatterReducer(group, pieces)

    def forward(self, local: Gaussians, num_gaussians: int, extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor,
                image_shape: tuple[int, int], depth_mode=None) -> DecoderOutput:
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        if world == 1:
            return self.decoder.forward(local, extrinsics, intrinsics, near, far, image_shape, depth_mode=depth_mode)
        _require_views(extrinsics.shape[1], world)
        # the NVLS kernel moves 16-byte vectors: every scene's tensors must start on one
        aligned = local.means.shape[0] == 1 or num_gaussians % 4 == 0
        by_kernel = self.reducer is not None and self.reducer.available and torch.is_grad_enabled() and aligned
        if hasattr(self.decoder, "grad_reducer"):
            self.decoder.grad_reducer = self.reducer if by_kernel else None
        m, c, h, o = _GatherRanges.apply(self.group, num_gaussians, by_kernel, local.means, local.covariances, local.harmonics, local.opacities)
        sl = [shard_views(t, world, rank, interleave=self.interleave) for t in (extrinsics, intrinsics, near, far)]
        return self.decoder.forward(Gaussians(m, c, h, o), *sl, image_shape, depth_mode=depth_mode)


def render_sharded(render_fn: Callable[..., Sequence[Tensor]], gaussians: Gaussians, cameras: Sequence[Tensor], *args,
                   group: Optional[dist.ProcessGroup] = None, **kw):
    """Functional form used by the tests: ``render_fn(gaussians, *camera_slices, *args, **kw)`` on this
    rank's view slice with gradient synchronisation."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    sl = [shard_views(t, world, rank) for t in cameras]
    return render_fn(sync_gaussian_grads(gaussians, group), *sl, *args, **kw)
