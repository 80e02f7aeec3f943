"""Multi-view differentiable rasterization: torch.autograd.Function over the C-ABI library.

One call renders ALL (scene, camera) views of a batch from the un-replicated ``[B,N,...]`` Gaussian
tensors.  It replaces, in one go, what the reference does per view in a Python loop
(src/model/decoder/cuda_splatting.py:90-125): GaussianRasterizer(settings)(means3D, means2D, shs |
colors_precomp, opacities, cov3D_precomp), including the autograd of the per-view replication
(decoder_splatting_cuda.py:53-56) and the second, depth-as-colour render
(cuda_splatting.py:250-263).

PyTorch here is plumbing only: device memory (caching allocator), the current stream, autograd.

Threading: the call pattern of the reference is one Python thread per process driving the default stream, plus the
autograd engine's thread for the backward.  The per-(device, stream) caches below (scratch buffer,
workspace pool, capacity hints, host status ring) follow it: calls on different streams or devices are independent,
two Python threads rendering on the SAME stream at the same time are not supported.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib

_DEPTH_MODES = {None: _lib.DEPTH_NONE, "depth": _lib.DEPTH_Z, "relative_disparity": _lib.DEPTH_Z,
                "disparity": _lib.DEPTH_DISPARITY, "log": _lib.DEPTH_LOG}


@dataclass
class ViewPack:
    """Per-view camera block (all fp32 CUDA tensors, VV = number of views of the call)."""
    scene_index: torch.Tensor          # [VV] int32
    viewmatrix: torch.Tensor           # [VV,4,4] transposed storage (cuda_splatting.py:85)
    projmatrix: torch.Tensor           # [VV,4,4] transposed storage (cuda_splatting.py:86)
    campos: torch.Tensor               # [VV,3]
    tanfov: torch.Tensor               # [VV,2]
    background: torch.Tensor           # [VV,3]
    height: int
    width: int
    scale: Optional[torch.Tensor] = None         # [VV,2] (s, s^2) or None
    depth_mode: Optional[str] = None
    depth_affine: Optional[torch.Tensor] = None  # [VV,4]
    depth_clamp: Optional[torch.Tensor] = None   # [VV,2]
    grad_reducer: Optional[object] = None        # dist.FusedGradReducer: sum the gradients over ranks inside the kernel
    mse_target: Optional[torch.Tensor] = None    # [VV,3,H,W] ground truth: fuse weight * mean((color - target)^2) into the forward
    mse_weight: float = 1.0
    mse_l1: bool = False                         # mean absolute error instead (l1_loss=True of loss_mse.py:41-43)
    mse_count: Optional[int] = None              # colour values the mean runs over (default: those of this call; view-sharded
                                                 # callers pass the count over all ranks)


@dataclass
class RawScene:
    """The Gaussians as the encoder head's raw output (gaussian_adapter.FusedAdapterDecoder): the adapter runs inside the
    projection kernels.  head [B,Vc,37,H*W] and depth [B,Vc,H*W] are differentiable; image [B,Vc,3,H*W]; camera [B,Vc,56]."""
    views: int
    height: int
    width: int
    scale_min: float
    scale_max: float
    image: torch.Tensor
    camera: torch.Tensor
    cooked_out: Optional[torch.Tensor] = None   # tests: [B,N,40] world-space Gaussians as the kernel built them


@dataclass
class RenderStats:
    num_pairs: int = 0
    num_visible: int = 0
    tested: int = 0
    blended: int = 0
    max_tile_len: int = 0
    pair_capacity: int = 0
    retries: int = 0


# "binned" (default): pairs scattered into their (view, tile) bin, one CTA orders each bin in shared memory;
# "global": one onesweep radix sort over 64-bit (view | tile | depth) keys.  Same lists and ranges, bit for bit.
sort_mode = os.environ.get("B200S_SORT_MODE", "binned")
_SORT_MODES = {"binned": _lib.SORT_BINNED, "global": _lib.SORT_GLOBAL}
_mode_hint: dict = {}   # problem shapes whose bins turned out too long for the BINNED mode (stage A said so): GLOBAL from then on

# "event" (default): the host waits for stage A of every forward (pair count known, overflow repaired transparently);
# "lazy": no host wait once a problem shape has been seen twice (see _Rasterize.forward)
sync_policy = os.environ.get("B200S_SYNC", "event")

_capacity_hint: dict = {}
_max_pairs: dict = {}
_seen: dict = {}
_pending: list = []


class PairCapacityOverflow(RuntimeError):
    """Lazy mode only: a forward that had already returned ran out of pair capacity; its images are invalid."""


def _note_pairs(key, num_pairs: int) -> None:
    _capacity_hint[key] = min(max(int(num_pairs * 1.25) + 4096, 1 << 16), _PAIR_LIMIT)
    _max_pairs[key] = max(_max_pairs.get(key, 0), num_pairs)
    _seen[key] = _seen.get(key, 0) + 1


class _Pending:
    """A lazy-mode forward whose status word has not been looked at yet."""
    __slots__ = ("ev", "slot", "key", "cap", "done")

    def __init__(self, ev, slot, key, cap):
        self.ev, self.slot, self.key, self.cap, self.done = ev, slot, key, cap, False

    def check(self, block: bool) -> bool:
        if self.done:
            return True
        if not block and not self.ev.query():
            return False
        self.ev.synchronize()
        self.done = True
        if self in _pending:
            _pending.remove(self)
        words = _status_ring.words
        num_pairs, flags = int(words[2 * self.slot]), int(words[2 * self.slot + 1])
        if not (flags >> 32):
            raise RuntimeError("stage A did not report its pair count (status word not written)")
        _note_pairs(self.key, num_pairs)
        last_stats.num_pairs = num_pairs
        if flags & 2:
            _mode_hint[self.key] = _lib.SORT_GLOBAL
        if flags & 0xFFFFFFFF:
            _seen[self.key] = 0  # back to the waiting protocol until the shape has settled again
            raise PairCapacityOverflow(
                f"a forward that had already returned needed {num_pairs} (tile, Gaussian) pairs but ran with a capacity of {self.cap}: "
                "its images are invalid. rasterizer.sync_policy = 'event' repairs this transparently at the cost of one host wait per forward")
        return True


def _drain_pending(block: bool) -> None:
    for p in list(_pending):
        p.check(block)
    # a status slot must not be handed out again while its call is still unchecked
    while len(_pending) >= _HostStatusRing.SLOTS - 2:
        _pending[0].check(True)
_scratch_cache: dict = {}
last_stats = RenderStats()
# parity tests set debug_keep to look at the stage outputs (records, sorted keys, ranges) of the last call
debug_keep = False
debug_last: Optional[dict] = None


_PAIR_LIMIT = (1 << 32) - 8193  # list positions are 32-bit


class PairLimitExceeded(RuntimeError):
    """One call would produce 2^32 or more (tile, Gaussian) pairs -- or need more workspace than ``max_workspace_bytes`` --;
    render fewer views per call (cuda_splatting.render_views splits the views and retries on its own)."""


# Bound on the two workspaces of ONE call (bytes).  The pair buffers take 24 bytes per (tile, Gaussian) pair; a stress scene
# (BASELINE config 5: 3 M Gaussians with scales up to 0.5, 1.3 * 10^9 pairs per 512x960 view) would otherwise take whatever
# the pair count asks for (110 GB for two views in round 1).  A call whose MEASURED pair count needs more than this and has
# more than one view raises PairLimitExceeded, so that the views are rendered in smaller groups; a single view is always
# rendered, whatever it needs.
max_workspace_bytes = int(float(os.environ.get("B200S_MAX_WORKSPACE_GB", "32")) * 1e9)
_BYTES_PER_PAIR = 24

# Rematerialisation.  A call that had to be split to stay inside max_workspace_bytes would otherwise keep the forward->backward
# workspace of EVERY part alive until the backward (sorted lists: 4 bytes per pair -- 5 GB per view of the stress config),
# so the bound on one call would not bound the step.  While ``remat_depth`` > 0 (cuda_splatting.render_views raises it
# around the parts of a split call) a differentiable forward gives its workspace back at once and the backward rebuilds it:
# binning, sort and compositing forward run again from the same inputs at the same capacity -- deterministic kernels, the
# lists, final_T and n_contrib come out bit for bit as they were -- and then the backward proper.  Peak memory of the step
# = one part's workspaces; cost = one more forward per part.
remat_depth = 0
remat_count = 0   # backward passes that rebuilt their workspace (tests, diagnostics)


class _HostStatusRing:
    """64 slots of 16 bytes of mapped pinned host memory that stage A writes the pair count into."""
    SLOTS = 64

    def __init__(self):
        L = _lib.load()
        self.base = L.b200s_host_alloc(self.SLOTS * 16)
        if not self.base:
            raise RuntimeError("b200s_host_alloc failed")
        self.words = (C.c_uint64 * (self.SLOTS * 2)).from_address(self.base)
        self.next = 0

    def take(self):
        i = self.next
        self.next = (i + 1) % self.SLOTS
        self.words[2 * i] = 0
        self.words[2 * i + 1] = 0
        return i, self.base + 16 * i


_status_ring: Optional[_HostStatusRing] = None


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the rasterizer has no CPU path")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


def _with_slack(nbytes: int, frac: float) -> int:
    """Buffers are allocated a little larger than asked so that a slightly bigger next call reuses them -- by a fraction for
    ordinary sizes, by at most 256 MB for the multi-GB workspaces of a stress scene (max_workspace_bytes bounds what is
    ASKED for; the slack must not add gigabytes on top)."""
    return nbytes + min(int(nbytes * frac), 256 << 20) + 4096


def _scratch(device, nbytes: int) -> torch.Tensor:
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    buf = _scratch_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = None
        _scratch_cache.pop(key, None)
        buf = torch.empty(_with_slack(nbytes, 0.10), dtype=torch.uint8, device=device)
        _scratch_cache[key] = buf
    return buf


class _SavedLease:
    """A forward->backward workspace taken from a small per-(device, stream) pool.  The caching allocator
    handles these ~1 GB blocks badly when the caller keeps results of earlier steps alive (fresh
    cudaMalloc / cudaFree every few steps, 10-200 ms each on a B200 box), so they are recycled here: the
    lease returns its buffer to the pool when the autograd context that holds it dies."""
    __slots__ = ("tensor", "key")

    def __init__(self, device, nbytes: int):
        self.key = (device, torch.cuda.current_stream(device).cuda_stream)
        pool = _saved_pool.setdefault(self.key, [])
        best = None
        for i, t in enumerate(pool):
            if t.numel() >= nbytes and (best is None or t.numel() < pool[best].numel()):
                best = i
        if best is not None and pool[best].numel() <= 2 * nbytes + (1 << 20):
            self.tensor = pool.pop(best)
        else:
            if nbytes > (256 << 20):
                # a large workspace that no pooled buffer serves: the pooled ones are leftovers of smaller attempts of the same
                # call (speculative capacities, the parts of a split) -- gigabytes that would sit next to the new buffer
                pool.clear()
            self.tensor = torch.empty(_with_slack(nbytes, 0.05), dtype=torch.uint8, device=device)

    def __del__(self):
        try:
            pool = _saved_pool.setdefault(self.key, [])
            if len(pool) < 4:
                pool.append(self.tensor)
        except Exception:
            pass


_saved_pool: dict = {}


def release_scratch() -> None:
    _scratch_cache.clear()
    _saved_pool.clear()


def _build_structs(means, covs, colors, opacities, use_sh, sh_degree, sh_layout, vp: ViewPack, raw: Optional[RawScene] = None):
    if raw is not None:  # (means, covs) carry (head, depth)
        head, depth = means, covs
        B, N = head.shape[0], raw.views * raw.height * raw.width
        sc = _lib.Scene()
        sc.num_scenes, sc.num_gaussians, sc.sh_degree, sc.sh_coeffs = B, N, 2, 9
        sc.cov_layout, sc.sh_layout = _lib.COV_3X3, _lib.SH_CHANNEL_MAJOR
        sc.raw_head, sc.raw_depth, sc.raw_image, sc.raw_camera = _ptr(head), _ptr(depth), _ptr(raw.image), _ptr(raw.camera)
        sc.raw_views, sc.raw_h, sc.raw_w = raw.views, raw.height, raw.width
        sc.raw_scale_min, sc.raw_scale_max = float(raw.scale_min), float(raw.scale_max)
        sc.raw_cooked_out = _ptr(raw.cooked_out)
        return sc, _build_views(vp)
    B, N = means.shape[0], means.shape[1]
    sc = _lib.Scene()
    sc.num_scenes, sc.num_gaussians = B, N
    sc.cov_layout = _lib.COV_UPPER6 if covs.shape[-1] == 6 and covs.dim() == 3 else _lib.COV_3X3
    sc.means, sc.covariances, sc.opacities = _ptr(means), _ptr(covs), _ptr(opacities)
    if use_sh:
        d_sh = colors.shape[-1] if sh_layout == _lib.SH_CHANNEL_MAJOR else colors.shape[-2]
        sc.sh_degree, sc.sh_coeffs, sc.sh_layout = sh_degree, d_sh, sh_layout
        sc.harmonics, sc.colors_precomp = _ptr(colors), None
    else:
        sc.sh_degree, sc.sh_coeffs, sc.sh_layout = 0, 1, _lib.SH_CHANNEL_MAJOR
        sc.harmonics, sc.colors_precomp = None, _ptr(colors)
    return sc, _build_views(vp)


def _build_views(vp: ViewPack):
    vw = _lib.Views()
    vw.num_views, vw.height, vw.width = vp.scene_index.shape[0], vp.height, vp.width
    vw.depth_mode = _DEPTH_MODES[vp.depth_mode]
    vw.scene_index, vw.viewmatrix, vw.projmatrix = _ptr(vp.scene_index), _ptr(vp.viewmatrix), _ptr(vp.projmatrix)
    vw.campos, vw.tanfov, vw.background = _ptr(vp.campos), _ptr(vp.tanfov), _ptr(vp.background)
    vw.scale, vw.depth_affine, vw.depth_clamp = _ptr(vp.scale), _ptr(vp.depth_affine), _ptr(vp.depth_clamp)
    return vw


class _Rasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means, covs, colors, opacities, means2d, vp: ViewPack, use_sh: bool, sh_degree: int, sh_layout: int,
                want_radii: bool, count_work: bool, raw: Optional[RawScene] = None):
        # the library launches on the CURRENT device: make it the tensors' device for the duration of the call
        with torch.cuda.device(means.device):
            return _Rasterize._forward(ctx, means, covs, colors, opacities, means2d, vp, use_sh, sh_degree, sh_layout, want_radii, count_work, raw)

    @staticmethod
    def backward(ctx, g_color, g_depth, _g_radii, g_loss=None, _g_sse=None):
        with torch.cuda.device(ctx.saved_tensors[0].device):
            color_scale = None
            if g_loss is not None and ctx.mse_grad is not None:
                # the fused loss: its dL/dcolor was written by the forward epilogue, and the upstream gradient of the loss (a
                # device scalar) is applied by the backward kernel as it loads the pixel -- no pass over the images, no sync
                if g_color is None:
                    g_color, color_scale = ctx.mse_grad, g_loss.to(torch.float32).reshape(1).contiguous()
                else:  # the caller ALSO used the colour
                    g_color = g_color + ctx.mse_grad * g_loss
            ctx.color_scale = color_scale
            return _Rasterize._backward(ctx, g_color, g_depth, _g_radii)

    @staticmethod
    def _forward(ctx, means, covs, colors, opacities, means2d, vp: ViewPack, use_sh: bool, sh_degree: int, sh_layout: int,
                 want_radii: bool, count_work: bool, raw: Optional[RawScene] = None):
        L = _lib.load()
        dev = means.device
        B, N = (means.shape[0], means.shape[1]) if raw is None else (means.shape[0], raw.views * raw.height * raw.width)
        ctx.raw = raw
        VV, H, W = vp.scene_index.shape[0], vp.height, vp.width
        stream = torch.cuda.current_stream(dev).cuda_stream
        sc, vw = _build_structs(means, covs, colors, opacities, use_sh, sh_degree, sh_layout, vp, raw)

        color = torch.empty((VV, 3, H, W), dtype=torch.float32, device=dev)
        depth = torch.empty((VV, H, W), dtype=torch.float32, device=dev) if vp.depth_mode is not None else None
        radii = torch.empty((VV, N), dtype=torch.int32, device=dev) if want_radii else None
        global _status_ring
        if _status_ring is None:
            _status_ring = _HostStatusRing()
        slot, slot_ptr = _status_ring.take()
        out = _lib.Out(_ptr(color), _ptr(depth), _ptr(radii), 1 if count_work else 0, slot_ptr)
        mse_grad = mse_partials = None
        if vp.mse_target is not None:
            tgt = _f32c(vp.mse_target, "mse_target")
            if tgt.shape != (VV, 3, H, W):
                raise ValueError(f"mse_target must be [{VV},3,{H},{W}], got {tuple(tgt.shape)}")
            tiles = ((H + 15) // 16) * ((W + 15) // 16)
            mse_grad = torch.empty_like(color)
            mse_partials = torch.empty((VV, tiles, 2), dtype=torch.float32, device=dev)
            out.mse_target, out.mse_grad, out.mse_partials = _ptr(tgt), _ptr(mse_grad), _ptr(mse_partials)
            out.mse_scale = float(vp.mse_weight) / float(vp.mse_count or VV * 3 * H * W)
            out.mse_l1 = 1 if vp.mse_l1 else 0

        key = (dev.index, B, N, VV, H, W)
        _drain_pending(block=False)
        # "lazy": once this problem shape has been seen, do not wait for stage A at all -- run at twice the largest pair count
        # seen so far and look at the status word later (next call, or this call's backward).  An overflow found that late
        # cannot be repaired (the images were already handed out), so it raises; "event" (default) waits for stage A -- and
        # for whatever the stream still had queued before it -- and re-runs the call transparently.
        lazy = sync_policy == "lazy" and _seen.get(key, 0) >= 2 and not (debug_keep or count_work)
        if lazy:
            cap = min(max(2 * _max_pairs[key] + 4096, 1 << 16), _PAIR_LIMIT)
        else:
            cap = _capacity_hint.get(key) or max(4 * N * VV, 1 << 16)
            cap = min(cap, _PAIR_LIMIT)
            # the first, speculative attempt stays inside the budget too (a single view: unless it is known to need more)
            budget_pairs = (max_workspace_bytes - VV * (N * 72 + H * W * 8)) // _BYTES_PER_PAIR
            if VV == 1:
                budget_pairs = max(budget_pairs, _max_pairs.get(key, 0) + 4096)
            cap = max(min(cap, budget_pairs), 1 << 16)
        words = _status_ring.words
        retries = 0
        while True:
            plan = _lib.plan(B, N, VV, H, W, cap, _mode_hint.get(key, _SORT_MODES[sort_mode]))
            lease = _SavedLease(dev, plan.saved_bytes)
            saved = lease.tensor
            scratch = _scratch(dev, plan.scratch_bytes)
            _lib.check(L.b200s_forward_bin(C.byref(sc), C.byref(vw), C.byref(plan), saved.data_ptr(), scratch.data_ptr(),
                                           C.byref(out), stream), "b200s_forward_bin")
            # stage A wrote the pair count straight into mapped host memory; the event marks its end, and the
            # GPU already sorts and composites (speculatively, at this capacity) while the host looks at it
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            _lib.check(L.b200s_forward_render(C.byref(sc), C.byref(vw), C.byref(plan), saved.data_ptr(), scratch.data_ptr(),
                                              C.byref(out), stream), "b200s_forward_render")
            if lazy:
                pending = _Pending(ev, slot, key, cap)
                _pending.append(pending)
                num_pairs = -1
                break
            pending = None
            ev.synchronize()
            num_pairs = int(words[2 * slot])
            flags = int(words[2 * slot + 1])
            if not (flags >> 32):
                raise RuntimeError("stage A did not report its pair count (status word not written)")
            overflow = flags & 0xFFFFFFFF
            if not overflow:
                break
            words[2 * slot] = 0
            words[2 * slot + 1] = 0
            lease = saved = None  # the failed attempt's workspace goes back before the next one (or the parts of a split) allocate
            if overflow & 2:  # bins too long for the BINNED mode: this shape renders in GLOBAL mode from now on
                _mode_hint[key] = _lib.SORT_GLOBAL
            if num_pairs >= _PAIR_LIMIT:
                raise PairLimitExceeded(f"{num_pairs} (tile, Gaussian) pairs in one call exceed the 2^32 limit; render fewer views per call")
            if VV > 1 and num_pairs * _BYTES_PER_PAIR + VV * (N * 72 + H * W * 8) > max_workspace_bytes:
                raise PairLimitExceeded(f"{num_pairs} (tile, Gaussian) pairs for {VV} views need {num_pairs * _BYTES_PER_PAIR / 1e9:.1f} GB of pair buffers, "
                                        f"more than max_workspace_bytes = {max_workspace_bytes / 1e9:.0f} GB; render fewer views per call")
            cap = min(int(num_pairs * 1.25) + 4096, _PAIR_LIMIT)
            # the measured count is exact: the 25 % head-room shrinks to whatever the budget leaves (a single view is rendered
            # whatever it needs, but asks for no more than it needs once the budget is exceeded)
            cap = min(cap, max((max_workspace_bytes - VV * (N * 72 + H * W * 8)) // _BYTES_PER_PAIR, num_pairs + 4096))
            retries += 1
        if not lazy:
            _note_pairs(key, num_pairs)

        st = last_stats
        st.num_pairs, st.pair_capacity, st.retries = num_pairs, cap, retries
        if count_work:
            torch.cuda.current_stream(dev).synchronize()
            h = saved[:64].view(torch.int64).cpu()
            st.num_visible = (int(h[1]) >> 32) & 0xFFFFFFFF
            st.tested, st.blended, st.max_tile_len = int(h[2]), int(h[3]), int(h[4]) & 0xFFFFFFFF

        if debug_keep:
            global debug_last
            debug_last = dict(plan=plan, saved=saved, scratch=scratch, num_pairs=num_pairs, N=N, VV=VV, H=H, W=W)
        ctx.save_for_backward(means, covs, colors, opacities)
        ctx.remat = None
        if remat_depth > 0 and not lazy and not debug_keep and any(ctx.needs_input_grad[:4]):
            ctx.remat = True
            lease = None   # back to the pool now; the backward takes a fresh one and rebuilds its contents
        ctx.b200 = (vp, use_sh, sh_degree, sh_layout, plan, lease, means2d is not None, pending)
        ctx.set_materialize_grads(False)
        ctx.mse_grad = mse_grad
        if mse_partials is not None:
            sums = mse_partials.sum(dim=1)                       # [VV, 2], fixed order
            loss = sums[:, 0].sum() * out.mse_scale              # weight * mean(...)
            sse_clipped = sums[:, 1]                             # per view, for the PSNR
        else:
            loss = color.new_empty(0)
            sse_clipped = color.new_empty(0)
        outs = [color]
        if depth is not None:
            outs.append(depth)
        else:
            outs.append(color.new_empty(0))
            ctx.mark_non_differentiable(outs[-1])
        if radii is None:
            radii = torch.empty(0, dtype=torch.int32, device=dev)
        ctx.mark_non_differentiable(radii, sse_clipped)
        if mse_partials is None:
            ctx.mark_non_differentiable(loss)
        return outs[0], outs[1], radii, loss, sse_clipped

    @staticmethod
    def _backward(ctx, g_color, g_depth, _g_radii):
        L = _lib.load()
        means, covs, colors, opacities = ctx.saved_tensors
        vp, use_sh, sh_degree, sh_layout, plan, lease, want_m2d, pending = ctx.b200
        if pending is not None:
            pending.check(block=True)  # lazy mode: the forward's status word, normally long since written
        dev = means.device
        raw = ctx.raw
        VV, N = vp.scene_index.shape[0], (means.shape[1] if raw is None else raw.views * raw.height * raw.width)
        stream = torch.cuda.current_stream(dev).cuda_stream
        sc, vw = _build_structs(means, covs, colors, opacities, use_sh, sh_degree, sh_layout, vp, raw)
        if ctx.remat is not None:
            global remat_count
            remat_count += 1
            # rebuild the forward->backward workspace: same inputs, same capacity, same sort mode -> same lists, final_T,
            # n_contrib (the images go to a throw-away buffer; the fused loss is not evaluated again)
            lease = _SavedLease(dev, plan.saved_bytes)
            tmp_color = torch.empty((VV, 3, vp.height, vp.width), dtype=torch.float32, device=dev)
            tmp_depth = torch.empty((VV, vp.height, vp.width), dtype=torch.float32, device=dev) if vp.depth_mode is not None else None
            _slot, slot_ptr = _status_ring.take()
            fout = _lib.Out(_ptr(tmp_color), _ptr(tmp_depth), None, 0, slot_ptr)
            scratch = _scratch(dev, plan.scratch_bytes)
            for fn, what in ((L.b200s_forward_bin, "b200s_forward_bin (remat)"), (L.b200s_forward_render, "b200s_forward_render (remat)")):
                _lib.check(fn(C.byref(sc), C.byref(vw), C.byref(plan), lease.tensor.data_ptr(), scratch.data_ptr(), C.byref(fout), stream), what)
            del tmp_color, tmp_depth
        saved = lease.tensor
        g_color = (torch.zeros((VV, 3, vp.height, vp.width), dtype=torch.float32, device=dev) if g_color is None
                   else g_color.to(torch.float32).contiguous())
        if vp.depth_mode is not None:
            g_depth = (torch.zeros((VV, vp.height, vp.width), dtype=torch.float32, device=dev) if g_depth is None
                       else g_depth.to(torch.float32).contiguous())
        else:
            g_depth = None
        d_m2d = torch.empty((VV, N, 3), dtype=torch.float32, device=dev) if want_m2d else None
        scratch = _scratch(dev, plan.scratch_bytes)
        gout = _lib.GradOut(_ptr(g_color), _ptr(g_depth), _ptr(getattr(ctx, "color_scale", None)))
        reducer = vp.grad_reducer if (vp.grad_reducer is not None and getattr(vp.grad_reducer, "available", False)) else None
        out = _lib.Out(None, None, None, 0)

        def call(gin):
            _lib.check(L.b200s_backward(C.byref(sc), C.byref(vw), C.byref(plan), saved.data_ptr(), scratch.data_ptr(), C.byref(out),
                                        C.byref(gout), C.byref(gin), stream), "b200s_backward")

        if raw is not None:
            # raw scene: gradients w.r.t. the head's channel planes and the depth, straight out of the projection backward
            d_head, d_depth = torch.empty_like(means), torch.empty_like(covs)
            gin = _lib.GradIn(None, None, None, None, None, _ptr(d_m2d), 0, 0, 0, 0, 0, 0, _ptr(d_head), _ptr(d_depth))
            call(gin)
            return d_head, d_depth, None, None, d_m2d, None, None, None, None, None, None, None
        if reducer is not None and getattr(reducer, "scatter", False):
            # reduce-scatter by Gaussian range: outputs are views of a symmetric-memory buffer; the projection backward runs
            # in pieces that each cover a part of EVERY rank's range, and the reducer pulls this rank's share of a finished
            # piece out of the NVSwitch on a side stream while the next piece is computed
            d_means, d_covs, d_colors, d_op = reducer.begin([means.shape, covs.shape, colors.shape, opacities.shape], dev)
            mk = lambda stages, pc: _lib.GradIn(_ptr(d_means), _ptr(d_covs), _ptr(d_colors) if use_sh else None,
                                                None if use_sh else _ptr(d_colors), _ptr(d_op), _ptr(d_m2d), 0, stages, *pc)
            call(mk(1, (0, 0, 0, 0)))
            for pc in reducer.pieces():
                call(mk(2, pc))
                reducer.piece_done(pc)
            d_means, d_covs, d_colors, d_op = reducer.end()
        elif reducer is not None and getattr(reducer, "in_place", False):
            # outputs are views of a symmetric-memory buffer; the reducer sums it over the ranks in place afterwards
            d_means, d_covs, d_colors, d_op = reducer.begin([means.shape, covs.shape, colors.shape, opacities.shape], dev)
            call(_lib.GradIn(_ptr(d_means), _ptr(d_covs), _ptr(d_colors) if use_sh else None, None if use_sh else _ptr(d_colors),
                             _ptr(d_op), _ptr(d_m2d), 0, 0, 0, 0))
            reducer.end()
        elif reducer is not None and not getattr(reducer, "chunked", False):
            # outputs are NVLS multicast addresses: the kernel ADDS into every rank's (zeroed) replica
            (d_means, d_covs, d_colors, d_op), mc = reducer.begin([means.shape, covs.shape, colors.shape, opacities.shape], dev)
            call(_lib.GradIn(mc[0], mc[1], mc[2] if use_sh else None, None if use_sh else mc[2], mc[3], _ptr(d_m2d), 1, 0, 0, 0))
            reducer.end()
        else:
            d_means, d_covs, d_colors, d_op = _grad_tensors(means, covs, colors, opacities)
            mk = lambda stages, c0, cn: _lib.GradIn(_ptr(d_means), _ptr(d_covs), _ptr(d_colors) if use_sh else None,
                                                    None if use_sh else _ptr(d_colors), _ptr(d_op), _ptr(d_m2d), 0, stages, c0, cn)
            if reducer is None:
                call(mk(0, 0, 0))
            else:
                # compositing backward once, then the projection backward range by range; every finished range
                # goes to an async NCCL all-reduce that runs under the next range's kernel
                call(mk(1, 0, 0))
                total = (N + 255) // 256
                nch = reducer.chunks if means.shape[0] == 1 else 1  # ranges are contiguous only within one scene
                per = (total + nch - 1) // nch
                works = []
                for c0 in range(0, total, per):
                    cn = min(per, total - c0)
                    call(mk(2, c0, cn))
                    g0, g1 = c0 * 256, min(N, (c0 + cn) * 256)
                    parts = [d_means, d_covs, d_colors, d_op] if nch == 1 else [t[:, g0:g1] for t in (d_means, d_covs, d_colors, d_op)]
                    works += reducer.reduce_async(parts)
                for w in works:
                    w.wait()
        return d_means, d_covs, d_colors, d_op, d_m2d, None, None, None, None, None, None, None


def _grad_tensors(*like):
    """Gradient tensors shaped like the inputs, carved back to back out of ONE allocation when every size keeps the next
    one 16-byte aligned (it does whenever N is a multiple of 4): a cross-rank sum of the per-Gaussian gradients
    (dist._SyncGrads) is then a single all-reduce over the flat buffer instead of one per tensor."""
    sizes = [t.numel() for t in like]
    if any(n % 4 for n in sizes[:-1]):
        return tuple(torch.empty_like(t) for t in like)
    flat = torch.empty(sum(sizes), dtype=torch.float32, device=like[0].device)
    out, o = [], 0
    for t, n in zip(like, sizes):
        out.append(flat[o:o + n].view(t.shape))
        o += n
    return tuple(out)


def rasterize(means: torch.Tensor, covariances: torch.Tensor, colors: torch.Tensor, opacities: torch.Tensor, views: ViewPack, *,
              use_sh: bool = True, sh_degree: Optional[int] = None, sh_layout: int = _lib.SH_CHANNEL_MAJOR,
              means2d: Optional[torch.Tensor] = None, want_radii: bool = False, count_work: bool = False):
    """means [B,N,3]; covariances [B,N,3,3] or [B,N,6]; colors = harmonics [B,N,3,d_sh] (channel-major,
    DepthSplat layout) / [B,N,d_sh,3] (coefficient-major, extension layout) when use_sh else [B,N,3];
    opacities [B,N].  Returns (color [VV,3,H,W], depth [VV,H,W] | None, radii [VV,N] | None)."""
    means = _f32c(means, "means"); covariances = _f32c(covariances, "covariances")
    colors = _f32c(colors, "colors"); opacities = _f32c(opacities, "opacities")
    if means.dim() != 3 or means.shape[-1] != 3:
        raise ValueError("means must be [B,N,3]")
    B, N = means.shape[:2]
    if covariances.shape not in ((B, N, 3, 3), (B, N, 6)):
        raise ValueError("covariances must be [B,N,3,3] or [B,N,6]")
    if opacities.shape != (B, N):
        raise ValueError("opacities must be [B,N]")
    if use_sh:
        if colors.dim() != 4 or colors.shape[:2] != (B, N):
            raise ValueError("harmonics must be [B,N,3,d_sh] or [B,N,d_sh,3]")
        d_sh = colors.shape[-1] if sh_layout == _lib.SH_CHANNEL_MAJOR else colors.shape[-2]
        if colors.shape[-2 if sh_layout == _lib.SH_CHANNEL_MAJOR else -1] != 3:
            raise ValueError("harmonics layout does not match sh_layout")
        if sh_degree is None:
            sh_degree = int(round(d_sh ** 0.5)) - 1
    else:
        if colors.shape != (B, N, 3):
            raise ValueError("colors_precomp must be [B,N,3]")
        sh_degree = 0
    if N == 0 or views.scene_index.shape[0] == 0:
        raise ValueError("empty scene or view list")
    color, depth, radii, loss, sse = _Rasterize.apply(means, covariances, colors, opacities, means2d, views, use_sh, sh_degree, sh_layout,
                                                      want_radii, count_work)
    if views.mse_target is not None:
        views.mse_result = (loss, sse)
    return color, (depth if views.depth_mode is not None else None), (radii if want_radii else None)


def rasterize_raw(head: torch.Tensor, depth: torch.Tensor, raw: RawScene, views: ViewPack, *, want_radii: bool = False,
                  count_work: bool = False):
    """The encoder head's raw output rendered directly: head [B,Vc,37,H*W], depth [B,Vc,H*W] (both differentiable), the
    rest in ``raw``.  Returns (color [VV,3,H,W], depth image [VV,H,W] | None, radii [VV,N] | None)."""
    head = _f32c(head, "head"); depth = _f32c(depth, "depth")
    raw.image = _f32c(raw.image, "image"); raw.camera = _f32c(raw.camera, "camera")
    B = head.shape[0]
    hw = raw.height * raw.width
    if head.shape != (B, raw.views, 37, hw) or depth.shape != (B, raw.views, hw) or raw.image.shape != (B, raw.views, 3, hw) or raw.camera.shape != (B, raw.views, 56):
        raise ValueError("raw scene tensors must be head [B,Vc,37,H*W], depth [B,Vc,H*W], image [B,Vc,3,H*W], camera [B,Vc,56]")
    if hw % 256:
        raise ValueError("the fused adapter needs H*W of the context views to be a multiple of 256")
    dummy = head.new_empty(0)
    color, dimg, radii, _, _ = _Rasterize.apply(head, depth, dummy, dummy, None, views, True, 2, _lib.SH_CHANNEL_MAJOR, want_radii, count_work, raw)
    return color, (dimg if views.depth_mode is not None else None), (radii if want_radii else None)
