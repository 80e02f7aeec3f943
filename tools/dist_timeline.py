"""Per-rank timeline of the view-sharded training step (torchrun --nproc-per-node N tools/dist_timeline.py [pieces]):
CUDA events on the main stream (forward end, compositing backward end, every projection-backward piece) and on the
reducer's side stream (barrier passed, pull finished), printed relative to the start of the step -- where does the time
between the last kernel of the backward and the next step go?  Also the step time of the same views with NO reduction
(what the rank's own kernels take)."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from my_depthsplat_b200 import _lib  # noqa: E402
from my_depthsplat_b200 import dist as D  # noqa: E402
from my_depthsplat_b200.decoder_splatting_cuda import DecoderSplattingCUDACfg, get_decoder  # noqa: E402
from my_depthsplat_b200.scenes import make_scene  # noqa: E402
from my_depthsplat_b200.types import Gaussians  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
pieces = int(sys.argv[1]) if len(sys.argv) > 1 else 4
V = 4
sc = make_scene("C2T", v_tgt=V * world).to(dev)
H, W = sc.image_shape
cfg = type("D", (), {"background_color": [0.0, 0.0, 0.0]})()
g = sc.gaussians
grad_color = sc.grad_color[:, rank::world].contiguous()
marks = []


def mark(name, stream=None):
    e = torch.cuda.Event(enable_timing=True)
    e.record(stream if stream is not None else torch.cuda.current_stream(dev))
    marks.append((name, e))


def step(decoder):
    leaves = [t.detach().requires_grad_() for t in (g.means, g.covariances, g.harmonics, g.opacities)]
    mark("step start")
    out = decoder.forward(Gaussians(*leaves), sc.extrinsics, sc.intrinsics, sc.near, sc.far, (H, W), depth_mode=None)
    mark("forward end")
    grads = torch.autograd.grad(out.color, leaves, grad_color)
    mark("backward end (main stream)")
    return grads


def run(decoder, label, steps=6):
    for _ in range(4):
        step(decoder)
    dist.barrier(); torch.cuda.synchronize()
    marks.clear()
    for _ in range(steps):
        step(decoder)
    mark("end")
    torch.cuda.synchronize()
    t0 = marks[0][1]
    total = t0.elapsed_time(marks[-1][1]) / steps
    # the last step in detail
    last = max(i for i, (n, _) in enumerate(marks) if n == "step start")
    base = marks[last][1]
    lines = [f"  {n:44s} {base.elapsed_time(e):8.3f} ms" for n, e in marks[last:]]
    for r in range(world):
        dist.barrier()
        if r == rank:
            print(f"rank {rank} [{label}] {total:.3f} ms / step\n" + "\n".join(lines), flush=True)
    dist.barrier()


# (a) no reduction at all: every rank renders and differentiates its own views
class Own(torch.nn.Module):
    def __init__(self, dec):
        super().__init__(); self.dec = dec

    def forward(self, gs, ext, K, near, far, shape, depth_mode=None):
        sl = [D.shard_views(t, world, rank, interleave=True) for t in (ext, K, near, far)]
        return self.dec.forward(gs, *sl, shape, depth_mode=depth_mode)


run(Own(get_decoder(DecoderSplattingCUDACfg(name="splatting_cuda"), cfg).to(dev)), "own views, no reduction")

# (b) the default: reduce-scatter in pieces, instrumented
dec = D.ViewShardedDecoder(get_decoder(DecoderSplattingCUDACfg(name="splatting_cuda"), cfg).to(dev), scatter_grads=True, pieces=pieces, interleave=True)
red = dec.reducer
orig_piece_done, orig_end, orig_begin = red.piece_done, red.end, red.begin
L = _lib.load()
orig_pull = L.b200s_nvls_reduce_segments
state = {"k": 0}


def pull(mc, local_, so, sn, n, stream):
    mark(f"  side: barrier {state['k']} passed", red._side)
    return orig_pull(mc, local_, so, sn, n, stream)


pull.argtypes = orig_pull.argtypes


def piece_done(piece):
    mark(f"piece {state['k']} kernel end (main)")
    L.b200s_nvls_reduce_segments = pull
    try:
        orig_piece_done(piece)
    finally:
        L.b200s_nvls_reduce_segments = orig_pull
    mark(f"  side: pull {state['k']} done", red._side)
    state["k"] += 1


def begin(shapes, device):
    state["k"] = 0
    r = orig_begin(shapes, device)
    mark("reducer.begin done (main)")
    return r


def end():
    r = orig_end()
    mark("reducer.end: main waited for side")
    return r


red.piece_done, red.end, red.begin = piece_done, end, begin
run(dec, f"reduce-scatter, {pieces} pieces")
dist.destroy_process_group()
