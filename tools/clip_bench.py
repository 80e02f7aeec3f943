"""Clip rendering throughput (BASELINE configs 2 and 3): python tools/clip_bench.py [C3] [frames] [chunk]
or torchrun --nproc-per-node N tools/clip_bench.py ...  Renders the `frames` target views of the config with
my_depthsplat_b200.video.render_clip (views sharded over the ranks, chunks of `chunk`, frames streamed to pinned
host memory) and, at N=1, the reference's schedule over the upstream-style comparator (baseline/) for the same
frames: per-view calls inside decoder.forward per chunk + torch.cat."""
import json
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from my_depthsplat_b200.decoder_splatting_cuda import DecoderSplattingCUDACfg, get_decoder  # noqa: E402
from my_depthsplat_b200.scenes import make_scene  # noqa: E402
from my_depthsplat_b200.types import Gaussians  # noqa: E402
from my_depthsplat_b200.video import render_clip  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 100
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 10
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
sc = make_scene(name, v_tgt=frames).to(dev)
H, W = sc.image_shape
g = Gaussians(sc.gaussians.means, sc.gaussians.covariances, sc.gaussians.harmonics, sc.gaussians.opacities)
dec = get_decoder(DecoderSplattingCUDACfg(name="splatting_cuda"), type("D", (), {"background_color": [0.0, 0.0, 0.0]})()).to(dev)


def timed(fn, reps):
    fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if world > 1:
        t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t)
    return ms


out = {"config": name, "frames": frames, "chunk": chunk, "n_gpus": world, "HxW": f"{H}x{W}", "gaussians": g.means.shape[1]}
ms_dev = timed(lambda: render_clip(dec, g, sc.extrinsics, sc.intrinsics, sc.near, sc.far, (H, W), chunk_size=chunk), 3)
ms_host = timed(lambda: render_clip(dec, g, sc.extrinsics, sc.intrinsics, sc.near, sc.far, (H, W), chunk_size=chunk, to_host=True), 3)
out.update({"clip_ms_device": round(ms_dev, 2), "frames_per_s_device": round(frames / ms_dev * 1e3, 1),
            "clip_ms_to_host": round(ms_host, 2), "frames_per_s_to_host": round(frames / ms_host * 1e3, 1),
            "Mpix_s_to_host": round(frames * H * W / ms_host / 1e3, 1)})
if world == 1:
    from baseline import per_view_glue, upstream_ext

    def reference_schedule():
        color = None
        with torch.no_grad():
            for i in range((frames + chunk - 1) // chunk):
                sl = slice(chunk * i, chunk * (i + 1))
                cur, _ = per_view_glue.decoder_forward(upstream_ext, g, sc.extrinsics[:, sl], sc.intrinsics[:, sl], sc.near[:, sl], sc.far[:, sl],
                                                       (H, W), sc.background, None)
                color = cur if color is None else torch.cat((color, cur), dim=1)
        return color

    ms_ref = timed(reference_schedule, 1)
    out.update({"upstream_style_clip_ms": round(ms_ref, 2), "upstream_style_frames_per_s": round(frames / ms_ref * 1e3, 1),
                "speedup_device": round(ms_ref / ms_dev, 2)})
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
