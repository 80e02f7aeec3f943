"""Throughput of the five BASELINE.json configs on one GPU (python tools/sweep.py [names...]).
Prints one line per config: rendered Mpix/s, ms per call, pairs, stage times."""
import json
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from my_depthsplat_b200 import _lib, rasterizer as R  # noqa: E402
from my_depthsplat_b200.cuda_splatting import render_views  # noqa: E402
from my_depthsplat_b200.scenes import CONFIGS, make_scene  # noqa: E402

RUNS = {  # name -> (config, views per call, depth_mode, backward)
    "C1": ("C1", 4, "depth", True),
    "C2": ("C2", 10, None, False),
    "C2T": ("C2T", 4, None, True),
    "C3": ("C3", 10, None, False),          # 10-view chunk of the 100-frame clip (views are sharded over GPUs)
    "C4": ("C4", 4, None, True),            # 8 scenes on ONE GPU (the 8-GPU run puts one scene per GPU)
    "C5": ("C5", 4, None, True),
}


def run(name, steps=5, warmup=2):
    cfgname, V, depth_mode, backward = RUNS[name]
    t0 = time.time()
    sc = make_scene(cfgname, v_tgt=V).to("cuda")
    g = sc.gaussians
    H, W = sc.image_shape
    B = g.means.shape[0]

    def step():
        leaves = [t.detach().requires_grad_(backward) for t in (g.means, g.covariances, g.harmonics, g.opacities)]
        with torch.set_grad_enabled(backward):
            color, depth = render_views(sc.extrinsics, sc.intrinsics, sc.near, sc.far, (H, W), sc.background, *leaves, depth_mode=depth_mode)
        if backward:
            outs, gouts = [color], [sc.grad_color]
            if depth is not None:
                outs.append(depth); gouts.append(sc.grad_depth)
            torch.autograd.grad(outs, leaves, gouts)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    _lib.profile_enable(True); _lib.profile_read()
    step(); torch.cuda.synchronize()
    st = {k: round(v, 3) for k, v in _lib.profile_read().items() if v > 0.0005 and k != "end"}
    _lib.profile_enable(False)
    pairs = R.last_stats.num_pairs
    # the upstream-style GPU comparator (baseline/) on the same scene: per-view calls, V-fold replication, CUB sort
    base_ms = None
    if name != "C5":  # the comparator's 32-bit pair offsets cannot hold the stress scene
        from baseline import per_view_glue, upstream_ext
        from my_depthsplat_b200.types import Gaussians

        def base_step():
            leaves = [t.detach().requires_grad_(backward) for t in (g.means, g.covariances, g.harmonics, g.opacities)]
            with torch.set_grad_enabled(backward):
                color, depth = per_view_glue.decoder_forward(upstream_ext, Gaussians(*leaves), sc.extrinsics, sc.intrinsics, sc.near, sc.far,
                                                             (H, W), sc.background, depth_mode)
            if backward:
                outs, gouts = [color], [sc.grad_color]
                if depth is not None:
                    outs.append(depth); gouts.append(sc.grad_depth)
                torch.autograd.grad(outs, leaves, gouts)

        base_step(); torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            base_step()
        e1.record(); torch.cuda.synchronize()
        base_ms = e0.elapsed_time(e1) / 3
    out = {"config": name, "scenes": B, "gaussians": g.means.shape[1], "views": V, "HxW": f"{H}x{W}", "mode": ("fwd+bwd" if backward else "fwd") + ("+depth" if depth_mode else ""),
           "ms_per_call": round(ms, 3), "Mpix_s": round(B * V * H * W / ms / 1e3, 1),
           "upstream_style_ms": None if base_ms is None else round(base_ms, 3), "speedup": None if base_ms is None else round(base_ms / ms, 2), "pairs_last_call": pairs, "stages_ms": st,
           "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 2), "setup_s": round(time.time() - t0, 1)}
    print(json.dumps(out), flush=True)
    del sc, g
    R.release_scratch(); torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(RUNS)):
        run(n)
