"""Timeline of the pipelined end-to-end loop (debug aid): host timestamps and CUDA-event times per step."""
import sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from my_depthsplat_b200.scenes import make_scene
from my_depthsplat_b200.cuda_splatting import render_views

from my_depthsplat_b200 import rasterizer as R
R.debug_keep = "timing"
dev = torch.device("cuda", 0)
sc = make_scene("C2T")
H, W = sc.image_shape
host = {"means": sc.gaussians.means, "covariances": sc.gaussians.covariances, "harmonics": sc.gaussians.harmonics,
        "opacities": sc.gaussians.opacities, "grad_color": sc.grad_color}
host = {k: v.pin_memory() for k, v in host.items()}
cams = [t.to(dev) for t in (sc.extrinsics, sc.intrinsics, sc.near, sc.far)]
bg = torch.zeros(3, device=dev)
s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
dev_in = [{k: torch.empty_like(v, device=dev) for k, v in host.items()} for _ in range(2)]
host_out = None
names = ("means", "covariances", "harmonics", "opacities")
mode = sys.argv[1] if len(sys.argv) > 1 else "pipe"

def h2d(slot, stream):
    with torch.cuda.stream(stream):
        for k, v in host.items():
            dev_in[slot][k].copy_(v, non_blocking=True)

evs = []
torch.cuda.synchronize()
for i in range(12):
    t0 = time.perf_counter()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    cur = torch.cuda.current_stream(dev)
    slot = i & 1
    if mode == "pipe":
        e[0].record(s_in); h2d(slot, s_in); e[1].record(s_in); cur.wait_event(e[1])
    else:
        e[0].record(cur); h2d(slot, cur); e[1].record(cur)
    t1 = time.perf_counter()
    leaves = [dev_in[slot][k].detach().requires_grad_() for k in names]
    e[2].record(cur)
    color, _ = render_views(*cams, (H, W), bg, *leaves)
    t2 = time.perf_counter()
    grads = torch.autograd.grad(color, leaves, dev_in[slot]["grad_color"])
    e[3].record(cur)
    t3 = time.perf_counter()
    outs = [color] + list(grads)
    if host_out is None:
        host_out = [torch.empty(o.shape, dtype=o.dtype, pin_memory=True) for o in outs]
    st = s_out if mode == "pipe" else cur
    if mode == "pipe":
        s_out.wait_event(e[3])
    with torch.cuda.stream(st):
        e[4].record(st)
        for ho, o in zip(host_out, outs):
            ho.copy_(o, non_blocking=True)
        e[5].record(st)
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    ms = lambda a, b: e[a].elapsed_time(e[b])
    print(f"step {i}: host enq h2d {1e3*(t1-t0):6.2f}  fwd {1e3*(t2-t1):6.2f}  bwd {1e3*(t3-t2):6.2f}  total {1e3*(t4-t0):6.2f} | gpu h2d {ms(0,1):6.2f} compute {ms(2,3):6.2f} d2h {ms(4,5):6.2f}")
