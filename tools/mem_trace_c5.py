"""Which allocations make the peak of the stress config (python tools/mem_trace_c5.py): every torch.empty / zeros of
more than 0.3 GB issued while one C5 training step runs, with the allocator's live total at that moment."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from my_depthsplat_b200 import rasterizer as R  # noqa: E402
from my_depthsplat_b200.cuda_splatting import render_views  # noqa: E402
from my_depthsplat_b200.scenes import make_scene  # noqa: E402

sc = make_scene("C5", v_tgt=4).to("cuda")
g = sc.gaussians
H, W = sc.image_shape
_empty, _zeros, _empty_like = torch.empty, torch.zeros, torch.empty_like


def traced(fn, name):
    def f(*a, **k):
        t = fn(*a, **k)
        if t.is_cuda and t.numel() * t.element_size() > 3e8:
            import traceback
            where = [fr for fr in traceback.extract_stack(limit=6) if "my_depthsplat_b200" in fr.filename]
            print(f"{name} {t.numel() * t.element_size() / 1e9:7.2f} GB  live {torch.cuda.memory_allocated() / 1e9:7.2f} GB  "
                  f"{where[-1].name if where else '?'}:{where[-1].lineno if where else 0}", flush=True)
        return t
    return f


torch.empty, torch.zeros, torch.empty_like = traced(_empty, "empty"), traced(_zeros, "zeros"), traced(_empty_like, "empty_like")
for it in range(2):
    print(f"--- step {it}", flush=True)
    leaves = [t.detach().requires_grad_() for t in (g.means, g.covariances, g.harmonics, g.opacities)]
    color, _ = render_views(sc.extrinsics, sc.intrinsics, sc.near, sc.far, (H, W), sc.background, *leaves)
    print(f"forward done: live {torch.cuda.memory_allocated() / 1e9:.2f} GB, peak {torch.cuda.max_memory_allocated() / 1e9:.2f} GB", flush=True)
    torch.autograd.grad([color], leaves, [sc.grad_color])
    torch.cuda.synchronize()
    print(f"backward done: live {torch.cuda.memory_allocated() / 1e9:.2f} GB, peak {torch.cuda.max_memory_allocated() / 1e9:.2f} GB, remat {R.remat_count}", flush=True)
