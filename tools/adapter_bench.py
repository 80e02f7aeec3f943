"""Adapter + decoder, forward + backward, two ways on one GPU (python tools/adapter_bench.py [C1|C2T]):
  unfused  the Gaussian adapter in PyTorch (the reference's tensor expressions, my_depthsplat_b200.gaussian_adapter) writes
           the world-space Gaussians (160 B each), the decoder reads them; the backward writes their gradients (160 B) and
           autograd runs the adapter's backward over them;
  fused    the projection kernels start from the head's raw channel planes (164 B per Gaussian read in the forward and
           again in the backward, 152 B of gradients written) -- the world-space tensors never exist.
Prints one JSON line with both times and the HBM bytes per Gaussian of the adapter <-> decoder interface."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from my_depthsplat_b200 import gaussian_adapter as GA  # noqa: E402
from my_depthsplat_b200.decoder_splatting_cuda import DecoderSplattingCUDACfg, get_decoder  # noqa: E402
from my_depthsplat_b200.scenes import make_scene  # noqa: E402


def main(name="C2T", steps=5):
    sc = make_scene(name).to("cuda")
    b, vc = sc.ctx_extrinsics.shape[:2]
    h, w = sc.image_shape
    g = torch.Generator(device="cuda").manual_seed(0)
    head = torch.randn(b, vc, 37, h, w, device="cuda", generator=g)
    head[:, :, 3:6] = head[:, :, 3:6] * 0.5 - 1.0
    head[:, :, 10:] *= 0.5
    depth = sc.near[0, 0] * (2.0 + 18.0 * torch.rand(b, vc, h, w, device="cuda", generator=g))
    images = torch.rand(b, vc, 3, h, w, device="cuda", generator=g)
    K = sc.intrinsics[:, :1].expand(b, vc, 3, 3).contiguous()
    adapter = GA.GaussianAdapter(GA.GaussianAdapterCfg(1e-10, 0.1 if "C2" in name else 3.0, 2)).cuda()
    dec = get_decoder(DecoderSplattingCUDACfg(name="splatting_cuda"), type("D", (), {"background_color": [0.0, 0.0, 0.0]})()).cuda()
    fused = GA.FusedAdapterDecoder(adapter, dec)
    cams = (sc.extrinsics, sc.intrinsics, sc.near, sc.far, sc.image_shape)

    def unfused_step():
        h0, d0 = head.detach().requires_grad_(), depth.detach().requires_grad_()
        gs = GA.adapt_head_output(adapter, h0, d0, images, sc.ctx_extrinsics, K, (h, w))
        out = dec.forward(gs, *cams)
        torch.autograd.grad(out.color, (h0, d0), sc.grad_color)

    def fused_step():
        h0, d0 = head.detach().requires_grad_(), depth.detach().requires_grad_()
        out = fused.forward(h0, d0, images, sc.ctx_extrinsics, K, *cams)
        torch.autograd.grad(out.color, (h0, d0), sc.grad_color)

    res = {}
    for nm, fn in (("unfused", unfused_step), ("fused", fused_step)):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record(); torch.cuda.synchronize()
        res[nm] = {"ms_per_step": round(e0.elapsed_time(e1) / steps, 3), "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 2)}
    n = vc * h * w
    print(json.dumps({"config": name, "gaussians": n, "views": int(sc.extrinsics.shape[1]), **res,
                      "interface_bytes_per_gaussian": {"unfused": {"fwd": "164 read + 160 written (adapter) + 160 read (projection)", "bwd": "160 read + 160 written (projection) + 160 + 164 read, 152 written (adapter backward)", "total": 1280},
                                                       "fused": {"fwd": "164 read", "bwd": "164 read + 152 written", "total": 480}}}))


if __name__ == "__main__":
    main(*(sys.argv[1:2] or ["C2T"]))
