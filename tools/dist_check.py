"""N-GPU equivalence check (torchrun --nproc-per-node N tools/dist_check.py): view-sharded gradients, summed
(a) by the fused NVLS reduce inside the backward kernel, (b) by a plain NCCL all-reduce, (c) chunked / overlapped,
(d) by the library's own two-shot NVLS all-reduce kernel on a symmetric buffer, against the single-GPU gradients over
all views."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from my_depthsplat_b200.decoder_splatting_cuda import DecoderSplattingCUDACfg, get_decoder  # noqa: E402
from my_depthsplat_b200.dist import RangeShardedDecoder, ViewShardedDecoder, range_bounds, shard_bounds  # noqa: E402
from my_depthsplat_b200.scenes import make_scene  # noqa: E402
from my_depthsplat_b200.types import Gaussians  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
name = sys.argv[1] if len(sys.argv) > 1 else "small_trained"
sc = make_scene(name, v_tgt=2 * world + 1).to(dev)  # uneven split
H, W = sc.image_shape
V = sc.extrinsics.shape[1]
cfg = type("D", (), {"background_color": [0.0, 0.0, 0.0]})()


def grads(decoder, cams_sliced):
    g = sc.gaussians
    leaves = [t.detach().clone().requires_grad_() for t in (g.means, g.covariances, g.harmonics, g.opacities)]
    out = decoder.forward(Gaussians(*leaves), sc.extrinsics, sc.intrinsics, sc.near, sc.far, (H, W), depth_mode="depth")
    lo, hi = cams_sliced
    loss = (out.color * sc.grad_color[:, lo:hi]).sum() + (out.depth * sc.grad_depth[:, lo:hi]).sum()
    return torch.autograd.grad(loss, leaves)


lo, hi = shard_bounds(V, world, rank)
single = grads(get_decoder(DecoderSplattingCUDACfg(name="splatting_cuda"), cfg).to(dev), (0, V))          # all views on this GPU
nccl = grads(ViewShardedDecoder(get_decoder(DecoderSplattingCUDACfg(name="splatting_cuda"), cfg).to(dev)), (lo, hi))
fused_dec = ViewShardedDecoder(get_decoder(DecoderSplattingCUDACfg(name="splatting_cuda"), cfg).to(dev), fused_reduce=True)
fused = [t.clone() for t in grads(fused_dec, (lo, hi))]
fused2 = [t.clone() for t in grads(fused_dec, (lo, hi))]   # second call: the other symmetric buffer
ovl_dec = ViewShardedDecoder(get_decoder(DecoderSplattingCUDACfg(name="splatting_cuda"), cfg).to(dev), overlap_reduce=True)
ovl = grads(ovl_dec, (lo, hi))
nvls_dec = ViewShardedDecoder(get_decoder(DecoderSplattingCUDACfg(name="splatting_cuda"), cfg).to(dev), nvls_reduce=True)
nvls = [[t.clone() for t in grads(nvls_dec, (lo, hi))] for _ in range(4)]  # four calls: the buffer ring wraps around
# (e) Gaussians sharded by range: all-gather in the forward, reduce-scatter in the backward -- by the library's NVLS kernel in
# pieces under the projection backward, and by NCCL
N = sc.gaussians.means.shape[1]
glo, ghi = range_bounds(N, world, rank)


def range_grads(dec):
    g = sc.gaussians
    leaves = [t[:, glo:ghi].detach().clone().requires_grad_() for t in (g.means, g.covariances, g.harmonics, g.opacities)]
    out = dec.forward(Gaussians(*leaves), N, sc.extrinsics, sc.intrinsics, sc.near, sc.far, (H, W), depth_mode="depth")
    loss = (out.color * sc.grad_color[:, lo:hi]).sum() + (out.depth * sc.grad_depth[:, lo:hi]).sum()
    return [t.clone() for t in torch.autograd.grad(loss, leaves)]


rk_dec = RangeShardedDecoder(get_decoder(DecoderSplattingCUDACfg(name="splatting_cuda"), cfg).to(dev), pieces=3)
rk = [range_grads(rk_dec) for _ in range(4)]   # four calls: the symmetric buffers rotate
rn = range_grads(RangeShardedDecoder(get_decoder(DecoderSplattingCUDACfg(name="splatting_cuda"), cfg).to(dev), kernel_reduce=False))
# (f) replicated Gaussians, gradients reduce-scattered: the own range holds the sum over all ranks, the rest is zero
sc_dec = ViewShardedDecoder(get_decoder(DecoderSplattingCUDACfg(name="splatting_cuda"), cfg).to(dev), scatter_grads=True, pieces=3)
sg = [[t.clone() for t in grads(sc_dec, (lo, hi))] for _ in range(4)]
torch.cuda.synchronize()
ok = True
for k, nm in enumerate(("means", "covariances", "harmonics", "opacities")):
    scale = float(single[k].abs().max())
    inside = max(float((sg[j][k][:, glo:ghi] - single[k][:, glo:ghi]).abs().max()) / scale for j in range(4)) if ghi > glo else 0.0
    outside = max(float(sg[j][k][:, :glo].abs().max() if glo else 0.0) + float(sg[j][k][:, ghi:].abs().max() if ghi < N else 0.0) for j in range(4))
    print(f"rank {rank} {nm:12s} scatter_grads: |own range - single| {inside:.2e}, largest value outside the range {outside:.1e} (4 calls)", flush=True)
    ok &= inside < 2e-4 and outside == 0.0
for k, nm in enumerate(("means", "covariances", "harmonics", "opacities")):
    ref = single[k][:, glo:ghi]
    scale = float(single[k].abs().max())
    if ref.numel() == 0:
        continue
    e_k = max(float((rk[j][k] - ref).abs().max()) / scale for j in range(4))
    e_n = float((rn[k] - ref).abs().max()) / scale
    print(f"rank {rank} {nm:12s} range [{glo}, {ghi}): |NVLS reduce-scatter - single| {e_k:.2e} (4 calls)  |NCCL reduce-scatter - single| {e_n:.2e}", flush=True)
    ok &= e_k < 2e-4 and e_n < 2e-4
print(f"rank {rank} range reducer active: {rk_dec.reducer is not None and rk_dec.reducer.available}", flush=True)
for nm, s, a, b, c, d in zip(("means", "covariances", "harmonics", "opacities"), single, nccl, fused, fused2, ovl):
    scale = float(s.abs().max())
    e_n, e_f, e_f2, e_o = (float((x - s).abs().max()) / scale for x in (a, b, c, d))
    print(f"rank {rank} {nm:12s} |nccl - single| {e_n:.2e}  |fused - single| {e_f:.2e}  |fused(2nd) - single| {e_f2:.2e}  |overlapped - single| {e_o:.2e}", flush=True)
    ok &= e_n < 2e-4 and e_f < 2e-4 and e_f2 < 2e-4 and e_o < 2e-4
for k, nm in enumerate(("means", "covariances", "harmonics", "opacities")):
    scale = float(single[k].abs().max())
    errs = [float((nvls[j][k] - single[k]).abs().max()) / scale for j in range(4)]
    print(f"rank {rank} {nm:12s} |NVLS kernel all-reduce - single| {max(errs):.2e} (4 calls)", flush=True)
    ok &= max(errs) < 2e-4
print(f"rank {rank} fused reducer active: {fused_dec.reducer is not None and fused_dec.reducer.available}  {'OK' if ok else 'MISMATCH'}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
