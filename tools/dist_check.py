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
from my_depthsplat_b200.dist import ViewShardedDecoder, shard_bounds  # noqa: E402
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
torch.cuda.synchronize()
ok = True
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
