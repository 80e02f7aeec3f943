"""Multi-process host logic of my_depthsplat_b200.dist on CPU: gloo backend, world_size 2.
The renderer behind the sharding wrapper is the CPU oracle (a checker standing in for the CUDA path,
which needs a GPU); what is under test is the partitioning, the gradient all-reduce and the gather."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, result_dir):
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from helpers import leaf_gaussians, oracle_decoder_forward
        from my_depthsplat_b200 import dist as D
        from my_depthsplat_b200.scenes import make_scene
        from my_depthsplat_b200.types import DecoderOutput

        scene = make_scene("tiny", v_tgt=3)  # 3 views over 2 ranks: uneven split (2 + 1)

        class OracleDecoder(torch.nn.Module):
            def forward(self, gaussians, extrinsics, intrinsics, near, far, image_shape, depth_mode=None):
                c, d = oracle_decoder_forward(gaussians, extrinsics, intrinsics, near, far, image_shape, scene.background, depth_mode)
                return DecoderOutput(c, d)

        # --- training: views sharded, gradients summed across ranks
        g = leaf_gaussians(scene)
        dec = D.ViewShardedDecoder(OracleDecoder())
        out = dec.forward(g, scene.extrinsics, scene.intrinsics, scene.near, scene.far, scene.image_shape, depth_mode="depth")
        lo, hi = D.shard_bounds(3, world, rank)
        assert out.color.shape[1] == hi - lo
        loss = (out.color * scene.grad_color[:, lo:hi]).sum() + (out.depth * scene.grad_depth[:, lo:hi]).sum()
        loss.backward()
        # --- inference: gather the frames
        with torch.no_grad():
            full = D.ViewShardedDecoder(OracleDecoder(), gather=True).forward(
                g, scene.extrinsics, scene.intrinsics, scene.near, scene.far, scene.image_shape, depth_mode="depth")
        # --- clip rendering: views sharded, chunks of 1 view, frames gathered
        from my_depthsplat_b200.video import render_clip
        clip = render_clip(OracleDecoder(), g, scene.extrinsics, scene.intrinsics, scene.near, scene.far, scene.image_shape,
                           chunk_size=1, depth_mode="depth", gather=True)
        local = render_clip(OracleDecoder(), g, scene.extrinsics, scene.intrinsics, scene.near, scene.far, scene.image_shape, chunk_size=None)
        assert local.color.shape[1] == hi - lo and local.depth is None
        assert torch.equal(local.color, clip.color[:, lo:hi])
        # --- gradients carved out of one allocation (what the CUDA rasterizer's backward returns): one flat all-reduce
        from my_depthsplat_b200.rasterizer import _grad_tensors
        xs = [torch.zeros(s, requires_grad=True) for s in ((1, 8, 3), (1, 8, 3, 3), (1, 8, 3, 9), (1, 8))]
        carved = _grad_tensors(*xs)
        assert D._flat_span(list(carved)) is not None
        for i, t in enumerate(carved):
            t.fill_(float((rank + 1) * (i + 1)))
        torch.autograd.backward(D._SyncGrads.apply(None, *xs), grad_tensors=list(carved))
        flat_sums = [float(x.grad.min()) for x in xs] + [float(x.grad.max()) for x in xs]
        # --- Gaussians sharded by range: all-gather in the forward, this rank's range of the summed gradients back
        from my_depthsplat_b200.types import Gaussians
        N = scene.gaussians.means.shape[1]
        glo, ghi = D.range_bounds(N, world, rank)
        mine = Gaussians(*(t[:, glo:ghi].detach().clone().requires_grad_() for t in
                           (scene.gaussians.means, scene.gaussians.covariances, scene.gaussians.harmonics, scene.gaussians.opacities)))
        rdec = D.RangeShardedDecoder(OracleDecoder())
        rout = rdec.forward(mine, N, scene.extrinsics, scene.intrinsics, scene.near, scene.far, scene.image_shape, depth_mode="depth")
        assert torch.equal(rout.color, out.color.detach())
        ((rout.color * scene.grad_color[:, lo:hi]).sum() + (rout.depth * scene.grad_depth[:, lo:hi]).sum()).backward()
        range_grads = [mine.means.grad, mine.covariances.grad, mine.harmonics.grad, mine.opacities.grad]
        # --- fewer views than ranks: every rank raises the same error instead of one of them hanging the others
        try:
            dec.forward(g, scene.extrinsics[:, :1], scene.intrinsics[:, :1], scene.near[:, :1], scene.far[:, :1], scene.image_shape)
            too_few = "no error"
        except ValueError as e:
            too_few = str(e)
        torch.save({"range": (glo, ghi), "range_grads": range_grads, "too_few": too_few, "color": full.color, "depth": full.depth, "flat_sums": flat_sums, "clip_color": clip.color, "clip_depth": clip.depth, "grads": [g.means.grad, g.covariances.grad, g.harmonics.grad, g.opacities.grad]},
                   Path(result_dir) / f"rank{rank}.pt")
    finally:
        dist.destroy_process_group()


def test_shard_bounds():
    from my_depthsplat_b200.dist import shard_bounds
    for n in (0, 1, 3, 4, 10, 100):
        for w in (1, 2, 4, 8):
            parts = [shard_bounds(n, w, r) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def test_view_sharded_decoder_two_ranks(tmp_path):
    sys.path.insert(0, str(ROOT / "tests"))
    from helpers import leaf_gaussians, oracle_decoder_forward
    from my_depthsplat_b200.scenes import make_scene
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(tmp_path / f"rank{r}.pt") for r in range(world)]
    # single-process result over all views
    scene = make_scene("tiny", v_tgt=3)
    g = leaf_gaussians(scene)
    c, d = oracle_decoder_forward(g, scene.extrinsics, scene.intrinsics, scene.near, scene.far, scene.image_shape, scene.background, "depth")
    ((c * scene.grad_color).sum() + (d * scene.grad_depth).sum()).backward()
    for r in res:
        assert torch.equal(r["color"], c.detach()) and torch.equal(r["depth"], d.detach())  # gathered frames = 1-process frames
        assert r["flat_sums"] == [3.0, 6.0, 9.0, 12.0] * 2  # (1 + 2) * (i + 1) on every element of tensor i
        assert torch.equal(r["clip_color"], c.detach()) and torch.equal(r["clip_depth"], d.detach())  # chunked clip too
        for got, ref in zip(r["grads"], (g.means.grad, g.covariances.grad, g.harmonics.grad, g.opacities.grad)):
            torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-6 * float(ref.abs().max()))  # sum order differs
    for a, b in zip(res[0]["grads"], res[1]["grads"]):
        assert torch.equal(a, b)  # every rank holds the same reduced gradients
    # range sharding: the ranges tile the Gaussians, every rank holds the summed gradient of its own range
    assert res[0]["range"][0] == 0 and res[0]["range"][1] == res[1]["range"][0] and res[1]["range"][1] == g.means.shape[1]
    for r in res:
        glo, ghi = r["range"]
        assert "cannot be sharded" in r["too_few"]
        for got, ref in zip(r["range_grads"], (g.means.grad, g.covariances.grad, g.harmonics.grad, g.opacities.grad)):
            torch.testing.assert_close(got, ref[:, glo:ghi], rtol=1e-4, atol=1e-6 * float(ref.abs().max()))


def test_flat_span_recognises_only_gapless_tilings():
    from my_depthsplat_b200.dist import _flat_span
    from my_depthsplat_b200.rasterizer import _grad_tensors
    like = [torch.zeros(2, 8, 3), torch.zeros(2, 8, 3, 3), torch.zeros(2, 8, 3, 9), torch.zeros(2, 8)]
    carved = _grad_tensors(*like)
    flat = _flat_span(list(carved))
    assert flat is not None and flat.numel() == sum(t.numel() for t in like) and flat.data_ptr() == carved[0].data_ptr()
    assert [t.shape for t in carved] == [t.shape for t in like]
    flat.fill_(1.5)
    assert all(float(t.min()) == 1.5 for t in carved)
    # sizes that would misalign the next tensor fall back to separate allocations -> no flat span
    odd = _grad_tensors(torch.zeros(1, 7, 3), torch.zeros(1, 7, 3, 3), torch.zeros(1, 7, 3, 9), torch.zeros(1, 7))
    assert _flat_span(list(odd)) is None
    # unrelated tensors, sub-ranges (gaps) and mixed dtypes are not spans
    assert _flat_span([torch.zeros(4), torch.zeros(4)]) is None
    assert _flat_span([carved[0][:, :4], carved[1][:, :4]]) is None
    assert _flat_span([carved[0], carved[1].double()]) is None
    assert _flat_span([carved[0]]) is None


def test_render_clip_rejects_host_streaming_without_cuda_and_with_gather():
    from my_depthsplat_b200.scenes import make_scene
    from my_depthsplat_b200.types import DecoderOutput
    from my_depthsplat_b200.video import render_clip
    scene = make_scene("tiny")

    class Dummy(torch.nn.Module):
        def forward(self, g, e, k, n, f, shape, depth_mode=None):
            return DecoderOutput(torch.zeros(e.shape[0], e.shape[1], 3, *shape), None)

    args = (Dummy(), scene.gaussians, scene.extrinsics, scene.intrinsics, scene.near, scene.far, scene.image_shape)
    out = render_clip(*args, chunk_size=1)
    assert out.color.shape == (1, scene.extrinsics.shape[1], 3, *scene.image_shape) and out.depth is None
    with pytest.raises(ValueError):
        render_clip(*args, to_host=True)             # CPU tensors: there is no CPU path to stream from
    with pytest.raises(ValueError):
        render_clip(*args, to_host=True, gather=True)


@pytest.mark.parametrize("n,world,k", [(2949120, 8, 4), (2949120, 2, 4), (46080, 2, 3), (300, 4, 4), (131072, 8, 8), (257, 2, 1)])
def test_reduce_scatter_pieces_tile_every_ranks_range(n, world, k):
    """RangeScatterReducer.pieces(): the pieces of the projection backward cover chunk offsets [0, per_rank) of every
    rank's range exactly once, in order, with non-increasing sizes (the last, un-hidden pull is the smallest)."""
    from my_depthsplat_b200.dist import RangeScatterReducer
    r = RangeScatterReducer.__new__(RangeScatterReducer)
    r._N, r.world, r.num_pieces = n, world, k
    ps = r.pieces()
    per_rank = r._per_rank
    assert per_rank == -(-(-(-n // 256)) // world)
    assert 1 <= len(ps) <= k
    pos = 0
    for c0, cn, stride, repeat in ps:
        assert c0 == pos and cn > 0 and stride == per_rank and repeat == world and stride >= cn
        pos += cn
    assert pos == per_rank
    sizes = [p[1] for p in ps]
    assert all(a >= b for a, b in zip(sizes[:-2], sizes[1:-1]))  # the last piece takes the remainder
    if k >= 2 and per_rank >= 4 * k:
        assert sizes[-1] <= sizes[0]
