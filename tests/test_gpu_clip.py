"""render_clip (my_depthsplat_b200/video.py) against the reference's chunk loop (model_wrapper.py:455-484): one
decoder.forward per chunk of render_chunk_size views, colours concatenated -- same frames, bit for bit, whatever the
chunk size, on the device or streamed to pinned host memory -- and against the ORACLE: the frames of the clip are the
frames the reference's decoder path (restated over the CPU oracle, helpers.oracle_decoder_forward) renders view by view."""
import pytest
import torch

from my_depthsplat_b200.scenes import make_scene
from my_depthsplat_b200.types import Gaussians

pytestmark = pytest.mark.gpu


def _decoder():
    from my_depthsplat_b200.decoder_splatting_cuda import DecoderSplattingCUDACfg, get_decoder
    cfg = type("DatasetCfg", (), {"background_color": [0.0, 0.0, 0.0]})()
    return get_decoder(DecoderSplattingCUDACfg(name="splatting_cuda"), cfg).to("cuda")


def _reference_chunk_loop(dec, g, sc, chunk_size):
    V = sc.extrinsics.shape[1]
    color = None
    for i in range((V + chunk_size - 1) // chunk_size):
        sl = slice(chunk_size * i, chunk_size * (i + 1))
        cur = dec.forward(g, sc.extrinsics[:, sl], sc.intrinsics[:, sl], sc.near[:, sl], sc.far[:, sl], sc.image_shape, depth_mode=None)
        color = cur.color if color is None else torch.cat((color, cur.color), dim=1)
    return color


@pytest.mark.parametrize("name,views,chunk", [("small", 7, 3), ("tiny", 5, 1), ("small", 4, None)])
def test_clip_equals_the_reference_chunk_loop(name, views, chunk):
    from my_depthsplat_b200.video import render_clip
    sc = make_scene(name, v_tgt=views).to("cuda")
    g = Gaussians(sc.gaussians.means, sc.gaussians.covariances, sc.gaussians.harmonics, sc.gaussians.opacities)
    dec = _decoder()
    with torch.no_grad():
        want = _reference_chunk_loop(dec, g, sc, chunk or views)
        one = dec.forward(g, sc.extrinsics, sc.intrinsics, sc.near, sc.far, sc.image_shape, depth_mode="depth")
    assert torch.equal(want, one.color)  # views are independent: chunking never changes a frame
    dev_out = render_clip(dec, g, sc.extrinsics, sc.intrinsics, sc.near, sc.far, sc.image_shape, chunk_size=chunk, depth_mode="depth")
    host_out = render_clip(dec, g, sc.extrinsics, sc.intrinsics, sc.near, sc.far, sc.image_shape, chunk_size=chunk, depth_mode="depth", to_host=True)
    assert torch.equal(dev_out.color, want) and torch.equal(dev_out.depth, one.depth)
    assert host_out.color.is_pinned() and not host_out.color.is_cuda
    assert torch.equal(host_out.color, want.cpu()) and torch.equal(host_out.depth, one.depth.cpu())
    assert not dev_out.color.requires_grad


@pytest.mark.parametrize("name,views,chunk,to_host", [("tiny", 5, 2, False), ("small_trained", 6, 4, True)])
def test_clip_frames_match_the_oracle(name, views, chunk, to_host):
    """Oracle parity of the clip path itself (not only self-consistency): colour and depth of every frame within 1e-5
    (depth relative to max(1, |d|)); a frame may hold a few threshold-flipped pixels (bounded as in test_gpu_parity)."""
    from helpers import leaf_gaussians, oracle_decoder_forward
    from my_depthsplat_b200.video import render_clip
    cpu = make_scene(name, v_tgt=views)
    with torch.no_grad():
        ref_c, ref_d = oracle_decoder_forward(leaf_gaussians(cpu), cpu.extrinsics, cpu.intrinsics, cpu.near, cpu.far, cpu.image_shape,
                                              cpu.background, "depth")
    sc = cpu.to("cuda")
    g = Gaussians(sc.gaussians.means, sc.gaussians.covariances, sc.gaussians.harmonics, sc.gaussians.opacities)
    out = render_clip(_decoder(), g, sc.extrinsics, sc.intrinsics, sc.near, sc.far, sc.image_shape, chunk_size=chunk, depth_mode="depth",
                      to_host=to_host)
    assert out.color.shape == ref_c.shape and out.depth.shape == ref_d.shape
    H, W = cpu.image_shape
    allowed = max(2, int(1e-3 * H * W))  # threshold-flipped PIXELS per frame (one flip moves all three channels of its pixel)
    for v in range(views):  # every frame on its own: a chunk boundary must not disturb the frames next to it
        cerr = (out.color[:, v].cpu() - ref_c[:, v]).abs().amax(dim=1)   # worst channel per pixel
        derr = (out.depth[:, v].cpu() - ref_d[:, v]).abs() / ref_d[:, v].abs().clamp(min=1.0)
        assert int((cerr > 1e-5).sum()) <= allowed, (v, int((cerr > 1e-5).sum()), float(cerr.max()))
        assert int((derr > 1e-5).sum()) <= allowed, (v, int((derr > 1e-5).sum()), float(derr.max()))
        assert float(cerr.median()) <= 1e-6
