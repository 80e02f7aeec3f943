"""render_clip (my_depthsplat_b200/video.py) against the reference's chunk loop (model_wrapper.py:455-484): one
decoder.forward per chunk of render_chunk_size views, colours concatenated -- same frames, bit for bit, whatever the
chunk size, on the device or streamed to pinned host memory."""
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
