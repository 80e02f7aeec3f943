"""The Gaussian adapter fused into the projection against the unfused route: the adapter in PyTorch
(my_depthsplat_b200.gaussian_adapter, pinned to the reference's unmodified file by tests/test_gaussian_adapter.py) followed
by the plain decoder.
  * the world-space Gaussians the kernel builds (exported through ``cooked_out``) equal the PyTorch adapter's to fp32
    rounding of a reassociated computation;
  * images agree (the two routes hand the projection means / covariances that differ in the last bits, so a radius or a
    threshold decision can differ on a few pixels: bounded fraction, as for CUDA-built camera blocks);
  * gradients w.r.t. the head's 37 channel planes and the depth agree with autograd through the unfused route."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(b=1, v=2, h=32, w=48, seed=0):
    from my_depthsplat_b200.scenes import SceneConfig, make_scene
    sc = make_scene(SceneConfig("adapter", 21 + seed, v, h, w, batch=b, v_tgt=3, scale_mode="init", scale_max=3.0)).to("cuda")
    g = torch.Generator(device="cuda").manual_seed(seed)
    head = torch.randn(b, v, 37, h, w, device="cuda", generator=g)
    head[:, :, 3:6] = head[:, :, 3:6] * 0.5 + 0.5          # scales around softplus(-3.5) ~ 0.03
    head[:, :, 10:] *= 0.5
    depth = 1.0 + 9.0 * torch.rand(b, v, h, w, device="cuda", generator=g)
    images = torch.rand(b, v, 3, h, w, device="cuda", generator=g)
    K = sc.intrinsics[:, :1].expand(b, v, 3, 3).contiguous()
    return sc, head, depth, images, K


@pytest.mark.parametrize("depth_mode", [None, "depth"])
def test_fused_adapter_matches_the_unfused_route(depth_mode):
    from my_depthsplat_b200 import gaussian_adapter as GA
    from my_depthsplat_b200.decoder_splatting_cuda import DecoderSplattingCUDACfg, get_decoder
    sc, head, depth, images, K = _setup()
    b, v, _, h, w = head.shape
    adapter = GA.GaussianAdapter(GA.GaussianAdapterCfg(1e-10, 3.0, 2)).cuda()
    dec = get_decoder(DecoderSplattingCUDACfg(name="splatting_cuda"), type("D", (), {"background_color": [0.0, 0.0, 0.0]})()).cuda()
    cams = (sc.extrinsics, sc.intrinsics, sc.near, sc.far, sc.image_shape)

    h0, d0 = head.clone().requires_grad_(), depth.clone().requires_grad_()
    gs = GA.adapt_head_output(adapter, h0, d0, images, sc.ctx_extrinsics, K, (h, w))
    ref = dec.forward(gs, *cams, depth_mode=depth_mode)
    loss = (ref.color * sc.grad_color).sum() + (0 if depth_mode is None else (ref.depth * sc.grad_depth).sum())
    loss.backward()

    h1, d1 = head.clone().requires_grad_(), depth.clone().requires_grad_()
    cooked = torch.empty(b, v * h * w, 40, device="cuda")
    fused = GA.FusedAdapterDecoder(adapter, dec)
    out = fused.forward(h1, d1, images, sc.ctx_extrinsics, K, *cams, depth_mode=depth_mode, cooked_out=cooked)
    loss = (out.color * sc.grad_color).sum() + (0 if depth_mode is None else (out.depth * sc.grad_depth).sum())
    loss.backward()

    # the Gaussians the kernel built
    want = torch.cat([gs.means, gs.covariances.reshape(b, -1, 9), gs.opacities[..., None], gs.harmonics.reshape(b, -1, 27)], dim=-1).detach()
    for name, sl, tol in (("means", slice(0, 3), 2e-6), ("covariances", slice(3, 12), 1e-5), ("opacities", slice(12, 13), 1e-6), ("harmonics", slice(13, 40), 2e-6)):
        a, r = cooked[..., sl], want[..., sl]
        err = float((a - r).abs().max() / r.abs().max())
        assert err <= tol, (name, err)
    # images
    cerr = (out.color - ref.color).abs()
    assert float((cerr > 1e-5).float().mean()) <= 2e-3, (float(cerr.max()), float((cerr > 1e-5).float().mean()))
    if depth_mode is not None:
        derr = (out.depth - ref.depth).abs() / ref.depth.abs().clamp(min=1.0)
        assert float((derr > 1e-5).float().mean()) <= 2e-3
    # gradients w.r.t. the raw channels and the depth: 99.9 % within 1e-4 of the scale (the rest sit next to a flipped decision)
    for name, a, r in (("head", h1.grad, h0.grad), ("depth", d1.grad, d0.grad)):
        groups = [("all", slice(None))] if name == "depth" else [("opacity", slice(0, 1)), ("offset", slice(1, 3)), ("scales", slice(3, 6)), ("quaternion", slice(6, 10)), ("sh", slice(10, 37))]
        for gname, sl in groups:
            aa, rr = (a, r) if name == "depth" else (a[:, :, sl], r[:, :, sl])
            e = (aa - rr).abs().flatten().cpu().numpy()
            scale = float(rr.abs().max())
            assert scale > 0, (name, gname)
            assert np.quantile(e, 0.999) <= 1e-4 * scale and e.max() <= 2e-2 * scale, (name, gname, e.max() / scale, np.quantile(e, 0.999) / scale)


def test_fused_adapter_rejects_what_it_does_not_cover():
    from my_depthsplat_b200 import gaussian_adapter as GA
    from my_depthsplat_b200.decoder_splatting_cuda import DecoderSplattingCUDACfg, get_decoder
    sc, head, depth, images, K = _setup(h=10, w=10)   # 100 pixels: not a multiple of the 256-Gaussian chunk
    adapter = GA.GaussianAdapter(GA.GaussianAdapterCfg(1e-10, 3.0, 2)).cuda()
    dec = get_decoder(DecoderSplattingCUDACfg(name="splatting_cuda"), type("D", (), {"background_color": [0.0, 0.0, 0.0]})()).cuda()
    with pytest.raises(ValueError, match="multiple of 256"):
        GA.FusedAdapterDecoder(adapter, dec).forward(head, depth, images, sc.ctx_extrinsics, K, sc.extrinsics, sc.intrinsics, sc.near, sc.far, sc.image_shape)
    with pytest.raises(NotImplementedError):
        GA.FusedAdapterDecoder(GA.GaussianAdapter(GA.GaussianAdapterCfg(1e-10, 3.0, 1)), dec)
