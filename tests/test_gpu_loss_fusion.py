"""Loss-side fusion: the MSE / L1 loss of loss_mse.py:33-44, its gradient and compute_psnr's
squared error computed in the compositing epilogue, against the same quantities computed from the rendered colour with the
reference's tensor expressions (my_depthsplat_b200.loss_mse, pinned to the unmodified reference file by
tests/test_reference_loss.py)."""
import pytest
import torch

from helpers import cuda_leaf_gaussians
from my_depthsplat_b200.scenes import make_scene

pytestmark = pytest.mark.gpu


def _decoder():
    from my_depthsplat_b200.decoder_splatting_cuda import DecoderSplattingCUDACfg, get_decoder
    cfg = type("DatasetCfg", (), {"background_color": [0.1, 0.2, 0.3]})()
    return get_decoder(DecoderSplattingCUDACfg(name="splatting_cuda"), cfg).cuda()


@pytest.mark.parametrize("name", ["small", "ragged"])
@pytest.mark.parametrize("l1", [False, True])
@pytest.mark.parametrize("scale", [1.0, 3.0])
def test_fused_loss_gradients_and_psnr(name, l1, scale):
    from my_depthsplat_b200 import loss_mse as LM
    sc = make_scene(name).to("cuda")
    B, V = sc.extrinsics.shape[:2]
    gen = torch.Generator(device="cuda").manual_seed(4)
    target = torch.rand(B, V, 3, *sc.image_shape, device="cuda", generator=gen) * 1.2 - 0.1
    batch = {"target": {"image": target}}
    dec = _decoder()
    loss_fn = LM.LossMse(LM.LossMseCfgWrapper(LM.LossMseCfg(0.5)))
    cams = (sc.extrinsics, sc.intrinsics, sc.near, sc.far, sc.image_shape)

    g0 = cuda_leaf_gaussians(sc)
    plain = dec.forward(g0, *cams)
    ref = loss_fn.forward(plain, batch, g0, 0, l1_loss=l1)
    (ref * scale).backward()

    g1 = cuda_leaf_gaussians(sc)
    fused = dec.forward_fused_mse(g1, *cams, mse_target=target, mse_weight=0.5, mse_l1=l1)
    assert torch.equal(fused.color, plain.color)
    got = loss_fn.forward(fused, batch, g1, 0, l1_loss=l1)
    assert got is fused.fused_mse.loss
    torch.testing.assert_close(got, ref, rtol=2e-5, atol=0)
    (got * scale).backward()
    for k in ("means", "covariances", "harmonics", "opacities"):
        a, b = getattr(g1, k).grad, getattr(g0, k).grad
        assert float((a - b).abs().max()) <= 2e-5 * float(b.abs().max()), k
    want = torch.stack([LM.compute_psnr(target[b], plain.color[b].detach()) for b in range(B)])
    torch.testing.assert_close(LM.fused_psnr(fused), want, rtol=1e-5, atol=1e-5)
    # anything the epilogue does not cover takes the reference's expressions on the colour
    other = loss_fn.forward(fused, batch, g1, 0, l1_loss=l1, clamp_large_error=0.2)
    assert other is not fused.fused_mse.loss


def test_fused_loss_next_to_a_loss_on_the_colour():
    """LPIPS-style second loss on the same colour: both gradients reach the Gaussians."""
    from my_depthsplat_b200 import loss_mse as LM
    sc = make_scene("small").to("cuda")
    target = torch.rand(*sc.grad_color.shape, device="cuda")
    dec = _decoder()
    cams = (sc.extrinsics, sc.intrinsics, sc.near, sc.far, sc.image_shape)
    g0, g1 = cuda_leaf_gaussians(sc), cuda_leaf_gaussians(sc)
    plain = dec.forward(g0, *cams)
    (((plain.color - target) ** 2).mean() + (plain.color * sc.grad_color).sum()).backward()
    fused = dec.forward_fused_mse(g1, *cams, mse_target=target, mse_weight=1.0)
    (fused.fused_mse.loss + (fused.color * sc.grad_color).sum()).backward()
    for k in ("means", "covariances", "harmonics", "opacities"):
        a, b = getattr(g1, k).grad, getattr(g0, k).grad
        assert float((a - b).abs().max()) <= 2e-5 * float(b.abs().max()), k
