"""my_depthsplat_b200.loss_mse (the unfused expressions) against the reference's UNMODIFIED src/loss/loss_mse.py and
src/evaluation/metrics.py::compute_psnr, imported with stub parents (SURVEY.md appendix A recipe).  Only in the build
container: /root/reference is not on the GPU box."""
import importlib
import sys
import types

import pytest
import torch

from helpers import REFERENCE_SRC, have_reference

pytestmark = pytest.mark.skipif(not have_reference(), reason="/root/reference not available")


def _load_reference():
    root = str(REFERENCE_SRC)
    for name, path in [("src", root), ("src.loss", root + "/loss"), ("src.evaluation", root + "/evaluation"), ("src.model", root + "/model"),
                       ("src.model.decoder", root + "/model/decoder"), ("src.dataset", root + "/dataset")]:
        pkg = types.ModuleType(name)
        pkg.__path__ = [path]
        sys.modules[name] = pkg
    # leaf modules the two files import only for type annotations / unrelated metrics
    for name, attrs in [("src.dataset.types", ["BatchedExample"]), ("src.model.decoder.decoder", ["DecoderOutput"]), ("src.model.types", ["Gaussians"]),
                        ("lpips", ["LPIPS"]), ("skimage", []), ("skimage.metrics", ["structural_similarity"])]:
        if name not in sys.modules or name.startswith("src."):
            m = types.ModuleType(name)
            for a in attrs:
                setattr(m, a, object)
            sys.modules[name] = m
    for name in ("src.loss.loss", "src.loss.loss_mse", "src.evaluation.metrics"):
        sys.modules.pop(name, None)
    return importlib.import_module("src.loss.loss_mse"), importlib.import_module("src.evaluation.metrics")


def test_unfused_loss_and_psnr_equal_the_reference():
    from my_depthsplat_b200 import loss_mse as ours
    ref_loss, ref_metrics = _load_reference()
    g = torch.Generator().manual_seed(3)
    color = torch.rand(2, 3, 3, 20, 24, generator=g) * 1.2 - 0.1
    target = torch.rand(2, 3, 3, 20, 24, generator=g)
    pred = types.SimpleNamespace(color=color, depth=None)
    batch = {"target": {"image": target}}
    mask = torch.rand(2, 3, 3, 20, 24, generator=g) > 0.5
    r = ref_loss.LossMse(ref_loss.LossMseCfgWrapper(ref_loss.LossMseCfg(0.7)))
    o = ours.LossMse(ours.LossMseCfgWrapper(ours.LossMseCfg(0.7)))
    assert o.name == r.name == "mse"
    for kw in (dict(l1_loss=False, clamp_large_error=0.0, valid_depth_mask=None), dict(l1_loss=True, clamp_large_error=0.0, valid_depth_mask=None),
               dict(l1_loss=False, clamp_large_error=0.3, valid_depth_mask=None), dict(l1_loss=False, clamp_large_error=0.0, valid_depth_mask=mask),
               dict(l1_loss=True, clamp_large_error=0.2, valid_depth_mask=mask)):
        assert torch.equal(o.forward(pred, batch, None, 0, **kw), r.forward(pred, batch, None, 0, **kw)), kw
    assert torch.equal(ours.compute_psnr(target[0], color[0]), ref_metrics.compute_psnr(target[0], color[0]))
