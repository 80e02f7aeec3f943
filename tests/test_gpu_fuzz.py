"""Seeded random scene shapes against the oracle: image sizes that are no multiple of the 16x16 tile (down to a single
partial tile), 1-3 scenes x 1-3 views, every scale regime (incl. near-plane clustering that makes tile lists long and
footprints larger than the image), coloured backgrounds, depth modes.  Same bars as tests/test_gpu_fullsize_parity.py
(helpers.strict_parity_check): stages bit-exact, colour / depth within 1e-5 on every pixel the oracle does not mark
fragile, gradients within 1e-4 of their scale off the flipped pixels."""
import numpy as np
import pytest
import torch

from helpers import strict_parity_check
from my_depthsplat_b200.scenes import SceneConfig, make_scene

pytestmark = pytest.mark.gpu


def _cases():
    rng = np.random.default_rng(2024)
    out = []
    for i in range(10):
        h, w = int(rng.integers(5, 75)), int(rng.integers(5, 90))
        mode = ["init", "trained", "stress"][i % 3]
        out.append(dict(
            cfg=SceneConfig(f"fuzz{i}", 500 + i, int(rng.integers(1, 4)), h, w, batch=int(rng.integers(1, 4)), v_tgt=int(rng.integers(1, 4)),
                            scale_mode=mode, scale_max=float([3.0, 0.1, 0.5][i % 3]), near_plane_fraction=0.2 if mode == "stress" else 0.0,
                            far=float(rng.choice([20.0, 100.0, 400.0]))),
            depth_mode=[None, "depth", "disparity", "log"][i % 4],
            bg=torch.tensor(rng.random(3), dtype=torch.float32)))
    return out


@pytest.mark.parametrize("case", _cases(), ids=lambda c: f"{c['cfg'].name}-{c['cfg'].height}x{c['cfg'].width}-b{c['cfg'].batch}v{c['cfg'].v_tgt}-{c['cfg'].scale_mode}-{c['depth_mode']}")
def test_random_scene_against_the_oracle(case, capsys):
    cpu = make_scene(case["cfg"])
    cpu.background = case["bg"]
    report = strict_parity_check(cpu, case["depth_mode"], case["cfg"].name)
    with capsys.disabled():
        print("\n" + "\n".join(report))
