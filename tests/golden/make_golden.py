"""Generate the golden fixtures of tests/golden/ (run in the BUILD container, where /root/reference exists):

    python tests/golden/make_golden.py

Each fixture is the output of the reference's UNMODIFIED src/model/decoder/cuda_splatting.py
(render_cuda / render_depth_cuda, imported from /root/reference) running on top of the CPU oracle
(oracle/ext_compat.py stands in for the third-party diff_gaussian_rasterization, which is absent), on a
seeded synthetic scene of my_depthsplat_b200.scenes, plus the gradients of
sum(color * grad_color) + sum(depth * grad_depth) w.r.t. every Gaussian tensor, plus digests of the
oracle's stage outputs (sorted keys, values, tile ranges) of every view.
The GPU box has no /root/reference: the fixtures are how the reference's glue travels there.
"""
import hashlib
import os
import sys
from pathlib import Path

os.environ.setdefault("MKL_CBWR", "COMPATIBLE")   # host-independent MKL code path (see tests/conftest.py)

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))
sys.path.insert(0, str(HERE.parent))

from helpers import input_digest, leaf_gaussians, load_reference_cuda_splatting, per_view_extension_inputs  # noqa: E402
from my_depthsplat_b200.scenes import make_scene  # noqa: E402
from test_reference_glue import _ref_render  # noqa: E402

FIXTURES = [("tiny", "depth"), ("ragged", None), ("tiny", "disparity")]


def stage_digests(scene):
    from oracle import splat_oracle as so
    B, V = scene.extrinsics.shape[:2]
    out = []
    for b in range(B):
        for v in range(V):
            st = so.forward_view(**per_view_extension_inputs(scene, b, v))
            h = hashlib.sha256()
            for a in (st.keys, st.vals, st.ranges, st.radii, st.n_contrib):
                h.update(np.ascontiguousarray(a).tobytes())
            out.append(h.hexdigest())
            st.close()
    return out


def main():
    from oracle import ext_compat
    cs = load_reference_cuda_splatting(ext_compat)
    for name, depth_mode in FIXTURES:
        scene = make_scene(name)
        g = leaf_gaussians(scene)
        color, depth = _ref_render(cs, scene, g, depth_mode)
        loss = (color * scene.grad_color).sum()
        if depth is not None:
            loss = loss + (depth * scene.grad_depth).sum()
        loss.backward()
        arrays = dict(color=color.detach().numpy(), d_means=g.means.grad.numpy(), d_covariances=g.covariances.grad.numpy(),
                      d_harmonics=g.harmonics.grad.numpy(), d_opacities=g.opacities.grad.numpy(),
                      stage_digests=np.array(stage_digests(scene)), input_digest=np.array(input_digest(scene)))
        if depth is not None:
            arrays["depth"] = depth.detach().numpy()
        path = HERE / f"{name}_{depth_mode or 'color'}.npz"
        np.savez_compressed(path, **arrays)
        print(path.name, {k: v.shape for k, v in arrays.items()}, f"{path.stat().st_size / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
