"""``.ply`` interchange (my_depthsplat_b200/ply.py): the vertex table equals, bit for bit, the one the reference's
UNMODIFIED src/model/ply_export.py::export_ply hands to plyfile (captured with a stand-in ``plyfile`` module, since that
package is not installed here); the file round-trips through ``load_ply``; ``gaussians_from_ply`` rebuilds covariances."""
import importlib
import sys
import types

import numpy as np
import pytest
import torch

from helpers import REFERENCE_SRC, have_reference


def _scene_arrays(n=257, seed=0):
    g = torch.Generator().manual_seed(seed)
    yaw = 0.3
    ext = torch.eye(4)
    ext[:3, :3] = torch.tensor([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]], dtype=torch.float32)
    ext[:3, 3] = torch.tensor([0.1, -0.2, 0.3])
    q = torch.randn(n, 4, generator=g)
    q = q / q.norm(dim=-1, keepdim=True)
    return dict(extrinsics=ext, means=torch.randn(n, 3, generator=g), scales=torch.rand(n, 3, generator=g) * 0.1 + 1e-3, rotations=q,
                harmonics=torch.randn(n, 3, 9, generator=g), opacities=torch.rand(n, generator=g) * 0.98 + 0.01)


@pytest.mark.skipif(not have_reference(), reason="/root/reference not available")
def test_vertex_table_equals_the_reference(tmp_path):
    from my_depthsplat_b200 import ply
    captured = {}

    class PlyElement:
        @staticmethod
        def describe(elements, name):
            captured["elements"], captured["name"] = elements, name
            return elements

    class PlyData:
        def __init__(self, elements):
            pass

        def write(self, path):
            captured["path"] = path

    stub = types.ModuleType("plyfile")
    stub.PlyData, stub.PlyElement = PlyData, PlyElement
    sys.modules["plyfile"] = stub
    try:
        for name, path in [("src", str(REFERENCE_SRC)), ("src.model", str(REFERENCE_SRC / "model"))]:
            pkg = types.ModuleType(name)
            pkg.__path__ = [path]
            sys.modules[name] = pkg
        sys.modules.pop("src.model.ply_export", None)
        ref = importlib.import_module("src.model.ply_export")
        a = _scene_arrays()
        ref.export_ply(a["extrinsics"], a["means"], a["scales"], a["rotations"], a["harmonics"], a["opacities"], tmp_path / "ref.ply")
    finally:
        sys.modules.pop("plyfile", None)
    el = captured["elements"]
    assert captured["name"] == "vertex" and list(el.dtype.names) == ply.construct_list_of_attributes(0)
    want = np.stack([el[n] for n in el.dtype.names], axis=1)
    got = ply._vertex_table(a["extrinsics"], a["means"], a["scales"], a["rotations"], a["harmonics"], a["opacities"])
    np.testing.assert_array_equal(got, want)


def test_file_round_trip_and_header(tmp_path):
    from my_depthsplat_b200 import ply
    a = _scene_arrays(n=100, seed=1)
    a["extrinsics"] = torch.eye(4)  # identity orientation: what comes back is what went in
    path = tmp_path / "scene" / "g.ply"
    ply.export_ply(a["extrinsics"], a["means"], a["scales"], a["rotations"], a["harmonics"], a["opacities"], path)
    head = path.read_bytes().split(b"end_header\n")[0].decode().splitlines()
    assert head[:3] == ["ply", "format binary_little_endian 1.0", "element vertex 100"]
    assert [ln.split()[-1] for ln in head[3:]] == ply.construct_list_of_attributes(0)
    assert path.stat().st_size == len(b"\n".join(map(str.encode, head))) + len(b"\nend_header\n") + 100 * 17 * 4
    d = ply.load_ply(path)
    torch.testing.assert_close(d["means"], a["means"], rtol=0, atol=1e-6)
    torch.testing.assert_close(d["scales"], a["scales"], rtol=1e-6, atol=0)
    torch.testing.assert_close(d["opacities"], a["opacities"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(d["harmonics"][..., 0], a["harmonics"][..., 0], rtol=0, atol=0)
    sign = torch.sign((d["rotations"] * a["rotations"]).sum(-1, keepdim=True))  # q and -q are the same rotation
    torch.testing.assert_close(d["rotations"] * sign, a["rotations"], rtol=0, atol=1e-6)
    g = ply.gaussians_from_ply(path, sh_coeffs=9)
    assert g.means.shape == (1, 100, 3) and g.covariances.shape == (1, 100, 3, 3) and g.harmonics.shape == (1, 100, 3, 9)
    assert float(g.harmonics[..., 1:].abs().max()) == 0.0
    evals = torch.linalg.eigvalsh(g.covariances[0].double())
    torch.testing.assert_close(evals, (a["scales"].double() ** 2).sort(dim=-1).values, rtol=1e-4, atol=1e-9)
    # an ascii file with reordered properties and higher SH bands loads too
    txt = tmp_path / "ascii.ply"
    names = ["opacity", "x", "y", "z", "f_dc_0", "f_dc_1", "f_dc_2"] + [f"f_rest_{i}" for i in range(9)] + ["scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]
    rows = np.arange(2 * len(names), dtype=np.float32).reshape(2, len(names)) / 50
    txt.write_text("ply\nformat ascii 1.0\nelement vertex 2\n" + "".join(f"property float {n}\n" for n in names) + "end_header\n" +
                   "\n".join(" ".join(repr(float(v)) for v in r) for r in rows) + "\n")
    d2 = ply.load_ply(txt)
    assert d2["harmonics"].shape == (2, 3, 4) and float(d2["means"][1, 0]) == float(rows[1, 1])
    np.testing.assert_allclose(d2["harmonics"][0, 1].numpy(), [rows[0, 5], rows[0, 10], rows[0, 11], rows[0, 12]], rtol=1e-6)
