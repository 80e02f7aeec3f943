"""CPU checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol the
header declares, its structs have the layout the ctypes mirror assumes, the workspace plan is sane,
and the product path refuses to run without CUDA (no fallback)."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest
import torch

from my_depthsplat_b200 import _lib
from my_depthsplat_b200.scenes import make_scene

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "b200splat.h"


def _declared_functions():
    txt = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(b200s_[a-z0-9_]+)\s*\(", txt)))


def test_library_loads_and_exports_every_declared_symbol():
    L = _lib.load()
    declared = _declared_functions()
    assert len(declared) >= 10
    for name in declared:
        assert hasattr(L, name), f"{name} is declared in include/b200splat.h but not exported"
    assert sorted(_lib.EXPORTS) == declared, "the ctypes mirror and the header disagree on the entry points"
    assert L.b200s_abi_version() == _lib.ABI_VERSION
    assert b"sm_100a" in L.b200s_build_info()


def test_struct_layouts_match_the_header(tmp_path):
    """sizeof / offsetof of every struct, compiled from the header with gcc, against the ctypes mirror."""
    structs = {"B200sScene": _lib.Scene, "B200sViews": _lib.Views, "B200sDims": _lib.Dims, "B200sPlan": _lib.Plan,
               "B200sStatus": _lib.Status, "B200sOut": _lib.Out, "B200sGradOut": _lib.GradOut, "B200sGradIn": _lib.GradIn}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for cname, cls in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["/usr/bin/gcc", str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"


@pytest.mark.parametrize("H,W,tiles,tile_bits", [(256, 256, 256, 8), (512, 960, 1920, 11), (50, 70, 20, 5)])
def test_plan(H, W, tiles, tile_bits):
    p = _lib.plan(1, 131072, 4, H, W, 1 << 20)
    assert (p.tiles, p.tile_bits, p.view_bits) == (tiles, tile_bits, 2)
    assert p.sort_bits == 32 + tile_bits + 2 and p.sort_passes == (p.sort_bits + 7) // 8
    assert p.bins == 4 << tile_bits and p.pre_tickets == 4 * 512
    offs = [p.off_status, p.off_rec, p.off_vals_a, p.off_ranges, p.off_final_T, p.off_n_contrib]
    assert offs == sorted(offs) and p.off_n_contrib + 4 * 4 * H * W <= p.saved_bytes
    assert all(o % 256 == 0 for o in offs)
    assert p.off_rec + 4 * 131072 * 64 <= p.off_vals_a
    assert p.scratch_bytes >= p.off_counters + 256 and p.scratch_bytes >= 4 * 131072 * 48


def test_plan_rejects_bad_arguments():
    with pytest.raises(ValueError):
        _lib.plan(1, 0, 1, 16, 16, 1024)
    with pytest.raises(ValueError):
        _lib.plan(1, 100, 1, 16, 16, 0)
    with pytest.raises(ValueError):
        _lib.plan(1, 100, 1, 16, 16, 1 << 32)       # list positions are 32-bit
    with pytest.raises(ValueError):
        _lib.plan(1, 100, 1, 16 * 300, 16, 1024)     # rect packing: at most 255 tiles per axis


def test_null_arguments_are_bad_arguments_not_crashes():
    L = _lib.load()
    assert L.b200s_plan(None, None) == _lib.B200S_EBADARG
    assert L.b200s_forward_bin(None, None, None, None, None, None, None) == _lib.B200S_EBADARG
    assert L.b200s_backward(None, None, None, None, None, None, None, None, None) == _lib.B200S_EBADARG
    assert L.b200s_sort_pairs(None, None, None, None, 10, 47, None, None) == _lib.B200S_EBADARG


def test_no_cpu_fallback():
    """CPU tensors are refused loudly; there is no eager / PyTorch path behind the API."""
    from my_depthsplat_b200 import cuda_splatting as cs
    from my_depthsplat_b200.decoder_splatting_cuda import DecoderSplattingCUDACfg, get_decoder
    sc = make_scene("tiny")
    g = sc.gaussians
    with pytest.raises(RuntimeError, match="no CPU path"):
        cs.render_views(sc.extrinsics, sc.intrinsics, sc.near, sc.far, sc.image_shape, sc.background, g.means, g.covariances,
                        g.harmonics, g.opacities)
    dec = get_decoder(DecoderSplattingCUDACfg(name="splatting_cuda"), type("D", (), {"background_color": [0, 0, 0]})())
    with pytest.raises(RuntimeError, match="no CPU path"):
        dec.forward(g, sc.extrinsics, sc.intrinsics, sc.near, sc.far, sc.image_shape)


def test_missing_library_is_an_error(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(_lib.LibraryMissing, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under my_depthsplat_b200/ may import, include or load it."""
    bad = re.compile(r"^\s*(from\s+oracle|import\s+oracle|from\s+\.+oracle|#\s*include\s+[\"<].*oracle)|libsplat_oracle|splat_oracle\.so", re.M)
    for f in (ROOT / "my_depthsplat_b200").rglob("*"):
        if f.suffix in (".py", ".cu", ".cuh", ".h") and f.is_file():
            assert not bad.search(f.read_text()), f


def test_product_never_imports_the_gpu_comparator():
    """baseline/ (the upstream-style restatement with CUB that bench.py times next to the product) is a comparator:
    nothing under my_depthsplat_b200/ may import, include or load it; and it uses none of the product's kernels."""
    bad = re.compile(r"^\s*(from\s+baseline|import\s+baseline|#\s*include\s+[\"<].*baseline)|libupstream_style|cub/", re.M)
    for f in (ROOT / "my_depthsplat_b200").rglob("*"):
        if f.suffix in (".py", ".cu", ".cuh", ".h") and f.is_file():
            assert not bad.search(f.read_text()), f
    src = (ROOT / "baseline" / "upstream_style.cu").read_text()
    assert "my_depthsplat_b200" not in src and "b200splat" not in src


def test_gpu_comparator_library_loads():
    from baseline import upstream_ext
    if not upstream_ext.available():
        pytest.skip("baseline/_build/libupstream_style.so not built (make -C baseline)")
    L = upstream_ext.load()
    for name in ("ups_preprocess", "ups_bin_render", "ups_backward", "ups_scan_temp_bytes", "ups_sort_temp_bytes"):
        assert hasattr(L, name)
