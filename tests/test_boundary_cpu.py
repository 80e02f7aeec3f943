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


def test_plan_regions_are_disjoint_aligned_and_in_bounds():
    """Every region of the two caller-owned workspaces, at its documented size (include/b200splat.h), lies inside the
    workspace, starts 256-byte aligned and overlaps no other region -- for random problem sizes, including pair
    capacities that are no multiple of the sort tile."""
    import random
    rnd = random.Random(7)
    SORT_TILE = 3072
    for _ in range(200):
        B = rnd.choice([1, 1, 2, 8]); N = rnd.randint(1, 3_000_000); V = rnd.randint(1, 12); VV = B * V
        H, W = rnd.randint(1, 2000), rnd.randint(1, 2000)
        if (H + 15) // 16 > 255 or (W + 15) // 16 > 255:
            continue
        cap = rnd.randint(1, 1 << rnd.randint(4, 31))
        mode = rnd.choice([_lib.SORT_BINNED, _lib.SORT_GLOBAL])
        p = _lib.plan(B, N, VV, H, W, cap, mode)
        assert p.sort_mode == mode
        tickets = p.pre_tickets
        assert tickets == VV * ((N + 255) // 256)
        sort_tiles = (cap + SORT_TILE - 1) // SORT_TILE + 1
        saved = [(p.off_status, 64), (p.off_rec, VV * N * 64), (p.off_vals_a, cap * 4), (p.off_ranges, p.bins * 8),
                 (p.off_final_T, VV * H * W * 4), (p.off_n_contrib, VV * H * W * 4)]
        scratch = [(p.off_keys_a, cap * 8), (p.off_keys_b, cap * 8), (p.off_vals_b, cap * 4), (p.off_scan_state, tickets * 8),
                   (p.off_ticket_totals, tickets * 4), (p.off_scan_blocks, (tickets // 2048 + 1) * 8), (p.off_bin_info, tickets * 256 * 8),
                   (p.off_hist, 8 * 256 * 4), (p.off_counters, 64 * 4)]
        if mode == _lib.SORT_GLOBAL:
            scratch += [(p.off_lookback, 2 * sort_tiles * 256 * 8)]
        else:  # per-bin counts, scatter cursors, the four size-class lists of the segment sort
            scratch += [(p.off_bin_count, p.bins * 4), (p.off_bin_cursor, p.bins * 4), (p.off_long_list, 4 * p.bins * 4)]
        for regions, total in ((saved, p.saved_bytes), (scratch, p.scratch_bytes)):
            regions = sorted(regions)
            for (o, n), (o2, _) in zip(regions, regions[1:] + [(total, 0)]):
                assert o % 256 == 0 and o + n <= o2, (B, N, VV, H, W, cap, mode, o, n, o2)
        # the backward's gradient records reuse the scratch from its start
        assert p.off_grad_rec + VV * N * 48 <= p.scratch_bytes
