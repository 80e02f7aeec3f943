"""ctypes front-end of the CPU oracle (oracle/splat_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of splat_oracle.c.  Nothing under my_depthsplat_b200/
imports this module.  PARITY UNPINNED for the per-view arithmetic (third-party extension absent);
the Python glue above it is pinned by tests/golden/.

Per-view API mirrors what the reference binds at src/model/decoder/cuda_splatting.py:98-123:
``forward_view`` ~ ``_C.rasterize_gaussians``; ``backward_view`` ~ ``_C.rasterize_gaussians_backward``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "libsplat_oracle.so"


def build(force: bool = False) -> Path:
    """Compile the oracle with the committed Makefile (gcc only)."""
    src = _HERE / "splat_oracle.c"
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_SO))
        vp, i32, f32, i64 = C.c_void_p, C.c_int, C.c_float, C.c_int64
        L.orc_forward.restype = vp
        L.orc_forward.argtypes = [i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, f32, f32]
        L.orc_backward.restype = None
        L.orc_backward.argtypes = [vp] * 13
        L.orc_set_parallel_backward.restype = None
        L.orc_set_parallel_backward.argtypes = [i32]
        L.orc_free.restype = None
        L.orc_free.argtypes = [vp]
        for name in ("orc_num_rendered", "orc_tested", "orc_blended"):
            getattr(L, name).restype = i64
            getattr(L, name).argtypes = [vp]
        for name in (
            "orc_out_color", "orc_radii", "orc_depths", "orc_xy", "orc_conic_opacity", "orc_rgb",
            "orc_clamped", "orc_tiles_touched", "orc_offsets", "orc_keys_unsorted", "orc_vals_unsorted",
            "orc_keys", "orc_vals", "orc_ranges", "orc_final_T", "orc_n_contrib", "orc_fragile",
        ):
            getattr(L, name).restype = vp
            getattr(L, name).argtypes = [vp]
        _lib = L
    return _lib


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _view(ptr, dtype, shape) -> np.ndarray:
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape, dtype=dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape).copy()


class ViewState:
    """Everything one forward of one view produced (all stage outputs)."""

    def __init__(self, handle, P, M, H, W):
        L = lib()
        self._h = handle
        self.P, self.M, self.H, self.W = P, M, H, W
        gx, gy = (W + 15) // 16, (H + 15) // 16
        self.grid = (gx, gy)
        R = self.num_rendered = int(L.orc_num_rendered(handle))
        self.tested = int(L.orc_tested(handle))
        self.blended = int(L.orc_blended(handle))
        self.color = _view(L.orc_out_color(handle), np.float32, (3, H, W))
        self.radii = _view(L.orc_radii(handle), np.int32, (P,))
        self.depths = _view(L.orc_depths(handle), np.float32, (P,))
        self.xy = _view(L.orc_xy(handle), np.float32, (P, 2))
        self.conic_opacity = _view(L.orc_conic_opacity(handle), np.float32, (P, 4))
        self.rgb = _view(L.orc_rgb(handle), np.float32, (P, 3))
        self.clamped = _view(L.orc_clamped(handle), np.uint8, (P, 3))
        self.tiles_touched = _view(L.orc_tiles_touched(handle), np.uint32, (P,))
        self.offsets = _view(L.orc_offsets(handle), np.uint32, (P,))
        self.keys_unsorted = _view(L.orc_keys_unsorted(handle), np.uint64, (R,))
        self.vals_unsorted = _view(L.orc_vals_unsorted(handle), np.uint32, (R,))
        self.keys = _view(L.orc_keys(handle), np.uint64, (R,))
        self.vals = _view(L.orc_vals(handle), np.uint32, (R,))
        self.ranges = _view(L.orc_ranges(handle), np.uint32, (gx * gy, 2))
        self.final_T = _view(L.orc_final_T(handle), np.float32, (H, W))
        self.n_contrib = _view(L.orc_n_contrib(handle), np.uint32, (H, W))
        self.fragile = _view(L.orc_fragile(handle), np.uint8, (H, W))

    def close(self):
        if self._h:
            lib().orc_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def forward_view(*, H, W, bg, means3D, opacities, cov3D, viewmatrix, projmatrix, campos, tanfovx, tanfovy,
                 shs=None, colors_precomp=None, sh_degree=0) -> ViewState:
    """shs [P,M,3] or colors_precomp [P,3]; cov3D [P,6]; matrices flat[16] in the transposed
    (column-major) storage the reference passes."""
    assert (shs is None) != (colors_precomp is None)
    means3D = _f32(means3D).reshape(-1, 3)
    P = means3D.shape[0]
    M = 0
    if shs is not None:
        shs = _f32(shs)
        M = shs.shape[-2]
        shs = shs.reshape(P, M, 3)
        assert (sh_degree + 1) ** 2 <= M
    else:
        colors_precomp = _f32(colors_precomp).reshape(P, 3)
    keep = [_f32(bg).reshape(3), means3D, shs, colors_precomp, _f32(opacities).reshape(P), _f32(cov3D).reshape(P, 6),
            _f32(viewmatrix).reshape(16), _f32(projmatrix).reshape(16), _f32(campos).reshape(3)]
    h = lib().orc_forward(P, int(sh_degree), M, int(H), int(W), *[_ptr(a) for a in keep], float(tanfovx), float(tanfovy))
    st = ViewState(h, P, M, int(H), int(W))
    st._inputs = dict(means3D=means3D, shs=shs, colors_precomp=colors_precomp, cov3D=keep[5])
    return st


def backward_view(st: ViewState, dL_dpix) -> dict:
    """dL_dpix [3,H,W] -> dict of per-Gaussian gradients (fp32)."""
    P, M = st.P, st.M
    g = _f32(dL_dpix).reshape(3, st.H, st.W)
    out = dict(
        means2D=np.zeros((P, 3), np.float32), conic=np.zeros((P, 4), np.float32), opacity=np.zeros((P,), np.float32),
        colors=np.zeros((P, 3), np.float32), means3D=np.zeros((P, 3), np.float32), cov3D=np.zeros((P, 6), np.float32),
        sh=np.zeros((P, max(M, 1), 3), np.float32),
    )
    i = st._inputs
    lib().orc_backward(st._h, _ptr(i["means3D"]), _ptr(i["shs"]), _ptr(i["colors_precomp"]), _ptr(i["cov3D"]), _ptr(g),
                       _ptr(out["means2D"]), _ptr(out["conic"]), _ptr(out["opacity"]), _ptr(out["colors"]),
                       _ptr(out["means3D"]), _ptr(out["cov3D"]), _ptr(out["sh"]) if M > 0 else None)
    if M == 0:
        out["sh"] = None
    return out


def set_parallel_backward(on: bool) -> None:
    """OpenMP + atomic adds in the compositing backward (timing only; the parity tests use the serial order)."""
    lib().orc_set_parallel_backward(1 if on else 0)


def set_threads(n: int | None = None) -> int:
    """Size of the OpenMP team for the next calls (default: every core this process may run on).  Goes through
    omp_set_num_threads, so it also works after libgomp was initialised with OMP_NUM_THREADS=1 (torchrun sets that)."""
    L = lib()
    L.orc_set_threads.restype = None
    L.orc_set_threads.argtypes = [C.c_int]
    L.orc_max_threads.restype = C.c_int
    L.orc_set_threads(int(n or len(os.sched_getaffinity(0))))
    return int(L.orc_max_threads())
