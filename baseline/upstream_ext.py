"""A `diff_gaussian_rasterization`-shaped module over baseline/upstream_style.cu (GPU comparator, NOT product code).

Same two names and call convention the reference uses (src/model/decoder/cuda_splatting.py:5-8, :98-123):
``GaussianRasterizationSettings`` and ``GaussianRasterizer(settings)(means3D=, means2D=, shs= | colors_precomp=,
opacities=, cov3D_precomp=) -> (image [3,H,W], radii [N])``, differentiable.  Like the extension it restates, one
call renders ONE view, allocates its three working buffers per call through the caching allocator, reads the
pair count back to the host in the middle of the forward and zero-fills every gradient tensor in the backward.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import NamedTuple

import torch
from torch import nn

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libupstream_style.so"
_lib = None


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


class _View(C.Structure):
    _fields_ = [("H", C.c_int), ("W", C.c_int), ("D", C.c_int), ("M", C.c_int), ("tanfovx", C.c_float), ("tanfovy", C.c_float),
                ("view", C.c_void_p), ("proj", C.c_void_p), ("campos", C.c_void_p), ("bg", C.c_void_p)]


_STATE_FIELDS = ["depths", "xy", "conic_opacity", "rgb", "clamped", "radii", "tiles_touched", "offsets", "scan_temp",
                 "scan_temp_bytes", "keys_unsorted", "keys", "vals_unsorted", "vals", "sort_temp", "sort_temp_bytes", "ranges",
                 "final_T", "n_contrib"]


class _State(C.Structure):
    _fields_ = [(n, C.c_size_t if n.endswith("_bytes") else C.c_void_p) for n in _STATE_FIELDS]


def available() -> bool:
    return _LIB_PATH.exists()


def load():
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise RuntimeError(f"{_LIB_PATH} missing: run `make -C baseline`")
        L = C.CDLL(str(_LIB_PATH))
        L.ups_scan_temp_bytes.restype = C.c_size_t
        L.ups_scan_temp_bytes.argtypes = [C.c_int]
        L.ups_sort_temp_bytes.restype = C.c_size_t
        L.ups_sort_temp_bytes.argtypes = [C.c_longlong, C.c_int]
        L.ups_preprocess.restype = C.c_int
        L.ups_preprocess.argtypes = [C.POINTER(_View), C.c_int] + [C.c_void_p] * 4 + [C.POINTER(_State), C.POINTER(C.c_longlong), C.c_void_p]
        L.ups_bin_render.restype = C.c_int
        L.ups_bin_render.argtypes = [C.POINTER(_View), C.c_int, C.c_longlong, C.POINTER(_State), C.c_void_p, C.c_void_p, C.c_void_p]
        L.ups_backward.restype = C.c_int
        L.ups_backward.argtypes = [C.POINTER(_View), C.c_int, C.c_longlong, C.POINTER(_State)] + [C.c_void_p] * 13
        L.ups_struct_sizes.restype = C.c_int
        L.ups_struct_sizes.argtypes = [C.c_int]
        assert L.ups_struct_sizes(0) == C.sizeof(_View) and L.ups_struct_sizes(1) == C.sizeof(_State)
        _lib = L
    return _lib


def _carve(buf: torch.Tensor, sizes):
    """128-byte aligned sub-ranges of one byte buffer -> list of addresses."""
    out, off = [], 0
    base = buf.data_ptr()
    for n in sizes:
        out.append(base + off)
        off += (n + 127) // 128 * 128
    return out


def _total(sizes):
    return sum((n + 127) // 128 * 128 for n in sizes) + 128


def _ptr(t):
    return None if t is None else t.data_ptr()


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, cov3D, rs: GaussianRasterizationSettings):
        L = load()
        dev = means3D.device
        P = means3D.shape[0]
        H, W = int(rs.image_height), int(rs.image_width)
        M = 0 if sh is None else sh.shape[1]
        means3D, opacities, cov3D = means3D.contiguous(), opacities.contiguous(), cov3D.contiguous()
        sh = None if sh is None else sh.contiguous()
        colors_precomp = None if colors_precomp is None else colors_precomp.contiguous()
        viewm, projm, campos, bg = (t.contiguous().float() for t in (rs.viewmatrix, rs.projmatrix, rs.campos, rs.bg))
        view = _View(H, W, int(rs.sh_degree), M, float(rs.tanfovx), float(rs.tanfovy), viewm.data_ptr(), projm.data_ptr(),
                     campos.data_ptr(), bg.data_ptr())
        stream = torch.cuda.current_stream(dev).cuda_stream
        tiles = ((W + 15) // 16) * ((H + 15) // 16)

        st = _State()
        scan_bytes = L.ups_scan_temp_bytes(max(P, 1))
        gsizes = [4 * P, 8 * P, 16 * P, 12 * P, 3 * P, 4 * P, 4 * P, 4 * P, scan_bytes]
        geom = torch.empty(_total(gsizes), dtype=torch.uint8, device=dev)
        (st.depths, st.xy, st.conic_opacity, st.rgb, st.clamped, st.radii, st.tiles_touched, st.offsets, st.scan_temp) = _carve(geom, gsizes)
        st.scan_temp_bytes = scan_bytes
        isizes = [8 * tiles, 4 * H * W, 4 * H * W]
        img = torch.empty(_total(isizes), dtype=torch.uint8, device=dev)
        st.ranges, st.final_T, st.n_contrib = _carve(img, isizes)

        color = torch.zeros((3, H, W), dtype=torch.float32, device=dev)
        radii = torch.zeros((P,), dtype=torch.int32, device=dev)
        st.radii = radii.data_ptr()
        R = C.c_longlong(0)
        rc = L.ups_preprocess(C.byref(view), P, _ptr(means3D), _ptr(sh), _ptr(opacities), _ptr(cov3D), C.byref(st), C.byref(R), stream)
        if rc:
            raise RuntimeError(f"upstream-style preprocess failed: cuda error {rc}")
        R = int(R.value)
        sort_bytes = L.ups_sort_temp_bytes(max(R, 1), tiles)
        bsizes = [8 * R, 8 * R, 4 * R, 4 * R, sort_bytes]
        binning = torch.empty(_total(bsizes), dtype=torch.uint8, device=dev)
        st.keys_unsorted, st.keys, st.vals_unsorted, st.vals, st.sort_temp = _carve(binning, bsizes)
        st.sort_temp_bytes = sort_bytes
        rc = L.ups_bin_render(C.byref(view), P, R, C.byref(st), _ptr(colors_precomp), color.data_ptr(), stream)
        if rc:
            raise RuntimeError(f"upstream-style render failed: cuda error {rc}")
        ctx.keep = (geom, img, binning, radii, viewm, projm, campos, bg, means3D, sh, colors_precomp, cov3D)
        ctx.view, ctx.st, ctx.R, ctx.P, ctx.M = view, st, R, P, M
        ctx.op_shape = opacities.shape
        ctx.mark_non_differentiable(radii)
        return color, radii

    @staticmethod
    def backward(ctx, grad_color, _grad_radii):
        L = load()
        geom, img, binning, radii, viewm, projm, campos, bg, means3D, sh, colors_precomp, cov3D = ctx.keep
        dev, P, M = means3D.device, ctx.P, ctx.M
        z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=dev)
        d_mean3D, d_mean2D, d_color, d_conic, d_opac, d_cov = z(P, 3), z(P, 3), z(P, 3), z(P, 2, 2), z(P, 1), z(P, 6)
        d_sh = z(P, max(M, 1), 3)
        grad_color = grad_color.contiguous()
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = L.ups_backward(C.byref(ctx.view), P, ctx.R, C.byref(ctx.st), _ptr(means3D), _ptr(sh), _ptr(colors_precomp), _ptr(cov3D),
                            grad_color.data_ptr(), d_mean2D.data_ptr(), d_conic.data_ptr(), d_opac.data_ptr(), d_color.data_ptr(),
                            d_mean3D.data_ptr(), d_cov.data_ptr(), d_sh.data_ptr(), stream)
        if rc:
            raise RuntimeError(f"upstream-style backward failed: cuda error {rc}")
        return (d_mean3D, d_mean2D, d_sh if sh is not None else None, None if sh is not None else d_color,
                d_opac.reshape(ctx.op_shape), d_cov, None)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None):
        if (shs is None) == (colors_precomp is None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if scales is not None or rotations is not None or cov3D_precomp is None:
            raise NotImplementedError("only cov3D_precomp is on the DepthSplat path (cuda_splatting.py:122)")
        return _RasterizeGaussians.apply(means3D, means2D, shs, colors_precomp, opacities, cov3D_precomp, self.raster_settings)
