"""ctypes binding of libb200splat.so (include/b200splat.h).

The product path has no fallback: if the CUDA library is missing or was built for another ABI,
importing the ops raises.  Build it with ``python -m my_depthsplat_b200.build`` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libb200splat.so"
ABI_VERSION = 11

B200S_OK, B200S_EBADARG, B200S_ECUDA = 0, 1, 3
COV_3X3, COV_UPPER6 = 0, 1
SH_CHANNEL_MAJOR, SH_COEFF_MAJOR = 0, 1
DEPTH_NONE, DEPTH_Z, DEPTH_DISPARITY, DEPTH_LOG = 0, 1, 2, 3
SORT_BINNED, SORT_GLOBAL = 0, 1

_f32p = C.c_void_p  # device pointers travel as integers


class Scene(C.Structure):
    _fields_ = [
        ("num_scenes", C.c_int32), ("num_gaussians", C.c_int32), ("sh_degree", C.c_int32), ("sh_coeffs", C.c_int32),
        ("cov_layout", C.c_int32), ("sh_layout", C.c_int32),
        ("means", _f32p), ("covariances", _f32p), ("harmonics", _f32p), ("colors_precomp", _f32p), ("opacities", _f32p),
        ("raw_head", _f32p), ("raw_depth", _f32p), ("raw_image", _f32p), ("raw_camera", _f32p),
        ("raw_views", C.c_int32), ("raw_h", C.c_int32), ("raw_w", C.c_int32), ("raw_scale_min", C.c_float), ("raw_scale_max", C.c_float),
        ("raw_cooked_out", _f32p),
    ]


class Views(C.Structure):
    _fields_ = [
        ("num_views", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("depth_mode", C.c_int32),
        ("scene_index", C.c_void_p), ("viewmatrix", _f32p), ("projmatrix", _f32p), ("campos", _f32p), ("tanfov", _f32p),
        ("background", _f32p), ("scale", _f32p), ("depth_affine", _f32p), ("depth_clamp", _f32p),
    ]


class Dims(C.Structure):
    _fields_ = [
        ("num_scenes", C.c_int32), ("num_gaussians", C.c_int32), ("num_views", C.c_int32), ("height", C.c_int32),
        ("width", C.c_int32), ("sort_mode", C.c_int32), ("pair_capacity", C.c_int64),
    ]


class Plan(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("tile_bits", C.c_int32), ("view_bits", C.c_int32), ("sort_bits", C.c_int32),
        ("sort_passes", C.c_int32), ("grid_x", C.c_int32), ("grid_y", C.c_int32), ("tiles", C.c_int32), ("bins", C.c_int32),
        ("pre_tickets", C.c_int32), ("sort_tiles_cap", C.c_int32), ("final_in_a", C.c_int32), ("pair_capacity", C.c_int64),
        ("saved_bytes", C.c_size_t), ("off_status", C.c_size_t), ("off_rec", C.c_size_t), ("off_vals_a", C.c_size_t),
        ("off_ranges", C.c_size_t), ("off_final_T", C.c_size_t), ("off_n_contrib", C.c_size_t),
        ("scratch_bytes", C.c_size_t), ("off_keys_a", C.c_size_t), ("off_keys_b", C.c_size_t), ("off_vals_b", C.c_size_t),
        ("off_scan_state", C.c_size_t), ("off_ticket_totals", C.c_size_t), ("off_scan_blocks", C.c_size_t),
        ("off_bin_info", C.c_size_t), ("off_hist", C.c_size_t), ("off_lookback", C.c_size_t), ("off_counters", C.c_size_t),
        ("off_grad_rec", C.c_size_t),
        ("sort_mode", C.c_int32), ("bin_sort_cap", C.c_int32), ("off_bin_count", C.c_size_t), ("off_bin_cursor", C.c_size_t),
        ("off_long_list", C.c_size_t),
    ]


class Status(C.Structure):
    _fields_ = [
        ("num_pairs", C.c_uint64), ("overflow", C.c_uint32), ("num_visible", C.c_uint32), ("tested", C.c_uint64),
        ("blended", C.c_uint64), ("max_tile_len", C.c_uint32), ("max_bin_len", C.c_uint32), ("reserved", C.c_uint32 * 4),
    ]


class Out(C.Structure):
    _fields_ = [("color", _f32p), ("depth", _f32p), ("radii", C.c_void_p), ("count_work", C.c_int32), ("status_host", C.c_void_p),
                ("mse_target", _f32p), ("mse_grad", _f32p), ("mse_partials", _f32p), ("mse_scale", C.c_float), ("mse_l1", C.c_int32)]


class GradOut(C.Structure):
    _fields_ = [("dL_dcolor", _f32p), ("dL_ddepth", _f32p), ("dL_dcolor_scale", _f32p)]


class GradIn(C.Structure):
    _fields_ = [
        ("dL_dmeans", _f32p), ("dL_dcovariances", _f32p), ("dL_dharmonics", _f32p), ("dL_dcolors", _f32p),
        ("dL_dopacities", _f32p), ("dL_dmeans2D", _f32p), ("multicast", C.c_int32),
        ("stages", C.c_int32), ("chunk_begin", C.c_int32), ("chunk_count", C.c_int32), ("chunk_stride", C.c_int32),
        ("chunk_repeat", C.c_int32), ("dL_draw_head", _f32p), ("dL_draw_depth", _f32p),
    ]


# every symbol include/b200splat.h declares (tests check the exports against this list)
EXPORTS = (
    "b200s_plan", "b200s_forward_bin", "b200s_forward_render", "b200s_backward", "b200s_sort_tmp_bytes",
    "b200s_sort_pairs", "b200s_abi_version", "b200s_last_cuda_error", "b200s_build_info", "b200s_profile_enable",
    "b200s_profile_read", "b200s_kernel_launches", "b200s_debug_set", "b200s_p2p_reduce_segments", "b200s_host_alloc", "b200s_host_free",
    "b200s_nvls_allreduce", "b200s_nvls_reduce_segments", "b200s_segment_sort", "b200s_segment_sort_tmp_bytes",
)
STAGES = ("pre_bin", "sort_hist", "sort_passes", "ranges", "comp_fwd", "grad_zero", "comp_bwd", "pre_bwd", "end", "bin_sort")

_lib = None


class LibraryMissing(RuntimeError):
    pass


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise LibraryMissing(
            f"{LIB_PATH} not found: the sm_100a CUDA library is not built. Run `python -m my_depthsplat_b200.build` "
            "(needs nvcc). There is no CPU or PyTorch fallback for the rasterizer."
        )
    L = C.CDLL(str(LIB_PATH))
    for name in EXPORTS:
        if not hasattr(L, name):
            raise LibraryMissing(f"{LIB_PATH} does not export {name}; rebuild it")
    P = C.POINTER
    vp = C.c_void_p
    L.b200s_abi_version.restype = C.c_int
    L.b200s_last_cuda_error.restype = C.c_int
    L.b200s_build_info.restype = C.c_char_p
    L.b200s_plan.restype = C.c_int
    L.b200s_plan.argtypes = [P(Dims), P(Plan)]
    L.b200s_forward_bin.restype = C.c_int
    L.b200s_forward_bin.argtypes = [P(Scene), P(Views), P(Plan), vp, vp, P(Out), vp]
    L.b200s_forward_render.restype = C.c_int
    L.b200s_forward_render.argtypes = [P(Scene), P(Views), P(Plan), vp, vp, P(Out), vp]
    L.b200s_backward.restype = C.c_int
    L.b200s_backward.argtypes = [P(Scene), P(Views), P(Plan), vp, vp, P(Out), P(GradOut), P(GradIn), vp]
    L.b200s_sort_tmp_bytes.restype = C.c_size_t
    L.b200s_sort_tmp_bytes.argtypes = [C.c_int64]
    L.b200s_sort_pairs.restype = C.c_int
    L.b200s_sort_pairs.argtypes = [vp, vp, vp, vp, C.c_int64, C.c_int32, vp, vp]
    L.b200s_profile_enable.restype = None
    L.b200s_profile_enable.argtypes = [C.c_int]
    L.b200s_profile_read.restype = C.c_int
    L.b200s_profile_read.argtypes = [P(C.c_float)]
    L.b200s_kernel_launches.restype = C.c_longlong
    L.b200s_host_alloc.restype = C.c_void_p
    L.b200s_host_alloc.argtypes = [C.c_size_t]
    L.b200s_host_free.restype = None
    L.b200s_host_free.argtypes = [C.c_void_p]
    L.b200s_nvls_allreduce.restype = C.c_int
    L.b200s_nvls_allreduce.argtypes = [C.c_void_p, C.c_ulonglong, C.c_int, C.c_int, C.c_void_p]
    L.b200s_nvls_reduce_segments.restype = C.c_int
    L.b200s_nvls_reduce_segments.argtypes = [vp, vp, P(C.c_ulonglong), P(C.c_ulonglong), C.c_int, vp]
    L.b200s_p2p_reduce_segments.restype = C.c_int
    L.b200s_p2p_reduce_segments.argtypes = [P(C.c_void_p), C.c_int, C.c_int, vp, P(C.c_ulonglong), P(C.c_ulonglong), C.c_int, vp]
    L.b200s_segment_sort_tmp_bytes.restype = C.c_size_t
    L.b200s_segment_sort_tmp_bytes.argtypes = [C.c_int64, C.c_int32]
    L.b200s_segment_sort.restype = C.c_int
    L.b200s_segment_sort.argtypes = [vp, C.c_int32, vp, C.c_int64, vp, vp, vp, vp]
    L.b200s_debug_set.restype = None
    L.b200s_debug_set.argtypes = [C.c_int, C.c_int]
    if L.b200s_abi_version() != ABI_VERSION:
        raise LibraryMissing(f"{LIB_PATH} has ABI {L.b200s_abi_version()}, expected {ABI_VERSION}; rebuild it")
    # debug only: kernel variant switches for A/B runs ("2=1,3=1" -> b200s_debug_set(2, 1), b200s_debug_set(3, 1));
    # variants never change results
    for kv in filter(None, os.environ.get("B200S_KNOBS", "").split(",")):
        k, v = kv.split("=")
        L.b200s_debug_set(int(k), int(v))
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc == B200S_OK:
        return
    if rc == B200S_EBADARG:
        raise ValueError(f"{what}: bad argument (B200S_EBADARG)")
    if rc == B200S_ECUDA:
        raise RuntimeError(f"{what}: CUDA error {load().b200s_last_cuda_error()} (B200S_ECUDA)")
    raise RuntimeError(f"{what}: unknown return code {rc}")


def plan(num_scenes: int, num_gaussians: int, num_views: int, height: int, width: int, pair_capacity: int,
         sort_mode: int = SORT_BINNED) -> Plan:
    d = Dims(num_scenes, num_gaussians, num_views, height, width, sort_mode, pair_capacity)
    p = Plan()
    check(load().b200s_plan(C.byref(d), C.byref(p)), "b200s_plan")
    return p


def profile_enable(on: bool) -> None:
    load().b200s_profile_enable(1 if on else 0)


def profile_read() -> dict:
    """Milliseconds per stage accumulated since the last read (synchronises on the recorded events)."""
    buf = (C.c_float * len(STAGES))()
    load().b200s_profile_read(buf)
    # "end" = time between the end of one library call and the first stage of the next one (host glue, gaps)
    return {name: float(buf[i]) for i, name in enumerate(STAGES)}
