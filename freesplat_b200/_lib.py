"""ctypes binding of libfreesplat_b200.so (C ABI: include/freesplat_b200.h).

There is NO fallback: if the shared library is missing or does not export the ABI the
import of any operator fails loudly.  PyTorch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfreesplat_b200.so")
ABI_VERSION = 4

vp = C.c_void_p


class FsRasterFwdArgs(C.Structure):
    _fields_ = [
        ("P", C.c_int32), ("V", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("sh_degree", C.c_int32), ("M", C.c_int32), ("scale_modifier", C.c_float),
        ("prefiltered", C.c_int32), ("stages", C.c_int32), ("sh_layout", C.c_int32), ("cov_stride", C.c_int32),
        ("capacity", C.c_int64),
        ("means3D", vp), ("shs", vp), ("colors_precomp", vp), ("opacities", vp),
        ("scales", vp), ("rotations", vp), ("cov3D_precomp", vp), ("views", vp),
        ("out_color", vp), ("out_depth", vp), ("final_T", vp), ("n_contrib", vp), ("radii", vp),
        ("rec", vp), ("cov3D", vp), ("tiles_touched", vp), ("clamped", vp),
        ("tile_count", vp), ("tile_cursor", vp), ("ranges", vp), ("keybuf", vp),
        ("point_list", vp), ("status", vp), ("bins", vp), ("bin_cap", C.c_int32),
    ]


class FsRasterBwdArgs(C.Structure):
    _fields_ = [
        ("P", C.c_int32), ("V", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("sh_degree", C.c_int32), ("M", C.c_int32), ("scale_modifier", C.c_float),
        ("has_depth_grad", C.c_int32), ("sh_layout", C.c_int32), ("cov_stride", C.c_int32),
        ("means3D", vp), ("shs", vp), ("colors_precomp", vp), ("opacities", vp),
        ("scales", vp), ("rotations", vp), ("cov3D_precomp", vp), ("views", vp),
        ("rec", vp), ("radii", vp), ("clamped", vp), ("ranges", vp),
        ("point_list", vp), ("final_T", vp), ("n_contrib", vp), ("status", vp),
        ("dL_dcolor", vp), ("dL_ddepth", vp), ("dL_dalpha", vp), ("dL_dscreen", vp),
        ("dL_dmeans2D", vp), ("dL_dmeans3D", vp), ("dL_dcov3D", vp), ("dL_dshs", vp),
        ("dL_dcolors", vp), ("dL_dopacities", vp), ("dL_dscales", vp), ("dL_drotations", vp),
        ("peer_delta", vp), ("shard_rows", C.c_int32), ("world", C.c_int32),
    ]


# every symbol include/freesplat_b200.h declares (tests/test_abi.py checks the header against this)
EXPORTS = [
    "fs_abi_version", "fs_struct_size", "fs_last_error", "fs_device_sm_count",
    "fs_raster_forward", "fs_raster_backward", "fs_mark_visible", "fs_camera_records",
    "fs_cost_volume_forward", "fs_cost_volume_backward",
    "fs_ptf_match", "fs_ptf_merge", "fs_ptf_gru_inputs", "fs_ptf_gru_update", "fs_ptf_gru_output",
    "fs_ptf_view_setup", "fs_ptf_pool_update", "fs_ptf_pool_order", "fs_ptf_pool_gather", "fs_ptf_merge_backward", "fs_ptf_gru_output_backward", "fs_ptf_gru_update_backward", "fs_ptf_gru_inputs_backward", "fs_gaussian_head_backward", "fs_depth_head_backward", "fs_backproject_backward", "fs_graph_capture_begin", "fs_graph_capture_end", "fs_graph_launch", "fs_graph_destroy",
    "fs_ptf_gru", "fs_ptf_gru_wscratch_bytes", "fs_ptf_gru_bwd_data", "fs_ptf_gru_bwd_weights", "fs_gaussian_head", "fs_depth_head", "fs_backproject", "fs_ply_vertices", "fs_image_u8",
]

_lib = None


class FreeSplatB200Error(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Loads the CUDA library; raises (never falls back) when it is unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FreeSplatB200Error(
            f"{LIB_PATH} not found: build it with `python -m freesplat_b200.build` "
            "(there is no CPU / PyTorch fallback for the hot path)")
    L = C.CDLL(LIB_PATH)
    L.fs_abi_version.restype = C.c_int
    L.fs_last_error.restype = C.c_char_p
    if L.fs_abi_version() != ABI_VERSION:
        raise FreeSplatB200Error(f"ABI mismatch: library {L.fs_abi_version()} != binding {ABI_VERSION}")
    for name in EXPORTS:
        if not hasattr(L, name):
            raise FreeSplatB200Error(f"{LIB_PATH} does not export {name}")
        getattr(L, name).restype = C.c_int if name != "fs_last_error" else C.c_char_p
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().fs_last_error().decode(errors="replace")
        raise FreeSplatB200Error(f"{what} failed ({rc}): {msg}")


def ptr(t) -> int | None:
    """Raw device pointer of a (contiguous) torch tensor, or None."""
    return None if t is None else t.data_ptr()
