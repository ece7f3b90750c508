"""TEST INFRASTRUCTURE -- numpy/ctypes front end of oracle/raster_oracle.c.

CPU restatement of the rasterizer FreeSplat calls at
/root/reference/src/model/decoder/cuda_splatting.py:100-127 (third-party
`diff_gaussian_rasterization_depth`, absent from /root/reference: PARITY
UNPINNED, see the header of raster_oracle.c and DESIGN.md).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libfs_oracle.so")
    src = os.path.join(_HERE, "raster_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-B", "-C", _HERE], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.fso_raster_preprocess.restype = C.c_int64
        _LIB.fso_num_threads.restype = C.c_int
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float32))


@dataclass
class RasterState:
    """Everything the forward produced (the comparables of SURVEY Appendix A.2)."""
    P: int
    H: int
    W: int
    R: int
    radii: np.ndarray
    depths: np.ndarray
    xy: np.ndarray
    conic_opacity: np.ndarray
    rgb: np.ndarray
    clamped: np.ndarray
    cov3D: np.ndarray
    tiles_touched: np.ndarray
    offsets: np.ndarray
    keys: np.ndarray
    point_list: np.ndarray
    ranges: np.ndarray
    color: np.ndarray
    depth: np.ndarray
    final_T: np.ndarray
    n_contrib: np.ndarray


def forward(*, H, W, tanfovx, tanfovy, bg, viewmatrix, projmatrix, campos, means3D, opacities,
            shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None,
            sh_degree=0, scale_modifier=1.0, render=True) -> RasterState:
    """viewmatrix/projmatrix: flat 16 floats exactly as the reference passes them
    (transposed, cuda_splatting.py:85-87)."""
    L = lib()
    means3D = _f32(means3D); P = means3D.shape[0]
    opacities = _f32(opacities).reshape(-1)
    shs = _f32(shs); colors_precomp = _f32(colors_precomp)
    scales = _f32(scales); rotations = _f32(rotations); cov3D_precomp = _f32(cov3D_precomp)
    assert (shs is None) != (colors_precomp is None)
    assert (cov3D_precomp is None) != (scales is None or rotations is None)
    M = 0 if shs is None else shs.shape[1]
    vm = _f32(viewmatrix).reshape(16); pm = _f32(projmatrix).reshape(16)
    cp = _f32(campos).reshape(3); bg = _f32(bg).reshape(3)
    radii = np.zeros(P, np.int32); depths = np.zeros(P, np.float32)
    xy = np.zeros((P, 2), np.float32); co = np.zeros((P, 4), np.float32)
    rgb = np.zeros((P, 3), np.float32); clamped = np.zeros((P, 3), np.int32)
    cov3D = np.zeros((P, 6), np.float32)
    tt = np.zeros(P, np.uint32); off = np.zeros(P, np.uint32)
    R = L.fso_raster_preprocess(
        C.c_int(P), C.c_int(sh_degree), C.c_int(M), C.c_int(H), C.c_int(W),
        C.c_float(tanfovx), C.c_float(tanfovy), C.c_float(scale_modifier),
        _p(means3D), _p(shs), _p(colors_precomp), _p(opacities), _p(scales), _p(rotations),
        _p(cov3D_precomp), _p(vm), _p(pm), _p(cp),
        _p(radii), _p(depths), _p(xy), _p(co), _p(rgb), _p(clamped), _p(cov3D), _p(tt), _p(off))
    gx, gy = (W + 15) // 16, (H + 15) // 16
    keys = np.zeros(max(R, 1), np.uint64); pl = np.zeros(max(R, 1), np.uint32)
    ranges = np.zeros((gx * gy, 2), np.uint32)
    L.fso_raster_bin(C.c_int(P), C.c_int(H), C.c_int(W), C.c_int64(R), _p(radii), _p(depths), _p(xy),
                     _p(off), _p(keys), _p(pl), _p(ranges))
    color = np.zeros((3, H, W), np.float32); depth = np.zeros((H, W), np.float32)
    fT = np.zeros((H, W), np.float32); nc = np.zeros((H, W), np.uint32)
    if render:
        L.fso_raster_render(C.c_int(H), C.c_int(W), _p(bg), _p(ranges), _p(pl), _p(xy), _p(co), _p(rgb),
                            _p(depths), _p(color), _p(depth), _p(fT), _p(nc))
    return RasterState(P, H, W, int(R), radii, depths, xy, co, rgb, clamped, cov3D, tt, off,
                       keys[:R], pl[:R], ranges, color, depth, fT, nc)


def backward(st: RasterState, *, tanfovx, tanfovy, bg, viewmatrix, projmatrix, campos, means3D,
             dL_dcolor, dL_ddepth=None, shs=None, colors_precomp=None, scales=None, rotations=None,
             sh_degree=0, scale_modifier=1.0, dL_dalpha=None) -> dict:
    L = lib()
    P, H, W = st.P, st.H, st.W
    means3D = _f32(means3D); shs = _f32(shs); colors_precomp = _f32(colors_precomp)
    scales = _f32(scales); rotations = _f32(rotations)
    M = 0 if shs is None else shs.shape[1]
    vm = _f32(viewmatrix).reshape(16); pm = _f32(projmatrix).reshape(16)
    cp = _f32(campos).reshape(3); bg = _f32(bg).reshape(3)
    dL_dcolor = _f32(dL_dcolor); dL_ddepth = _f32(dL_ddepth); dL_dalpha = _f32(dL_dalpha)
    g = dict(
        means2D=np.zeros((P, 3), np.float32), conic=np.zeros((P, 4), np.float32),
        opacities=np.zeros((P, 1), np.float32), colors=np.zeros((P, 3), np.float32),
        means3D=np.zeros((P, 3), np.float32), cov3D=np.zeros((P, 6), np.float32),
        shs=None if shs is None else np.zeros((P, M, 3), np.float32),
        scales=None if scales is None else np.zeros((P, 3), np.float32),
        rotations=None if rotations is None else np.zeros((P, 4), np.float32),
    )
    pl = st.point_list if st.R else np.zeros(1, np.uint32)
    L.fso_raster_backward(
        C.c_int(P), C.c_int(sh_degree), C.c_int(M), C.c_int(H), C.c_int(W),
        C.c_float(tanfovx), C.c_float(tanfovy), C.c_float(scale_modifier),
        _p(bg), _p(means3D), _p(shs), _p(colors_precomp), _p(scales), _p(rotations), _p(st.cov3D),
        _p(vm), _p(pm), _p(cp), _p(st.radii), _p(st.xy), _p(st.conic_opacity), _p(st.rgb), _p(st.depths),
        _p(st.clamped), _p(st.ranges), _p(pl), _p(st.final_T), _p(st.n_contrib),
        _p(dL_dcolor), _p(dL_ddepth), _p(dL_dalpha),
        _p(g["means2D"]), _p(g["conic"]), _p(g["opacities"]), _p(g["colors"]), _p(g["means3D"]),
        _p(g["cov3D"]), _p(g["shs"]), _p(g["scales"]), _p(g["rotations"]))
    return g


def num_threads() -> int:
    return lib().fso_num_threads()


def set_num_threads(n: int) -> None:
    lib().fso_set_num_threads(C.c_int(n))
