""".ply export of a Gaussian set on the B200 kernels.

`export_ply` has the signature and the file contents of the reference's function of the same name
(/root/reference/src/model/ply_export.py:26-92; called from src/model/model_wrapper.py's test step): Gaussians shifted to
their median, rescaled by the 95 % quantile, rotated into the viewer frame of the given camera, DC band only.  The
per-Gaussian table is built in one launch (fs_ply_vertices); the two order statistics are torch reductions on the device;
the host only writes the header and the bytes (no plyfile dependency).  CPU tensors raise."""
from __future__ import annotations

import ctypes as C
import math
from pathlib import Path

import torch

from . import _lib
from ._lib import check, ptr

PROPERTIES = ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2", "opacity", "scale_0", "scale_1", "scale_2",
              "rot_0", "rot_1", "rot_2", "rot_3"]


class FsPlyArgs(C.Structure):
    _fields_ = [("N", C.c_int32), ("d_sh", C.c_int32), ("scale_factor", C.c_float), ("shift", C.c_float * 3), ("R", C.c_float * 9),
                ("means", C.c_void_p), ("scales", C.c_void_p), ("rotations", C.c_void_p), ("harmonics", C.c_void_p),
                ("opacities", C.c_void_p), ("table", C.c_void_p)]


def viewer_rotation(extrinsics: torch.Tensor) -> torch.Tensor:
    """ply_export.py:43-63, fp32 on the host: Rz(-45 deg) @ [[0,0,1],[-1,0,0],[0,-1,0]] @ inv(extrinsics[:3,:3])."""
    base = torch.tensor([[0, 0, 1], [-1, 0, 0], [0, -1, 0]], dtype=torch.float32)
    a = math.radians(-45.0)
    adj = torch.tensor([[math.cos(a), -math.sin(a), 0], [math.sin(a), math.cos(a), 0], [0, 0, 1]], dtype=torch.float64).float()
    return (adj @ base) @ extrinsics[:3, :3].detach().float().cpu().inverse()


def vertex_table(extrinsics, means, scales, rotations, harmonics, opacities) -> torch.Tensor:
    """[N,17] float32 device tensor, columns = PROPERTIES."""
    if not means.is_cuda:
        raise _lib.FreeSplatB200Error("export_ply needs CUDA tensors (no CPU fallback exists)")
    L = _lib.lib()
    dev = means.device
    f = lambda t: t.detach().float().contiguous()
    means, scales, rotations, harmonics, opacities = f(means), f(scales), f(rotations), f(harmonics), f(opacities)
    N = means.shape[0]
    table = torch.empty((N, 17), dtype=torch.float32, device=dev)
    if N == 0:
        return table
    shift = means.median(dim=0).values
    sf = float((means - shift).abs().quantile(0.95, dim=0).max())
    R = viewer_rotation(extrinsics).reshape(-1).tolist()
    a = FsPlyArgs(N=N, d_sh=harmonics.shape[-1], scale_factor=sf, shift=(C.c_float * 3)(*shift.tolist()), R=(C.c_float * 9)(*R),
                  means=ptr(means), scales=ptr(scales), rotations=ptr(rotations), harmonics=ptr(harmonics), opacities=ptr(opacities),
                  table=ptr(table))
    with torch.cuda.device(dev):
        check(L.fs_ply_vertices(C.byref(a), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "fs_ply_vertices")
    return table


def header(n: int) -> bytes:
    lines = ["ply", "format binary_little_endian 1.0", f"element vertex {n}"] + [f"property float {p}" for p in PROPERTIES] + ["end_header"]
    return ("\n".join(lines) + "\n").encode("ascii")


def export_ply(extrinsics, means, scales, rotations, harmonics, opacities, path: Path):
    table = vertex_table(extrinsics, means, scales, rotations, harmonics, opacities)
    path = Path(path)
    path.parent.mkdir(exist_ok=True, parents=True)
    with open(path, "wb") as fh:
        fh.write(header(table.shape[0]))
        fh.write(table.cpu().numpy().astype("<f4", copy=False).tobytes())
