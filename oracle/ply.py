"""TEST INFRASTRUCTURE -- CPU restatement (numpy) of the reference's Gaussian .ply export
(/root/reference/src/model/ply_export.py:26-92): median shift, 95 %-quantile rescale, the fixed
viewer rotation composed with the inverse camera rotation, quaternion re-orientation (scipy's matrix -> quaternion rule,
restated below so that no scipy call remains), DC band of the harmonics, log scales; and of the file layout plyfile
writes for it (binary_little_endian, 17 float properties).
Pinned against the vertex table the reference's own export_ply produces: tests/golden/ply_*.npz (make_ply_golden.py).
The byte layout of the file is NOT pinned (plyfile is absent here): header restated from the PLY specification.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module."""
from __future__ import annotations

import numpy as np

F32 = np.float32
PROPERTIES = ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2", "opacity", "scale_0", "scale_1", "scale_2",
              "rot_0", "rot_1", "rot_2", "rot_3"]


def viewer_rotation(extrinsics) -> np.ndarray:
    """ply_export.py:43-63: fp32 3x3 = Rz(-45 deg) @ [[0,0,1],[-1,0,0],[0,-1,0]] @ inv(extrinsics[:3,:3])."""
    base = np.array([[0, 0, 1], [-1, 0, 0], [0, -1, 0]], F32)
    a = np.deg2rad(-45.0)
    adj = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]], np.float64).astype(F32)
    return ((adj @ base) @ np.linalg.inv(np.asarray(extrinsics, F32)[:3, :3]).astype(F32)).astype(F32)


def quat_to_matrix(q) -> np.ndarray:
    """scipy Rotation.from_quat(q).as_matrix(): scalar-last, normalised first, fp64."""
    q = np.asarray(q, np.float64)
    q = q / np.linalg.norm(q, axis=-1, keepdims=True)
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    m = np.empty((q.shape[0], 3, 3))
    m[:, 0, 0] = x * x - y * y - z * z + w * w; m[:, 1, 0] = 2 * (x * y + z * w); m[:, 2, 0] = 2 * (x * z - y * w)
    m[:, 0, 1] = 2 * (x * y - z * w); m[:, 1, 1] = -x * x + y * y - z * z + w * w; m[:, 2, 1] = 2 * (y * z + x * w)
    m[:, 0, 2] = 2 * (x * z + y * w); m[:, 1, 2] = 2 * (y * z - x * w); m[:, 2, 2] = -x * x - y * y + z * z + w * w
    return m


def matrix_to_quat(m) -> np.ndarray:
    """scipy Rotation.from_matrix(m).as_quat() (scalar-last, sign as computed): the largest of (m00, m11, m22, trace)
    selects the branch."""
    m = np.asarray(m, np.float64)
    out = np.empty((m.shape[0], 4))
    for n in range(m.shape[0]):
        a = m[n]
        dec = [a[0, 0], a[1, 1], a[2, 2], a[0, 0] + a[1, 1] + a[2, 2]]
        c = int(np.argmax(dec))
        q = np.empty(4)
        if c != 3:
            i, j, k = c, (c + 1) % 3, (c + 2) % 3
            q[i] = 1 - dec[3] + 2 * a[i, i]; q[j] = a[j, i] + a[i, j]; q[k] = a[k, i] + a[i, k]; q[3] = a[k, j] - a[j, k]
        else:
            q[0] = a[2, 1] - a[1, 2]; q[1] = a[0, 2] - a[2, 0]; q[2] = a[1, 0] - a[0, 1]; q[3] = 1 + dec[3]
        out[n] = q / np.linalg.norm(q)
    return out


def lower_median(x) -> np.ndarray:
    """torch.median(dim=0).values: the lower of the two middle elements."""
    s = np.sort(np.asarray(x, F32), axis=0)
    return s[(s.shape[0] - 1) // 2]


def quantile95_max(x) -> np.float32:
    """means.abs().quantile(0.95, dim=0).max() (torch: linear interpolation)."""
    s = np.sort(np.abs(np.asarray(x, F32)), axis=0)
    pos = F32(0.95) * F32(s.shape[0] - 1)
    lo = int(np.floor(pos)); hi = min(lo + 1, s.shape[0] - 1)
    t = F32(pos - F32(lo))
    return F32((s[lo] + (s[hi] - s[lo]) * t).max())


def vertex_table(extrinsics, means, scales, rotations, harmonics, opacities) -> np.ndarray:
    """[N,17] float32 rows in the order of PROPERTIES."""
    means = np.asarray(means, F32); scales = np.asarray(scales, F32)
    means = means - lower_median(means)
    sf = quantile95_max(means)
    means = (means / sf).astype(F32); scales = (scales / sf).astype(F32)
    R = viewer_rotation(extrinsics)
    means = (means @ R.T).astype(F32)
    q = matrix_to_quat(R.astype(np.float64)[None] @ quat_to_matrix(np.asarray(rotations, F32)))
    rot = np.stack([q[:, 3], q[:, 0], q[:, 1], q[:, 2]], axis=-1)
    N = means.shape[0]
    return np.concatenate([means, np.zeros((N, 3), F32), np.asarray(harmonics, F32)[..., 0], np.asarray(opacities, F32)[:, None],
                           np.log(scales), rot.astype(F32)], axis=1).astype(F32)


def header(n: int) -> bytes:
    lines = ["ply", "format binary_little_endian 1.0", f"element vertex {n}"] + [f"property float {p}" for p in PROPERTIES] + ["end_header"]
    return ("\n".join(lines) + "\n").encode("ascii")


def file_bytes(table: np.ndarray) -> bytes:
    return header(table.shape[0]) + np.ascontiguousarray(table, dtype="<f4").tobytes()
