"""CPU: .ply export restatement (oracle/ply.py) against the vertex table of the REFERENCE's own export_ply."""
import glob
import os

import numpy as np
import pytest

from oracle import ply as oply

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ply_*.npz")))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_vertex_table_matches_reference(path):
    z = np.load(path)
    assert list(z["names"]) == oply.PROPERTIES
    got = oply.vertex_table(z["ext"], z["means"], z["scales"], z["rotations"], z["harmonics"], z["opacities"])
    np.testing.assert_allclose(got, z["table"], rtol=2e-5, atol=2e-6)


def test_quaternion_rule_matches_scipy():
    from scipy.spatial.transform import Rotation as R
    g = np.random.default_rng(0)
    q = g.standard_normal((500, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    m = R.from_quat(q).as_matrix()
    np.testing.assert_allclose(oply.quat_to_matrix(q), m, atol=1e-14)
    np.testing.assert_allclose(oply.matrix_to_quat(m), R.from_matrix(m).as_quat(), atol=1e-12)


def test_file_layout():
    t = np.arange(34, dtype=np.float32).reshape(2, 17)
    b = oply.file_bytes(t)
    head, body = b.split(b"end_header\n")
    assert head.startswith(b"ply\nformat binary_little_endian 1.0\nelement vertex 2\nproperty float x\n")
    assert head.count(b"property float ") == 17 and len(body) == 2 * 17 * 4
    assert np.array_equal(np.frombuffer(body, "<f4").reshape(2, 17), t)
