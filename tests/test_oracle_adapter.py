"""CPU: Gaussian-head restatement (oracle/adapter.py) against golden outputs of the REFERENCE's own GaussianAdapter."""
import glob
import os

import numpy as np
import pytest

from oracle import adapter as oad

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "adapter_*.npz")))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_matches_reference(path):
    z = np.load(path)
    _, N, h, w = [int(x) for x in z["meta"]]
    out = oad.forward(z["raw"], z["depths"], z["opac"], z["coords"], z["ext"], z["K"], (h, w))
    for k in ("means", "covariances", "harmonics", "opacities", "scales", "rotations"):
        np.testing.assert_allclose(out[k], z[k], rtol=2e-5, atol=1e-9 if k == "covariances" else 1e-7, err_msg=k)


BP = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "backproject_*.npz")))


@pytest.mark.parametrize("path", BP, ids=[os.path.basename(p) for p in BP])
def test_backproject_matches_reference_bit_for_bit(path):
    """The world coordinates feed PTF's index decisions: the restatement must reproduce the reference's fp32 bits."""
    z = np.load(path)
    _, V, h, w = [int(x) for x in z["meta"]]
    got = oad.backproject(z["depths"], z["K"], z["c2w"], (h, w))
    assert np.array_equal(got.view(np.uint32), z["means"].view(np.uint32))
