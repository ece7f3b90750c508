"""CPU: the PTF restatement (oracle/ptf.py) against golden outputs of the REFERENCE's own
fuse_gaussians (tests/golden/ptf_*.npz).  Index decisions must coincide exactly (same output length and
ordering); coordinates / depths / extrinsics bit-for-bit; GRU features within fp32 tolerance."""
import glob
import os

import numpy as np
import pytest

from freesplat_b200 import synth
from oracle import ptf as optf
from tests.ptf_helpers import flat_inputs, torch_inverses

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ptf_*.npz")))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_matches_reference(path):
    z = np.load(path)
    seed = int(z["meta"][0])
    feats, coords, dens, wemb, depths, ext, K, hw = flat_inputs(z)
    gru = optf.torch_gru_fn(synth.gru_state(seed))
    F_, X_, E_, Z_ = optf.fuse(feats, coords, dens, wemb, depths, ext, K, hw, gru, E_invs=torch_inverses(ext))
    assert F_.shape[0] == z["out_feats"].shape[1], (F_.shape, z["out_feats"].shape)
    bits = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)
    assert np.array_equal(bits(X_), bits(z["out_coords"][0]))
    assert np.array_equal(bits(Z_), bits(z["out_depths"][0]))
    assert np.array_equal(bits(E_), bits(z["out_ext"][0]))
    np.testing.assert_allclose(F_, z["out_feats"][0], rtol=1e-4, atol=1e-5)


def test_positional_encoding_layout():
    x = np.array([[0.3, 1.7]], np.float32)
    pe = optf.positional_encoding(x, 6)
    assert pe.shape == (1, 24)
    np.testing.assert_allclose(pe[0, 0], np.sin(0.3), rtol=1e-6)
    np.testing.assert_allclose(pe[0, 1], np.cos(0.3), rtol=1e-6)
    np.testing.assert_allclose(pe[0, 2], np.sin(0.6), rtol=1e-6)
    np.testing.assert_allclose(pe[0, 12], np.sin(1.7), rtol=1e-6)
