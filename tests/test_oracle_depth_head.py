"""CPU: depth-head restatement (oracle/depth_head.py) against golden outputs of the REFERENCE's own DepthDecoder."""
import glob
import os

import numpy as np
import pytest

from oracle import depth_head as odh

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "depth_head_*.npz")))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_matches_reference(path):
    z = np.load(path)
    logp = bool(z["meta"][5])
    for s in range(4):
        out = odh.forward(z[f"logits_s{s}"], z["candi"], logp, upsample=(s == 0))
        np.testing.assert_allclose(out["expect"], z[f"log_depth_s{s}"], rtol=2e-6, atol=2e-6, err_msg=f"E s{s}")
        np.testing.assert_allclose(out["depth"], z[f"depth_s{s}"], rtol=1e-5, err_msg=f"depth s{s}")
        if s == 0:
            np.testing.assert_allclose(out["depth_up"], z["depth_up"], rtol=1e-5, err_msg="depth_up")
            np.testing.assert_allclose(out["weights_up"], z["weights_up"], rtol=1e-5, atol=1e-7, err_msg="weights_up")


def test_upsample_matches_torch_on_odd_sizes():
    import torch
    import torch.nn.functional as F
    g = np.random.default_rng(0)
    for h, w in [(1, 1), (1, 5), (7, 3), (33, 41), (240, 320)]:
        x = g.standard_normal((1, 2, h, w)).astype(np.float32)
        want = F.interpolate(torch.from_numpy(x), scale_factor=2, mode="bilinear", align_corners=True).numpy()
        np.testing.assert_allclose(odh.upsample2(x), want, rtol=1e-5, atol=1e-6)
