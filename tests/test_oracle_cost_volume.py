"""CPU: the cost-volume restatement (oracle/cost_volume.py) against golden outputs of the
REFERENCE's own code (tests/golden/cost_volume_*.npz, made by make_cost_volume_golden.py)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import cost_volume as ocv

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "cost_volume_*.npz")))


def load(path):
    z = np.load(path)
    t = lambda k: torch.from_numpy(z[k])
    seed, V, K, C, Hf, Wf, D = [int(x) for x in z["meta"]]
    inp = {k: t(k) for k in ("cur_feats", "src_feats", "src_extrinsics", "src_poses", "src_Ks", "cur_invK", "min_depth", "max_depth")}
    mlp = [t(f"mlp{i}") for i in range(6)]
    return z, inp, mlp, D


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_forward_matches_reference(path):
    z, inp, mlp, D = load(path)
    out = ocv.forward(inp["cur_feats"], inp["src_feats"], inp["src_extrinsics"], inp["src_Ks"], inp["cur_invK"],
                      inp["min_depth"], inp["max_depth"], mlp, D)
    ref = z["out"]
    assert out.shape == ref.shape
    err = np.abs(out.numpy() - ref)
    assert err.max() <= 1e-4 * np.abs(ref).max() + 1e-5, err.max()
    assert np.isclose(out.numpy(), ref, rtol=1e-4, atol=2e-6).mean() > 0.999
    np.testing.assert_allclose(ocv.depth_planes(inp["min_depth"], inp["max_depth"], D).numpy(), z["planes"][0], rtol=1e-6)


@pytest.mark.parametrize("path", GOLD[:2], ids=[os.path.basename(p) for p in GOLD[:2]])
def test_backward_matches_reference(path):
    """autograd through the restatement == the reference's autograd (grid_sample backward etc.)."""
    z, inp, mlp, D = load(path)
    cur = inp["cur_feats"].clone().requires_grad_(True); src = inp["src_feats"].clone().requires_grad_(True)
    mlp = [w.clone().requires_grad_(True) for w in mlp]
    out = ocv.forward(cur, src, inp["src_extrinsics"], inp["src_Ks"], inp["cur_invK"], inp["min_depth"], inp["max_depth"], mlp, D)
    (out * torch.from_numpy(z["wts"])).sum().backward()
    for name, got, want in [("cur", cur.grad, z["g_cur"]), ("src", src.grad, z["g_src"])] + \
            [(f"mlp{i}", mlp[i].grad, z[f"g_mlp{i}"]) for i in range(6)]:
        scale = np.abs(want).max() + 1e-12
        assert np.abs(got.numpy() - want).max() / scale < 2e-4, name
