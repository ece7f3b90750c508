"""Loader of the compact mid-size reference goldens (tests/golden/mid_*.npz, made by tests/golden/make_mid_golden.py from
the reference's own code): re-generates the seeded inputs and verifies their stored checksums."""
import os

import numpy as np
import torch

from freesplat_b200 import synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _chk(t) -> float:
    return float(torch.as_tensor(t).double().sum())


def cost_volume():
    z = np.load(os.path.join(GOLDEN, "mid_cost_volume_v3k2.npz"))
    seed, V, K, C, Hf, Wf, D, cs = (int(x) for x in z["meta"])
    inp = synth.cost_volume_inputs(seed, V, K, C, Hf, Wf)
    mlp = synth.cost_volume_mlp(seed, C)
    wts = torch.randn((V, D, Hf, Wf), generator=torch.Generator().manual_seed(77 + seed))
    got = [_chk(inp[k]) for k in ("cur_feats", "src_feats", "src_extrinsics", "src_Ks", "cur_invK")] + [_chk(w) for w in mlp] + [_chk(wts)]
    assert np.allclose(got, z["checksums"], rtol=1e-12, atol=1e-9), "synthetic inputs drifted from the ones the golden was made with"
    kink = torch.from_numpy(np.unpackbits(z["kink_packed"])[:V * D * Hf * Wf].reshape(V, D, Hf, Wf).astype(bool))
    assert int(kink.sum()) == int(z["kink_rows"])
    # rows next to a LeakyReLU kink carry zero loss weight (counted: z["kink_rows"]; see make_mid_golden.py)
    return z, inp, mlp, wts * (~kink), (V, K, C, Hf, Wf, D, cs)


def ptf():
    z = np.load(os.path.join(GOLDEN, "mid_ptf_v4.npz"))
    seed, V, h, w, S, N = (int(x) for x in z["meta"])
    inp = synth.ptf_inputs(seed, V, h, w)
    gen = torch.Generator().manual_seed(500 + seed)
    wF, wX = torch.randn((1, N, 64), generator=gen), torch.randn((1, N, 3), generator=gen)
    wE, wZ = torch.randn((1, N, 4, 4), generator=gen), torch.randn((1, N), generator=gen)
    got = [_chk(inp["gaussians"][0]), _chk(inp["coords"][0]), _chk(inp["densities"]), _chk(inp["weight_emb"]), _chk(inp["depths"]),
           _chk(inp["extrinsics"]), _chk(wF)]
    assert np.allclose(got, z["checksums"], rtol=1e-12, atol=1e-9), "synthetic inputs drifted from the ones the golden was made with"
    return z, inp, (wF[0], wX[0], wE[0], wZ[0]), (seed, V, h, w, S, N)
