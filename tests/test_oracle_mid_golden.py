"""CPU: the cost-volume and PTF restatements against the MID-SIZE reference goldens (60x80xD32, K = 2; 120x160, V = 4):
the oracles are pinned on outputs of the reference's own code beyond toy sizes."""
import numpy as np
import torch

from freesplat_b200 import synth
from tests import mid_golden

bits = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_cost_volume_oracle_mid():
    from oracle import cost_volume as ocv
    z, inp, mlp, wts, (V, K, C, Hf, Wf, D, cs) = mid_golden.cost_volume()
    cur = inp["cur_feats"].clone().requires_grad_(True); src = inp["src_feats"].clone().requires_grad_(True)
    m = [w.clone().requires_grad_(True) for w in mlp]
    out = ocv.forward(cur, src, inp["src_extrinsics"], inp["src_Ks"], inp["cur_invK"], inp["min_depth"], inp["max_depth"], m, D)
    ref = z["out"]
    assert np.isclose(out.detach().numpy(), ref, rtol=1e-4, atol=1e-5 * np.abs(ref).max()).all()
    (out * wts).sum().backward()
    # rows next to a LeakyReLU kink (counted: z["kink_rows"]) carry zero loss weight; everything else element-wise, no outliers
    from tests.helpers import grad_report
    assert 0 < int(z["kink_rows"]) < 0.1 * V * D * Hf * Wf
    for name, got, want in [("cur", cur.grad[:, ::2], z["g_cur_sub"]), ("src", src.grad[:, :, ::cs], z["g_src_sub"])] + \
            [(f"mlp{i}", m[i].grad, z[f"g_mlp{i}"]) for i in range(6)]:
        rep = grad_report(got.numpy(), want, max_outlier_frac=0.0)
        assert rep["ok"], (name, rep)


def test_ptf_oracle_mid():
    from oracle import ptf as optf
    from tests.ptf_helpers import flat_inputs
    z, inp, _, (seed, V, h, w, S, N) = mid_golden.ptf()
    feats, coords, dens, wemb, depths, ext, K, hw = flat_inputs(inp)
    F_, X_, E_, Z_ = optf.fuse(feats, coords, dens, wemb, depths, ext, K, hw, optf.torch_gru_fn(synth.gru_state(seed)), E_invs=z["E_inv"])
    assert F_.shape[0] == N
    assert np.array_equal(bits(X_), bits(z["out_coords"])) and np.array_equal(bits(Z_), bits(z["out_depths"]))
    assert np.array_equal(bits(E_[::S]), bits(z["out_ext_sub"]))
    assert abs(float(torch.from_numpy(E_).double().sum()) - float(z["out_ext_sum"])) < 1e-6 * N
    np.testing.assert_allclose(F_[::S], z["out_feats_sub"], rtol=1e-4, atol=1e-5)
