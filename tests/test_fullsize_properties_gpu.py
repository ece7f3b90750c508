"""GPU, BASELINE full sizes: comparisons with the oracles where they finish in seconds (PTF config 4, raster backward
config 3, one view of the cost volume) plus size-independent properties."""
import numpy as np
import pytest
import torch

from freesplat_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _cv_module(mlp, Hf=120, Wf=160, D=128):
    from freesplat_b200.cost_volume import AVGFeatureVolumeManager
    m = AVGFeatureVolumeManager(Hf, Wf, num_depth_bins=D, matching_dim_size=48).to(DEV)
    with torch.no_grad():
        for p, w in zip([m.mlp.net[0].weight, m.mlp.net[0].bias, m.mlp.net[2].weight, m.mlp.net[2].bias,
                         m.mlp.net[4].weight, m.mlp.net[4].bias], mlp):
            p.copy_(w)
    return m


def test_cost_volume_full_size_config3():
    """[3 views, K=2, 48x120x160, D=128]: (1) source-order invariance, (2) exact linearity in the last layer,
    (3) tensor-core vs fp32 MLP, (4) agreement with the oracle on a random subset of planes / one view."""
    from freesplat_b200 import cost_volume as cvm
    from oracle import cost_volume as ocv
    inp = synth.cost_volume_inputs(11, 3, 2, 48, 120, 160)
    mlp = synth.cost_volume_mlp(11)
    g = {k: v.to(DEV) for k, v in inp.items()}
    with torch.no_grad():
        out = _cv_module(mlp)(**g)
        assert out.shape == (3, 128, 120, 160) and torch.isfinite(out).all()
        perm = dict(g)
        for k in ("src_feats", "src_extrinsics", "src_poses", "src_Ks"):
            perm[k] = g[k].flip(1).contiguous()
        out_p = _cv_module(mlp)(**perm)
        scale = out.abs().max()
        assert (out - out_p).abs().max() / scale < 2e-5
        mlp2 = [w.clone() for w in mlp]; mlp2[4] *= 2; mlp2[5] *= 2
        assert torch.equal(_cv_module(mlp2)(**g), 2 * out)           # power-of-two scaling is exact in fp32
        old = cvm.MLP_MODE
        try:
            cvm.MLP_MODE = 1
            out_f = _cv_module(mlp)(**g)
        finally:
            cvm.MLP_MODE = old
        assert (out - out_f).abs().max() / scale < 5e-5
    # oracle on view 1 only (a few seconds on the CPU)
    sub = {k: v[1:2] for k, v in inp.items() if k not in ("min_depth", "max_depth")}
    want = ocv.forward(sub["cur_feats"], sub["src_feats"], sub["src_extrinsics"], sub["src_Ks"], sub["cur_invK"],
                       inp["min_depth"], inp["max_depth"], mlp, 128, plane_chunk=8)
    assert (out[1:2].cpu() - want).abs().max() / want.abs().max() < 1e-4


def test_ptf_full_size_identical_views_fuse_completely():
    """640x480: fusing a view with an exact copy of itself matches every pixel with its twin: N stays HW, densities and
    weights double, coordinates / depths stay put (weighted mean of equal values), order preserved."""
    from freesplat_b200 import ptf
    from tests.test_ptf_gpu import _gru
    from tests.ptf_helpers import flat_inputs
    inp = synth.ptf_inputs(5, 1, 480, 640, noise=0.0)
    feats, coords, dens, wemb, depths, ext, K, hw = flat_inputs(inp)
    dup = lambda a: np.concatenate([a, a], 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    with torch.no_grad():
        (F_, X_, E_, Z_), dbg, (D_, W_) = ptf.fuse_views(_gru(5, DEV), t(dup(feats)), t(dup(coords)), t(dup(dens)), t(dup(wemb)),
                                                         t(dup(depths)), t(dup(ext)), t(dup(K)), hw, return_debug=True)
    HW = 480 * 640
    c = dbg[0]["counts"]
    assert c[0] == HW and c[1] + c[2] + c[3] == c[4]
    assert c[2] >= 0.999 * HW and F_.shape[0] <= 1.001 * HW        # (a handful of pixels project onto a rounding tie)
    full = c[2] == HW
    if full:
        assert torch.allclose(X_, t(coords[0]), rtol=0, atol=2e-6) and torch.allclose(Z_, t(depths[0]), rtol=1e-6, atol=0)
        assert torch.allclose(D_, 2 * t(dens[0])) and torch.allclose(W_, 2 * t(wemb[0]))


def test_ptf_full_size_config4_counts():
    """FVT-style 10 views at 640x480: per-step bookkeeping is consistent and the fused set is much smaller than 10*HW."""
    from freesplat_b200 import ptf
    from tests.test_ptf_gpu import _gru
    from tests.ptf_helpers import flat_inputs
    V, h, w = 10, 480, 640
    inp = synth.ptf_inputs(6, V, h, w)
    feats, coords, dens, wemb, depths, ext, K, hw = flat_inputs(inp)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    with torch.no_grad():
        (F_, X_, E_, Z_), dbg, (D_, W_) = ptf.fuse_views(_gru(6, DEV), t(feats), t(coords), t(dens), t(wemb), t(depths), t(ext), t(K),
                                                         hw, return_debug=True)
    n = h * w
    for d in dbg:
        c = d["counts"]
        assert c[0] == n and c[1] + c[2] == c[0] and c[1] + c[2] + c[3] == c[4]
        assert int(d["match"].sum()) == c[2] and int(d["append"].sum()) == c[3]
        n = c[4]
    assert F_.shape[0] == n < 0.6 * V * h * w
    assert torch.isfinite(F_).all() and torch.isfinite(X_).all() and torch.isfinite(E_).all() and (Z_ > 0).all()
    # density is conserved: every view pixel is either appended or added to (at least) one global Gaussian
    assert float(D_.double().sum()) >= float(torch.from_numpy(dens).double().sum()) * (1 - 1e-6)


def test_ptf_full_size_config4_vs_oracle():
    """BASELINE config 4 (10 views of 640x480) through the PUBLIC path (in-kernel canonical inverse, sync-free fold) against
    the CPU restatement: order, coordinates, depths, extrinsics bit-exact; fused latents 1e-4."""
    from freesplat_b200 import ptf
    from oracle import ptf as optf
    from tests.test_ptf_gpu import _gru
    from tests.ptf_helpers import flat_inputs
    V, h, w = 10, 480, 640
    inp = synth.ptf_inputs(6, V, h, w)
    feats, coords, dens, wemb, depths, ext, K, hw = flat_inputs(inp)
    oF, oX, oE, oZ = optf.fuse(feats, coords, dens, wemb, depths, ext, K, hw, optf.torch_gru_fn(synth.gru_state(6)))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    with torch.no_grad():
        F_, X_, E_, Z_ = ptf.fuse_views(_gru(6, DEV), t(feats), t(coords), t(dens), t(wemb), t(depths), t(ext), t(K), hw)
    bits = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)
    assert F_.shape[0] == oF.shape[0]
    assert np.array_equal(bits(X_.cpu().numpy()), bits(oX))
    assert np.array_equal(bits(Z_.cpu().numpy()), bits(oZ))
    assert np.array_equal(bits(E_.cpu().numpy()), bits(oE))
    np.testing.assert_allclose(F_.cpu().numpy(), oF, rtol=1e-4, atol=2e-5)


def test_raster_backward_full_size_config3_vs_oracle():
    """BASELINE config 3 raster size: P = 460 800 (3 context views, one Gaussian per pixel before fusion), 4 target views,
    gradients of an MSE-like loss against the oracle's (fp64-accumulated) backward."""
    from tests import raster_compare as rc
    sc = synth.pixel_aligned_scene(seed=3, h=480, w=640, n_context=3, n_target=4, keep=460800)
    st, views = rc.run_cuda(sc)
    assert st.P == 460800 and st.V == 4
    g = torch.Generator().manual_seed(21)
    dC = torch.randn((st.V, 3, st.H, st.W), generator=g)
    m = rc.compare_backward(sc, st, views, dC)
    assert not rc.backward_ok(m), m
