"""GPU parity tests of Pixel-wise Triplet Fusion: CUDA kernels (through the C ABI) vs (a) golden outputs of
the reference's own fuse_gaussians and (b) the CPU restatement at a larger size.  Index decisions,
output ordering, merged coordinates / depths / extrinsics / densities: bit-exact.  GRU features: 1e-4
(cuBLAS vs CPU GEMM summation order)."""
import glob
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from freesplat_b200 import synth
from tests.ptf_helpers import flat_inputs, torch_inverses

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ptf_*.npz")))
bits = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)


class GRU(torch.nn.Module):
    """Same structure / parameter names as networks.py:188-214 (the reference module is not on the GPU box)."""

    def __init__(self, input_channel=64, hidden_channel=64, weights_dim=24):
        super().__init__()
        mk = lambda din: torch.nn.Sequential(torch.nn.Linear(din, hidden_channel), torch.nn.ReLU(),
                                             torch.nn.Linear(hidden_channel, hidden_channel))
        self.mlp_z = mk(hidden_channel + input_channel + 2 * weights_dim)
        self.mlp_r = mk(hidden_channel + input_channel + 2 * weights_dim)
        self.mlp_n = mk(hidden_channel + input_channel + weights_dim)

    def forward(self, input_feat, hidden_feat, input_weights_emb, hidden_weights_emb):
        input_feat_1 = torch.cat((input_feat, input_weights_emb), dim=-1)
        hidden_feat_1 = torch.cat((hidden_feat, hidden_weights_emb), dim=-1)
        concat_input = torch.cat((hidden_feat_1, input_feat_1), dim=-1)
        r = torch.sigmoid(self.mlp_r(concat_input))
        z = torch.sigmoid(self.mlp_z(concat_input))
        q = torch.tanh(self.mlp_n(torch.cat((r * hidden_feat, input_feat_1), dim=-1)))
        return (1 - z) * hidden_feat + z * q


def _gru(seed, dev):
    g = GRU()
    g.load_state_dict(synth.gru_state(seed))
    return g.to(dev)


def _run(feats, coords, dens, wemb, depths, ext, K, hw, seed, debug=False):
    from freesplat_b200 import ptf
    dev = "cuda:0"
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    with torch.no_grad():
        return ptf.fuse_views(_gru(seed, dev), t(feats), t(coords), t(dens), t(wemb), t(depths), t(ext), t(K), hw,
                              E_inv=t(torch_inverses(ext)), return_debug=debug)


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_vs_reference_golden(path):
    z = np.load(path)
    seed = int(z["meta"][0])
    F_, X_, E_, Z_ = _run(*flat_inputs(z), seed)
    assert F_.shape[0] == z["out_feats"].shape[1]
    assert np.array_equal(bits(X_.cpu().numpy()), bits(z["out_coords"][0]))
    assert np.array_equal(bits(Z_.cpu().numpy()), bits(z["out_depths"][0]))
    assert np.array_equal(bits(E_.cpu().numpy()), bits(z["out_ext"][0]))
    np.testing.assert_allclose(F_.cpu().numpy(), z["out_feats"][0], rtol=1e-4, atol=2e-5)


def test_vs_oracle_medium_and_reference_signature():
    from freesplat_b200 import ptf
    from oracle import ptf as optf
    seed, V, h, w = 3, 5, 60, 80
    inp = synth.ptf_inputs(seed, V, h, w)
    feats, coords, dens, wemb, depths, ext, K, hw = flat_inputs(inp)
    (oF, oX, oE, oZ, oD, oW), steps = optf.fuse(feats, coords, dens, wemb, depths, ext, K, hw,
                                                optf.torch_gru_fn(synth.gru_state(seed)), E_invs=torch_inverses(ext),
                                                return_steps=True)
    (F_, X_, E_, Z_), dbg, (D_, W_) = _run(feats, coords, dens, wemb, depths, ext, K, hw, seed, debug=True)
    for st, d in zip(steps, dbg):
        n = len(st["pix"])
        assert np.array_equal(d["pix"].cpu().numpy()[:n], st["pix"])
        assert np.array_equal(d["match"].cpu().numpy()[:n].astype(bool), st["match"])
        assert np.array_equal(d["append"].cpu().numpy().astype(bool), ~st["fuse_pix"])
        assert np.array_equal(d["zbuf"].cpu().numpy().view(np.uint32), bits(st["zbuf"]))
    assert F_.shape[0] == oF.shape[0] and F_.shape[0] < V * h * w
    for got, want in ((X_, oX), (Z_, oZ), (E_, oE), (D_, oD), (W_, oW)):
        assert np.array_equal(bits(got.cpu().numpy()), bits(want))
    np.testing.assert_allclose(F_.cpu().numpy(), oF, rtol=1e-4, atol=2e-5)
    # the reference's call signature (encoder_freesplat.py:364-368): the PUBLIC path computes extrinsic.inverse() itself, in the
    # canonical arithmetic (fs_ptf_view_setup == oracle.canonical_inverse) -> bit-exact against the oracle's default inverse
    (cF, cX, cE, cZ, cD, cW), _ = optf.fuse(feats, coords, dens, wemb, depths, ext, K, hw, optf.torch_gru_fn(synth.gru_state(seed)),
                                            return_steps=True)
    dev = "cuda:0"
    self = SimpleNamespace(gru=_gru(seed, dev))
    mv = lambda v: [x.to(dev) for x in v] if isinstance(v, list) else (v.to(dev) if isinstance(v, torch.Tensor) else v)
    with torch.no_grad():
        r = ptf.fuse_gaussians(self, *[mv(inp[k]) for k in ("gaussians", "coords", "densities", "weight_emb", "depths",
                                                            "extrinsics", "intrinsics")], inp["image_shape"])
    assert r[0].shape[0] == 1 and r[0].shape[2] == 64 and r[1].shape[-1] == 3 and r[2].shape[-2:] == (4, 4) and r[3].dim() == 2
    assert r[0].shape[1] == cF.shape[0]
    for got, want in ((r[1][0], cX), (r[3][0], cZ), (r[2][0], cE)):
        assert np.array_equal(bits(got.cpu().numpy()), bits(want))
    np.testing.assert_allclose(r[0][0].cpu().numpy(), cF, rtol=1e-4, atol=2e-5)


def test_canonical_inverse_kernel_is_bit_exact():
    """fs_ptf_view_setup: E^-1 (fp64 cofactors, one rounding) and the pixel-space intrinsics, against the oracle's restatement."""
    import ctypes as C
    from freesplat_b200 import _lib
    from oracle import ptf as optf
    g = torch.Generator().manual_seed(0)
    V, h, w = 37, 480, 640
    ext = synth.camera_path(V, spacing=0.31)
    ext[:, :3, :3] = ext[:, :3, :3] + 0.01 * torch.randn((V, 3, 3), generator=g)     # not exactly rigid
    K = synth.intrinsics(V)
    dev = "cuda:0"
    e, k = ext.to(dev).contiguous(), K.to(dev).contiguous()
    Ei, Kp = torch.empty((V, 4, 4), device=dev), torch.empty((V, 3, 3), device=dev)
    L = _lib.lib()
    _lib.check(L.fs_ptf_view_setup(C.c_int32(V), C.c_int32(h), C.c_int32(w), C.c_void_p(e.data_ptr()), C.c_void_p(k.data_ptr()),
                                   C.c_void_p(Ei.data_ptr()), C.c_void_p(Kp.data_ptr()), C.c_void_p(0)), "fs_ptf_view_setup")
    torch.cuda.synchronize()
    want = np.stack([optf.canonical_inverse(ext[v].numpy()) for v in range(V)])
    assert np.array_equal(bits(Ei.cpu().numpy()), bits(want))
    kp = K.numpy().copy(); kp[:, 0, :] *= np.float32(w); kp[:, 1, :] *= np.float32(h)
    assert np.array_equal(bits(Kp.cpu().numpy()), bits(kp))
    # and it is an inverse: within fp32 rounding of LAPACK's
    assert np.abs(want - np.linalg.inv(ext.numpy().astype(np.float64))).max() < 1e-6


def test_no_match_and_single_view():
    """Views that do not overlap at all: every pixel is appended, nothing is fused (the `if mask.sum() > 0` branch
    of encoder_freesplat.py:484 is skipped)."""
    inp = synth.ptf_inputs(4, 2, 16, 16)
    feats, coords, dens, wemb, depths, ext, K, hw = flat_inputs(inp)
    ext = ext.copy(); ext[1, :3, 3] += np.array([50.0, 0.0, 0.0], np.float32)      # far away
    F_, X_, E_, Z_ = _run(feats, coords, dens, wemb, depths, ext, K, hw, 4)
    assert F_.shape[0] == 2 * 256
    assert np.array_equal(bits(X_.cpu().numpy()), bits(np.concatenate([coords[0], coords[1]])))
    F1, X1, _, _ = _run(feats[:1], coords[:1], dens[:1], wemb[:1], depths[:1], ext[:1], K[:1], hw, 4)
    assert F1.shape[0] == 256 and np.array_equal(bits(X1.cpu().numpy()), bits(coords[0]))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_backward_vs_reference_golden(path):
    """d(sum(out*w))/d(feats, coords, densities, weights, depths, GRU parameters) against the reference's own autograd."""
    from freesplat_b200 import ptf
    z = np.load(path)
    seed = int(z["meta"][0])
    feats, coords, dens, wemb, depths, ext, K, hw = flat_inputs(z)
    dev = "cuda:0"
    t = lambda a, g=True: torch.from_numpy(np.ascontiguousarray(a)).to(dev).requires_grad_(g)
    tf, tx, td, tw, tz = t(feats), t(coords), t(dens), t(wemb), t(depths)
    gru = _gru(seed, dev)
    F_, X_, E_, Z_ = ptf.fuse_views(gru, tf, tx, td, tw, tz, t(ext, False), t(K, False), hw, E_inv=t(torch_inverses(ext), False))
    w = lambda k: torch.from_numpy(z[k]).to(dev)
    loss = (F_ * w("wF")[0]).sum() + (X_ * w("wX")[0]).sum() + (E_ * w("wE")[0]).sum() + (Z_ * w("wZ")[0]).sum()
    loss.backward()
    V = feats.shape[0]
    pairs = [("feats", tf.grad, z["g_in_feats"][0]), ("coords", tx.grad, z["g_in_coords"][0, :, :, 0, 0, :]),
             ("dens", td.grad, z["g_in_dens"][0, :, :, 0, 0]), ("wemb", tw.grad, z["g_in_wemb"][0, :, :, 0, 0]),
             ("depths", tz.grad, z["g_in_depths"].reshape(V, -1))]
    pairs += [("gru." + n, p.grad, z["g_gru." + n]) for n, p in gru.named_parameters()]
    for name, got, want in pairs:
        assert got is not None, name
        scale = np.abs(want).max() + 1e-12
        err = np.abs(got.cpu().numpy() - want).max() / scale
        assert err < 2e-4, (name, err)


def test_gru_paths_agree():
    """The three GRU evaluations -- tensor-core kernel (fs_ptf_gru), glue kernels + cuBLAS, plain module call -- on the
    same matched pairs."""
    from freesplat_b200 import ptf
    inp = synth.ptf_inputs(7, 3, 64, 96)
    feats, coords, dens, wemb, depths, ext, K, hw = flat_inputs(inp)
    outs = {}
    old = ptf.GRU_MODE
    try:
        for mode in ("tc", "cublas"):
            ptf.GRU_MODE = mode
            outs[mode] = _run(feats, coords, dens, wemb, depths, ext, K, hw, 7)[0].cpu().numpy()
    finally:
        ptf.GRU_MODE = old
    from oracle import ptf as optf
    want = optf.fuse(feats, coords, dens, wemb, depths, ext, K, hw, optf.torch_gru_fn(synth.gru_state(7)),
                     E_invs=torch_inverses(ext))[0]
    assert outs["tc"].shape == want.shape
    np.testing.assert_allclose(outs["cublas"], want, rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(outs["tc"], want, rtol=1e-4, atol=2e-5)


def test_mid_size_reference_golden_forward_and_backward():
    """120x160, 4 views: outputs and autograd gradients of the REFERENCE's own fuse_gaussians (tests/golden/mid_ptf_v4.npz;
    the reference's LAPACK inverse is passed in, as stored).  Order / coordinates / depths / extrinsics bit-exact."""
    from freesplat_b200 import ptf
    from tests import mid_golden
    from tests.helpers import grad_report
    z, inp, (wF, wX, wE, wZ), (seed, V, h, w, S, N) = mid_golden.ptf()
    feats, coords, dens, wemb, depths, ext, K, hw = flat_inputs(inp)
    dev = "cuda:0"
    t = lambda a, g=True: torch.from_numpy(np.ascontiguousarray(a)).to(dev).requires_grad_(g)
    with torch.no_grad():
        F_, X_, E_, Z_ = ptf.fuse_views(_gru(seed, dev), t(feats, False), t(coords, False), t(dens, False), t(wemb, False),
                                        t(depths, False), t(ext, False), t(K, False), hw, E_inv=t(z["E_inv"], False))
    assert F_.shape[0] == N
    assert np.array_equal(bits(X_.cpu().numpy()), bits(z["out_coords"])) and np.array_equal(bits(Z_.cpu().numpy()), bits(z["out_depths"]))
    assert np.array_equal(bits(E_.cpu().numpy()[::S]), bits(z["out_ext_sub"]))
    assert abs(float(E_.double().sum()) - float(z["out_ext_sum"])) < 1e-6 * N
    np.testing.assert_allclose(F_.cpu().numpy()[::S], z["out_feats_sub"], rtol=1e-4, atol=2e-5)
    # training path: gradients against the reference's own autograd
    tf, tx, td, tw, tz = t(feats), t(coords), t(dens), t(wemb), t(depths)
    gru = _gru(seed, dev)
    F_, X_, E_, Z_ = ptf.fuse_views(gru, tf, tx, td, tw, tz, t(ext, False), t(K, False), hw, E_inv=t(z["E_inv"], False))
    ((F_ * wF.to(dev)).sum() + (X_ * wX.to(dev)).sum() + (E_ * wE.to(dev)).sum() + (Z_ * wZ.to(dev)).sum()).backward()
    pairs = [("feats", tf.grad[:, ::S], z["g_in_feats_sub"]), ("coords", tx.grad, z["g_in_coords"]), ("dens", td.grad, z["g_in_dens"]),
             ("wemb", tw.grad, z["g_in_wemb"]), ("depths", tz.grad, z["g_in_depths"])]
    pairs += [("gru." + n, p.grad, z["g_gru." + n]) for n, p in gru.named_parameters()]
    report, bad = [], []
    for name, got, want in pairs:
        assert got is not None, name
        # The GRU's ReLU units have a kink: a pre-activation within rounding distance of 0 takes slope 0 in one fp32 evaluation
        # and 1 in another, which changes that PAIR's whole gradient by O(1).  Even the reference's own op sequence deviates from
        # its fp64 evaluation by 3e-2 of the maximum on single elements and 2.5e-3 on the first-layer weight sums at this size
        # (tests/test_gru_grad_conditioning.py demonstrates it on the CPU).  Element-wise criterion for everything else; the
        # flipped pairs are COUNTED (<= 1e-3 of the elements, nothing beyond 5e-2); parameter sums: additive term 5e-3 of max
        # (measured: 3.5e-3 on two elements of mlp_n.0.weight, everything else below 1.5e-3).
        is_param = name.startswith("gru.")
        rep = grad_report(got.cpu().numpy(), want, atol_of_max=5e-3 if is_param else 1e-5, max_outlier_frac=0.0 if is_param else 1e-3,
                          gross=5e-2)
        report.append((name, rep))
        if not rep["ok"]:
            bad.append((name, rep))
    out_dir = os.path.join(os.path.dirname(os.path.dirname(__file__)), "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "ptf_mid_golden_report.txt"), "w") as f:
            f.write("\n".join(map(str, report)))
    assert not bad, bad


@pytest.mark.parametrize("case", ["ties_golden", "five_views"])
def test_pool_fold_is_bit_identical_to_compacting_fold(case, monkeypatch):
    """The append-only pool (in-place fused rows, appended rows, order index, one gather) must return exactly what the compacting
    fold returns: same length, same ORDER, same bits in every field -- incl. the golden case with exact z-buffer ties."""
    from freesplat_b200 import ptf
    if case == "ties_golden":
        z = np.load([p for p in GOLD if "ties" in p][0])
        seed = int(z["meta"][0])
        args = flat_inputs(z)
    else:
        seed = 9
        args = flat_inputs(synth.ptf_inputs(seed, 5, 96, 128))
    outs = []
    for pool in (False, True):
        monkeypatch.setattr(ptf, "POOL", pool)
        outs.append([t_.cpu().numpy() for t_ in _run(*args, seed)])
    assert outs[0][0].shape == outs[1][0].shape and outs[0][0].shape[0] > args[0].shape[1]
    for a_, b_ in zip(outs[0], outs[1]):
        assert np.array_equal(bits(a_), bits(b_))


@pytest.mark.parametrize("M", [1, 127, 128, 1000, 33333])
def test_gru_backward_data_product(M):
    """fs_ptf_gru_bwd_data (tcgen05, 3xTF32): C (op)= A[M,64] @ W[64,N] against an fp64 product -- the three epilogues (store,
    ReLU mask, accumulate), the three widths the GRU uses (64, 152 = not a multiple of 16, 176) and strided operands."""
    import ctypes as C
    from freesplat_b200 import _lib, ptf
    L = _lib.lib()
    dev = "cuda:0"
    g = torch.Generator().manual_seed(M)
    st = C.c_void_p(torch.cuda.current_stream(torch.device(dev)).cuda_stream)
    for N, mode in ((64, 1), (152, 0), (176, 2)):
        wide = torch.randn(M, 3 * 64, generator=g).to(dev)
        A = wide[:, 64:128]                                      # ld = 192
        W = torch.randn(64, N, generator=g).to(dev)
        mask = (torch.randn(M, N, generator=g) * (torch.rand(M, N, generator=g) > 0.4)).to(dev)
        out = torch.randn(M, N, generator=g).to(dev)
        want = A.double() @ W.double()
        if mode == 1:
            want = want * (mask > 0)
        if mode == 2:
            want = want + out.double()
        with torch.cuda.device(dev):
            ptf._bwd_data(L, st, A, W, N, out, mode=mode, mask=mask if mode == 1 else None)
        torch.cuda.synchronize()
        err = float((out.double() - want).abs().max()) / max(float(want.abs().max()), 1e-30)
        assert err < 2e-6, (M, N, mode, err)


@pytest.mark.parametrize("M", [1, 31, 32, 33, 5000, 77777])
def test_gru_backward_weight_product(M):
    """fs_ptf_gru_bwd_weights: G = [Y0 | Y1]^T @ [X0 | X1 | 1] over the pairs against fp64 -- the three shapes of the GRU backward
    (two layers with separate inputs, two layers sharing one input, Y1 absent), tails that are not a multiple of the 32-pair step."""
    import ctypes as C
    from freesplat_b200 import _lib, ptf
    L = _lib.lib()
    dev = "cuda:0"
    g = torch.Generator().manual_seed(1000 + M)
    st = C.c_void_p(torch.cuda.current_stream(torch.device(dev)).cuda_stream)
    r = lambda *s: torch.randn(*s, generator=g).to(dev)
    cases = [(r(M, 64), r(M, 64), r(M, 64), r(M, 152), 224), (r(M, 64), r(M, 64), r(M, 176), None, 192),
             (r(M, 128)[:, 64:], None, r(M, 64), r(M, 64), 144)]
    for Y0, Y1, X0, X1, ldg in cases:
        G = torch.full((128, ldg), float("nan"), device=dev)
        with torch.cuda.device(dev):
            ptf._bwd_weights(L, st, Y0, Y1, X0, X1, G)
        torch.cuda.synchronize()
        Y = torch.cat([Y0, Y1 if Y1 is not None else torch.zeros_like(Y0)], 1).double()
        X = torch.cat([X0] + ([X1] if X1 is not None else []) + [torch.ones(M, 1, device=dev)], 1).double()
        want = Y.t() @ X
        n = X.shape[1]
        err = float((G[:, :n].double() - want).abs().max()) / max(float(want.abs().max()), 1e-30)
        assert err < 5e-6, (M, ldg, err)
        assert float(G[:, n:].abs().max()) == 0.0 if n < ldg else True


def test_gru_backward_tensor_core_path_matches_fp32_gemm_path(monkeypatch):
    """The training fold's gradients with the tensor-core GRU backward (saved activations, fs_ptf_gru_bwd_data / _weights) against the
    recompute + fp32 cuBLAS path on the same inputs: every input gradient within 2e-4 of its maximum except ReLU-kink flips (a saved
    3xTF32 pre-activation and an fp32-recomputed one can land on different sides of 0; counted, <= 1e-3)."""
    from freesplat_b200 import ptf
    from tests.helpers import grad_report
    dev = "cuda:0"
    feats, coords, dens, wemb, depths, ext, K, hw = flat_inputs(synth.ptf_inputs(3, 3, 96, 128))
    grads = {}
    for mode in ("tc", "cublas"):
        monkeypatch.setattr(ptf, "GRU_BWD", mode)
        t = lambda a, g=True: torch.from_numpy(np.ascontiguousarray(a)).to(dev).requires_grad_(g)
        tf, tx, td, tw, tz = t(feats), t(coords), t(dens), t(wemb), t(depths)
        gru = _gru(3, dev)
        F_, X_, E_, Z_ = ptf.fuse_views(gru, tf, tx, td, tw, tz, t(ext, False), t(K, False), hw)
        gen = torch.Generator().manual_seed(5)
        (F_ * torch.randn(F_.shape, generator=gen).to(dev)).sum().backward()
        grads[mode] = {"feats": tf.grad, "dens": td.grad, "wemb": tw.grad, **{"gru." + n: p.grad for n, p in gru.named_parameters()}}
    bad = []
    for name, got in grads["tc"].items():
        # parameter gradients are sums over all pairs incl. the flipped ones: same additive 5e-3 of max as the reference-golden test
        is_param = name.startswith("gru.")
        rep = grad_report(got.cpu().numpy(), grads["cublas"][name].cpu().numpy(), atol_of_max=5e-3 if is_param else 2e-4,
                          max_outlier_frac=0.0 if is_param else 1e-3, gross=5e-2)
        if not rep["ok"]:
            bad.append((name, rep))
    assert not bad, bad


@pytest.mark.parametrize("M", [1, 129, 20000])
def test_gru_backward_update_gate_epilogue(M):
    """fs_ptf_gru_bwd_data mode 3: the product dU = dHn @ W_n0 with the chain rule of U = [sigmoid(r_lin) h | x | e_in] in its
    epilogue (dr_lin, dA1[:, :64] +=, dA1[:, 64:88] = 0, dA1[:, 88:] = dU[:, 64:]) against the same expressions in fp64."""
    import ctypes as C
    from freesplat_b200 import _lib, ptf
    L = _lib.lib()
    dev = "cuda:0"
    g = torch.Generator().manual_seed(77 + M)
    r = lambda *s: torch.randn(*s, generator=g).to(dev)
    st = C.c_void_p(torch.cuda.current_stream(torch.device(dev)).cuda_stream)
    dHn, Wn0, A1, r_lin, dA1 = r(M, 64), r(64, 152), r(M, 176), r(M, 64), r(M, 176)
    dr = torch.full((M, 64), float("nan"), device=dev)
    dU = dHn.double() @ Wn0.double()
    rr = torch.sigmoid(r_lin.double())
    want = torch.cat([dA1[:, :64].double() + dU[:, :64] * rr, torch.zeros(M, 24, device=dev, dtype=torch.float64), dU[:, 64:]], 1)
    want_dr = dU[:, :64] * A1[:, :64].double() * rr * (1 - rr)
    with torch.cuda.device(dev):
        ptf._bwd_data(L, st, dHn, Wn0, 152, dA1, mode=3, h=A1, r_lin=r_lin, dr_lin=dr)
    torch.cuda.synchronize()
    for got, ref in ((dA1, want), (dr, want_dr)):
        assert float((got.double() - ref).abs().max()) / float(ref.abs().max()) < 2e-6
