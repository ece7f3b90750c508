"""GPU: fused depth-head tail (TMA-staged and LDG-staged tiles) vs golden outputs of the reference's DepthDecoder
(tolerance 1e-4 relative) and vs the oracle on odd shapes and at the full 640x480 size."""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "depth_head_*.npz")))
RTOL = 1e-4


def _close(got, want, name):
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=RTOL * float(np.abs(want).max()) * 1e-2, err_msg=name)


@pytest.mark.parametrize("tile_mode", [0, 1], ids=["tma", "ldg"])
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_vs_reference_golden(path, tile_mode):
    from freesplat_b200.depth_head import depth_head_tail
    z = np.load(path)
    logits = {s: torch.from_numpy(z[f"logits_s{s}"]).to("cuda:0") for s in range(4)}
    with torch.no_grad():
        out = depth_head_tail(logits, torch.from_numpy(z["candi"]), bool(z["meta"][5]), tile_mode=tile_mode)
    for s in range(4):
        _close(out[f"depth_pred_s{s}_b1hw"].cpu().numpy(), z[f"depth_s{s}"], f"depth s{s}")
        _close(out[f"log_depth_pred_s{s}_b1hw"].cpu().numpy(), z[f"log_depth_s{s}"], f"log depth s{s}")
    _close(out["depth_pred_s-1_b1hw"].cpu().numpy(), z["depth_up"], "depth_up")
    _close(out["depth_weights"].cpu().numpy(), z["weights_up"], "weights_up")


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_backward_vs_reference_autograd_golden(path):
    """fs_depth_head_backward: gradient of a seeded loss on every output of the tail (four scales, the x2 depth and the
    depth weights) w.r.t. the plane logits, against the REFERENCE's own autograd (make_depth_head_golden.py)."""
    from freesplat_b200.depth_head import depth_head_tail
    from tests.helpers import grad_report
    z = np.load(path)
    dev = "cuda:0"
    logits = {s: torch.from_numpy(z[f"logits_s{s}"]).to(dev).requires_grad_(True) for s in range(4)}
    out = depth_head_tail(logits, torch.from_numpy(z["candi"]), bool(z["meta"][5]))
    w = lambda k: torch.from_numpy(z[k]).to(dev)
    loss = (out["depth_pred_s-1_b1hw"] * w("w_depth_up")).sum() + (out["depth_weights"] * w("w_weights_up")).sum()
    for s in range(4):
        loss = loss + (out[f"depth_pred_s{s}_b1hw"] * w(f"w_depth_s{s}")).sum() + (out[f"log_depth_pred_s{s}_b1hw"] * w(f"w_log_depth_s{s}")).sum()
    loss.backward()
    for s in range(4):
        # arg-max ties between planes of the upsampled softmax (scale 0) are measure-zero but possible: counted, <= 0.1 %
        rep = grad_report(logits[s].grad.cpu().numpy(), z[f"g_logits_s{s}"], max_outlier_frac=1e-3 if s == 0 else 0.0)
        assert rep["ok"], (s, rep)


@pytest.mark.parametrize("shape", [(1, 1, 1, 1), (2, 5, 3, 7), (1, 33, 9, 40), (3, 128, 37, 68), (1, 100, 61, 130)])
def test_vs_oracle_ragged_shapes(shape):
    """Odd sizes: one plane, tiles that hang over the border, widths that are not a multiple of 4 (LDG staging),
    plane counts that are not a multiple of the 32-plane TMA box."""
    from freesplat_b200.depth_head import depth_regression
    from oracle import depth_head as odh
    B, D, h, w = shape
    g = torch.Generator().manual_seed(B * 1000 + D)
    logits = torch.randn(shape, generator=g) * 4
    candi = torch.linspace(-0.7, 2.7, D)
    want = odh.forward(logits.numpy(), candi.numpy(), True, upsample=True)
    with torch.no_grad():
        got = depth_regression(logits.to("cuda:0"), candi, True, upsample=True)
        got1 = depth_regression(logits.to("cuda:0"), candi, True, upsample=False)
    for k in ("expect", "depth", "depth_up", "weights_up"):
        _close(got[k].cpu().numpy(), want[k], k)
    _close(got1["expect"].cpu().numpy(), want["expect"], "expect (no upsampling)")


def test_full_size_tma_equals_ldg_and_properties():
    """640x480 (scale 0 = 240x320, D = 128, 3 views): both staging variants agree bit-for-bit; weights are the max of
    a convex combination of probabilities (in (1/D, 1]); depth_up lies between the plane extremes."""
    from freesplat_b200.depth_head import depth_regression
    g = torch.Generator().manual_seed(0)
    logits = (torch.randn((3, 128, 240, 320), generator=g) * 5).to("cuda:0")
    candi = torch.log(torch.tensor(0.5)) + torch.linspace(0, 1, 128) * torch.log(torch.tensor(30.0))
    with torch.no_grad():
        a = depth_regression(logits, candi, True, upsample=True, tile_mode=0)
        b = depth_regression(logits, candi, True, upsample=True, tile_mode=1)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    assert float(a["weights_up"].min()) > 1.0 / 128 and float(a["weights_up"].max()) <= 1.0 + 1e-6
    assert float(a["depth_up"].min()) >= 0.5 * (1 - 1e-5) and float(a["depth_up"].max()) <= 15.0 * (1 + 1e-5)
    # scale 0 of the fused kernel == the plain expectation kernel
    with torch.no_grad():
        c = depth_regression(logits, candi, True, upsample=False)
    torch.testing.assert_close(a["expect"], c["expect"], rtol=1e-5, atol=1e-6)
