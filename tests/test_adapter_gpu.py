"""GPU: fused Gaussian head vs golden outputs of the reference's GaussianAdapter (1e-4 relative) and vs the oracle."""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "adapter_*.npz")))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_vs_reference_golden(path):
    from freesplat_b200.adapter import gaussian_head
    z = np.load(path)
    _, N, h, w = [int(x) for x in z["meta"]]
    t = lambda k: torch.from_numpy(z[k]).to("cuda:0")
    with torch.no_grad():
        g = gaussian_head(t("raw"), t("depths"), t("opac"), t("coords"), t("ext"), t("K"), (h, w))
    for k in ("means", "covariances", "harmonics", "opacities", "scales", "rotations"):
        got = getattr(g, k).cpu().numpy(); want = z[k]
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4 * np.abs(want).max() * 1e-3 + 1e-12, err_msg=k)


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_backward_vs_reference_autograd_golden(path):
    """fs_gaussian_head_backward against the gradients the REFERENCE's own autograd produced for a seeded loss on all six
    outputs (make_adapter_golden.py): raw features, depths, opacities, coordinates and the per-Gaussian c2w matrices."""
    from freesplat_b200.adapter import gaussian_head
    from tests.helpers import grad_report
    z = np.load(path)
    _, N, h, w = [int(x) for x in z["meta"]]
    t = lambda k, g=True: torch.from_numpy(z[k]).to("cuda:0").requires_grad_(g)
    raw, depths, opac, coords, ext = t("raw"), t("depths"), t("opac"), t("coords"), t("ext")
    g = gaussian_head(raw, depths, opac, coords, ext, t("K", False), (h, w))
    loss = sum((getattr(g, k) * torch.from_numpy(z["w_" + k]).to("cuda:0")).sum()
               for k in ("means", "covariances", "harmonics", "opacities", "scales", "rotations"))
    loss.backward()
    for name, got, want in (("raw", raw.grad, z["g_raw"]), ("depths", depths.grad, z["g_depths"]), ("opac", opac.grad, z["g_opac"]),
                            ("coords", coords.grad, z["g_coords"]), ("ext", ext.grad, z["g_ext"])):
        rep = grad_report(got.cpu().numpy().reshape(want.shape), want, max_outlier_frac=0.0)
        assert rep["ok"], (name, rep)


def test_feeds_the_rasterizer_in_place():
    """The head's outputs go straight into render_views (layouts [N,3,3] / [N,3,d_sh])."""
    from freesplat_b200 import decoder, synth
    from freesplat_b200.adapter import gaussian_head
    sc = synth.pixel_aligned_scene(seed=0, h=96, w=128, n_context=1, n_target=2, keep=None).to("cuda:0")
    N = sc.means.shape[0]
    g = torch.Generator().manual_seed(0)
    raw = torch.randn((N, 34), generator=g).to("cuda:0"); raw[:, :3] -= 3.0
    ext = sc.context_extrinsics[0][None].expand(N, 4, 4).contiguous()
    depth = (ext[0].inverse() @ torch.cat([sc.means, torch.ones_like(sc.means[:, :1])], 1).T)[2]
    with torch.no_grad():
        gs = gaussian_head(raw, depth, sc.opacities, sc.means, ext, sc.intrinsics[0], (96, 128))
        col, dep = decoder.render_views(sc.extrinsics, sc.intrinsics, sc.near, sc.far, (96, 128), torch.zeros((2, 3), device="cuda:0"),
                                        gs.means, gs.covariances, gs.harmonics, gs.opacities)
    assert torch.isfinite(col).all() and float(col.abs().max()) > 0.05


BP = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "backproject_*.npz")))


@pytest.mark.parametrize("path", BP, ids=[os.path.basename(p) for p in BP])
def test_backproject_vs_reference_golden_bit_exact(path):
    from freesplat_b200.adapter import backproject_depth
    z = np.load(path)
    _, V, h, w = [int(x) for x in z["meta"]]
    t = lambda k: torch.from_numpy(z[k]).to("cuda:0")
    with torch.no_grad():
        got = backproject_depth(t("depths"), t("K"), t("c2w"), (h, w)).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), z["means"].view(np.uint32))


def test_backproject_full_size_vs_oracle():
    """640x480, 10 views: bit-identical to the CPU restatement."""
    from freesplat_b200 import synth
    from freesplat_b200.adapter import backproject_depth
    from oracle import adapter as oad
    g = torch.Generator().manual_seed(3)
    V, h, w = 10, 480, 640
    depths = 0.5 + 8 * torch.rand((V, h, w), generator=g)
    K = synth.intrinsics(1)[0]; c2w = synth.camera_path(V)
    want = oad.backproject(depths.numpy(), K.numpy(), c2w.numpy(), (h, w))
    with torch.no_grad():
        got = backproject_depth(depths.to("cuda:0"), K.to("cuda:0"), c2w.to("cuda:0"), (h, w)).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_backproject_backward_vs_autograd():
    """fs_backproject_backward against torch autograd through a plain restatement of Create_from_depth_map.project."""
    from freesplat_b200 import synth
    from freesplat_b200.adapter import backproject_depth
    from tests.helpers import grad_report
    dev = "cuda:0"
    V, h, w = 3, 24, 40
    g = torch.Generator().manual_seed(2)
    depth = (0.5 + 4 * torch.rand((V, h * w), generator=g)).to(dev).requires_grad_(True)
    Kn, c2w = synth.intrinsics(1)[0].to(dev), synth.camera_path(V).to(dev)
    wts = torch.randn((V, h * w, 3), generator=g).to(dev)
    (backproject_depth(depth, Kn, c2w, (h, w)) * wts).sum().backward()
    got = depth.grad.clone()
    d2 = depth.detach().clone().requires_grad_(True)
    ii, jj = torch.meshgrid(torch.arange(h, device=dev, dtype=torch.float32), torch.arange(w, device=dev, dtype=torch.float32), indexing="ij")
    x = ((jj.reshape(-1) - Kn[0, 2] * w) / (Kn[0, 0] * w))[None] * d2
    y = ((ii.reshape(-1) - Kn[1, 2] * h) / (Kn[1, 1] * h))[None] * d2
    cam = torch.stack([x, y, d2, torch.ones_like(d2)], -1)                     # [V,HW,4]
    world = torch.einsum("vrc,vnc->vnr", c2w[:, :3, :], cam)
    (world * wts).sum().backward()
    rep = grad_report(got.cpu().numpy(), d2.grad.cpu().numpy(), max_outlier_frac=0.0)
    assert rep["ok"], rep
