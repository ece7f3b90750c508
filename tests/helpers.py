"""Shared test helpers: turn a synthetic Scene into per-view oracle inputs exactly the way the
reference adapter (cuda_splatting.py:47-127) prepares the rasterizer call."""
from __future__ import annotations

import numpy as np
import torch

from freesplat_b200 import decoder, synth


def view_inputs(scene: synth.Scene, i: int, scale_invariant: bool = True, bg=(0.0, 0.0, 0.0)):
    """dict of numpy arrays for oracle.raster.forward() for target view i."""
    V = scene.extrinsics.shape[0]
    bgc = torch.tensor(bg, dtype=torch.float32)[None].expand(V, 3)
    views, tanfov = decoder.camera_records(scene.extrinsics, scene.intrinsics, scene.near, scene.far, bgc,
                                           scale_invariant)
    rec = views[i]
    s = rec[40]
    d_sh = scene.harmonics.shape[-1]
    row, col = torch.triu_indices(3, 3)
    return dict(
        H=scene.image_shape[0], W=scene.image_shape[1],
        tanfovx=float(rec[38]), tanfovy=float(rec[39]), bg=rec[35:38].numpy(),
        viewmatrix=rec[0:16].numpy(), projmatrix=rec[16:32].numpy(), campos=rec[32:35].numpy(),
        means3D=(scene.means * s).numpy(),
        opacities=scene.opacities.numpy(),
        shs=scene.harmonics.transpose(1, 2).contiguous().numpy(),
        cov3D_precomp=(scene.covariances * (s * s))[:, row, col].contiguous().numpy(),
        sh_degree=int(round(d_sh ** 0.5)) - 1,
    ), views


def rel_err(a, b, eps=1e-6):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / (np.abs(b) + eps)


def frac_within(a, b, rtol=1e-4, atol=1e-5):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    ok = np.abs(a - b) <= atol + rtol * np.abs(b)
    return float(ok.mean())


def grad_report(got, want, rtol=1e-4, atol_of_max=1e-5, max_outlier_frac=5e-4, gross=1e-2) -> dict:
    """Element-wise gradient check with an explicit, counted exclusion list.

    An element passes if |got - want| <= rtol * |want| + atol_of_max * max|want| (the additive term covers elements that are
    sums of cancelling contributions: fp32 atomics accumulate them in another order than the fp64 oracle).  Elements outside
    are COUNTED (threshold decisions -- alpha >= 1/255, T < 1e-4, LeakyReLU / ReLU kinks -- can flip between MUFU and libm
    arithmetic and change single contributions); their share must stay below max_outlier_frac and no element may be off by
    more than `gross` of the tensor's maximum."""
    a = np.asarray(got, dtype=np.float64); b = np.asarray(want, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    scale = np.abs(b).max() + 1e-30
    err = np.abs(a - b)
    bad = err > rtol * np.abs(b) + atol_of_max * scale
    rep = dict(n=int(b.size), outliers=int(bad.sum()), outlier_frac=float(bad.mean()), max_err_of_max=float(err.max() / scale),
               scale=float(scale))
    rep["ok"] = bool(rep["outlier_frac"] <= max_outlier_frac and rep["max_err_of_max"] <= gross)
    return rep


def render_per_view_upstream_layout(extrinsics, intrinsics, near, far, image_shape, background_color, means, covariances,
                                    harmonics, opacities):
    """Test-side driver of the drop-in op in the UPSTREAM layouts, one call per view: [P,M,3] harmonics, [P,6] covariances
    pre-scaled by 1/near (what the reference adapter hands to the op, cuda_splatting.py:64-127).  Used to cross-check the
    batched in-place path; returns (color [V,3,H,W], depth [V,H,W])."""
    from diff_gaussian_rasterization_depth import GaussianRasterizationSettings, GaussianRasterizer
    V = extrinsics.shape[0]
    h, w = image_shape
    views, tanfov = decoder.camera_records(extrinsics, intrinsics, near, far, background_color, True)
    iu = torch.triu_indices(3, 3)
    shs = harmonics.permute(0, 2, 1)
    d_sh = harmonics.shape[-1]
    colors, depths = [], []
    for i in range(V):
        s = 1.0 / near[i]
        cfg = GaussianRasterizationSettings(h, w, float(tanfov[i, 0]), float(tanfov[i, 1]), background_color[i], 1.0,
                                            views[i, 0:16].reshape(4, 4), views[i, 16:32].reshape(4, 4),
                                            int(round(d_sh ** 0.5)) - 1, views[i, 32:35], False, False)
        c, _, d, _ = GaussianRasterizer(cfg)(means3D=means * s, means2D=torch.zeros_like(means, requires_grad=True), shs=shs,
                                             opacities=opacities[:, None], cov3D_precomp=(covariances * (s * s))[:, iu[0], iu[1]])
        colors.append(c); depths.append(d)
    return torch.stack(colors), torch.stack(depths)
