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
