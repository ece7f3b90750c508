"""TEST INFRASTRUCTURE -- CPU restatement (numpy fp32) of the Gaussian head the encoder applies to the fused
latents: GaussianAdapter.forward with fusion=False and `coords` given
(/root/reference/src/model/encoder/common/gaussian_adapter.py:136-200, build_covariance / quaternion_to_matrix
common/gaussians.py:8-44), called at src/model/encoder/encoder_freesplat.py:376-386.
Pinned against outputs of the reference code itself: tests/golden/adapter_*.npz (make_adapter_golden.py).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module."""
from __future__ import annotations

import numpy as np

F32 = np.float32


def sh_mask(sh_degree: int) -> np.ndarray:
    m = np.ones(((sh_degree + 1) ** 2,), F32)
    for d in range(1, sh_degree + 1):
        m[d * d:(d + 1) * (d + 1)] = F32(0.1 * 0.25 ** d)
    return m


def forward(raw, depths, opacities, coords, ext, K, image_shape, scale_min=0.5, scale_max=15.0, sh_degree=2, eps=1e-8):
    """raw [N, 7+3*d_sh], depths/opacities [N], coords [N,3], ext [N,4,4], K [3,3] normalised."""
    raw = np.asarray(raw, F32); depths = np.asarray(depths, F32); ext = np.asarray(ext, F32); K = np.asarray(K, F32)
    h, w = image_shape
    d_sh = (sh_degree + 1) ** 2
    s_raw, q_raw, sh_raw = raw[:, :3], raw[:, 3:7], raw[:, 7:7 + 3 * d_sh]
    scales = F32(scale_min) + F32(scale_max - scale_min) * (1.0 / (1.0 + np.exp(-s_raw.astype(np.float64)))).astype(F32)
    pix = np.array([1.0 / w, 1.0 / h], F32)
    mult = F32(0.1) * (np.linalg.inv(K[:2, :2].astype(np.float64)).astype(F32) @ pix)
    mult = F32(mult.sum())
    scales = (scales * depths[:, None] * mult).astype(F32)
    rot = (q_raw / (np.linalg.norm(q_raw, axis=-1, keepdims=True) + F32(eps))).astype(F32)
    sh = (sh_raw.reshape(-1, 3, d_sh) * sh_mask(sh_degree)).astype(F32)
    i, j, k, r = rot[:, 0], rot[:, 1], rot[:, 2], rot[:, 3]
    two_s = 2.0 / ((rot * rot).sum(-1) + eps)
    R = np.stack([1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                  two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                  two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)], -1).reshape(-1, 3, 3).astype(F32)
    S = scales[:, None, :] * np.eye(3, dtype=F32)[None]
    cov = R @ S @ np.swapaxes(S, 1, 2) @ np.swapaxes(R, 1, 2)
    C = ext[:, :3, :3]
    cov = (C @ cov @ np.swapaxes(C, 1, 2)).astype(F32)
    return dict(means=np.asarray(coords, F32), covariances=cov, harmonics=sh, opacities=np.asarray(opacities, F32), scales=scales,
                rotations=rot)


def _fma(a, b, c):
    """fp32 fused multiply-add (one rounding): the product of two floats is exact in fp64."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(F32)


def backproject(depths, K_norm, c2w, image_shape):
    """GaussianAdapter.forward(fusion=True) (gaussian_adapter.py:175-189) -> Create_from_depth_map.project (:48-68):
    depths [V,H,W], K_norm [3,3] (view 0), c2w [V,4,4] -> world coordinates [V,H*W,3].  Canonical order: intrinsics rows
    scaled by w / h in fp32, (j - cx) / fx * z, then the fma chain of torch's CPU sgemm for [4,4] @ [4,N]."""
    h, w = image_shape
    depths = np.asarray(depths, F32).reshape(-1, h, w); K = np.asarray(K_norm, F32); c2w = np.asarray(c2w, F32)
    fx, cx = F32(K[0, 0] * F32(w)), F32(K[0, 2] * F32(w))
    fy, cy = F32(K[1, 1] * F32(h)), F32(K[1, 2] * F32(h))
    jj, ii = np.meshgrid(np.arange(w, dtype=F32), np.arange(h, dtype=F32))
    px = ((jj - cx) / fx).astype(F32); py = ((ii - cy) / fy).astype(F32)
    out = np.empty((depths.shape[0], h * w, 3), F32)
    for v in range(depths.shape[0]):
        z = depths[v]; x = (px * z).astype(F32); y = (py * z).astype(F32); E = c2w[v]
        for r in range(3):
            t = (E[r, 0] * x).astype(F32)
            t = _fma(E[r, 1], y, t); t = _fma(E[r, 2], z, t); t = _fma(E[r, 3], F32(1.0), t)
            out[v, :, r] = t.reshape(-1)
    return out
