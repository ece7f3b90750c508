"""TEST INFRASTRUCTURE -- CPU restatement (numpy fp32, explicit operation order) of Pixel-wise
Triplet Fusion, EncoderFreeSplat.fuse_gaussians
(/root/reference/src/model/encoder/encoder_freesplat.py:431-522; positional_encoding :62-77; GRU
/root/reference/src/model/encoder/modules/networks.py:188-214), written from SURVEY.md Appendix C.
Pinned against outputs of the reference code itself: tests/golden/ptf_*.npz
(tests/golden/make_ptf_golden.py).

Canonical arithmetic (what the CUDA kernels reproduce bit for bit, because the merged coordinates
feed the index decisions of the next view):
  q_r   = fma(E3,1, fma(E2,z, fma(E1,y, E0*x)))      (torch's CPU sgemm order for [4,4]@[4,N])
  uv    = (q_xy / q_z) * f + c                        (div, mul, add: three roundings)
  merge = (a*w0 + b*w1) / (w0 + w1)                   (mul, mul, add, add, div)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def _fma(a, b, c):
    # exact product in float64 (24x24 bits fit), one rounding to float32 after the add
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F32)


def positional_encoding(x: np.ndarray, freqs: int = 6) -> np.ndarray:
    """encoder_freesplat.py:62-77 with ori=False: x [...,D] -> [..., 2*D*F], (sin, cos) interleaved."""
    bands = (2.0 ** np.arange(freqs)).astype(F32)
    pts = (x[..., None].astype(F32) * bands).reshape(x.shape[:-1] + (freqs * x.shape[-1],))
    out = np.stack([np.sin(pts), np.cos(pts)], axis=-1)
    return out.reshape(pts.shape[:-1] + (pts.shape[-1] * 2,)).astype(F32)


def canonical_inverse(E: np.ndarray) -> np.ndarray:
    """extrinsic.inverse() (encoder_freesplat.py:454) in the canonical arithmetic of the CUDA path: cofactor expansion over
    2x2 sub-determinants in float64 (plain multiplies / adds, no fma), one rounding to float32.  The reference's own result
    depends on the LAPACK / cuSOLVER build it runs on (last-bit differences); tests that compare against reference golden
    outputs pass the reference's inverse in explicitly (E_invs)."""
    a = [np.float64(x) for x in np.asarray(E, dtype=F32).reshape(16)]
    s0 = a[0] * a[5] - a[4] * a[1]; s1 = a[0] * a[6] - a[4] * a[2]; s2 = a[0] * a[7] - a[4] * a[3]
    s3 = a[1] * a[6] - a[5] * a[2]; s4 = a[1] * a[7] - a[5] * a[3]; s5 = a[2] * a[7] - a[6] * a[3]
    c5 = a[10] * a[15] - a[14] * a[11]; c4 = a[9] * a[15] - a[13] * a[11]; c3 = a[9] * a[14] - a[13] * a[10]
    c2 = a[8] * a[15] - a[12] * a[11]; c1 = a[8] * a[14] - a[12] * a[10]; c0 = a[8] * a[13] - a[12] * a[9]
    det = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0
    i = np.float64(1.0) / det
    b = [(a[5] * c5 - a[6] * c4 + a[7] * c3) * i, (-a[1] * c5 + a[2] * c4 - a[3] * c3) * i,
         (a[13] * s5 - a[14] * s4 + a[15] * s3) * i, (-a[9] * s5 + a[10] * s4 - a[11] * s3) * i,
         (-a[4] * c5 + a[6] * c2 - a[7] * c1) * i, (a[0] * c5 - a[2] * c2 + a[3] * c1) * i,
         (-a[12] * s5 + a[14] * s2 - a[15] * s1) * i, (a[8] * s5 - a[10] * s2 + a[11] * s1) * i,
         (a[4] * c4 - a[5] * c2 + a[7] * c0) * i, (-a[0] * c4 + a[1] * c2 - a[3] * c0) * i,
         (a[12] * s4 - a[13] * s2 + a[15] * s0) * i, (-a[8] * s4 + a[9] * s2 - a[11] * s0) * i,
         (-a[4] * c3 + a[5] * c1 - a[6] * c0) * i, (a[0] * c3 - a[1] * c1 + a[2] * c0) * i,
         (-a[12] * s3 + a[13] * s1 - a[14] * s0) * i, (a[8] * s3 - a[9] * s1 + a[10] * s0) * i]
    return np.array(b, dtype=np.float64).astype(F32).reshape(4, 4)


def project(coords: np.ndarray, E_inv: np.ndarray, fx, fy, cx, cy, h: int, w: int):
    """Appendix C steps 1-2.  coords [N,3] -> (pix [N] int64 (-1 if invalid), zeta [N] f32, valid [N])."""
    x, y, z = coords[:, 0], coords[:, 1], coords[:, 2]
    one = np.ones_like(x)
    q = []
    for r in range(3):
        e = [np.full_like(x, E_inv[r, k]) for k in range(4)]
        q.append(_fma(e[3], one, _fma(e[2], z, _fma(e[1], y, (e[0] * x).astype(F32)))))
    zeta = q[2]
    with np.errstate(divide="ignore", invalid="ignore"):
        u = ((q[0] / zeta).astype(F32) * F32(fx)).astype(F32) + F32(cx)
        v = ((q[1] / zeta).astype(F32) * F32(fy)).astype(F32) + F32(cy)
        u = u.astype(F32); v = v.astype(F32)
        col = np.rint(u); row = np.rint(v)             # round half to even, like torch.round
    finite = np.isfinite(col) & np.isfinite(row)
    coli = np.where(finite, col, -1).astype(np.int64); rowi = np.where(finite, row, -1).astype(np.int64)
    valid = finite & (rowi >= 0) & (rowi < h) & (coli >= 0) & (coli < w) & (zeta > 0)
    pix = np.where(valid, coli + rowi * w, -1)
    return pix, zeta, valid


def fuse(feats, coords, dens, wemb, depths, extrinsics, intrinsics, image_shape, gru_fn, depth_thres=0.1,
         E_invs=None, return_steps=False):
    """feats [V,HW,F], coords [V,HW,3], dens/wemb [V,HW], depths [V,HW] (predicted depth maps),
    extrinsics [V,4,4] (c2w), intrinsics [V,3,3] (normalised).  gru_fn(input_feat [M,F], hidden [M,F],
    e_in [M,24], e_h [M,24]) -> [M,F].  Returns (feats [N,F], coords [N,3], extr [N,4,4], depths [N])."""
    f32 = lambda a: np.ascontiguousarray(a, dtype=F32)
    feats, coords, dens, wemb, depths, extrinsics, intrinsics = map(f32, (feats, coords, dens, wemb, depths, extrinsics, intrinsics))
    V = feats.shape[0]
    h, w = image_shape
    HW = h * w
    gF, gX, gD, gW = feats[0].copy(), coords[0].copy(), dens[0].copy(), wemb[0].copy()
    gE = np.repeat(extrinsics[0][None], HW, 0)
    gZ = depths[0].copy()
    steps = []
    for i in range(1, V):
        Kpx = intrinsics[i].copy()
        Kpx[0, :] *= F32(w); Kpx[1, :] *= F32(h)
        E_inv = canonical_inverse(extrinsics[i]) if E_invs is None else f32(E_invs[i])
        pix, zeta, valid = project(gX, E_inv, Kpx[0, 0], Kpx[1, 1], Kpx[0, 2], Kpx[1, 2], h, w)
        zbuf = np.full(HW, 1e4, F32)
        np.minimum.at(zbuf, pix[valid], zeta[valid])
        d_i = depths[i]
        fuse_pix = np.abs(zbuf - d_i) < np.maximum((d_i * F32(0.05)).astype(F32), F32(depth_thres))
        match = valid.copy()
        match[valid] = (zeta[valid] == zbuf[pix[valid]]) & fuse_pix[pix[valid]]
        mi = np.nonzero(match)[0]
        p = pix[mi]
        if len(mi) > 0:
            e_in = positional_encoding(np.stack([gD[mi], wemb[i][p]], -1))
            e_h = positional_encoding(np.stack([dens[i][p], gW[mi]], -1))
            newF = gru_fn(feats[i][p], gF[mi], e_in, e_h).astype(F32)
            w0, w1 = gD[mi], dens[i][p]
            ws = (w0 + w1).astype(F32)
            mrg = lambda a, b: (((a * w0[(...,) + (None,) * (a.ndim - 1)]).astype(F32) + (b * w1[(...,) + (None,) * (b.ndim - 1)]).astype(F32)).astype(F32)
                                / ws[(...,) + (None,) * (a.ndim - 1)]).astype(F32)
            newX = mrg(gX[mi], coords[i][p])
            newZ = mrg(gZ[mi], depths[i][p])
            newE = mrg(gE[mi], np.repeat(extrinsics[i][None], len(mi), 0))
            newD = ws
            newW = (gW[mi] + wemb[i][p]).astype(F32)
            keep = ~match
            gF = np.concatenate([gF[keep], newF]); gX = np.concatenate([gX[keep], newX])
            gD = np.concatenate([gD[keep], newD]); gW = np.concatenate([gW[keep], newW])
            gE = np.concatenate([gE[keep], newE]); gZ = np.concatenate([gZ[keep], newZ])
        app = ~fuse_pix
        na = int(app.sum())
        gF = np.concatenate([gF, feats[i][app]]); gX = np.concatenate([gX, coords[i][app]])
        gD = np.concatenate([gD, dens[i][app]]); gW = np.concatenate([gW, wemb[i][app]])
        gE = np.concatenate([gE, np.repeat(extrinsics[i][None], na, 0)]); gZ = np.concatenate([gZ, depths[i][app]])
        steps.append(dict(pix=pix, valid=valid, zeta=zeta, zbuf=zbuf, fuse_pix=fuse_pix, match=match))
    if return_steps:
        return (gF, gX, gE, gZ, gD, gW), steps
    return gF, gX, gE, gZ


def torch_gru_fn(state_dict):
    """GRU of networks.py:188-214 evaluated with torch on the CPU (fp32)."""
    import torch

    sd = {k: torch.as_tensor(v).float() for k, v in state_dict.items()}

    def mlp(name, x):
        x = torch.relu(x @ sd[f"{name}.0.weight"].T + sd[f"{name}.0.bias"])
        return x @ sd[f"{name}.2.weight"].T + sd[f"{name}.2.bias"]

    def fn(inp, hid, e_in, e_h):
        inp, hid, e_in, e_h = (torch.from_numpy(np.ascontiguousarray(a)).float() for a in (inp, hid, e_in, e_h))
        in1 = torch.cat([inp, e_in], -1); h1 = torch.cat([hid, e_h], -1)
        cat = torch.cat([h1, in1], -1)
        r = torch.sigmoid(mlp("mlp_r", cat)); z = torch.sigmoid(mlp("mlp_z", cat))
        q = torch.tanh(mlp("mlp_n", torch.cat([r * hid, in1], -1)))
        return ((1 - z) * hid + z * q).numpy()

    return fn
