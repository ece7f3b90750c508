"""TEST INFRASTRUCTURE -- CPU restatement (numpy fp32) of the tail of the reference's DepthDecoder.forward
(/root/reference/src/model/encoder/modules/networks.py:130-152): per scale
    planes = softmax(logits, dim=1); E = sum_d candi[d] * planes[d]; depth = exp(E) (log_planes) or 1/E,
and for scale 0 additionally
    depth_up   = f(bilinear x2 (align_corners=True) of E),   weights_up = max_d bilinear x2 of planes[d].
The bilinear rule is torch's (aten UpSampleKernel.cpp `compute_indices_weights`, align_corners=True): scale =
(in-1)/(out-1) in fp32, src = scale*dst, i0 = min(int(src), in-1), lam = clamp(src-i0, 0, 1), i1 = min(i0+1, in-1).
Pinned against outputs of the reference code itself: tests/golden/depth_head_*.npz (make_depth_head_golden.py).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module."""
from __future__ import annotations

import numpy as np

F32 = np.float32


def expectation(logits, candi, log_planes=True):
    """logits [B,D,h,w], candi [D] -> planes [B,D,h,w], E [B,1,h,w], depth [B,1,h,w]."""
    logits = np.asarray(logits, F32); candi = np.asarray(candi, F32).reshape(1, -1, 1, 1)
    m = logits.max(axis=1, keepdims=True)
    e = np.exp(logits - m, dtype=F32)
    planes = (e / e.sum(axis=1, keepdims=True, dtype=F32)).astype(F32)
    E = (candi * planes).sum(axis=1, keepdims=True, dtype=F32)
    depth = np.exp(E, dtype=F32) if log_planes else (F32(1.0) / E).astype(F32)
    return planes, E, depth


def _axis_taps(n_in: int, n_out: int):
    scale = F32(n_in - 1) / F32(n_out - 1) if n_out > 1 else F32(0.0)
    src = (scale * np.arange(n_out, dtype=F32)).astype(F32)
    i0 = np.minimum(src.astype(np.int64), n_in - 1)
    lam = np.clip(src - i0.astype(F32), F32(0.0), F32(1.0)).astype(F32)
    i1 = np.minimum(i0 + 1, n_in - 1)
    return i0, i1, (F32(1.0) - lam).astype(F32), lam


def upsample2(x):
    """F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=True) for x [B,C,h,w]."""
    x = np.asarray(x, F32)
    h, w = x.shape[-2:]
    y0, y1, wy0, wy1 = _axis_taps(h, 2 * h)
    x0, x1, wx0, wx1 = _axis_taps(w, 2 * w)
    top = x[..., y0, :][..., x0] * wx0 + x[..., y0, :][..., x1] * wx1
    bot = x[..., y1, :][..., x0] * wx0 + x[..., y1, :][..., x1] * wx1
    return (top * wy0[:, None] + bot * wy1[:, None]).astype(F32)


def forward(logits, candi, log_planes=True, upsample=True):
    planes, E, depth = expectation(logits, candi, log_planes)
    out = dict(expect=E, depth=depth)
    if upsample:
        fine = upsample2(E)
        out["depth_up"] = np.exp(fine, dtype=F32) if log_planes else (F32(1.0) / fine).astype(F32)
        out["weights_up"] = upsample2(planes).max(axis=1, keepdims=True)
    return out
