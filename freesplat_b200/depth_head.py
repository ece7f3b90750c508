"""Tail of the depth-regression head on the B200 kernels.

`depth_head_tail(...)` returns what the last lines of the reference's DepthDecoder.forward
(/root/reference/src/model/encoder/modules/networks.py:130-152) put into `depth_outputs` -- `depth_pred_s{i}_b1hw`,
`log_depth_pred_s{i}_b1hw` for the four scales, `depth_pred_s-1_b1hw` and `depth_weights` -- from the plane logits
`conv_depth[i](output_pred_s{i})`, reading the logits once (csrc/depth_head.cu; scale 0 through TMA-staged tiles).
Differentiable w.r.t. the logits (fs_depth_head_backward); CPU tensors raise."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, ptr


class FsDepthHeadArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("D", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("log_planes", C.c_int32), ("upsample", C.c_int32), ("tile_mode", C.c_int32), ("reserved", C.c_int32),
                ("logits", C.c_void_p), ("candi", C.c_void_p), ("expect", C.c_void_p), ("depth", C.c_void_p),
                ("depth_up", C.c_void_p), ("weights_up", C.c_void_p)]


class FsDepthHeadBwdArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("D", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("log_planes", C.c_int32),
                ("upsample", C.c_int32), ("reserved0", C.c_int32), ("reserved1", C.c_int32)] + \
        [(n, C.c_void_p) for n in ("logits", "candi", "g_expect", "g_depth", "g_depth_up", "g_weights_up", "stats", "argmax_up",
                                   "d_logits")]


def _depth_forward(logits, candi, log_planes, upsample, tile_mode):
    L = _lib.lib()
    dev = logits.device
    B, D, h, w = logits.shape
    e = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
    expect, depth = e(B, 1, h, w), e(B, 1, h, w)
    depth_up = e(B, 1, 2 * h, 2 * w) if upsample else None
    weights_up = e(B, 1, 2 * h, 2 * w) if upsample else None
    a = FsDepthHeadArgs(B=B, D=D, h=h, w=w, log_planes=int(bool(log_planes)), upsample=int(bool(upsample)), tile_mode=tile_mode,
                        logits=ptr(logits), candi=ptr(candi), expect=ptr(expect), depth=ptr(depth),
                        depth_up=ptr(depth_up), weights_up=ptr(weights_up))
    with torch.cuda.device(dev):
        check(L.fs_depth_head(C.byref(a), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "fs_depth_head")
    return expect, depth, depth_up, weights_up


class _DepthRegression(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, candi, log_planes, upsample, tile_mode):
        ctx.cfg = (bool(log_planes), bool(upsample))
        ctx.save_for_backward(logits, candi)
        ctx.set_materialize_grads(False)
        out = _depth_forward(logits, candi, log_planes, upsample, tile_mode)
        return out if upsample else out[:2]

    @staticmethod
    def backward(ctx, g_expect, g_depth, g_depth_up=None, g_weights_up=None):
        L = _lib.lib()
        logits, candi = ctx.saved_tensors
        log_planes, upsample = ctx.cfg
        dev = logits.device
        B, D, h, w = logits.shape
        c = lambda g: None if g is None else g.contiguous()
        g_expect, g_depth, g_depth_up, g_weights_up = map(c, (g_expect, g_depth, g_depth_up, g_weights_up))
        d_logits = torch.empty_like(logits)
        stats = torch.empty((B, h, w, 3), dtype=torch.float32, device=dev)
        argmax = torch.empty((B, 2 * h, 2 * w), dtype=torch.uint8, device=dev) if (upsample and g_weights_up is not None) else None
        a = FsDepthHeadBwdArgs(B=B, D=D, h=h, w=w, log_planes=int(log_planes), upsample=int(upsample), logits=ptr(logits),
                               candi=ptr(candi), g_expect=ptr(g_expect), g_depth=ptr(g_depth), g_depth_up=ptr(g_depth_up),
                               g_weights_up=ptr(g_weights_up), stats=ptr(stats), argmax_up=ptr(argmax), d_logits=ptr(d_logits))
        with torch.cuda.device(dev):
            check(L.fs_depth_head_backward(C.byref(a), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "fs_depth_head_backward")
        return d_logits, None, None, None, None


def depth_regression(logits: torch.Tensor, candi: torch.Tensor, log_planes: bool = True, upsample: bool = False,
                     tile_mode: int = 0) -> dict:
    """logits [B,D,h,w], candi [D] (or DepthDecoder.depth_candi_curr [1,D,1,1]).  Returns expect / depth [B,1,h,w] and,
    with `upsample`, depth_up / weights_up [B,1,2h,2w].  tile_mode: 0 = TMA box loads, 1 = LDG staging (validation).
    Differentiable w.r.t. `logits`."""
    if not logits.is_cuda:
        raise _lib.FreeSplatB200Error("depth_regression needs CUDA tensors (no CPU fallback exists)")
    dev = logits.device
    logits = logits.float().contiguous()
    candi = candi.detach().to(dev).float().reshape(-1).contiguous()
    B, D, h, w = logits.shape
    if candi.numel() != D:
        raise _lib.FreeSplatB200Error(f"candi has {candi.numel()} entries for {D} planes")
    if torch.is_grad_enabled() and logits.requires_grad:
        o = _DepthRegression.apply(logits, candi, log_planes, upsample, tile_mode)
    else:
        o = _depth_forward(logits.detach(), candi, log_planes, upsample, tile_mode)
    out = dict(expect=o[0], depth=o[1])
    if upsample:
        out["depth_up"] = o[2]; out["weights_up"] = o[3]
    return out


def depth_head_tail(plane_logits: dict, depth_candi_curr: torch.Tensor, log_planes: bool = True, tile_mode: int = 0) -> dict:
    """plane_logits {scale i: conv_depth[i](output_pred_s{i}) [B,D,h_i,w_i]} -> the reference's depth_outputs entries."""
    res = {}
    for i in sorted(plane_logits, reverse=True):
        o = depth_regression(plane_logits[i], depth_candi_curr, log_planes, upsample=(i == 0), tile_mode=tile_mode)
        res[f"depth_pred_s{i}_b1hw"] = o["depth"]
        res[f"log_depth_pred_s{i}_b1hw"] = o["expect"]
        if i == 0:
            res["depth_pred_s-1_b1hw"] = o["depth_up"]
            res["depth_weights"] = o["weights_up"]
    return res
