"""Gaussian head on the B200 kernels.

`gaussian_head(...)` computes what the reference's GaussianAdapter.forward(fusion=False, coords=...) returns for the
fused Gaussians (/root/reference/src/model/encoder/common/gaussian_adapter.py:136-200; call site
src/model/encoder/encoder_freesplat.py:376-386) -- means, covariances [N,3,3], harmonics [N,3,d_sh], opacities, scales,
rotations -- in ONE launch, already in the layouts `freesplat_b200.decoder.render_views` reads in place.
Differentiable (fs_gaussian_head_backward: one kernel for the whole chain rule); CPU tensors raise."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import check, ptr


class FsAdapterArgs(C.Structure):
    _fields_ = [("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("sh_degree", C.c_int32),
                ("scale_min", C.c_float), ("scale_max", C.c_float), ("eps", C.c_float),
                ("raw", C.c_void_p), ("depths", C.c_void_p), ("opacities", C.c_void_p), ("coords", C.c_void_p), ("ext", C.c_void_p),
                ("K", C.c_void_p), ("means", C.c_void_p), ("covariances", C.c_void_p), ("harmonics", C.c_void_p),
                ("opacities_out", C.c_void_p), ("scales", C.c_void_p), ("rotations", C.c_void_p)]


class FsAdapterBwdArgs(C.Structure):
    _fields_ = [("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("sh_degree", C.c_int32),
                ("scale_min", C.c_float), ("scale_max", C.c_float), ("eps", C.c_float)] + \
        [(n, C.c_void_p) for n in ("raw", "depths", "ext", "K", "g_means", "g_cov", "g_harmonics", "g_opacities", "g_scales",
                                   "g_rotations", "d_raw", "d_depths", "d_opacities", "d_coords", "d_ext")]


class FsBackprojectArgs(C.Structure):
    _fields_ = [("V", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("reserved", C.c_int32),
                ("depth", C.c_void_p), ("K", C.c_void_p), ("c2w", C.c_void_p), ("means", C.c_void_p)]


class _Backproject(torch.autograd.Function):
    @staticmethod
    def forward(ctx, d, K, E, h, w):
        L = _lib.lib()
        dev = d.device
        V = E.shape[0]
        means = torch.empty((V, h * w, 3), dtype=torch.float32, device=dev)
        a = FsBackprojectArgs(V=V, H=h, W=w, depth=ptr(d), K=ptr(K), c2w=ptr(E), means=ptr(means))
        with torch.cuda.device(dev):
            check(L.fs_backproject(C.byref(a), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "fs_backproject")
        ctx.save_for_backward(K, E)
        ctx.hw = (h, w)
        return means

    @staticmethod
    def backward(ctx, g):
        L = _lib.lib()
        K, E = ctx.saved_tensors
        h, w = ctx.hw
        dev = g.device
        V = E.shape[0]
        g = g.contiguous()
        d_depth = torch.empty((V, h * w), dtype=torch.float32, device=dev)
        a = FsBackprojectArgs(V=V, H=h, W=w, depth=None, K=ptr(K), c2w=ptr(E), means=None)
        with torch.cuda.device(dev):
            check(L.fs_backproject_backward(C.byref(a), C.c_void_p(ptr(g)), C.c_void_p(ptr(d_depth)),
                                            C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "fs_backproject_backward")
        return d_depth, None, None, None, None


def backproject_depth(depths, intrinsics, extrinsics, image_shape) -> torch.Tensor:
    """World coordinates of every pixel of the context views: what GaussianAdapter.forward(..., fusion=True) returns
    (/root/reference/src/model/encoder/common/gaussian_adapter.py:175-189, Create_from_depth_map.project :48-68; call site
    src/model/encoder/encoder_freesplat.py:318-327).  depths [V,H*W] or [V,H,W], intrinsics [3,3] NORMALISED (view 0's, as
    the reference uses), extrinsics [V,4,4] camera-to-world -> means [V,H*W,3].  Differentiable w.r.t. the depths; CPU tensors
    raise."""
    if not depths.is_cuda:
        raise _lib.FreeSplatB200Error("backproject_depth needs CUDA tensors (no CPU fallback exists)")
    h, w = image_shape
    V = extrinsics.shape[0]
    d = depths.float().reshape(V, h * w).contiguous()
    K = intrinsics.detach().float().reshape(9).contiguous()
    E = extrinsics.detach().float().reshape(V, 16).contiguous()
    if torch.is_grad_enabled() and d.requires_grad:
        return _Backproject.apply(d, K, E, h, w)
    return _Backproject.forward(_NoCtx(), d.detach(), K, E, h, w)


class _NoCtx:
    def save_for_backward(self, *a):
        pass


@dataclass
class Gaussians:
    """Same fields as the reference dataclass (gaussian_adapter.py:100-107)."""
    means: torch.Tensor
    covariances: torch.Tensor
    scales: torch.Tensor
    rotations: torch.Tensor
    harmonics: torch.Tensor
    opacities: torch.Tensor


def _gaussian_head_forward(raw, depths, opacities, coords, ext, K, image_shape, scale_min, scale_max, sh_degree, eps):
    L = _lib.lib()
    dev = raw.device
    N = raw.shape[0]
    dsh = (sh_degree + 1) ** 2
    e = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
    out = (e(N, 3), e(N, 3, 3), e(N, 3, dsh), e(N), e(N, 3), e(N, 4))      # means, covariances, harmonics, opacities, scales, rotations
    a = FsAdapterArgs(N=N, H=image_shape[0], W=image_shape[1], sh_degree=sh_degree, scale_min=scale_min, scale_max=scale_max, eps=eps,
                      raw=ptr(raw), depths=ptr(depths), opacities=ptr(opacities), coords=ptr(coords), ext=ptr(ext), K=ptr(K),
                      means=ptr(out[0]), covariances=ptr(out[1]), harmonics=ptr(out[2]),
                      opacities_out=ptr(out[3]), scales=ptr(out[4]), rotations=ptr(out[5]))
    with torch.cuda.device(dev):
        check(L.fs_gaussian_head(C.byref(a), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "fs_gaussian_head")
    return out


class _GaussianHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raw, depths, opacities, coords, ext, K, image_shape, scale_min, scale_max, sh_degree, eps):
        ctx.cfg = (image_shape, scale_min, scale_max, sh_degree, eps)
        ctx.save_for_backward(raw, depths, ext, K)
        return _gaussian_head_forward(raw, depths, opacities, coords, ext, K, image_shape, scale_min, scale_max, sh_degree, eps)

    @staticmethod
    def backward(ctx, g_means, g_cov, g_harm, g_opac, g_scales, g_rot):
        L = _lib.lib()
        raw, depths, ext, K = ctx.saved_tensors
        image_shape, scale_min, scale_max, sh_degree, eps = ctx.cfg
        dev = raw.device
        N = raw.shape[0]
        c = lambda g: None if g is None else g.contiguous()
        g_means, g_cov, g_harm, g_opac, g_scales, g_rot = map(c, (g_means, g_cov, g_harm, g_opac, g_scales, g_rot))
        e = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        d_raw, d_depths, d_opac, d_coords = e(*raw.shape), e(N), e(N), e(N, 3)
        d_ext = e(N, 16) if ctx.needs_input_grad[4] else None
        a = FsAdapterBwdArgs(N=N, H=image_shape[0], W=image_shape[1], sh_degree=sh_degree, scale_min=scale_min, scale_max=scale_max,
                             eps=eps, raw=ptr(raw), depths=ptr(depths), ext=ptr(ext), K=ptr(K), g_means=ptr(g_means), g_cov=ptr(g_cov),
                             g_harmonics=ptr(g_harm), g_opacities=ptr(g_opac), g_scales=ptr(g_scales), g_rotations=ptr(g_rot),
                             d_raw=ptr(d_raw), d_depths=ptr(d_depths), d_opacities=ptr(d_opac), d_coords=ptr(d_coords), d_ext=ptr(d_ext))
        with torch.cuda.device(dev):
            check(L.fs_gaussian_head_backward(C.byref(a), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "fs_gaussian_head_backward")
        return d_raw, d_depths, d_opac, d_coords, d_ext, None, None, None, None, None, None


def gaussian_head(raw, depths, opacities, coords, extrinsics, intrinsics, image_shape, scale_min=0.5, scale_max=15.0,
                  sh_degree=2, eps=1e-8) -> Gaussians:
    """raw [N, 7+3*d_sh], depths / opacities [N], coords [N,3], extrinsics [N,4,4] (per-Gaussian c2w), intrinsics [3,3].
    Differentiable w.r.t. raw, depths, opacities, coords and extrinsics (one backward kernel)."""
    if not raw.is_cuda:
        raise _lib.FreeSplatB200Error("gaussian_head needs CUDA tensors (no CPU fallback exists)")
    f = lambda t: t.float().contiguous()
    raw, depths, opacities, coords, K = f(raw), f(depths), f(opacities), f(coords), f(intrinsics.detach())
    ext = f(extrinsics).reshape(-1, 16)
    dsh = (sh_degree + 1) ** 2
    assert raw.shape[1] == 7 + 3 * dsh
    cfg = (tuple(image_shape), float(scale_min), float(scale_max), int(sh_degree), float(eps))
    if torch.is_grad_enabled() and any(t.requires_grad for t in (raw, depths, opacities, coords, ext)):
        o = _GaussianHead.apply(raw, depths, opacities, coords, ext, K, *cfg)
    else:
        o = _gaussian_head_forward(raw.detach(), depths.detach(), opacities.detach(), coords.detach(), ext.detach(), K, *cfg)
    return Gaussians(means=o[0], covariances=o[1], harmonics=o[2], opacities=o[3], scales=o[4], rotations=o[5])
