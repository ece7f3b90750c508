"""Gaussian head on the B200 kernels.

`gaussian_head(...)` computes what the reference's GaussianAdapter.forward(fusion=False, coords=...) returns for the
fused Gaussians (/root/reference/src/model/encoder/common/gaussian_adapter.py:136-200; call site
src/model/encoder/encoder_freesplat.py:376-386) -- means, covariances [N,3,3], harmonics [N,3,d_sh], opacities, scales,
rotations -- in ONE launch, already in the layouts `freesplat_b200.decoder.render_views` reads in place.
Inference path (no autograd); CPU tensors raise."""
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


@dataclass
class Gaussians:
    """Same fields as the reference dataclass (gaussian_adapter.py:100-107)."""
    means: torch.Tensor
    covariances: torch.Tensor
    scales: torch.Tensor
    rotations: torch.Tensor
    harmonics: torch.Tensor
    opacities: torch.Tensor


def gaussian_head(raw, depths, opacities, coords, extrinsics, intrinsics, image_shape, scale_min=0.5, scale_max=15.0,
                  sh_degree=2, eps=1e-8) -> Gaussians:
    """raw [N, 7+3*d_sh], depths / opacities [N], coords [N,3], extrinsics [N,4,4] (per-Gaussian c2w), intrinsics [3,3]."""
    if not raw.is_cuda:
        raise _lib.FreeSplatB200Error("gaussian_head needs CUDA tensors (no CPU fallback exists)")
    if torch.is_grad_enabled() and any(t.requires_grad for t in (raw, depths, opacities, coords, extrinsics)):
        raise _lib.FreeSplatB200Error("gaussian_head is the inference path; wrap the call in torch.no_grad()")
    L = _lib.lib()
    dev = raw.device
    f = lambda t: t.detach().float().contiguous()
    raw, depths, opacities, coords, ext, K = f(raw), f(depths), f(opacities), f(coords), f(extrinsics).reshape(-1, 16), f(intrinsics)
    N = raw.shape[0]
    dsh = (sh_degree + 1) ** 2
    assert raw.shape[1] == 7 + 3 * dsh
    e = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
    out = Gaussians(means=e(N, 3), covariances=e(N, 3, 3), scales=e(N, 3), rotations=e(N, 4), harmonics=e(N, 3, dsh), opacities=e(N))
    a = FsAdapterArgs(N=N, H=image_shape[0], W=image_shape[1], sh_degree=sh_degree, scale_min=scale_min, scale_max=scale_max, eps=eps,
                      raw=ptr(raw), depths=ptr(depths), opacities=ptr(opacities), coords=ptr(coords), ext=ptr(ext), K=ptr(K),
                      means=ptr(out.means), covariances=ptr(out.covariances), harmonics=ptr(out.harmonics),
                      opacities_out=ptr(out.opacities), scales=ptr(out.scales), rotations=ptr(out.rotations))
    with torch.cuda.device(dev):
        check(L.fs_gaussian_head(C.byref(a), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "fs_gaussian_head")
    return out
