"""Plane-sweep feature volume on the B200 kernels.

`AVGFeatureVolumeManager` mirrors the reference class of the same name
(/root/reference/src/model/encoder/modules/cost_volume.py:384-619, base class :13-381): same
constructor arguments, same parameter names (`mlp.net.{0,2,4}.{weight,bias}`) and the same persistent
buffers (`linear_ramp_1d11`, `backprojector.pix_coords_13N`, `projector.eps`), so a FreeSplat
checkpoint loads into it unchanged, and the encoder's call
`self.cost_volume(cur_feats=..., src_feats=..., src_extrinsics=..., src_poses=..., src_Ks=...,
cur_invK=..., min_depth=..., max_depth=...)` (encoder_freesplat.py:280-288) works as is.

forward/backward run in libfreesplat_b200.so (fs_cost_volume_forward / _backward).  CPU tensors
raise: there is no PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, ptr


class FsCostVolumeArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("K", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("D", C.c_int32),
        ("cur_feats", C.c_void_p), ("src_feats", C.c_void_p), ("proj", C.c_void_p), ("cur_invK", C.c_void_p),
        ("planes", C.c_void_p), ("mlp", C.c_void_p), ("out", C.c_void_p),
        ("dL_dout", C.c_void_p), ("dL_dcur", C.c_void_p), ("dL_dsrc", C.c_void_p), ("dL_dmlp", C.c_void_p),
        ("src_packed", C.c_void_p), ("dsrc_packed", C.c_void_p), ("mlp_mode", C.c_int32),
    ]


def _cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.FreeSplatB200Error(f"{name} must be a CUDA tensor (no CPU fallback exists)")
    return t.float().contiguous()


MLP_SIZES = [(32, 49), (32,), (32, 32), (32,), (1, 32), (1,)]


def pack_mlp(params) -> torch.Tensor:
    return torch.cat([p.reshape(-1) for p in params]).float().contiguous()


def unpack_mlp(flat: torch.Tensor):
    out, o = [], 0
    for shp in MLP_SIZES:
        n = 1
        for s in shp:
            n *= s
        out.append(flat[o:o + n].reshape(shp))
        o += n
    return out


# 0: tensor cores (tcgen05, 3xTF32) ; 1: fp32 CUDA-core MLP (kept to validate mode 0)
MLP_MODE = int(os.environ.get("FREESPLAT_B200_CV_MLP_MODE", "0"))


def _args(cur, src, proj, invk, planes, mlp):
    B, K, Cc, H, W = src.shape
    return FsCostVolumeArgs(B=B, K=K, C=Cc, H=H, W=W, D=planes.numel(), cur_feats=ptr(cur), src_feats=ptr(src),
                            proj=ptr(proj), cur_invK=ptr(invk), planes=ptr(planes), mlp=ptr(mlp), mlp_mode=MLP_MODE)


class _CostVolumeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cur, src, proj, invk, planes, mlp_flat):
        L = _lib.lib()
        B, K, Cc, H, W = src.shape
        out = torch.empty((B, planes.numel(), H, W), dtype=torch.float32, device=cur.device)
        a = _args(cur, src, proj, invk, planes, mlp_flat)
        packed = torch.empty_like(src)               # channel-packed scratch, filled by the call
        a.out = ptr(out); a.src_packed = ptr(packed)
        with torch.cuda.device(cur.device):
            check(L.fs_cost_volume_forward(C.byref(a), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                  "fs_cost_volume_forward")
        ctx.save_for_backward(cur, src, proj, invk, planes, mlp_flat)
        return out

    @staticmethod
    def backward(ctx, g):
        cur, src, proj, invk, planes, mlp_flat = ctx.saved_tensors
        L = _lib.lib()
        g = g.float().contiguous()
        d_cur = torch.empty_like(cur); d_src = torch.empty_like(src); d_mlp = torch.empty_like(mlp_flat)
        a = _args(cur, src, proj, invk, planes, mlp_flat)
        packed = torch.empty_like(src); dpacked = torch.empty_like(src)
        a.dsrc_packed = ptr(dpacked)
        a.dL_dout = ptr(g); a.dL_dcur = ptr(d_cur); a.dL_dsrc = ptr(d_src); a.dL_dmlp = ptr(d_mlp); a.src_packed = ptr(packed)
        with torch.cuda.device(cur.device):
            check(L.fs_cost_volume_backward(C.byref(a), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                  "fs_cost_volume_backward")
        return d_cur, d_src, None, None, None, d_mlp


def cost_volume(cur_feats, src_feats, src_extrinsics, src_Ks, cur_invK, planes, mlp_params):
    """Functional form.  cur_feats [B,C,H,W], src_feats [B,K,C,H,W], src_extrinsics / src_Ks [B,K,4,4],
    cur_invK [B,4,4], planes [D], mlp_params = (W0,b0,W1,b1,W2,b2) -> [B,D,H,W]."""
    cur = _cuda_f32(cur_feats, "cur_feats"); src = _cuda_f32(src_feats, "src_feats")
    B, K = src.shape[:2]
    # Project3D: P = K @ cam_T_world (geometry_utils.py:78); tiny batched matmul, stays on the device
    P = (src_Ks.reshape(-1, 4, 4).float() @ src_extrinsics.reshape(-1, 4, 4).float())[:, :3, :].reshape(B, K, 3, 4).contiguous()
    invk = cur_invK[:, :3, :3].float().contiguous()
    return _CostVolumeFn.apply(cur, src, P, invk, planes.float().contiguous(), pack_mlp(mlp_params))


class _Backprojector(nn.Module):
    """Holds the reference's persistent buffer `backprojector.pix_coords_13N`
    (sr_utils/geometry_utils.py:34-48); the kernel regenerates the pixel grid itself."""

    def __init__(self, height: int, width: int):
        super().__init__()
        xx, yy = torch.meshgrid(torch.arange(width), torch.arange(height), indexing="xy")
        pix = torch.stack((xx, yy), dim=0) + 0.5
        pix = torch.cat([pix, torch.ones_like(pix[:1])], 0).flatten(1).unsqueeze(0)
        self.register_buffer("pix_coords_13N", pix)


class _Projector(nn.Module):
    def __init__(self, eps: float = 1e-8):
        super().__init__()
        self.register_buffer("eps", torch.tensor(eps).view(1, 1, 1))


class _MLP(nn.Module):
    """networks.py:218-236 with channel_list [49,32,32,1], disable_final_activation=True."""

    def __init__(self, channel_list):
        super().__init__()
        layers = []
        for i in range(len(channel_list) - 1):
            layers.append(nn.Linear(channel_list[i], channel_list[i + 1]))
            layers.append(nn.LeakyReLU(inplace=True))
        self.net = nn.Sequential(*layers[:-1])


class AVGFeatureVolumeManager(nn.Module):
    def __init__(self, matching_height, matching_width, num_depth_bins=64, mlp_channels=None, matching_dim_size=16):
        super().__init__()
        mlp_channels = list(mlp_channels) if mlp_channels is not None else [202, 32, 32, 1]
        self.num_depth_bins = num_depth_bins
        self.matching_height = matching_height
        self.matching_width = matching_width
        self.register_buffer("linear_ramp_1d11", torch.linspace(0, 1, num_depth_bins).view(1, num_depth_bins, 1, 1))
        self.backprojector = _Backprojector(matching_height, matching_width)
        self.projector = _Projector()
        mlp_channels[0] = matching_dim_size + 1
        if mlp_channels != [49, 32, 32, 1]:
            raise _lib.FreeSplatB200Error(f"the fused kernel implements the MLP 49->32->32->1 (got {mlp_channels})")
        self.mlp = _MLP(mlp_channels)

    def generate_depth_planes(self, batch_size, min_depth, max_depth):
        """cost_volume.py:98-134."""
        ramp = self.linear_ramp_1d11.expand(batch_size, self.num_depth_bins, 1, 1)
        inv_min, inv_max = 1 / min_depth, 1 / max_depth
        planes = 1 / (inv_min + (inv_max - inv_min) * ramp)
        self.depth_planes_bdhw = planes.expand(batch_size, self.num_depth_bins, self.matching_height, self.matching_width)
        return self.depth_planes_bdhw

    def build_cost_volume(self, cur_feats, src_feats, src_extrinsics, src_poses, src_Ks, cur_invK, min_depth, max_depth,
                          depth_planes_bdhw=None, return_mask=False):
        if return_mask:
            raise _lib.FreeSplatB200Error("return_mask=True is not on FreeSplat's path (encoder_freesplat.py:280-288)")
        B = src_feats.shape[0]
        numel = lambda x: x.numel() if torch.is_tensor(x) else 1
        # FreeSplat's own call passes scalar near / far (encoder_freesplat.py:280-288): the planes are then fronto-parallel and
        # shared by the batch by construction.  Caller-supplied planes or per-sample depth ranges are honoured by the reference
        # per batch element and per pixel; the fused kernel takes ONE [D] plane vector, so anything else is refused (one host
        # read, off FreeSplat's path) instead of silently producing a different volume.
        static_planes = depth_planes_bdhw is None and numel(min_depth) == 1 and numel(max_depth) == 1
        if depth_planes_bdhw is None:
            depth_planes_bdhw = self.generate_depth_planes(B, min_depth, max_depth)
        if not static_planes and not bool((depth_planes_bdhw == depth_planes_bdhw[:1, :, :1, :1]).all()):
            raise _lib.FreeSplatB200Error("depth planes must be constant over the batch and over the pixels (one [D] plane "
                                          "vector): per-sample / per-pixel planes are not implemented by the fused kernel")
        planes = depth_planes_bdhw[0, :, 0, 0]
        net = self.mlp.net
        vol = cost_volume(cur_feats, src_feats, src_extrinsics, src_Ks, cur_invK, planes,
                          (net[0].weight, net[0].bias, net[2].weight, net[2].bias, net[4].weight, net[4].bias))
        return vol, depth_planes_bdhw, None

    def forward(self, cur_feats, src_feats, src_extrinsics, src_poses, src_Ks, cur_invK, min_depth, max_depth,
                depth_planes_bdhw=None, return_mask=False):
        vol, _, _ = self.build_cost_volume(cur_feats=cur_feats, src_feats=src_feats, src_extrinsics=src_extrinsics,
                                           src_poses=src_poses, src_Ks=src_Ks, cur_invK=cur_invK, min_depth=min_depth,
                                           max_depth=max_depth, depth_planes_bdhw=depth_planes_bdhw, return_mask=return_mask)
        return vol
