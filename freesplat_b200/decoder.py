"""Host-side mirror of the reference's raster adapter.

  (`render_cuda` itself, /root/reference/src/model/decoder/cuda_splatting.py:47-132, needs no mirror: with the top-level
  module `diff_gaussian_rasterization_depth` of this repo on PYTHONPATH the reference's own function runs unchanged --
  tests/test_reference_adapter_gpu.py feeds the calls it makes through that module.)
  render_views(...)     what DecoderSplattingCUDA.forward (decoder_splatting_cuda.py:35-75) needs:
                        ONE Gaussian set per scene rendered into all its target views by a single
                        launch sequence; the `scale_invariant` rescale (cuda_splatting.py:64-71) is
                        folded into the preprocess kernel through FsView.scene_scale, and there is no
                        `.item()` / per-view host sync.

Camera helpers restate src/geometry/projection.py:233-247 (get_fov) and
cuda_splatting.py:17-44 (get_projection_matrix).
"""
from __future__ import annotations

from math import isqrt

import torch

from .rasterizer import pack_views, rasterize_views


def get_fov(intrinsics: torch.Tensor) -> torch.Tensor:
    """[B,3,3] normalised intrinsics -> [B,2] (fov_x, fov_y); geometry/projection.py:233-247."""
    inv = intrinsics.inverse()

    def process(vec):
        v = torch.tensor(vec, dtype=torch.float32, device=intrinsics.device)
        v = torch.einsum("bij,j->bi", inv, v)
        return v / v.norm(dim=-1, keepdim=True)

    left, right = process([0, 0.5, 1]), process([1, 0.5, 1])
    top, bottom = process([0.5, 0, 1]), process([0.5, 1, 1])
    fov_x = (left * right).sum(dim=-1).acos()
    fov_y = (top * bottom).sum(dim=-1).acos()
    return torch.stack((fov_x, fov_y), dim=-1)


def get_projection_matrix(near, far, fov_x, fov_y) -> torch.Tensor:
    """cuda_splatting.py:17-44."""
    tan_fov_x = (0.5 * fov_x).tan()
    tan_fov_y = (0.5 * fov_y).tan()
    top = tan_fov_y * near
    bottom = -top
    right = tan_fov_x * near
    left = -right
    (b,) = near.shape
    result = torch.zeros((b, 4, 4), dtype=torch.float32, device=near.device)
    result[:, 0, 0] = 2 * near / (right - left)
    result[:, 1, 1] = 2 * near / (top - bottom)
    result[:, 0, 2] = (right + left) / (right - left)
    result[:, 1, 2] = (top + bottom) / (top - bottom)
    result[:, 3, 2] = 1
    result[:, 2, 2] = far / (far - near)
    result[:, 2, 3] = -(far * near) / (far - near)
    return result


def camera_records(extrinsics, intrinsics, near, far, background_color, scale_invariant=True):
    """[V,4,4] c2w, [V,3,3] normalised K, [V] near/far, [V,3] bg -> ([V,48] records, tanfov[V,2]).

    Everything stays on the device (no .item()).  With scale_invariant the per-view factor
    1/near goes into the record; the kernel multiplies means by it and covariances by its square,
    which is exactly what cuda_splatting.py:64-71 does on full tensors."""
    if scale_invariant:
        scale = 1 / near
        extrinsics = extrinsics.clone()
        extrinsics[..., :3, 3] = extrinsics[..., :3, 3] * scale[:, None]
        near = near * scale
        far = far * scale
    else:
        scale = torch.ones_like(near)
    fov_x, fov_y = get_fov(intrinsics).unbind(dim=-1)
    tan_fov_x = (0.5 * fov_x).tan()
    tan_fov_y = (0.5 * fov_y).tan()
    projection_matrix = get_projection_matrix(near, far, fov_x, fov_y).transpose(1, 2)
    view_matrix = extrinsics.inverse().transpose(1, 2)
    full_projection = view_matrix @ projection_matrix
    views = pack_views(view_matrix, full_projection, extrinsics[:, :3, 3], background_color, tan_fov_x, tan_fov_y,
                       scene_scale=scale)
    return views, torch.stack((tan_fov_x, tan_fov_y), dim=-1)


def camera_records_fused(extrinsics, intrinsics, near, far, background_color, scale_invariant=True):
    """Same result as camera_records (to fp32 rounding: evaluated in fp64 on the device) in ONE kernel launch
    (fs_camera_records) instead of ~30 host-driven torch launches."""
    import ctypes as C
    from . import _lib
    L = _lib.lib()
    V = extrinsics.shape[0]
    dev = extrinsics.device
    if not extrinsics.is_cuda:
        raise _lib.FreeSplatB200Error("camera_records_fused needs CUDA tensors")
    f = lambda t: t.detach().float().contiguous()
    e, k, n, fr, bg = f(extrinsics), f(intrinsics), f(near), f(far), f(background_color)
    views = torch.empty((V, 48), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.fs_camera_records(C.c_int32(V), C.c_void_p(e.data_ptr()), C.c_void_p(k.data_ptr()), C.c_void_p(n.data_ptr()),
                                       C.c_void_p(fr.data_ptr()), C.c_void_p(bg.data_ptr()), C.c_int32(int(scale_invariant)),
                                       C.c_void_p(views.data_ptr()), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                   "fs_camera_records")
    return views


def render_views(extrinsics, intrinsics, near, far, image_shape, background_color, gaussian_means,
                 gaussian_covariances, gaussian_sh_coefficients, gaussian_opacities, scale_invariant=True,
                 use_sh=True, depth_grad=False, check_overflow=None, grad_reduce=None):
    """One scene, V target views.

    extrinsics [V,4,4], intrinsics [V,3,3], near/far [V], background_color [V,3],
    gaussian_means [G,3], gaussian_covariances [G,3,3], gaussian_sh_coefficients [G,3,d_sh],
    gaussian_opacities [G]  ->  (color [V,3,H,W], depth [V,H,W]) in the SCALED scene units
    (the caller divides by 1/near as decoder_splatting_cuda.py:62 does).

    The Gaussian tensors are consumed in the reference's own layouts ([G,3,3] covariances, [G,3,d_sh]
    harmonics): the kernels index them in place, so none of the per-call `rearrange` / `triu` gather
    copies of cuda_splatting.py:75,116,126 exist here, nor their mirror images in the backward pass."""
    assert use_sh or gaussian_sh_coefficients.shape[-1] == 1
    h, w = image_shape
    n = gaussian_sh_coefficients.shape[-1]
    degree = isqrt(n) - 1
    views = camera_records_fused(extrinsics, intrinsics, near, far, background_color, scale_invariant)
    color, radii, depth, _ = rasterize_views(
        gaussian_means, gaussian_opacities, views, h, w,
        shs=gaussian_sh_coefficients if use_sh else None,
        colors_precomp=None if use_sh else gaussian_sh_coefficients[:, :, 0],
        cov3D_precomp=gaussian_covariances, sh_degree=degree, depth_grad=depth_grad, sh_layout=1, cov_stride=9,
        check_overflow=check_overflow, grad_reduce=grad_reduce)
    return color, depth


class DecoderSplattingB200(torch.nn.Module):
    """forward() of DecoderSplattingCUDA (decoder_splatting_cuda.py:35-75) on the batched op.

    `gaussians` needs .means [b,G,3] .covariances [b,G,3,3] .harmonics [b,G,3,d_sh] .opacities [b,G]
    (src/model/types.py Gaussians).  Returns (color [b,v,3,h,w], depth [b,v,h,w] or None)."""

    def __init__(self, background_color=(0.0, 0.0, 0.0)):
        super().__init__()
        self.register_buffer("background_color", torch.tensor(background_color, dtype=torch.float32), persistent=False)

    def forward(self, gaussians, extrinsics, intrinsics, near, far, image_shape, depth_mode=None):
        b, v = extrinsics.shape[:2]
        colors, depths = [], []
        for i in range(b):
            bg = self.background_color.to(extrinsics.device)[None].expand(v, 3)
            c, d = render_views(extrinsics[i], intrinsics[i], near[i], far[i], image_shape, bg, gaussians.means[i],
                                gaussians.covariances[i], gaussians.harmonics[i], gaussians.opacities[i])
            colors.append(c)
            depths.append(d)
        color = torch.stack(colors)
        depth = torch.stack(depths) / 2   # decoder_splatting_cuda.py:62 (1/near = 2 at near = 0.5)
        return color, (None if depth_mode is None else depth)
