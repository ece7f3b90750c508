"""Operator surface of the B200 rasterizer.

Mirrors the module FreeSplat imports at
/root/reference/src/model/decoder/cuda_splatting.py:5-8 and calls at :100-127
(`GaussianRasterizationSettings`, `GaussianRasterizer`; 4-tuple return
`(color[3,H,W], radii[P], depth[H,W], alpha[H,W])`), and adds the batched form
`rasterize_views` that renders all V target views of one scene in a single launch
sequence (the reference loops over views in Python and `repeat`s the
Gaussians per view, decoder_splatting_cuda.py:55-58).

All compute happens in libfreesplat_b200.so (hand-written sm_100a CUDA) through the C ABI of
include/freesplat_b200.h.  There is no CPU or PyTorch fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import NamedTuple, Optional

import torch

from . import _lib
from ._lib import FsRasterBwdArgs, FsRasterFwdArgs, check, ptr

VIEW_FLOATS = 48
REC_FLOATS = 12


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


def pack_views(viewmatrix, projmatrix, campos, bg, tanfovx, tanfovy, scene_scale=None) -> torch.Tensor:
    """Builds the [V, 48] device-side camera records of include/freesplat_b200.h.

    viewmatrix/projmatrix: [V,4,4] (already transposed the way the reference passes them),
    campos/bg: [V,3], tanfovx/tanfovy: [V] tensors (device) -- no host sync."""
    V = viewmatrix.shape[0]
    dev = viewmatrix.device
    out = torch.zeros((V, VIEW_FLOATS), dtype=torch.float32, device=dev)
    out[:, 0:16] = viewmatrix.reshape(V, 16)
    out[:, 16:32] = projmatrix.reshape(V, 16)
    out[:, 32:35] = campos
    out[:, 35:38] = bg
    out[:, 38] = tanfovx
    out[:, 39] = tanfovy
    out[:, 40] = 1.0 if scene_scale is None else scene_scale
    return out


# ----------------------------------------------------------------------------------------------
# capacity bookkeeping: the number of (Gaussian, tile) instances R is only known on the device.
_capacity_hint: dict = {}
_ASYNC = os.environ.get("FREESPLAT_B200_DEFERRED_CHECK", "0") == "1"


# status word of the most recent forward issued with check_overflow="deferred"
_deferred: dict = {"status": None}


def last_deferred_status() -> Optional[torch.Tensor]:
    """Device tensor [4] {R_lo, R_hi, overflow, 0} of the last deferred-check forward (None if there was none)."""
    return _deferred["status"]


def _initial_capacity(P: int, V: int) -> int:
    return max(8 * P * V, 1 << 18)


class RasterState:
    """Tensors a forward call leaves behind (saved for backward; the parity comparables)."""
    __slots__ = ("P", "V", "H", "W", "M", "sh_degree", "scale_modifier", "views", "rec", "cov3D", "radii",
                 "clamped", "tiles_touched", "ranges", "point_list", "keybuf", "final_T", "n_contrib", "status",
                 "capacity", "color", "depth", "sh_layout", "cov_stride")

    def num_rendered(self) -> int:
        s = self.status.cpu()
        return int(s[0]) | (int(s[1]) << 32)

    def overflowed(self) -> bool:
        return bool(int(self.status.cpu()[2]))


def _f32c(t: Optional[torch.Tensor], name: str) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.FreeSplatB200Error(f"{name} must be a CUDA tensor (no CPU fallback exists)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def raster_forward_raw(means3D, opacities, views, H, W, *, shs=None, colors_precomp=None, scales=None,
                       rotations=None, cov3D_precomp=None, sh_degree=0, scale_modifier=1.0, prefiltered=False,
                       capacity: Optional[int] = None, check_overflow: str = "sync",
                       stage_events=None, debug_buffers: bool = False, sh_layout: int = 0,
                       cov_stride: int = 6) -> RasterState:
    """One launch sequence for V views (views: [V,48]).  Returns the RasterState.

    check_overflow: "sync"     read the device status word after enqueueing everything; re-run once
                               with a larger workspace if R exceeded the capacity;
                    "deferred" no host sync; caller must call state.overflowed() before trusting it."""
    L = _lib.lib()
    means3D = _f32c(means3D, "means3D"); opacities = _f32c(opacities, "opacities").reshape(-1)
    shs = _f32c(shs, "shs"); colors_precomp = _f32c(colors_precomp, "colors_precomp")
    scales = _f32c(scales, "scales"); rotations = _f32c(rotations, "rotations")
    cov3D_precomp = _f32c(cov3D_precomp, "cov3D_precomp"); views = _f32c(views, "views")
    dev = means3D.device
    P = means3D.shape[0]
    V = views.shape[0]
    M = 0 if shs is None else (shs.shape[2] if sh_layout else shs.shape[1])
    gx, gy = (W + 15) // 16, (H + 15) // 16
    nt = V * gx * gy
    key = (dev.index, P, V, H, W)
    if capacity is None:
        capacity = _capacity_hint.get(key, _initial_capacity(P, V))
    stream = torch.cuda.current_stream(dev).cuda_stream

    with torch.cuda.device(dev):
        while True:
            st = RasterState()
            st.P, st.V, st.H, st.W, st.M, st.sh_degree, st.scale_modifier = P, V, H, W, M, sh_degree, scale_modifier
            st.views = views
            st.capacity = capacity
            st.sh_layout, st.cov_stride = sh_layout, cov_stride
            e = lambda *shape, dtype=torch.float32: torch.empty(shape, dtype=dtype, device=dev)
            st.color = e(V, 3, H, W); st.depth = e(V, H, W); st.final_T = e(V, H, W)
            st.n_contrib = e(V, H, W, dtype=torch.int32)
            st.radii = e(V, P, dtype=torch.int32)
            st.rec = e(V, P, REC_FLOATS); st.clamped = e(V, P, dtype=torch.uint8)
            # only the parity tests ask for these (the hot path neither writes nor reads them)
            st.cov3D = e(V, P, 6) if debug_buffers else None
            st.tiles_touched = e(V, P, dtype=torch.int32) if debug_buffers else None
            tile_buf = e(2 * nt, dtype=torch.int32)          # [counters | cursors]: adjacent, zeroed by ONE memset in the call
            tile_count, tile_cursor = tile_buf[:nt], tile_buf[nt:]
            st.ranges = e(nt, 2, dtype=torch.int32)
            st.keybuf = e(max(capacity, 1), dtype=torch.int64)
            st.point_list = e(max(capacity, 1), dtype=torch.int32)
            st.status = e(4, dtype=torch.int32)
            a = FsRasterFwdArgs(
                P=P, V=V, H=H, W=W, sh_degree=sh_degree, M=M, scale_modifier=scale_modifier,
                prefiltered=int(prefiltered), stages=0, sh_layout=sh_layout, cov_stride=cov_stride, capacity=capacity,
                means3D=ptr(means3D), shs=ptr(shs), colors_precomp=ptr(colors_precomp), opacities=ptr(opacities),
                scales=ptr(scales), rotations=ptr(rotations), cov3D_precomp=ptr(cov3D_precomp), views=ptr(views),
                out_color=ptr(st.color), out_depth=ptr(st.depth), final_T=ptr(st.final_T), n_contrib=ptr(st.n_contrib),
                radii=ptr(st.radii), rec=ptr(st.rec), cov3D=ptr(st.cov3D), tiles_touched=ptr(st.tiles_touched),
                clamped=ptr(st.clamped), tile_count=ptr(tile_count), tile_cursor=ptr(tile_cursor),
                ranges=ptr(st.ranges), keybuf=ptr(st.keybuf), point_list=ptr(st.point_list), status=ptr(st.status))
            if stage_events is None:
                check(L.fs_raster_forward(C.byref(a), C.c_void_p(stream)), "fs_raster_forward")
            elif len(stage_events) == 2:
                # bracket exactly the one library call (workspace allocation above is host work, not part of the path)
                stage_events[0].record()
                check(L.fs_raster_forward(C.byref(a), C.c_void_p(stream)), "fs_raster_forward")
                stage_events[1].record()
            else:
                # stage_events: list of 4 torch.cuda.Event recorded around preprocess | binning | render
                for k, bit in enumerate((1, 2, 4)):
                    stage_events[k].record()
                    a.stages = bit
                    check(L.fs_raster_forward(C.byref(a), C.c_void_p(stream)), "fs_raster_forward")
                stage_events[3].record()
            if check_overflow != "sync":
                _deferred["status"] = st.status          # the caller (or the next call) must look at it
                return st
            s = st.status.cpu()
            R = int(s[0]) | (int(s[1]) << 32)
            if not int(s[2]):
                _capacity_hint[key] = max(int(R * 1.25) + 4096, 1 << 16)
                return st
            if R > 0xFFFFFFFF:
                raise _lib.FreeSplatB200Error(f"{R} tile instances exceed the 32-bit index space")
            capacity = int(R * 1.25) + 4096


def raster_backward_raw(st: RasterState, means3D, opacities, dL_dcolor, *, shs=None, colors_precomp=None,
                        scales=None, rotations=None, cov3D_precomp=None, dL_ddepth=None):
    """Gradients w.r.t. the op inputs, summed over the V views of `st`."""
    L = _lib.lib()
    dev = means3D.device
    P, V, H, W, M = st.P, st.V, st.H, st.W, st.M
    means3D = _f32c(means3D, "means3D"); opacities = _f32c(opacities, "opacities").reshape(-1)
    shs = _f32c(shs, "shs"); colors_precomp = _f32c(colors_precomp, "colors_precomp")
    scales = _f32c(scales, "scales"); rotations = _f32c(rotations, "rotations")
    cov3D_precomp = _f32c(cov3D_precomp, "cov3D_precomp")
    dL_dcolor = _f32c(dL_dcolor, "dL_dcolor"); dL_ddepth = _f32c(dL_ddepth, "dL_ddepth")
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        e = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
        use_sr = scales is not None and rotations is not None
        g = dict(
            means2D=e(V, P, 3), means3D=e(P, 3), opacities=e(P, 1),
            cov3D=None if use_sr else e(P, st.cov_stride if st.cov_stride == 9 else 6),
            shs=None if shs is None else e(*shs.shape),
            colors=None if colors_precomp is None else e(P, 3),
            scales=e(P, 3) if use_sr else None, rotations=e(P, 4) if use_sr else None)
        # the kernel needs somewhere to accumulate dL/dcov even when it is not returned
        cov_buf = g["cov3D"]
        dscreen = e(V, P, 12)
        a = FsRasterBwdArgs(
            P=P, V=V, H=H, W=W, sh_degree=st.sh_degree, M=M, scale_modifier=st.scale_modifier,
            has_depth_grad=int(dL_ddepth is not None), sh_layout=st.sh_layout, cov_stride=st.cov_stride,
            means3D=ptr(means3D), shs=ptr(shs), colors_precomp=ptr(colors_precomp), opacities=ptr(opacities),
            scales=ptr(scales), rotations=ptr(rotations), cov3D_precomp=ptr(cov3D_precomp), views=ptr(st.views),
            rec=ptr(st.rec), radii=ptr(st.radii), clamped=ptr(st.clamped), ranges=ptr(st.ranges),
            point_list=ptr(st.point_list), final_T=ptr(st.final_T), n_contrib=ptr(st.n_contrib), status=ptr(st.status),
            dL_dcolor=ptr(dL_dcolor), dL_ddepth=ptr(dL_ddepth), dL_dscreen=ptr(dscreen),
            dL_dmeans2D=ptr(g["means2D"]), dL_dmeans3D=ptr(g["means3D"]), dL_dcov3D=ptr(cov_buf), dL_dshs=ptr(g["shs"]),
            dL_dcolors=ptr(g["colors"]), dL_dopacities=ptr(g["opacities"]), dL_dscales=ptr(g["scales"]),
            dL_drotations=ptr(g["rotations"]))
        check(L.fs_raster_backward(C.byref(a), C.c_void_p(stream)), "fs_raster_backward")
    g["screen"] = dscreen
    return g


class _RasterizeViews(torch.autograd.Function):
    """autograd wrapper: V views of one Gaussian set (V=1 is the reference op)."""

    @staticmethod
    def forward(ctx, means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp, views,
                H, W, sh_degree, scale_modifier, prefiltered, depth_grad, sh_layout=0, cov_stride=6, check_overflow=None):
        if check_overflow is None:
            check_overflow = "deferred" if _ASYNC else "sync"
        st = raster_forward_raw(means3D, opacities, views, H, W, shs=shs, colors_precomp=colors_precomp,
                                scales=scales, rotations=rotations, cov3D_precomp=cov3D_precomp, sh_degree=sh_degree,
                                scale_modifier=scale_modifier, prefiltered=prefiltered,
                                check_overflow=check_overflow, sh_layout=sh_layout, cov_stride=cov_stride)
        color, depth = st.color, st.depth
        # the OUTPUT tensors must not stay reachable from ctx: output -> grad_fn -> ctx -> st -> output is a reference
        # cycle that only Python's cyclic GC breaks, i.e. every step's workspace (hundreds of MB) would outlive the step
        # and the caching allocator would cudaMalloc afresh each iteration
        st.color = None; st.depth = None
        st.keybuf = None  # sorted keys are only needed by the parity tests (raster_forward_raw)
        ctx.st = st
        ctx.depth_grad = depth_grad
        ctx.save_for_backward(means3D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp)
        ctx.mark_non_differentiable(st.radii)
        alpha = 1.0 - st.final_T
        return color, st.radii, depth, alpha

    @staticmethod
    def backward(ctx, g_color, g_radii, g_depth, g_alpha):
        means3D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp = ctx.saved_tensors
        st = ctx.st
        if g_color is None:
            g_color = torch.zeros((st.V, 3, st.H, st.W), dtype=torch.float32, device=means3D.device)
        dd = g_depth if (ctx.depth_grad and g_depth is not None) else None
        g = raster_backward_raw(st, means3D, opacities, g_color, shs=shs, colors_precomp=colors_precomp, scales=scales,
                                rotations=rotations, cov3D_precomp=cov3D_precomp, dL_ddepth=dd)
        gm2d = g["means2D"]
        return (g["means3D"], gm2d if st.V > 1 else gm2d[0], g["shs"], g["colors"],
                g["opacities"].reshape(opacities.shape), g["scales"], g["rotations"],
                None if g["cov3D"] is None else g["cov3D"].reshape(cov3D_precomp.shape), None,
                None, None, None, None, None, None, None, None, None)


def rasterize_views(means3D, opacities, views, image_height, image_width, *, shs=None, colors_precomp=None,
                    scales=None, rotations=None, cov3D_precomp=None, means2D=None, sh_degree=0, scale_modifier=1.0,
                    prefiltered=False, depth_grad=False, sh_layout=0, cov_stride=6, check_overflow=None):
    """Batched op: -> (color[V,3,H,W], radii[V,P], depth[V,H,W], alpha[V,H,W]).

    sh_layout=1 reads `shs` as [P,3,M] and cov_stride=9 reads `cov3D_precomp` as full [P,3,3] matrices -- the
    layouts of the reference's Gaussians dataclass -- in place (gradients come back in the same layouts).
    check_overflow="deferred" skips the host read of the device status word (see raster_forward_raw)."""
    if (shs is None) == (colors_precomp is None):
        raise Exception("Please provide excatly one of either SHs or precomputed colors!")
    if ((scales is None or rotations is None) and cov3D_precomp is None) or \
            ((scales is not None or rotations is not None) and cov3D_precomp is not None):
        raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
    if means2D is None:
        V = views.shape[0]
        means2D = torch.zeros((V, means3D.shape[0], 3), dtype=torch.float32, device=means3D.device)
    return _RasterizeViews.apply(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                 views, image_height, image_width, sh_degree, scale_modifier, prefiltered, depth_grad,
                                 sh_layout, cov_stride, check_overflow)


class GaussianRasterizer(torch.nn.Module):
    """Same constructor / forward signature as the module the reference imports
    (cuda_splatting.py:114-127)."""

    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def _views(self) -> torch.Tensor:
        rs = self.raster_settings
        dev = rs.viewmatrix.device
        tf = torch.tensor([float(rs.tanfovx), float(rs.tanfovy)], dtype=torch.float32).to(dev, non_blocking=True)
        return pack_views(rs.viewmatrix.reshape(1, 4, 4).float(), rs.projmatrix.reshape(1, 4, 4).float(),
                          rs.campos.reshape(1, 3).float(), rs.bg.reshape(1, 3).float(), tf[0:1], tf[1:2])

    def markVisible(self, positions: torch.Tensor) -> torch.Tensor:
        with torch.no_grad():
            L = _lib.lib()
            positions = _f32c(positions, "positions")
            views = self._views()
            vis = torch.empty(positions.shape[0], dtype=torch.uint8, device=positions.device)
            stream = torch.cuda.current_stream(positions.device).cuda_stream
            with torch.cuda.device(positions.device):
                check(L.fs_mark_visible(C.c_int32(positions.shape[0]), C.c_void_p(ptr(positions)), C.c_void_p(ptr(views)),
                                        C.c_void_p(ptr(vis)), C.c_void_p(stream)), "fs_mark_visible")
            return vis.bool()

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        rs = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        views = self._views()
        color, radii, depth, alpha = _RasterizeViews.apply(
            means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp, views,
            int(rs.image_height), int(rs.image_width), int(rs.sh_degree), float(rs.scale_modifier),
            bool(rs.prefiltered), False)
        return color[0], radii[0], depth[0], alpha[0]
