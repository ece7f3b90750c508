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
# FS_STAGE_RENDER_PACKED: blend with the two-pixels-per-lane packed-fp32 kernel (bit-identical results; A/B switch)
RENDER_PACKED = os.environ.get("FREESPLAT_B200_RENDER_PACKED", "0") == "1"
# tile scan inside the last preprocess CTA instead of its own one-block launch: built and bit-identical, but measured SLOWER
# (preprocess 47 -> 61 us against 9 us saved in the binning stage: a 256-thread CTA scans 3600 counters in ~14 us at the tail of
# the kernel, the 1024-thread launch in ~6 us + launch latency that the CUDA graph already hides) -> off by default
FUSED_SCAN = os.environ.get("FREESPLAT_B200_FUSED_SCAN", "0") == "1"
# direct binning (FsRasterFwdArgs.bins): keys per tile bin; 0 = off (preprocess counts, a separate scatter pass appends the keys).
# A tile with more instances than this makes the call fall back to the scatter pass on the device: results are identical.
BIN_CAP = int(os.environ.get("FREESPLAT_B200_BIN_CAP", "4096"))


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
#
# The hint is MONOTONE: it starts at 8 instances per (view, Gaussian) and only ever grows (to 1.25 R + 4096 after an
# overflow).  A hint that shrank to "what the last scene needed" would make a deferred-check step of a denser scene overflow
# with nobody looking (every kernel early-returns on the overflow flag, the outputs stay uninitialised).
_capacity_hint: dict = {}
# shapes whose tiles overflowed their bins (status[3] of an earlier call): no bins for them any more -- the device-side fallback
# keeps every call correct, this only stops paying for bins that are not used
_bins_off: set = set()
# default overflow check of the public calls: "deferred" = no host sync (the status word is copied to pinned memory behind
# the kernels and inspected by the NEXT call, which raises and grows the workspace), "sync" = read it before returning.
DEFAULT_CHECK = os.environ.get("FREESPLAT_B200_CHECK", "deferred")
if os.environ.get("FREESPLAT_B200_DEFERRED_CHECK") == "0":      # round-1 switch, still honoured
    DEFAULT_CHECK = "sync"
assert DEFAULT_CHECK in ("sync", "deferred")


def _initial_capacity(P: int, V: int) -> int:
    return max(8 * P * V, 1 << 18)


def _grow_hint(key, R: int) -> int:
    cap = max(_capacity_hint.get(key, 0), int(R * 1.25) + 4096, 1 << 16)
    _capacity_hint[key] = cap
    return cap


def reset_capacity_hints() -> None:
    _capacity_hint.clear()
    _bins_off.clear()


class _Pending:
    """Status word of a deferred-check forward on its way to pinned host memory."""
    __slots__ = ("host", "event", "key", "capacity")


_pending: list = []
_pinned_free: list = []
_deferred: dict = {"status": None}


def last_deferred_status() -> Optional[torch.Tensor]:
    """Device tensor [4] {R_lo, R_hi, overflow, 0} of the last deferred-check forward (None if there was none)."""
    return _deferred["status"]


def _defer_status(status: torch.Tensor, key, capacity: int) -> None:
    p = _Pending()
    p.host = _pinned_free.pop() if _pinned_free else torch.empty(4, dtype=torch.int32).pin_memory()
    p.host.copy_(status, non_blocking=True)          # 16 bytes behind the kernels, on the caller's stream
    p.event = torch.cuda.Event()
    p.event.record()
    p.key, p.capacity = key, capacity
    _pending.append(p)
    _deferred["status"] = status


def poll_deferred(block: bool = False) -> None:
    """Looks at the status words of earlier deferred-check forwards that have reached the host (all of them with
    block=True).  An overflowed one grows the capacity hint and raises: the outputs of THAT call were invalid."""
    bad = None
    keep = []
    for p in _pending:
        if block:
            p.event.synchronize()
        elif not p.event.query():
            keep.append(p)
            continue
        R = (int(p.host[0]) & 0xFFFFFFFF) | ((int(p.host[1]) & 0xFFFFFFFF) << 32)
        if int(p.host[3]):
            _note_bin_fallback(p.key)
        if int(p.host[2]):
            _grow_hint(p.key, R)
            bad = (R, p.capacity)
        _pinned_free.append(p.host)
    _pending[:] = keep
    if bad is not None:
        raise _lib.FreeSplatB200Error(
            f"an earlier rasterizer call (check_overflow='deferred') needed {bad[0]} tile instances but its workspace held "
            f"{bad[1]}: its outputs were invalid.  The workspace has been grown; re-run that step "
            "(or use check_overflow='sync' / FREESPLAT_B200_CHECK=sync).")


def _note_bin_fallback(key) -> None:
    _bins_off.add(key)
    for ck in [c for c in _scratch_cache if c[0] == key]:
        del _scratch_cache[ck]


class RasterState:
    """Tensors a forward call leaves behind (saved for backward; the parity comparables)."""
    __slots__ = ("P", "V", "H", "W", "M", "sh_degree", "scale_modifier", "views", "rec", "cov3D", "radii",
                 "clamped", "tiles_touched", "ranges", "point_list", "keybuf", "final_T", "n_contrib", "status",
                 "capacity", "color", "depth", "sh_layout", "cov_stride", "tile_buf", "bins", "bin_cap")

    def num_rendered(self) -> int:
        s = self.status.cpu()
        return (int(s[0]) & 0xFFFFFFFF) | ((int(s[1]) & 0xFFFFFFFF) << 32)

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


def _alloc_state(dev, P, V, H, W, M, sh_degree, scale_modifier, capacity, sh_layout, cov_stride, debug_buffers=False,
                 outputs=True, bins_ok=True) -> RasterState:
    """Every buffer one forward call writes (outputs, saved-for-backward state, binning scratch)."""
    gx, gy = (W + 15) // 16, (H + 15) // 16
    nt = V * gx * gy
    st = RasterState()
    st.P, st.V, st.H, st.W, st.M, st.sh_degree, st.scale_modifier = P, V, H, W, M, sh_degree, scale_modifier
    st.capacity = capacity
    st.sh_layout, st.cov_stride = sh_layout, cov_stride
    e = lambda *shape, dtype=torch.float32: torch.empty(shape, dtype=dtype, device=dev)
    st.color = e(V, 3, H, W) if outputs else None
    st.depth = e(V, H, W) if outputs else None
    st.final_T = e(V, H, W)
    st.n_contrib = e(V, H, W, dtype=torch.int32)
    st.radii = e(V, P, dtype=torch.int32)
    st.rec = e(V, P, REC_FLOATS); st.clamped = e(V, P, dtype=torch.uint8)
    # only the parity tests ask for these (the hot path neither writes nor reads them)
    st.cov3D = e(V, P, 6) if debug_buffers else None
    st.tiles_touched = e(V, P, dtype=torch.int32) if debug_buffers else None
    # [counters | cursors | status]: adjacent -> ONE memset per call, and the tile scan runs in the last preprocess CTA
    st.tile_buf = e(2 * nt + 4, dtype=torch.int32)
    st.ranges = e(nt, 2, dtype=torch.int32)
    st.keybuf = e(max(capacity, 1), dtype=torch.int64)
    st.point_list = e(max(capacity, 1), dtype=torch.int32)
    # FREESPLAT_B200_FUSED_SCAN=0: a separate status buffer -> the stand-alone one-block scan kernel runs (A/B measurements)
    st.status = st.tile_buf[2 * nt:] if FUSED_SCAN else e(4, dtype=torch.int32)
    # tile bins of the direct-binning path (16 KB per tile at 2048 keys; only the occupied head of a bin is ever touched);
    # tile_buf[2 nt] is the bin-overflow flag, zeroed with the counters
    st.bin_cap = BIN_CAP if (bins_ok and 0 < BIN_CAP <= 4096 and not FUSED_SCAN and not RENDER_PACKED and nt * BIN_CAP * 8 <= (1 << 30)) else 0
    st.bins = e(nt * st.bin_cap, dtype=torch.int64) if st.bin_cap else None
    return st


def _fwd_args(st: RasterState, means3D, opacities, views, shs, colors_precomp, scales, rotations, cov3D_precomp,
              prefiltered=False) -> FsRasterFwdArgs:
    nt = (st.tile_buf.numel() - 4) // 2
    tile_count, tile_cursor = st.tile_buf[:nt], st.tile_buf[nt:2 * nt]
    return FsRasterFwdArgs(
        P=st.P, V=st.V, H=st.H, W=st.W, sh_degree=st.sh_degree, M=st.M, scale_modifier=st.scale_modifier,
        prefiltered=int(prefiltered), stages=8 if RENDER_PACKED else 0, sh_layout=st.sh_layout, cov_stride=st.cov_stride, capacity=st.capacity,
        means3D=ptr(means3D), shs=ptr(shs), colors_precomp=ptr(colors_precomp), opacities=ptr(opacities),
        scales=ptr(scales), rotations=ptr(rotations), cov3D_precomp=ptr(cov3D_precomp), views=ptr(views),
        out_color=ptr(st.color), out_depth=ptr(st.depth), final_T=ptr(st.final_T), n_contrib=ptr(st.n_contrib),
        radii=ptr(st.radii), rec=ptr(st.rec), cov3D=ptr(st.cov3D), tiles_touched=ptr(st.tiles_touched),
        clamped=ptr(st.clamped), tile_count=ptr(tile_count), tile_cursor=ptr(tile_cursor),
        ranges=ptr(st.ranges), keybuf=ptr(st.keybuf), point_list=ptr(st.point_list), status=ptr(st.status),
        bins=ptr(st.bins), bin_cap=st.bin_cap)


# scratch of inference calls (no autograd graph keeps it alive), reused by the next call with the same shape on the same
# stream: ~12 allocator round trips less per call.  The tensors a caller gets back (colour, depth, radii, alpha) are fresh.
_scratch_cache: dict = {}


def raster_forward_raw(means3D, opacities, views, H, W, *, shs=None, colors_precomp=None, scales=None,
                       rotations=None, cov3D_precomp=None, sh_degree=0, scale_modifier=1.0, prefiltered=False,
                       capacity: Optional[int] = None, check_overflow: str = "sync",
                       stage_events=None, debug_buffers: bool = False, sh_layout: int = 0,
                       cov_stride: int = 6, reuse_scratch: bool = False) -> RasterState:
    """One launch sequence for V views (views: [V,48]).  Returns the RasterState.

    check_overflow: "sync"     read the device status word after enqueueing everything; re-run once
                               with a larger workspace if R exceeded the capacity;
                    "deferred" no host sync: the status word travels to pinned memory behind the kernels and the next
                               call (or poll_deferred()) raises if this one overflowed.
    reuse_scratch:  the state buffers (everything except colour / depth / radii) come from a per-(shape, stream) cache and
                    are overwritten by the next such call: inference only."""
    L = _lib.lib()
    means3D = _f32c(means3D, "means3D"); opacities = _f32c(opacities, "opacities").reshape(-1)
    shs = _f32c(shs, "shs"); colors_precomp = _f32c(colors_precomp, "colors_precomp")
    scales = _f32c(scales, "scales"); rotations = _f32c(rotations, "rotations")
    cov3D_precomp = _f32c(cov3D_precomp, "cov3D_precomp"); views = _f32c(views, "views")
    dev = means3D.device
    P = means3D.shape[0]
    V = views.shape[0]
    M = 0 if shs is None else (shs.shape[2] if sh_layout else shs.shape[1])
    key = (dev.index, P, V, H, W)
    explicit_capacity = capacity
    if capacity is None:
        capacity = _capacity_hint.setdefault(key, _initial_capacity(P, V))
    if _pending:
        poll_deferred()
    ts = torch.cuda.current_stream(dev)
    stream = ts.cuda_stream

    with torch.cuda.device(dev):
        while True:
            if reuse_scratch and not debug_buffers:
                ck = (key, capacity, M, stream)
                st = _scratch_cache.get(ck)
                if st is None:
                    if len(_scratch_cache) >= 8:
                        _scratch_cache.clear()
                    st = _scratch_cache[ck] = _alloc_state(dev, P, V, H, W, M, sh_degree, scale_modifier, capacity, sh_layout,
                                                           cov_stride, outputs=False, bins_ok=key not in _bins_off)
                st.sh_degree, st.scale_modifier, st.sh_layout, st.cov_stride = sh_degree, scale_modifier, sh_layout, cov_stride
                st.color = torch.empty((V, 3, H, W), dtype=torch.float32, device=dev)
                st.depth = torch.empty((V, H, W), dtype=torch.float32, device=dev)
                st.radii = torch.empty((V, P), dtype=torch.int32, device=dev)
            else:
                st = _alloc_state(dev, P, V, H, W, M, sh_degree, scale_modifier, capacity, sh_layout, cov_stride, debug_buffers,
                                  bins_ok=key not in _bins_off)
            st.views = views
            a = _fwd_args(st, means3D, opacities, views, shs, colors_precomp, scales, rotations, cov3D_precomp, prefiltered)
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
                    a.stages = bit | (8 if RENDER_PACKED else 0)
                    check(L.fs_raster_forward(C.byref(a), C.c_void_p(stream)), "fs_raster_forward")
                stage_events[3].record()
            if check_overflow != "sync":
                _defer_status(st.status, key, capacity)   # inspected by the next call (poll_deferred)
                return st
            s = st.status.cpu()
            R = (int(s[0]) & 0xFFFFFFFF) | ((int(s[1]) & 0xFFFFFFFF) << 32)
            if int(s[3]) and st.bin_cap:
                _note_bin_fallback(key)
            if not int(s[2]):
                return st
            if R > 0xFFFFFFFF:
                raise _lib.FreeSplatB200Error(f"{R} tile instances exceed the 32-bit index space")
            capacity = int(R * 1.25) + 4096
            if explicit_capacity is None:
                capacity = _grow_hint(key, R)


class RasterPlan:
    """A forward step on a STATIC workspace, recorded once as a CUDA graph (fs_graph_*) and replayed with one launch.

    For steady-state rendering where the input tensors keep their addresses (a serving loop that copies each scene into
    the same device buffers: HostRenderPipeline; bench.py): the memset + kernels of fs_raster_forward (and, when camera
    tensors are given, fs_camera_records) go to the GPU back to back with no host work in between -- on a busy host the
    launch gaps of the eager call sequence otherwise land inside the step.

        plan = RasterPlan(means, opacities, H, W, shs=..., cov3D_precomp=..., views=views)          # or cameras=(E,K,near,far,bg)
        plan.run(); plan.color, plan.depth      # results of the last run, overwritten by the next one
        plan.check()                            # (R, overflowed): one host read; grow() + run() again if it overflowed
    """

    def __init__(self, means3D, opacities, H, W, *, shs=None, colors_precomp=None, scales=None, rotations=None,
                 cov3D_precomp=None, views=None, cameras=None, scale_invariant=True, sh_degree=0, scale_modifier=1.0,
                 sh_layout=0, cov_stride=6, capacity: Optional[int] = None, graph: bool = True):
        self.inputs = tuple(_f32c(t, n) for t, n in ((means3D, "means3D"), (opacities, "opacities"), (shs, "shs"),
                                                      (colors_precomp, "colors_precomp"), (scales, "scales"),
                                                      (rotations, "rotations"), (cov3D_precomp, "cov3D_precomp")))
        for a, b in zip(self.inputs, (means3D, opacities, shs, colors_precomp, scales, rotations, cov3D_precomp)):
            if a is not None and a.data_ptr() != b.data_ptr():
                raise _lib.FreeSplatB200Error("RasterPlan needs contiguous fp32 CUDA inputs (it records their addresses)")
        dev = self.inputs[0].device
        self.dev, self.H, self.W = dev, H, W
        self.cameras = None
        if cameras is not None:
            self.cameras = tuple(_f32c(t, "camera tensor") for t in cameras)       # (extrinsics, intrinsics, near, far, bg)
            V = self.cameras[0].shape[0]
            with torch.cuda.device(dev):
                views = torch.empty((V, VIEW_FLOATS), dtype=torch.float32, device=dev)
        self.scale_invariant = bool(scale_invariant)
        self.views = _f32c(views, "views")
        P, V = self.inputs[0].shape[0], self.views.shape[0]
        shs_ = self.inputs[2]
        self.M = 0 if shs_ is None else (shs_.shape[2] if sh_layout else shs_.shape[1])
        self.key = (dev.index, P, V, H, W)
        if capacity is None:
            capacity = _capacity_hint.setdefault(self.key, _initial_capacity(P, V))
        self.meta = (P, V, sh_degree, scale_modifier, sh_layout, cov_stride)
        self.use_graph = graph
        self.graph = None
        self._build(capacity)

    def _enqueue(self, stream_ptr):
        L = _lib.lib()
        if self.cameras is not None:
            e, k, n, f, bg = self.cameras
            check(L.fs_camera_records(C.c_int32(self.views.shape[0]), C.c_void_p(ptr(e)), C.c_void_p(ptr(k)), C.c_void_p(ptr(n)),
                                      C.c_void_p(ptr(f)), C.c_void_p(ptr(bg)), C.c_int32(int(self.scale_invariant)),
                                      C.c_void_p(ptr(self.views)), C.c_void_p(stream_ptr)), "fs_camera_records")
        check(L.fs_raster_forward(C.byref(self.args), C.c_void_p(stream_ptr)), "fs_raster_forward")

    def _build(self, capacity):
        L = _lib.lib()
        P, V, sh_degree, scale_modifier, sh_layout, cov_stride = self.meta
        self.destroy()
        with torch.cuda.device(self.dev):
            self.st = _alloc_state(self.dev, P, V, self.H, self.W, self.M, sh_degree, scale_modifier, capacity, sh_layout, cov_stride,
                                   bins_ok=self.key not in _bins_off)
            self.st.views = self.views
            m, o, shs, cp, sc, ro, cov = self.inputs
            self.args = _fwd_args(self.st, m, o.reshape(-1), self.views, shs, cp, sc, ro, cov)
            if self.use_graph:
                torch.cuda.current_stream(self.dev).synchronize()       # the buffers above exist before the capture starts
                cs, g = C.c_void_p(), C.c_void_p()
                check(L.fs_graph_capture_begin(C.byref(cs)), "fs_graph_capture_begin")
                try:
                    self._enqueue(cs.value)
                finally:
                    rc = L.fs_graph_capture_end(cs, C.byref(g))
                check(rc, "fs_graph_capture_end")
                self.graph = g

    @property
    def color(self):
        return self.st.color

    @property
    def depth(self):
        return self.st.depth

    @property
    def status(self):
        return self.st.status

    @property
    def capacity(self):
        return self.st.capacity

    def run(self, stream: Optional[torch.cuda.Stream] = None):
        s = (stream or torch.cuda.current_stream(self.dev)).cuda_stream
        with torch.cuda.device(self.dev):
            if self.graph is not None:
                check(_lib.lib().fs_graph_launch(self.graph, C.c_void_p(s)), "fs_graph_launch")
            else:
                self._enqueue(s)
        return self

    def launches_per_run(self) -> int:
        # preprocess, tile scan (its own launch unless FUSED_SCAN), scatter, sort + render (+ camera records); with direct binning
        # the scan and the fallback scatter are one launch
        return (3 if (FUSED_SCAN or self.st.bin_cap) else 4) + int(self.cameras is not None)

    def check(self):
        """(R, overflowed) of the last run: ONE host read (waits for the stream)."""
        s = self.st.status.cpu()
        if int(s[3]) and self.st.bin_cap:
            _note_bin_fallback(self.key)          # the next (re)build of a plan for this shape runs without bins
        return (int(s[0]) & 0xFFFFFFFF) | ((int(s[1]) & 0xFFFFFFFF) << 32), bool(int(s[2]))

    def grow(self, R: int):
        """Re-allocates the workspace for R instances (and re-records the graph)."""
        self._build(_grow_hint(self.key, R))

    def run_checked(self, stream: Optional[torch.cuda.Stream] = None):
        """run() + check(); grows and re-runs once if the workspace was too small."""
        self.run(stream)
        if stream is not None:
            stream.synchronize()
        R, over = self.check()
        if over:
            self.grow(R)
            self.run(stream)
            if stream is not None:
                stream.synchronize()
            R, over = self.check()
            assert not over
        return R

    def destroy(self):
        if getattr(self, "graph", None) is not None:
            _lib.lib().fs_graph_destroy(self.graph)
            self.graph = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def raster_backward_raw(st: RasterState, means3D, opacities, dL_dcolor, *, shs=None, colors_precomp=None,
                        scales=None, rotations=None, cov3D_precomp=None, dL_ddepth=None, dL_dalpha=None, grad_reduce=None):
    """Gradients w.r.t. the op inputs, summed over the V views of `st`.  grad_reduce: a parallel.FusedGradReduce -- the
    per-Gaussian sums are then reduced over the RANKS inside the kernel (peer reductions over NVLink) and the returned
    tensors are the fully summed, replicated gradients (views of the symmetric buffer, valid until its next begin())."""
    L = _lib.lib()
    dev = means3D.device
    P, V, H, W, M = st.P, st.V, st.H, st.W, st.M
    means3D = _f32c(means3D, "means3D"); opacities = _f32c(opacities, "opacities").reshape(-1)
    shs = _f32c(shs, "shs"); colors_precomp = _f32c(colors_precomp, "colors_precomp")
    scales = _f32c(scales, "scales"); rotations = _f32c(rotations, "rotations")
    cov3D_precomp = _f32c(cov3D_precomp, "cov3D_precomp")
    dL_dcolor = _f32c(dL_dcolor, "dL_dcolor"); dL_ddepth = _f32c(dL_ddepth, "dL_ddepth")
    dL_dalpha = _f32c(dL_dalpha, "dL_dalpha")
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
        fused = grad_reduce is not None
        if fused:
            if use_sr or shs is None or st.cov_stride != 9 or not st.sh_layout:
                raise _lib.FreeSplatB200Error("grad_reduce supports the in-place layouts of render_views ([G,3,3] covariances, [G,3,d] harmonics)")
            grad_reduce.begin()
            g.update(means3D=grad_reduce.means, cov3D=grad_reduce.cov, shs=grad_reduce.sh, opacities=grad_reduce.opac.view(P, 1))
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
            dL_dcolor=ptr(dL_dcolor), dL_ddepth=ptr(dL_ddepth), dL_dalpha=ptr(dL_dalpha), dL_dscreen=ptr(dscreen),
            dL_dmeans2D=ptr(g["means2D"]), dL_dmeans3D=ptr(g["means3D"]), dL_dcov3D=ptr(cov_buf), dL_dshs=ptr(g["shs"]),
            dL_dcolors=ptr(g["colors"]), dL_dopacities=ptr(g["opacities"]), dL_dscales=ptr(g["scales"]),
            dL_drotations=ptr(g["rotations"]), peer_delta=ptr(grad_reduce.peer_delta) if fused else None,
            shard_rows=grad_reduce.shard_rows if fused else 0, world=grad_reduce.world if fused else 0)
        check(L.fs_raster_backward(C.byref(a), C.c_void_p(stream)), "fs_raster_backward")
        if fused:
            grad_reduce.finish()
            m_, c_, s_, o_ = grad_reduce.result()
            g.update(means3D=m_, cov3D=c_, shs=s_, opacities=o_.view(P, 1))
    g["screen"] = dscreen
    return g


class _RasterizeViews(torch.autograd.Function):
    """autograd wrapper: V views of one Gaussian set (V=1 is the reference op).  All four outputs that carry floats are
    differentiable: colour, depth (depth_grad=True) and the accumulated alpha 1 - final_T."""

    @staticmethod
    def forward(ctx, means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp, views,
                H, W, sh_degree, scale_modifier, prefiltered, depth_grad, sh_layout=0, cov_stride=6, check_overflow=None,
                grad_reduce=None):
        ctx.grad_reduce = grad_reduce
        if check_overflow is None:
            check_overflow = DEFAULT_CHECK
        needs_bwd = any(ctx.needs_input_grad)
        st = raster_forward_raw(means3D, opacities, views, H, W, shs=shs, colors_precomp=colors_precomp,
                                scales=scales, rotations=rotations, cov3D_precomp=cov3D_precomp, sh_degree=sh_degree,
                                scale_modifier=scale_modifier, prefiltered=prefiltered,
                                check_overflow=check_overflow, sh_layout=sh_layout, cov_stride=cov_stride,
                                reuse_scratch=not needs_bwd)
        color, depth = st.color, st.depth
        if not needs_bwd:
            # inference: nothing is saved; the scratch goes back to the cache (fresh tensors for final_T: alpha reads it)
            radii = st.radii
            alpha = 1.0 - st.final_T
            st.color = st.depth = st.radii = None
            ctx.mark_non_differentiable(radii)
            return color, radii, depth, alpha
        # the OUTPUT tensors must not stay reachable from ctx: output -> grad_fn -> ctx -> st -> output is a reference
        # cycle that only Python's cyclic GC breaks, i.e. every step's workspace (hundreds of MB) would outlive the step
        # and the caching allocator would cudaMalloc afresh each iteration
        st.color = None; st.depth = None
        st.keybuf = None  # sorted keys are only needed by the parity tests (raster_forward_raw)
        ctx.st = st
        ctx.depth_grad = depth_grad
        ctx.save_for_backward(means3D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp)
        ctx.mark_non_differentiable(st.radii)
        ctx.set_materialize_grads(False)       # outputs the loss does not touch arrive as None, not as zero tensors
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
                                rotations=rotations, cov3D_precomp=cov3D_precomp, dL_ddepth=dd, dL_dalpha=g_alpha,
                                grad_reduce=ctx.grad_reduce)
        gm2d = g["means2D"]
        return (g["means3D"], gm2d if st.V > 1 else gm2d[0], g["shs"], g["colors"],
                g["opacities"].reshape(opacities.shape), g["scales"], g["rotations"],
                None if g["cov3D"] is None else g["cov3D"].reshape(cov3D_precomp.shape), None,
                None, None, None, None, None, None, None, None, None, None)


def rasterize_views(means3D, opacities, views, image_height, image_width, *, shs=None, colors_precomp=None,
                    scales=None, rotations=None, cov3D_precomp=None, means2D=None, sh_degree=0, scale_modifier=1.0,
                    prefiltered=False, depth_grad=False, sh_layout=0, cov_stride=6, check_overflow=None, grad_reduce=None):
    """Batched op: -> (color[V,3,H,W], radii[V,P], depth[V,H,W], alpha[V,H,W]).

    sh_layout=1 reads `shs` as [P,3,M] and cov_stride=9 reads `cov3D_precomp` as full [P,3,3] matrices -- the
    layouts of the reference's Gaussians dataclass -- in place (gradients come back in the same layouts).
    check_overflow: None = DEFAULT_CHECK ("deferred": no host read of the device status word; an overflowed workspace is
    reported by the next call, see poll_deferred), "sync" = check before returning."""
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
                                 sh_layout, cov_stride, check_overflow, grad_reduce)


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
            bool(rs.prefiltered), True)         # depth and alpha are differentiable outputs, as in the module this replaces
        return color[0], radii[0], depth[0], alpha[0]
