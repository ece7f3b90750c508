"""Multi-GPU plumbing of the hot path: one process per GPU, torch.distributed.

  * target views are independent units -> `shard_views` partitions them round-robin, every rank renders
    its share of the replicated Gaussian set, no data-path collective ("weak" scaling in bench.py);
  * context views (backbone + cost volume) shard the same way after one all-gather of the stride-4
    matching features;
  * PTF is an order-dependent fold -> "replicas only": `all_gather_views` collects the per-view
    candidates of all ranks (the "cross-view PTF gather" of BASELINE.json) in global view order and every
    rank runs the identical deterministic fusion.

  * training: every rank back-propagates the loss of ITS target views into the replicated Gaussian set, so the
    gradients w.r.t. means / covariances / harmonics / opacities are partial sums: `sync_gaussian_grads` is the
    identity in forward and ONE all-reduce(SUM) of the packed [G, 3+9+d+1] gradient buffer in backward (the only
    data-path collective of the training step; NCCL rings over NVLink / NVSwitch), placed between the rasterizer
    backward and the PTF / encoder backward.  `render_views_sharded` bundles sharding, rendering and that collective.

Works with the NCCL backend on GPUs and with gloo on the CPU (tests/test_parallel_gloo.py).
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_views(num_views: int, rank: int, world: int) -> List[int]:
    """Round-robin ownership: view v belongs to rank v % world."""
    return [v for v in range(num_views) if v % world == rank]


def owner_of(view: int, world: int) -> int:
    return view % world


def all_gather_views(local: torch.Tensor, num_views: int, group=None) -> torch.Tensor:
    """local: [n_local, ...] holding this rank's views (in increasing global view index).
    Returns [num_views, ...] in global view order on every rank.  Ranks may own different counts.
    ONE collective into one flat buffer (all_gather_into_tensor: NCCL's native all-gather, no per-rank staging copies)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = [len(shard_views(num_views, r, world)) for r in range(world)]
    assert local.shape[0] == counts[rank], (local.shape, counts, rank)
    mx = max(counts)
    pad = local.contiguous()
    if local.shape[0] < mx:
        pad = torch.cat([pad, local.new_zeros((mx - local.shape[0],) + tuple(local.shape[1:]))], 0)
    flat = local.new_empty((world * mx,) + tuple(local.shape[1:]))
    try:
        dist.all_gather_into_tensor(flat, pad, group=group)
    except (RuntimeError, NotImplementedError):          # backends without the flat variant
        bufs = list(flat.view((world, mx) + tuple(local.shape[1:])).unbind(0))
        dist.all_gather(bufs, pad, group=group)
    # rank r's slot k holds global view r + k*world  ->  global view v sits at flat[(v % world) * mx + v // world]
    src = torch.tensor([(v % world) * mx + v // world for v in range(num_views)], dtype=torch.long, device=flat.device)
    return flat.index_select(0, src)


def all_gather_candidates(feats, coords, dens, wemb, depths, num_views: int, group=None):
    """The cross-view PTF gather: the per-view candidates of this rank -- feats [n,HW,F], coords [n,HW,3],
    dens / wemb / depths [n,HW] -- packed into ONE [n, HW, F+6] buffer (70 floats per candidate at F = 64) and gathered
    with one collective; returns the five tensors for all `num_views` views in global view order."""
    n, HW, F = feats.shape
    packed = torch.cat([feats, coords, dens.reshape(n, HW, 1), wemb.reshape(n, HW, 1), depths.reshape(n, HW, 1)], dim=-1)
    full = all_gather_views(packed, num_views, group)
    return (full[..., :F].contiguous(), full[..., F:F + 3].contiguous(), full[..., F + 3].contiguous(),
            full[..., F + 4].contiguous(), full[..., F + 5].contiguous())


class _AllGatherViewsGrad(torch.autograd.Function):
    """all_gather_views with a gradient: every rank back-propagates into ALL views' features (its cost volumes read the
    other ranks' maps as sources), so the backward is one all-reduce(SUM) of the [V, ...] gradient, of which each rank
    keeps the rows of the views it owns."""

    @staticmethod
    def forward(ctx, local, num_views, group):
        ctx.num_views, ctx.group = num_views, group
        return all_gather_views(local, num_views, group)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        world = dist.get_world_size(ctx.group)
        if world > 1:
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
        mine = shard_views(ctx.num_views, dist.get_rank(ctx.group), world)
        return g[torch.tensor(mine, dtype=torch.long, device=g.device)], None, None


def source_view_indices(num_views: int) -> torch.Tensor:
    """[V, V-1]: the sources of reference view v are all other views in index order (encoder_freesplat.py:236-238, the
    `not use_local` branch that ScanNet 2/3-view and the 10-view FVT configs with num_views >= V take)."""
    full = torch.arange(num_views)[None].repeat(num_views, 1)
    return full[full != torch.arange(num_views)[:, None]].view(num_views, num_views - 1)


def cost_volume_sharded(cost_volume_fn, local_feats, extrinsics, feat_intrinsics, near, far, group=None, src_indices=None):
    """context views shard over the ranks (view v on rank v % world); ONE all-gather of the stride-4
    matching features (3.7 MB per 640x480 view) and every rank builds the cost volumes of ITS reference views against
    sources that may live anywhere.

    local_feats [n_local, C, H', W'] (this rank's views, increasing global index); extrinsics [V,4,4] (c2w) and
    feat_intrinsics [V,3,3] (pixel intrinsics at feature resolution) of ALL views; near / far: scalars or [1,1,1,1] tensors.
    cost_volume_fn(cur_feats=, src_feats=, src_extrinsics=, src_poses=, src_Ks=, cur_invK=, min_depth=, max_depth=) is the
    operator (freesplat_b200.cost_volume.AVGFeatureVolumeManager instance).  The geometry follows
    encoder_freesplat.py:248-273.  Returns (volume [n_local, D, H', W'], owned view ids)."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    V = extrinsics.shape[0]
    mine = shard_views(V, rank, world)
    if world > 1:
        feats = _AllGatherViewsGrad.apply(local_feats, V, group) if local_feats.requires_grad else all_gather_views(local_feats, V, group)
    else:
        feats = local_feats
    if src_indices is None:
        src_indices = source_view_indices(V)
    dev = feats.device
    cur = torch.tensor(mine, dtype=torch.long, device=dev)
    if not mine:
        return feats.new_zeros((0,)), mine
    src = src_indices.to(dev)[cur]                                         # [n, K]
    src_ext = extrinsics[src]                                              # [n, K, 4, 4]
    cur_ext = extrinsics[cur]
    src_cam_T_cur_cam = src_ext.inverse() @ cur_ext[:, None]
    cur_cam_T_src_cam = cur_ext.inverse()[:, None] @ src_ext
    n, K = src.shape
    src_K = torch.eye(4, device=dev)[None, None].repeat(n, K, 1, 1)
    src_K[:, :, :3, :3] = feat_intrinsics[src]
    cur_invK = torch.eye(4, device=dev)[None].repeat(n, 1, 1)
    cur_invK[:, :3, :3] = feat_intrinsics[cur].inverse()
    vol = cost_volume_fn(cur_feats=feats[cur], src_feats=feats[src], src_extrinsics=src_cam_T_cur_cam, src_poses=cur_cam_T_src_cam,
                         src_Ks=src_K, cur_invK=cur_invK, min_depth=near, max_depth=far)
    return vol, mine


class ViewExchange:
    """The cross-view PTF gather without staging copies and without a bulk barrier.

    Every rank keeps ONE packed block per context view, [feats HW x F | coords HW x 3 | dens HW | wemb HW | depth HW]
    (F + 6 floats per candidate: 86 MB per 640 x 480 view at F = 64), and hands the PTF kernels strided views of those blocks:
    the fields of a view are dense sub-arrays, so nothing is re-packed on either side.  `exchange` copies the views this rank
    owns (view v belongs to rank v % world) into their blocks and enqueues, in view order on a side stream, one in-place all-gather
    per full round of `world` consecutive views (one broadcast per view for the tail); `ready_events[v]` fires when view v has
    landed.  (A scatter + all-gather per view was tried instead of the broadcasts: slower, 9.2 vs 7.85 ms at 4 ranks.)  The fold is sequential in v and step v needs view v only
    (encoder_freesplat.py:443-519), so `ptf.fuse_views(..., view_ready=ready_events)` folds view v while views v+1.. are
    still in flight over NVLink: the exchange costs its first two views, not all ten."""

    def __init__(self, num_views: int, HW: int, F: int, device, group=None, rounds: Optional[bool] = None):
        self.V, self.HW, self.F, self.dev, self.group = num_views, HW, F, device, group
        if rounds is None:
            rounds = os.environ.get("FREESPLAT_B200_EXCHANGE_ROUNDS", "0") == "1"
        self.rounds = rounds      # False: one broadcast per view (measured at 4 / 8 ranks: 7.85 / 8.33 ms for config 5)
        self.block = torch.empty((num_views, HW * (F + 6)), dtype=torch.float32, device=device)
        self.ready_events = [torch.cuda.Event() for _ in range(num_views)]
        self.stream = torch.cuda.Stream(device) if (isinstance(device, torch.device) and device.type == "cuda") or \
            str(device).startswith("cuda") else None
        o = 0
        self._fields = []
        for width in (F, 3, 1, 1, 1):
            self._fields.append((o, width))
            o += HW * width

    def fields(self):
        """(feats [V,HW,F], coords [V,HW,3], dens [V,HW], wemb [V,HW], depths [V,HW]): views into the blocks."""
        out = []
        for o, width in self._fields:
            t = self.block[:, o:o + self.HW * width]
            out.append(t.view(self.V, self.HW, width) if width > 1 else t)
        return tuple(out)

    def exchange(self, feats, coords, dens, wemb, depths):
        """Arguments: this rank's views ([n_local, HW, ...], increasing global index).  Returns fields()."""
        world = dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if world > 1 else 0
        mine = shard_views(self.V, rank, world)
        full = self.fields()
        for k, v in enumerate(mine):
            for dst, src in zip(full, (feats, coords, dens, wemb, depths)):
                dst[v].copy_(src[k].reshape(dst[v].shape))
        if world == 1:
            if self.stream is not None:
                for e in self.ready_events:
                    e.record()
            return full
        cuda = self.stream is not None
        if cuda:
            self.stream.wait_stream(torch.cuda.current_stream(self.dev))
        ctx = torch.cuda.stream(self.stream) if cuda else _null()
        with ctx:
            v = 0
            while v < self.V:
                if self.rounds and v + world <= self.V:
                    # a full round of `world` consecutive views, one per rank: ONE in-place all-gather (every NVLink port of
                    # every GPU busy: ~5x the rate of `world` ring broadcasts); all views of the round become ready together
                    out = self.block[v:v + world]
                    try:
                        w = dist.all_gather_into_tensor(out.view(-1), self.block[v + rank], group=self.group, async_op=True)
                        w.wait()
                    except (RuntimeError, NotImplementedError):      # backends without the flat variant (gloo)
                        dist.all_gather(list(out.unbind(0)), self.block[v + rank].clone(), group=self.group)
                    if cuda:
                        for k in range(world):
                            self.ready_events[v + k].record(self.stream)
                    v += world
                else:
                    src = dist.get_global_rank(self.group, owner_of(v, world)) if self.group is not None else owner_of(v, world)
                    w = dist.broadcast(self.block[v], src=src, group=self.group, async_op=True)
                    w.wait()                              # the side stream (not the host) waits for NCCL
                    if cuda:
                        self.ready_events[v].record(self.stream)
                    v += 1
        return full


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def max_over_ranks(values: Sequence[float], device, group=None) -> List[float]:
    """Timing reduction used by bench.py (device-timed milliseconds, max over ranks)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return [float(x) for x in t]


class _SyncGaussianGrads(torch.autograd.Function):
    """Identity on the replicated Gaussian tensors; backward sums their gradients over the ranks with ONE all-reduce
    of a packed buffer (a bucket sized for launch latency: 40 floats per Gaussian at SH degree 2, 120 MB at G = 750 k)."""

    @staticmethod
    def forward(ctx, group, *tensors):
        ctx.group = group
        ctx.shapes = [t.shape for t in tensors]
        return tuple(t.view_as(t) for t in tensors)

    @staticmethod
    def backward(ctx, *grads):
        n0 = ctx.shapes[0][0]
        cols = [int(torch.Size(s).numel() // max(n0, 1)) for s in ctx.shapes]
        ref = next(g for g in grads if g is not None)
        packed = ref.new_zeros((n0, sum(cols)))
        o = 0
        for g, c in zip(grads, cols):
            if g is not None:
                packed[:, o:o + c] = g.reshape(n0, c)
            o += c
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(ctx.group) > 1:
            dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=ctx.group)
        out, o = [], 0
        for s, c in zip(ctx.shapes, cols):
            out.append(packed[:, o:o + c].reshape(s))
            o += c
        return (None, *out)


def sync_gaussian_grads(*tensors, group=None):
    """tensors: per-Gaussian tensors that are replicated on every rank ([G, ...] each).  Returns them unchanged; their
    gradients are summed over the ranks in backward."""
    return _SyncGaussianGrads.apply(group, *tensors)


class FusedGradReduce:
    """Reduce-scatter of the Gaussian gradients FUSED into the rasterizer's backward kernel (training).

    The four gradient tensors (means [G,3], covariances [G,3,3], harmonics [G,3,d], opacities [G]) live in ONE symmetric
    allocation that every rank maps (torch.distributed._symmetric_memory: CUDA VMM over NVLink / NVSwitch; PyTorch is the
    plumbing that exchanges the handles).  `preprocess_bwd_kernel` of every rank adds its per-Gaussian partial sums directly
    into the owner rank's slice (red.global.add on peer addresses) while it is still computing the next Gaussians, so the
    reduction traffic overlaps the kernel instead of following it; after a cross-rank barrier each rank holds the complete
    sums of its slice, and one in-place all-gather per tensor replicates them.  Compared with sync_gaussian_grads (pack,
    NCCL all-reduce of 160 B per Gaussian, unpack) the all-reduce's reduce half costs no separate pass over HBM.

    Use: `rasterize_views(..., grad_reduce=obj)` / `render_views_sharded(..., grad_reduce=obj)`; every rank must render
    at least one view per step (all ranks run the same barriers / collectives)."""

    def __init__(self, num_gaussians: int, d_sh: int, device, group=None):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        G = num_gaussians
        self.G, self.d_sh = G, d_sh
        self.shard_rows = (G + self.world - 1) // self.world
        if G % self.world:
            raise ValueError(f"FusedGradReduce needs the Gaussian count ({G}) to be a multiple of the world size ({self.world})")
        cols = (3, 9, 3 * d_sh, 1)
        self.buf = symm.empty(G * sum(cols), dtype=torch.float32, device=device)
        self.hdl = symm.rendezvous(self.buf, self.group)
        o, views = 0, []
        for c in cols:
            views.append(self.buf[o:o + G * c])
            o += G * c
        self.means, self.cov, self.sh, self.opac = (views[0].view(G, 3), views[1].view(G, 3, 3), views[2].view(G, 3, d_sh), views[3])
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.peer_delta = torch.tensor([p - ptrs[self.rank] for p in ptrs], dtype=torch.int64, device=device)
        self.nvlink_bytes_per_step = G * sum(cols) * 4 * (self.world - 1) // self.world    # reds this rank sends to peers

    def begin(self):
        """Zero the buffers and wait until every rank has done so (device-side barrier on the current stream)."""
        self.buf.zero_()
        self.hdl.barrier(channel=0)

    def finish(self):
        """Barrier (all peers' reductions have landed), then replicate the owner slices: in-place all-gathers."""
        self.hdl.barrier(channel=1)
        n = self.shard_rows
        for t in (self.means, self.cov, self.sh, self.opac):
            dist.all_gather_into_tensor(t, t[self.rank * n:(self.rank + 1) * n], group=self.group)
        return self.means, self.cov, self.sh, self.opac

    def result(self):
        """Private copies of the summed gradients (autograd may keep a returned tensor as a leaf's .grad; the symmetric
        buffer itself is zeroed by the next begin())."""
        return self.means.clone(), self.cov.clone(), self.sh.clone(), self.opac.clone()


def render_views_sharded(extrinsics, intrinsics, near, far, image_shape, background_color, means, covariances, harmonics,
                         opacities, group=None, render_fn=None, grad_reduce=None, **kw):
    """View-sharded rendering of one scene: all arguments are the FULL [V, ...] camera tensors and the
    replicated Gaussian set; this rank renders views `shard_views(V, rank, world)` and returns (color, depth, view_ids)
    for them.  Under autograd the Gaussian gradients of all ranks are summed (sync_gaussian_grads)."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    ids = shard_views(extrinsics.shape[0], rank, world)
    if render_fn is None:
        from .decoder import render_views as render_fn
    if grad_reduce is not None:
        if not ids:
            raise ValueError("grad_reduce (fused reduce-scatter) needs at least one view on every rank")
        kw = dict(kw, grad_reduce=grad_reduce)      # the reduction happens inside the rasterizer's backward kernel
    elif torch.is_grad_enabled() and any(t.requires_grad for t in (means, covariances, harmonics, opacities)):
        means, covariances, harmonics, opacities = sync_gaussian_grads(means, covariances, harmonics, opacities, group=group)
    sel = torch.tensor(ids, dtype=torch.long, device=extrinsics.device)
    if not ids:
        # no view for this rank: empty outputs that still hang off the synced tensors, so that this rank's backward
        # reaches the all-reduce (a rank that skipped the collective would dead-lock the others)
        h, w = image_shape
        z = (means.sum() + covariances.sum() + harmonics.sum() + opacities.sum()) * 0.0
        return z.reshape(1, 1, 1, 1).expand(0, 3, h, w), z.reshape(1, 1, 1).expand(0, h, w), ids
    color, depth = render_fn(extrinsics[sel], intrinsics[sel], near[sel], far[sel], image_shape, background_color[sel], means,
                             covariances, harmonics, opacities, **kw)
    return color, depth, ids
