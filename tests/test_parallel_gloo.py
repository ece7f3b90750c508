"""CPU, world_size 2 over gloo: the host-side logic of the N>1 path (view sharding, the cross-view PTF
gather, timing reduction)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from freesplat_b200 import parallel


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _spawn(target, world, extra=(), attempts=3):
    """Runs `target(rank, world, port, *extra, q)` in `world` processes; a failed rendezvous (port taken between probing and
    binding on a busy host) is retried on a fresh port."""
    last = None
    for _ in range(attempts):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        ps = [ctx.Process(target=target, args=(r, world, port, *extra, q)) for r in range(world)]
        for p in ps:
            p.start()
        try:
            res = [q.get(timeout=180) for _ in ps]
        except Exception as exc:             # queue.Empty: a worker died or hung
            last = exc
            for p in ps:
                p.kill()
            continue
        ok = True
        for p in ps:
            p.join(timeout=60)
            ok = ok and p.exitcode == 0
        if ok:
            return res
        last = RuntimeError("worker exit codes: %s" % [p.exitcode for p in ps])
    raise last


def _worker(rank, world, port, num_views, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = parallel.shard_views(num_views, rank, world)
        # every view's candidates are a deterministic function of its global index
        local = torch.stack([torch.full((4, 3), float(v)) + torch.arange(3.0) for v in mine]) if mine else torch.zeros((0, 4, 3))
        full = parallel.all_gather_views(local, num_views)
        want = torch.stack([torch.full((4, 3), float(v)) + torch.arange(3.0) for v in range(num_views)])
        ok = torch.equal(full, want)
        # packed candidate gather: five tensors, one collective
        n = len(mine); HW, F = 6, 4
        mk = lambda v, c: torch.full((HW, c), float(v)) + torch.arange(float(c))
        loc = (torch.stack([mk(v, F) for v in mine]) if n else torch.zeros((0, HW, F)),
               torch.stack([mk(v, 3) * 2 for v in mine]) if n else torch.zeros((0, HW, 3)),
               *[torch.stack([mk(v, 1)[:, 0] * k for v in mine]) if n else torch.zeros((0, HW)) for k in (3, 5, 7)])
        gf, gc, gd, gw, gz = parallel.all_gather_candidates(*loc, num_views)
        ok = ok and torch.equal(gf, torch.stack([mk(v, F) for v in range(num_views)]))
        ok = ok and torch.equal(gc, torch.stack([mk(v, 3) * 2 for v in range(num_views)]))
        ok = ok and all(torch.equal(g_, torch.stack([mk(v, 1)[:, 0] * k for v in range(num_views)])) for g_, k in ((gd, 3), (gw, 5), (gz, 7)))
        # ViewExchange: per-view broadcasts into the packed blocks the PTF kernels read (strided field views, no re-packing)
        ex = parallel.ViewExchange(num_views, HW, F, "cpu")
        ef, ec, ed, ew, ez = ex.exchange(*loc)
        ok = ok and torch.equal(ef, gf) and torch.equal(ec, gc) and torch.equal(ed, gd) and torch.equal(ew, gw) and torch.equal(ez, gz)
        ok = ok and all(ef[v].is_contiguous() and ec[v].is_contiguous() for v in range(num_views)) and ef.data_ptr() == ex.block.data_ptr()
        t = parallel.max_over_ranks([1.0 + rank, 5.0 - rank], "cpu")
        q.put((rank, mine, ok, t))
    finally:
        dist.destroy_process_group()


def _run(num_views, world=2):
    return sorted(_spawn(_worker, world, (num_views,)))


def test_shard_partition_is_exact():
    for n in (1, 3, 9, 18):
        for world in (1, 2, 4, 8):
            owned = [parallel.shard_views(n, r, world) for r in range(world)]
            flat = sorted(v for o in owned for v in o)
            assert flat == list(range(n))
            assert all(parallel.owner_of(v, world) == r for r, o in enumerate(owned) for v in o)
            assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1


def test_gather_world2_uneven_and_even():
    for n in (5, 6):
        res = _run(n)
        assert [r[1] for r in res] == [parallel.shard_views(n, 0, 2), parallel.shard_views(n, 1, 2)]
        assert all(r[2] for r in res)
        assert all(r[3] == [2.0, 5.0] for r in res)


def _fake_render(ext, intr, near, far, image_shape, bg, means, cov, sh, op):
    """Differentiable stand-in for the rasterizer on the CPU: every view mixes all Gaussians with view-dependent weights."""
    h, w = image_shape
    wv = ext[:, 0, 3].reshape(-1, 1)                                         # [v,1]
    val = (means.sum(1) + cov.sum((1, 2)) * 0.5 + sh.sum((1, 2)) * 0.25 + op)[None] * wv      # [v,G]
    col = val.sum(1).reshape(-1, 1, 1, 1).expand(-1, 3, h, w) * torch.arange(1.0, 4.0).reshape(1, 3, 1, 1)
    return col, col[:, 0]


def _grad_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        G, V = 7, 5
        leaf = [torch.randn(s, generator=g).requires_grad_(True) for s in ((G, 3), (G, 3, 3), (G, 3, 9), (G,))]
        ext = torch.eye(4).repeat(V, 1, 1); ext[:, 0, 3] = torch.arange(1.0, V + 1)
        z = torch.zeros(V)
        col, dep, ids = parallel.render_views_sharded(ext, torch.eye(3).repeat(V, 1, 1), z, z, (2, 3), torch.zeros(V, 3), *leaf,
                                                      render_fn=_fake_render)
        (col.sum() + dep.sum()).backward()                                   # loss of the LOCAL views only
        # one view, two ranks: rank 1 renders nothing but must still take part in the gradient all-reduce
        leaf1 = [t.detach().clone().requires_grad_(True) for t in leaf]
        c1, d1, ids1 = parallel.render_views_sharded(ext[:1], torch.eye(3).repeat(1, 1, 1), z[:1], z[:1], (2, 3), torch.zeros(1, 3),
                                                     *leaf1, render_fn=_fake_render)
        (c1.sum() + d1.sum()).backward()
        assert len(ids1) == (1 if rank == 0 else 0) and c1.shape[0] == len(ids1)
        ref1 = [t.detach().clone().requires_grad_(True) for t in leaf]
        cr, dr = _fake_render(ext[:1], None, None, None, (2, 3), None, *ref1)
        (cr.sum() + dr.sum()).backward()
        for a_, b_ in zip(leaf1, ref1):
            torch.testing.assert_close(a_.grad, b_.grad, rtol=1e-6, atol=1e-6)
        q.put((rank, ids, [t.grad.clone() for t in leaf]))
    finally:
        dist.destroy_process_group()


def test_sharded_render_sums_gaussian_gradients_over_ranks():
    res = sorted(_spawn(_grad_worker, 2), key=lambda r: r[0])
    assert res[0][1] == [0, 2, 4] and res[1][1] == [1, 3]
    # single-process reference: all views on one rank
    g = torch.Generator().manual_seed(0)
    G, V = 7, 5
    leaf = [torch.randn(s, generator=g).requires_grad_(True) for s in ((G, 3), (G, 3, 3), (G, 3, 9), (G,))]
    ext = torch.eye(4).repeat(V, 1, 1); ext[:, 0, 3] = torch.arange(1.0, V + 1)
    col, dep = _fake_render(ext, None, None, None, (2, 3), None, *leaf)
    (col.sum() + dep.sum()).backward()
    for r in res:
        for got, want in zip(r[2], leaf):
            torch.testing.assert_close(got, want.grad, rtol=1e-6, atol=1e-6)


def _fake_cost_volume(cur_feats, src_feats, src_extrinsics, src_poses, src_Ks, cur_invK, min_depth, max_depth):
    """Stand-in operator (the real one needs a GPU): mixes the reference map with every source map and the relative pose."""
    w = src_extrinsics[:, :, 0, 3].abs() + 1.0                                        # [n, K]
    return (cur_feats[:, None] * src_feats * w[:, :, None, None, None]).sum(2).cumsum(1) + cur_invK[:, 0, 0].reshape(-1, 1, 1, 1)


def _cv_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        V, C, Hf, Wf = 5, 4, 3, 6
        g = torch.Generator().manual_seed(1)
        feats = torch.randn((V, C, Hf, Wf), generator=g)
        ext = torch.eye(4).repeat(V, 1, 1); ext[:, 0, 3] = torch.arange(V) * 0.3
        Kf = torch.eye(3).repeat(V, 1, 1); Kf[:, 0, 0] = 40.0; Kf[:, 1, 1] = 30.0; Kf[:, 0, 2] = Wf / 2; Kf[:, 1, 2] = Hf / 2
        mine = parallel.shard_views(V, rank, world)
        local = feats[mine].clone().requires_grad_(True)
        vol, ids = parallel.cost_volume_sharded(_fake_cost_volume, local, ext, Kf, 0.5, 15.0)
        vol.sum().backward()                                                          # loss over the LOCAL volumes only
        q.put((rank, ids, vol.detach(), local.grad.clone()))
    finally:
        dist.destroy_process_group()


def test_sharded_cost_volume_matches_single_rank_forward_and_backward():
    res = sorted(_spawn(_cv_worker, 2), key=lambda r: r[0])
    V, C, Hf, Wf = 5, 4, 3, 6
    g = torch.Generator().manual_seed(1)
    feats = torch.randn((V, C, Hf, Wf), generator=g).requires_grad_(True)
    ext = torch.eye(4).repeat(V, 1, 1); ext[:, 0, 3] = torch.arange(V) * 0.3
    Kf = torch.eye(3).repeat(V, 1, 1); Kf[:, 0, 0] = 40.0; Kf[:, 1, 1] = 30.0; Kf[:, 0, 2] = Wf / 2; Kf[:, 1, 2] = Hf / 2
    vol, ids = parallel.cost_volume_sharded(_fake_cost_volume, feats, ext, Kf, 0.5, 15.0)      # no process group: one rank owns all
    vol.sum().backward()
    assert ids == list(range(V)) and tuple(parallel.source_view_indices(3).tolist()) == ([1, 2], [0, 2], [0, 1])
    for rank, rids, rvol, rgrad in res:
        torch.testing.assert_close(rvol, vol.detach()[rids], rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(rgrad, feats.grad[rids], rtol=1e-5, atol=1e-6)
