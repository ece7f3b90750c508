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
        t = parallel.max_over_ranks([1.0 + rank, 5.0 - rank], "cpu")
        q.put((rank, mine, ok, t))
    finally:
        dist.destroy_process_group()


def _run(num_views, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, num_views, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(res)


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
