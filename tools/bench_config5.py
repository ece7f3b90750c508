"""BASELINE config 5 in miniature ("10 views 640x480, view-sharded PTF + raster across the box via NCCL"):
every rank owns context views v % world == rank, all-gathers the per-view candidates over NCCL (the cross-view PTF
gather), runs the identical deterministic fusion, then renders ITS share of the target views.
    torchrun --nproc-per-node N tools/bench_config5.py
Checks that all ranks obtain bit-identical fused Gaussian sets; prints timings (CUDA events, max over ranks)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from freesplat_b200 import decoder, parallel, ptf, synth  # noqa: E402
from ptf_helpers import flat_inputs  # noqa: E402
from test_ptf_gpu import GRU  # noqa: E402

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
V, T, h, w = 10, 18, 480, 640
inp = synth.ptf_inputs(0, V, h, w)                      # every rank can regenerate the synthetic scene; it only USES its own views
feats, coords, dens, wemb, depths, ext, K, hw = flat_inputs(inp)
mine = parallel.shard_views(V, rank, world)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
local_c = [t(x[mine]) for x in (feats, coords, dens, wemb, depths)]
gru = GRU(); gru.load_state_dict(synth.gru_state(0)); gru = gru.to(dev)


def step():
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record()
    full = list(parallel.all_gather_candidates(*local_c, V)) if world > 1 else local_c     # 70 fp32 per candidate, ONE collective
    e[1].record()
    with torch.no_grad():
        F_, X_, E_, Z_ = ptf.fuse_views(gru, full[0], full[1], full[2], full[3], full[4], t(ext), t(K), hw)
    e[2].record()
    N = F_.shape[0]
    # Gaussian head stand-in (the reference's to_gaussians + adapter are out of scope): isotropic covariances, SH from features
    means = X_
    s = (Z_ * 2.0e-3)[:, None, None]
    cov = torch.eye(3, device=dev)[None] * (s * s)
    sh = F_[:, :27].reshape(N, 3, 9).contiguous() * 0.2
    op = torch.sigmoid(F_[:, 27])
    tv = parallel.shard_views(T, rank, world)
    cams = synth.camera_path(T, spacing=0.1, t0=0.3)[tv].to(dev)
    n = len(tv)
    with torch.no_grad():
        col, dep = decoder.render_views(cams, synth.intrinsics(n).to(dev), torch.full((n,), 0.5, device=dev),
                                        torch.full((n,), 15.0, device=dev), (h, w), torch.zeros((n, 3), device=dev), means, cov, sh, op)
    e[3].record()
    torch.cuda.synchronize()
    return [e[i].elapsed_time(e[i + 1]) for i in range(3)], X_, N, col


for _ in range(2):
    ms, X_, N, col = step()
ms, X_, N, col = step()
chk = torch.tensor([float(N), float(X_.double().sum())], dtype=torch.float64, device=dev)
same = True
if world > 1:
    lst = [torch.empty_like(chk) for _ in range(world)]
    dist.all_gather(lst, chk)
    same = all(torch.equal(lst[0], x) for x in lst)
ms = parallel.max_over_ranks(ms, dev)
if rank == 0:
    res = {"world": world, "context_views": V, "target_views": T, "fused_gaussians": N, "identical_on_all_ranks": same,
           "ms_gather": ms[0], "ms_ptf": ms[1], "ms_render_share": ms[2], "finite": bool(torch.isfinite(col).all())}
    print(json.dumps(res))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"config5_world{world}.json"), "w"))
if world > 1:
    dist.destroy_process_group()
