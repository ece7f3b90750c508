"""BASELINE config 3's raster part, view-sharded over the ranks WITH autograd (training): every rank renders
its share of the target views of the replicated Gaussian set, back-propagates an MSE loss on them, and the Gaussian
gradients are summed by one NCCL all-reduce (parallel.sync_gaussian_grads).  Checks the summed gradients against a
single-rank render of all views (done redundantly on every rank); prints timings (CUDA events, max over ranks).
    torchrun --nproc-per-node N tools/bench_train_sharded.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from freesplat_b200 import decoder, parallel, synth  # noqa: E402

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
T, h, w = 8, 480, 640
sc = synth.pixel_aligned_scene(seed=0, h=h, w=w, n_context=3, n_target=T, keep=460800).to(dev)
bg = torch.zeros((T, 3), device=dev)
target = torch.rand((T, 3, h, w), generator=torch.Generator().manual_seed(1)).to(dev)
leaf = lambda: [x.detach().clone().requires_grad_(True) for x in (sc.means, sc.covariances, sc.harmonics, sc.opacities)]


FUSED = "--fused" in sys.argv and world > 1
reducer = parallel.FusedGradReduce(int(sc.means.shape[0]), int(sc.harmonics.shape[-1]), dev) if FUSED else None


def nvlink_tx_kib():
    """Sum of the NVLink data TX counters of this rank's GPU (nvidia-smi nvlink -gt d), or None."""
    import re
    import subprocess
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(local)], capture_output=True, text=True, timeout=20).stdout
        return sum(int(x) for x in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out))
    except Exception:
        return None


def sharded(params):
    col, dep, ids = parallel.render_views_sharded(sc.extrinsics, sc.intrinsics, sc.near, sc.far, (h, w), bg, *params,
                                                  grad_reduce=reducer)
    loss = ((col - target[ids]) ** 2).sum() / (T * 3 * h * w)
    loss.backward()
    return loss.detach()


def full(params):
    col, dep = decoder.render_views(sc.extrinsics, sc.intrinsics, sc.near, sc.far, (h, w), bg, *params)
    loss = ((col - target) ** 2).sum() / (T * 3 * h * w)
    loss.backward()
    return loss.detach()


pf = leaf(); full(pf)
ps = leaf(); sharded(ps)
err = max(float((a.grad - b.grad).abs().max() / (b.grad.abs().max() + 1e-30)) for a, b in zip(ps, pf))
ms = []
p = leaf()
tx0 = nvlink_tx_kib()
for it in range(12):                  # same parameter tensors every step (as in a training loop): the allocator reaches a steady state
    for x in p:
        x.grad = None
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sharded(p); e1.record(); torch.cuda.synchronize()
    if it >= 6:
        ms.append(e0.elapsed_time(e1))
tx1 = nvlink_tx_kib()
t = parallel.max_over_ranks([sum(ms) / len(ms)], dev)[0]
errs = parallel.max_over_ranks([err], dev)[0]
if rank == 0:
    res = {"world": world, "target_views": T, "gaussians": int(sc.means.shape[0]), "ms_fwd_bwd_allreduce": t,
           "views_per_s_train": T / (t * 1e-3), "max_rel_grad_diff_vs_single_rank": errs,
           "allreduce_bytes": int(sc.means.shape[0]) * (3 + 9 + 27 + 1) * 4,
           "exchange": "reduce-scatter fused into preprocess_bwd_kernel (peer red.global.add over NVLink) + in-place all-gathers"
                       if FUSED else "pack + NCCL all-reduce + unpack (parallel.sync_gaussian_grads)",
           "nvlink_tx_kib_rank0_over_12_steps": None if tx0 is None or tx1 is None else tx1 - tx0}
    print(json.dumps(res))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"train_sharded_world{world}{'_fused' if FUSED else ''}.json"), "w"))
if world > 1:
    dist.destroy_process_group()
