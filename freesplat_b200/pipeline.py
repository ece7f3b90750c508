"""Host-to-host rendering pipeline: the public entry point for callers whose Gaussians and cameras live in
(pinned) host memory.  Three CUDA streams -- H2D copies, the kernels, D2H copies -- and double-buffered
device / host staging, so the copy engines of step k+1 / k-1 run under the kernels of step k.
(The reference has no counterpart: it renders view by view with a blocking D2H inside the op.)"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from . import _lib, decoder, rasterizer


def bind_to_gpu_numa(device_index: int) -> dict:
    """Pins the calling process to the CPUs that are local to the GPU's PCIe root (NVML's ideal CPU affinity, intersected
    with the CPUs this process may use), so that the pinned staging buffers allocated AFTERWARDS are first-touched on the
    GPU-local NUMA node: on a two-socket host a remote node costs a third of the host-to-device bandwidth (34 instead of
    50 GB/s measured), and the host-to-host path is PCIe-bound.  Returns what was done; never raises."""
    import os
    info = {"bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        phys = device_index
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x.strip() for x in vis.split(",") if x.strip()]
            if device_index < len(ids) and ids[device_index].isdigit():
                phys = int(ids[device_index])
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        local = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1}
        allowed = set(os.sched_getaffinity(0))
        both = sorted(local & allowed)
        info.update(gpu_local_cpus=len(local), allowed_cpus=len(allowed))
        if both and len(both) < len(allowed):
            os.sched_setaffinity(0, both)
            info.update(bound=True, cpus=len(both))
        elif both:
            info.update(bound=False, note="already local")
    except Exception as exc:       # NVML missing, no permission, exotic topology: run unbound
        info["note"] = f"{type(exc).__name__}: {exc}"[:120]
    return info


class HostRenderPipeline:
    """submit(host tensors) -> slot; wait(slot) -> (colour, depth) in pinned host memory.

    Every slot owns static device input buffers, a static rasterizer workspace and ONE CUDA graph (camera records + memset +
    preprocess + tile scan + scatter + sort/render: rasterizer.RasterPlan), so a step costs the host three async copies in,
    one graph launch and two async copies out.  The instance count R stays on the device; the status word rides back with the
    results and is looked at when the slot is next touched (wait() or the submit that reuses it): an overflowed step is
    re-run on a grown workspace from the inputs still resident in the slot, transparently."""
    KEYS = ("extrinsics", "intrinsics", "near", "far", "means", "covariances", "harmonics", "opacities")
    GAUSSIAN_KEYS = ("means", "covariances", "harmonics", "opacities")

    def __init__(self, device, image_shape: Tuple[int, int], n_views: int, depth: int = 2, background=(0.0, 0.0, 0.0),
                 graph: bool = True, shard_group="none"):
        """shard_group: a torch.distributed process group (None = the default group) whose ranks all render views of the SAME
        scene (target views shard over GPUs, the Gaussian set is replicated).  Each rank then uploads only its
        1/world slice of the Gaussian tensors over PCIe and one in-place all-gather per tensor replicates the set over
        NVLink / NVSwitch -- the host link carries the scene once per step instead of once per rank.  submit() becomes a
        collective call (all ranks, same order).  "none": every rank uploads everything."""
        self.dev = torch.device(device)
        self.h, self.w = image_shape
        self.V = n_views
        self.depth = depth
        self.graph = graph
        self.s_h2d, self.s_run, self.s_d2h = (torch.cuda.Stream(self.dev) for _ in range(3))
        self.bg = torch.tensor(background, dtype=torch.float32, device=self.dev)[None].expand(n_views, 3).contiguous()
        self.dev_in: List[Optional[Dict[str, torch.Tensor]]] = [None] * depth
        self.plans: List[Optional[rasterizer.RasterPlan]] = [None] * depth
        self.out_c = [torch.empty((n_views, 3, self.h, self.w), dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.out_d = [torch.empty((n_views, self.h, self.w), dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.out_status = [torch.zeros(4, dtype=torch.int32).pin_memory() for _ in range(depth)]
        self.ev_in_free = [torch.cuda.Event() for _ in range(depth)]     # device inputs of slot may be overwritten
        self.ev_copied = [torch.cuda.Event() for _ in range(depth)]
        self.ev_done = [torch.cuda.Event() for _ in range(depth)]        # kernels of slot finished
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]         # host outputs of slot are valid
        self.group, self.world, self.rank = None, 1, 0
        if shard_group != "none":
            import torch.distributed as dist
            self.group = shard_group
            self.world, self.rank = dist.get_world_size(shard_group), dist.get_rank(shard_group)
        self.h2d_bytes = 0                 # bytes the last submit() moved over the host link
        self.unresolved = [False] * depth  # slot holds a step whose status word has not been looked at yet
        self.reruns = 0
        self.n = 0

    def _plan(self, slot: int, host) -> "rasterizer.RasterPlan":
        d = self.dev_in[slot]
        n = d["harmonics"].shape[-1]
        from math import isqrt
        return rasterizer.RasterPlan(d["means"], d["opacities"], self.h, self.w, shs=d["harmonics"],
                                     cov3D_precomp=d["covariances"].reshape(-1, 9),
                                     cameras=(d["extrinsics"], d["intrinsics"], d["near"], d["far"], self.bg),
                                     sh_degree=isqrt(n) - 1, sh_layout=1, cov_stride=9, graph=self.graph)

    def _copy_out(self, slot: int):
        plan = self.plans[slot]
        with torch.cuda.stream(self.s_d2h):
            self.s_d2h.wait_event(self.ev_done[slot])
            self.out_c[slot].copy_(plan.color, non_blocking=True)
            self.out_d[slot].copy_(plan.depth, non_blocking=True)
            self.out_status[slot].copy_(plan.status, non_blocking=True)
            self.ev_out[slot].record(self.s_d2h)

    def _resolve(self, slot: int):
        """Waits for the slot's results and looks at its status word; re-runs the step if the workspace overflowed."""
        if not self.unresolved[slot]:
            return
        self.ev_out[slot].synchronize()
        self.unresolved[slot] = False
        st = self.out_status[slot]
        if int(st[2]):
            R = (int(st[0]) & 0xFFFFFFFF) | ((int(st[1]) & 0xFFFFFFFF) << 32)
            plan = self.plans[slot]
            self.s_run.synchronize()
            with torch.cuda.stream(self.s_run):
                plan.grow(R)                                  # new workspace + graph; the inputs are still in the slot
                plan.run(self.s_run)
                self.ev_done[slot].record(self.s_run)
            self._copy_out(slot)
            self.ev_out[slot].synchronize()
            self.reruns += 1
            if int(self.out_status[slot][2]):
                raise _lib.FreeSplatB200Error("tile-instance workspace overflowed again after growing it")

    def submit(self, host: Dict[str, torch.Tensor]) -> int:
        """host: pinned CPU tensors named as KEYS (one scene, V target views).  Returns the slot index;
        call wait(slot) before reading the host outputs `out_c[slot]`, `out_d[slot]`."""
        slot = self.n % self.depth
        first_use = self.n < self.depth
        self.n += 1
        self._resolve(slot)        # the step that used this slot `depth` submits ago (long finished in steady state)
        d = self.dev_in[slot]
        if d is not None and any(d[k].shape != host[k].shape for k in self.KEYS):
            d = None               # a scene of another size: new buffers, new plan
            for s in (self.s_h2d, self.s_run, self.s_d2h):
                s.synchronize()
            first_use = True
        with torch.cuda.stream(self.s_h2d):
            if not first_use:
                self.s_h2d.wait_event(self.ev_in_free[slot])
            if d is None:
                d = self.dev_in[slot] = {k: torch.empty(host[k].shape, dtype=torch.float32, device=self.dev) for k in self.KEYS}
                self.plans[slot] = None
            G = host["means"].shape[0]
            sharded = self.world > 1 and G % self.world == 0
            moved = 0
            works = []
            for k in self.KEYS:
                if sharded and k in self.GAUSSIAN_KEYS:
                    n = G // self.world
                    part = d[k][self.rank * n:(self.rank + 1) * n]
                    part.copy_(host[k][self.rank * n:(self.rank + 1) * n], non_blocking=True)
                    moved += part.numel() * 4
                else:
                    d[k].copy_(host[k], non_blocking=True)
                    moved += d[k].numel() * 4
            if sharded:
                import torch.distributed as dist
                for k in self.GAUSSIAN_KEYS:             # in place: this rank's slice already sits where it belongs
                    n = G // self.world
                    works.append(dist.all_gather_into_tensor(d[k], d[k][self.rank * n:(self.rank + 1) * n], group=self.group,
                                                             async_op=True))
                for wk in works:
                    wk.wait()                            # s_h2d waits for the NCCL stream (no host block)
            self.h2d_bytes = moved
            self.ev_copied[slot].record(self.s_h2d)
        if self.plans[slot] is None:
            with torch.cuda.stream(self.s_run):
                self.plans[slot] = self._plan(slot, host)
        with torch.cuda.stream(self.s_run):
            self.s_run.wait_event(self.ev_copied[slot])
            if not first_use:
                self.s_run.wait_event(self.ev_out[slot])          # previous results of this slot were copied out
            self.plans[slot].run(self.s_run)
            self.ev_done[slot].record(self.s_run)
            self.ev_in_free[slot].record(self.s_run)
        self._copy_out(slot)
        self.unresolved[slot] = True
        return slot

    def wait(self, slot: int):
        self._resolve(slot)
        self.ev_out[slot].synchronize()
        return self.out_c[slot], self.out_d[slot]

    def drain(self):
        for slot in range(self.depth):
            self._resolve(slot)
        for s in (self.s_h2d, self.s_run, self.s_d2h):
            s.synchronize()
