#!/usr/bin/env python
"""bench.py -- rendered views/s @ 640x480 of the B200 rasterizer hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the CPU restatement of the reference path

Workload (BASELINE config 2, "ScanNet 2-views 640x480, ~300k Gaussians, forward raster only"):
one scene of 307 200 pixel-aligned Gaussians (SH degree 2, precomputed covariances) rendered into
T=3 target views (assets/evaluation_index_scannet_2views.json holds 3 targets per scene) at 640x480.
A step = one pass of the hot path over that batch: preprocess -> tile binning -> render, all views.

value  : views/s with the scene and camera records resident in HBM (CUDA events, per-step, L2 flushed
         between steps by writing a 256 MiB buffer outside the timed events).
e2e    : the same metric through the public call a user makes (`freesplat_b200.decoder.render_views`)
         starting from PINNED HOST tensors: H2D of Gaussians + cameras, render, D2H of colour + depth.
Multi-GPU: target views shard over ranks (each rank renders its own T views of the replicated
Gaussian set; no data-path collective) -> weak scaling; value = all views / max-over-ranks time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, P, T_VIEWS = 480, 640, 307200, 3
WORKLOAD = f"scannet_2views_{W}x{H}_P{P}_targets{T_VIEWS}_raster_fwd"
METRIC = "rendered views/sec @640x480"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def _ncu_traffic(kernel_substr: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of `kernel` from the committed ncu export (profiles/), bytes per launch."""
    import csv
    path = os.path.join(ROOT, "profiles", "r1_fwd_step_ncu_raw.csv")
    try:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in rows[2:]:
            if kernel_substr in r[ik]:
                return float(r[ir]) * mult[units[ir]] + float(r[iw]) * mult[units[iw]]
    except Exception:
        return None
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def host_threads() -> int:
    """Host threads the CPU reference may use: CPU affinity, capped by the cgroup CPU quota (oversubscribing
    the quota made the 128-thread run on the GPU box 2.4x SLOWER than 64 threads)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = max(1, min(n, int(int(q) / int(per))))
    except Exception:
        pass
    return n


def make_scene(rank: int):
    from freesplat_b200 import synth
    sc = synth.pixel_aligned_scene(seed=0, h=H, w=W, n_context=2, n_target=T_VIEWS, keep=P)
    if rank:
        # view sharding: rank r renders a different set of target cameras of the same Gaussians
        sc.extrinsics = synth.camera_path(T_VIEWS, spacing=0.08, t0=0.37 + 0.11 * rank)
    return sc


def cpu_reference_views_per_s(sc, n_views: int, repeats: int):
    """Times the CPU restatement of the reference rasterizer (oracle/raster_oracle.c, all host cores)."""
    from oracle import raster as oracle
    from tests.helpers import view_inputs
    inps = [view_inputs(sc, v)[0] for v in range(n_views)]
    best = None
    for n in sorted({host_threads(), max(1, host_threads() // 2)}, reverse=True):
        oracle.set_num_threads(n)
        oracle.forward(**inps[0])  # warm (page in, OpenMP pool)
        t0 = time.perf_counter(); oracle.forward(**inps[0]); dt = time.perf_counter() - t0
        if best is None or dt < best[1]:
            best = (n, dt)
    oracle.set_num_threads(best[0])
    t0 = time.perf_counter()
    for _ in range(repeats):
        for inp in inps:
            oracle.forward(**inp)
    dt = time.perf_counter() - t0
    return n_views * repeats / dt, oracle.num_threads(), dt


def ops_section(dev):
    """Cost volume and PTF at BASELINE config-3 sizes: GPU time of the product kernels next to the CPU restatements
    (oracle/, the cpu_baseline leg) on bounded samples, scaled to the full size (the sample is stated)."""
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from freesplat_b200 import ptf, synth
    from freesplat_b200.cost_volume import AVGFeatureVolumeManager
    from oracle import cost_volume as ocv, ptf as optf
    from ptf_helpers import flat_inputs, torch_inverses
    from test_ptf_gpu import GRU
    out = {}
    torch.set_num_threads(host_threads())

    def gpu_ms(fn, n=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    # ---- cost volume: 3 reference views, K = 2, 48 x 120 x 160, D = 128 ----
    V, K, Hf, Wf, D = 3, 2, 120, 160, 128
    inp = synth.cost_volume_inputs(0, V, K, 48, Hf, Wf)
    mlp = synth.cost_volume_mlp(0)
    m = AVGFeatureVolumeManager(Hf, Wf, num_depth_bins=D, matching_dim_size=48).to(dev)
    ginp = {k: v.to(dev) for k, v in inp.items()}
    with torch.no_grad():
        g_ms = gpu_ms(lambda: m(**ginp))
    Dsub = 32                                                  # CPU sample: 1 reference view, 32 of the 128 planes
    sub = {k: v[:1] for k, v in inp.items() if k not in ("min_depth", "max_depth")}
    t0 = time.perf_counter()
    ocv.forward(sub["cur_feats"], sub["src_feats"], sub["src_extrinsics"], sub["src_Ks"], sub["cur_invK"], inp["min_depth"],
                inp["max_depth"], mlp, Dsub, plane_chunk=8)
    c_s = time.perf_counter() - t0
    out["cost_volume_cfg3_fwd"] = {"gpu_ms": g_ms, "cpu_port_ms": c_s * 1e3 * (D / Dsub) * V, "cores": host_threads(),
                                   "sample": f"1 of {V} reference views, {Dsub} of {D} planes ({c_s:.1f} s), scaled"}
    # ---- PTF: 3 views of 640 x 480 ----
    pin = synth.ptf_inputs(0, 3, 480, 640)
    feats, coords, dens, wemb, depths, ext, Kn, hw = flat_inputs(pin)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    gru = GRU(); gru.load_state_dict(synth.gru_state(0)); gru = gru.to(dev)
    gargs = [t(x) for x in (feats, coords, dens, wemb, depths, ext, Kn)]
    with torch.no_grad():
        g_ms = gpu_ms(lambda: ptf.fuse_views(gru, *gargs, hw), n=3)
    t0 = time.perf_counter()
    optf.fuse(feats, coords, dens, wemb, depths, ext, Kn, hw, optf.torch_gru_fn(synth.gru_state(0)), E_invs=torch_inverses(ext))
    c_s = time.perf_counter() - t0
    out["ptf_3views_640x480"] = {"gpu_ms": g_ms, "cpu_port_ms": c_s * 1e3, "cores": host_threads(),
                                 "sample": f"the full 3-view fold ({c_s:.1f} s)"}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sc = make_scene(0)
    from oracle import raster as oracle
    from tests.helpers import view_inputs
    # torchrun exports OMP_NUM_THREADS=1: the reference arm uses every host thread it is allowed to
    # torchrun exports OMP_NUM_THREADS=1; pick the best of {all allowed threads, half of them (SMT siblings)}
    best = None
    for n in sorted({host_threads(), max(1, host_threads() // 2)}, reverse=True):
        oracle.set_num_threads(n)
        oracle.forward(**view_inputs(sc, 0)[0])
        t0 = time.perf_counter(); oracle.forward(**view_inputs(sc, 0)[0]); dt = time.perf_counter() - t0
        if best is None or dt < best[1]:
            best = (n, dt)
    oracle.set_num_threads(best[0])
    inps = [view_inputs(sc, v)[0] for v in range(T_VIEWS)]
    for _ in range(max(args.warmup, 1)):
        oracle.forward(**inps[0])
    # bounded sample: a step renders all T views of the scene; beyond 120 steps only view (k mod T) of step k, so that
    # the whole run stays within a few minutes at ~0.1 s per view (views/s is a per-view rate either way)
    per_step = T_VIEWS if args.steps <= 120 else 1
    times = []
    for k in range(args.steps):
        t0 = time.perf_counter()
        for inp in (inps if per_step == T_VIEWS else [inps[k % T_VIEWS]]):
            oracle.forward(**inp)
        times.append(time.perf_counter() - t0)
    total = sum(times)
    val = per_step * args.steps / total
    cores = oracle.num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "views/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU restatement of the reference's (CUDA-only, un-vendored) "
                   "rasterizer: oracle/raster_oracle.c, OpenMP; the reference has no CPU implementation"},
        "cpu_baseline": {"value": val, "unit": "views/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps x {per_step} view(s) of the full workload"},
        "e2e": {"value": val, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ops", action="store_true",
                    help="also time the cost volume and PTF (BASELINE config 3 sizes) on the GPU and their CPU restatements "
                         "on bounded samples; adds an `ops` object to the JSON line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from freesplat_b200 import decoder, rasterizer

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # NUMA placement of the pinned staging buffers (host-to-host path): stay on the CPUs next to this rank's GPU
    from freesplat_b200.pipeline import bind_to_gpu_numa
    numa = bind_to_gpu_numa(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    sc_cpu = make_scene(rank)
    sc = sc_cpu.to(dev)
    V = T_VIEWS
    bg = torch.zeros((V, 3), device=dev)
    row, col = torch.triu_indices(3, 3)
    shs = sc.harmonics.transpose(1, 2).contiguous()
    cov6 = sc.covariances[:, row, col].contiguous()
    views, _ = decoder.camera_records(sc.extrinsics, sc.intrinsics, sc.near, sc.far, bg, True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step(stage_events=None, check="deferred"):
        return rasterizer.raster_forward_raw(sc.means, sc.opacities, views, H, W, shs=shs, cov3D_precomp=cov6,
                                             sh_degree=2, check_overflow=check, stage_events=stage_events)

    st = step(check="sync")          # sizes the workspace (R is only known on the device)
    R = st.num_rendered()
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()

    # ---- device-resident timing: per-step CUDA events, L2 flushed between steps -----------------
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # (a) the timed region proper: one fs_raster_forward per step, events around the whole step
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        st = step(stage_events=list(evs[k]))
    barrier()
    assert not st.overflowed()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = sum(step_ms)
    # (b) the same steps again with events BETWEEN the three stages of the ABI (roofline of the render kernel)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        st = step(stage_events=ev[k])
    barrier()
    render_ms = [e[2].elapsed_time(e[3]) for e in ev]
    pre_ms = [e[0].elapsed_time(e[1]) for e in ev]
    bin_ms = [e[1].elapsed_time(e[2]) for e in ev]

    # ---- end to end from pinned host memory through the public API ------------------------------
    pin = lambda t: t.contiguous().pin_memory()
    h = dict(ext=pin(sc_cpu.extrinsics), K=pin(sc_cpu.intrinsics), near=pin(sc_cpu.near), far=pin(sc_cpu.far),
             means=pin(sc_cpu.means), cov=pin(sc_cpu.covariances), sh=pin(sc_cpu.harmonics), op=pin(sc_cpu.opacities))
    h2d = sum(t.numel() * t.element_size() for t in h.values())
    d2h = (V * 3 * H * W + V * H * W) * 4

    # public API for host-resident data: freesplat_b200.pipeline.HostRenderPipeline (3 streams, double buffering).
    # Every step moves its inputs H2D and its results D2H; copies of neighbouring steps overlap the kernels.
    from freesplat_b200.pipeline import HostRenderPipeline
    host = dict(extrinsics=h["ext"], intrinsics=h["K"], near=h["near"], far=h["far"], means=h["means"], covariances=h["cov"],
                harmonics=h["sh"], opacities=h["op"])
    pipe = HostRenderPipeline(dev, (H, W), V, depth=2)
    for _ in range(4):
        pipe.submit(host)
    pipe.drain()
    # K steps, timed three times; the median is reported (host-to-host transfers share the PCIe switch / host memory
    # with other tenants of the box: single runs scatter by +-30 %)
    cur_s = torch.cuda.current_stream()
    e2e_runs = []
    last = 0
    for _rep in range(3):
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(cur_s)
        for s_ in (pipe.s_h2d, pipe.s_run, pipe.s_d2h):
            s_.wait_event(t0)
        for k in range(args.steps):
            last = pipe.submit(host)
        for s_ in (pipe.s_h2d, pipe.s_run, pipe.s_d2h):
            ev_ = torch.cuda.Event(); ev_.record(s_); cur_s.wait_event(ev_)
        t1.record(cur_s)
        barrier()
        e2e_runs.append(t0.elapsed_time(t1))
    clocks = sampler.stop()          # sampled across the timed regions (device-resident, per-stage, host-to-host)
    out_c, out_d = pipe.wait(last)
    assert torch.isfinite(out_c).all()

    # ---- max over ranks -------------------------------------------------------------------------
    t = torch.tensor([total_ms] + e2e_runs, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t[0])
    e2e_all = sorted(float(x) for x in t[1:])
    e2e_ms = e2e_all[1]
    if rank == 0:
        peak, peak_src = _peaks()
        HW = H * W
        # render fwd = 44 R + 24 HW per view; the kernel also depth-sorts its tile first (8 B key read,
        # 8 B key + 4 B index written per instance): + 20 R
        alg_bytes = 64.0 * R + 24.0 * HW * V
        rd = sum(render_ms) / len(render_ms)
        achieved = alg_bytes / (rd * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": world * V * args.steps / (total_ms * 1e-3), "unit": "views/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "views_per_step_per_gpu": V, "gaussians": P, "tile_instances_R": R,
                       "l2": "flushed between steps (256 MiB write)", "parallelism": f"view-sharded x{world}"},
            "e2e": {"value": world * V * args.steps / (e2e_ms * 1e-3), "unit": "views/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps,
                    "api": "freesplat_b200.pipeline.HostRenderPipeline (3 streams, depth 2)",
                    "protocol": "median of 3 timings of K steps", "ms_per_step_all": [x / args.steps for x in e2e_all],
                    "h2d_gbs": h2d / (e2e_ms / args.steps * 1e-3) / 1e9, "numa": numa},
            "gpu_launches": 4 * args.steps,      # preprocess, tile scan, scatter, sort+render
            "stage_ms": {"preprocess": sum(pre_ms) / len(pre_ms), "binning": sum(bin_ms) / len(bin_ms), "render": rd},
            "roofline": {"kernel": "render_fwd_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": _ncu_traffic("render_fwd_kernel"),
                         "traffic_source": "profiles/r1_fwd_step_ncu_raw.csv (ncu --set full, same workload)", "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "per-tile depth sort + alpha blend in one kernel; FP32/SFU-issue bound at this size (SURVEY §7), reported against HBM as BASELINE asks"},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline:
            reps = 25                  # ~10 s of host work at ~8 views/s
            v, cores, dt = cpu_reference_views_per_s(sc_cpu, n_views=T_VIEWS, repeats=reps)
            line["cpu_baseline"] = {"value": v, "unit": "views/s", "cores": cores, "kind": "port",
                                    "sample": f"{reps} x {T_VIEWS} views of the full workload ({dt:.1f} s)"}
        if args.ops:
            line["ops"] = ops_section(dev)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
