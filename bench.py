#!/usr/bin/env python
"""bench.py -- rendered views/s @ 640x480 of the B200 rasterizer hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the CPU restatement of the reference path

Workload (BASELINE config 2, "ScanNet 2-views 640x480, ~300k Gaussians, forward raster only"):
one scene of 307 200 pixel-aligned Gaussians (SH degree 2, precomputed covariances) rendered into
T=3 target views (assets/evaluation_index_scannet_2views.json holds 3 targets per scene) at 640x480.
A step = one pass of the hot path over that batch: preprocess -> tile binning -> render, all views.

value  : views/s with the scene resident in HBM: one CUDA-graph launch of the step per iteration
         (rasterizer.RasterPlan -> fs_graph_launch), per-step CUDA events, L2 flushed between steps by a
         256 MiB write outside the events.  `per_rank_ms` carries every rank's min / median / max step.
api    : the same steps through the public device-resident call (`decoder.render_views`, default
         deferred overflow check: no host sync), K calls bracketed by two events: host overhead included.
e2e    : the same metric host-to-host through `pipeline.HostRenderPipeline`: H2D of Gaussians + cameras
         from pinned memory, render, D2H of colour + depth, every step.  With N > 1 ranks the Gaussian set
         (the same scene for all ranks: SURVEY 8e, target views shard over GPUs) is uploaded ONCE in total
         -- each rank copies its 1/N slice over PCIe -- and replicated over NVLink by in-place all-gathers.
ops    : (N = 1) the other operators of the path at BASELINE config 3 / 4 sizes, each with its own roofline
         line and CPU baseline: cost volume fwd / bwd (tensor pipe), PTF folds (HBM), raster fwd+bwd.
config5: (N > 1) BASELINE config 5: 10 context views -> cross-view candidate exchange over NCCL -> PTF fold
         (replicated) -> 18 target views sharded over the ranks.
Multi-GPU: target views shard over ranks (each rank renders its own T views of the replicated
Gaussian set; no data-path collective in `value`) -> weak scaling; value = all views / max-over-ranks time.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, P, T_VIEWS = 480, 640, 307200, 3
WORKLOAD = f"scannet_2views_{W}x{H}_P{P}_targets{T_VIEWS}_raster_fwd"
METRIC = "rendered views/sec @640x480"


def base_config(world: int) -> dict:
    """The `config` object BOTH arms print (the driver compares them key by key)."""
    return {"workload": WORKLOAD, "views_per_step_per_gpu": T_VIEWS, "gaussians": P, "image": [H, W], "sh_degree": 2,
            "l2": "flushed between steps (256 MiB write)", "parallelism": f"view-sharded x{world}"}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return d, "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


def _ncu_traffic(kernel_substr: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of `kernel` from the newest committed ncu export (profiles/), bytes per launch."""
    import csv
    for name in ("r2_fwd_step_ncu_raw.csv", "r1_fwd_step_ncu_raw.csv"):
        path = os.path.join(ROOT, "profiles", name)
        try:
            rows = list(csv.reader(open(path)))
            hdr, units = rows[0], rows[1]
            ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            for r in rows[2:]:
                if kernel_substr in r[ik]:
                    return float(r[ir]) * mult[units[ir]] + float(r[iw]) * mult[units[iw]], f"profiles/{name}"
        except Exception:
            continue
    return None, None


class ClockSampler:
    """SM clock and throttle reasons sampled every 100 ms while the timed regions run: NVML in-process (the same counters
    `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` prints; a forked nvidia-smi per rank costs the host a core
    each, which at 8 ranks showed up inside the timed region), nvidia-smi as the fallback."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.stop_flag, self.thread, self.mode = index, [], None, False, None, None
        self.mx = None

    def _phys_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x.strip() for x in vis.split(",") if x.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self._phys_index())
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            names = {"hw_slowdown": pynvml.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": pynvml.nvmlClocksEventReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": pynvml.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": pynvml.nvmlClocksEventReasonSwPowerCap}

            def loop():
                while not self.stop_flag:
                    try:
                        sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        bits = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                        self.rows.append((sm, [n for n, b in names.items() if bits & b]))
                    except Exception:
                        pass
                    time.sleep(0.1)
            self.mode = "nvml"
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.mode = "nvidia-smi"
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            try:
                self.mx = float(r[1])
                self.rows.append((float(r[0]), [n for n, v in zip(names, r[2:6]) if v.lower().startswith("active")]))
            except Exception:
                continue

    def stop(self) -> dict:
        if self.mode is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock source available"]}
        time.sleep(0.15)
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        sm = sorted(r[0] for r in self.rows)
        reasons = sorted({n for r in self.rows for n in r[1]})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx, "reasons": reasons, "samples": len(sm),
                "source": self.mode}


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


def oracle_view_inputs(sc, i: int):
    """Oracle inputs of target view i prepared the way the reference adapter prepares the rasterizer call
    (cuda_splatting.py:64-127: scene rescale by 1/near, transposed matrices, [P,M,3] harmonics, upper-triangle covariances)."""
    import torch
    from freesplat_b200 import decoder
    V = sc.extrinsics.shape[0]
    views, _ = decoder.camera_records(sc.extrinsics, sc.intrinsics, sc.near, sc.far, torch.zeros((V, 3)), True)
    rec = views[i]
    s = rec[40]
    iu = torch.triu_indices(3, 3)
    d_sh = sc.harmonics.shape[-1]
    return dict(H=sc.image_shape[0], W=sc.image_shape[1], tanfovx=float(rec[38]), tanfovy=float(rec[39]), bg=rec[35:38].numpy(),
                viewmatrix=rec[0:16].numpy(), projmatrix=rec[16:32].numpy(), campos=rec[32:35].numpy(), means3D=(sc.means * s).numpy(),
                opacities=sc.opacities.numpy(), shs=sc.harmonics.transpose(1, 2).contiguous().numpy(),
                cov3D_precomp=(sc.covariances * (s * s))[:, iu[0], iu[1]].contiguous().numpy(), sh_degree=int(round(d_sh ** 0.5)) - 1)


def _pick_oracle_threads(oracle, inp):
    # torchrun exports OMP_NUM_THREADS=1; pick the best of {all allowed threads, half of them (SMT siblings)}
    best = None
    for n in sorted({host_threads(), max(1, host_threads() // 2)}, reverse=True):
        oracle.set_num_threads(n)
        oracle.forward(**inp)                      # warm (page in, OpenMP pool)
        t0 = time.perf_counter(); oracle.forward(**inp); dt = time.perf_counter() - t0
        if best is None or dt < best[1]:
            best = (n, dt)
    oracle.set_num_threads(best[0])
    return best[0]


def cpu_reference_views_per_s(sc, n_views: int, repeats: int):
    """Times the CPU restatement of the reference rasterizer (oracle/raster_oracle.c, all host cores)."""
    from oracle import raster as oracle
    inps = [oracle_view_inputs(sc, v) for v in range(n_views)]
    _pick_oracle_threads(oracle, inps[0])
    t0 = time.perf_counter()
    for _ in range(repeats):
        for inp in inps:
            oracle.forward(**inp)
    dt = time.perf_counter() - t0
    return n_views * repeats / dt, oracle.num_threads(), dt


def PlainGRU(state, dev):
    """The GRU of networks.py:188-214 (module / parameter names as in the reference, so that its state dict loads): the
    module fs_ptf_gru reads its weights from in inference, and the torch module the training fold calls."""
    import torch

    class GRU(torch.nn.Module):
        def __init__(self):
            super().__init__()
            mk = lambda din: torch.nn.Sequential(torch.nn.Linear(din, 64), torch.nn.ReLU(), torch.nn.Linear(64, 64))
            self.mlp_z, self.mlp_r, self.mlp_n = mk(176), mk(176), mk(152)

        def forward(self, input_feat, hidden_feat, input_weights_emb, hidden_weights_emb):
            x1 = torch.cat((input_feat, input_weights_emb), dim=-1)
            cat = torch.cat((hidden_feat, hidden_weights_emb, x1), dim=-1)
            r, z = torch.sigmoid(self.mlp_r(cat)), torch.sigmoid(self.mlp_z(cat))
            q = torch.tanh(self.mlp_n(torch.cat((r * hidden_feat, x1), dim=-1)))
            return (1 - z) * hidden_feat + z * q
    g = GRU()
    g.load_state_dict(state)
    for p_ in g.parameters():
        p_.requires_grad_(False)
    return g.to(dev)


def _flat_ptf(inp):
    """synth.ptf_inputs (encoder-shaped) -> flat per-view tensors (feats [V,HW,F], coords [V,HW,3], dens, wemb, depths [V,HW])."""
    V = inp["gaussians"][0].shape[1]
    return (inp["gaussians"][0][0], inp["coords"][0][0, :, :, 0, 0, :], inp["densities"][0, :, :, 0, 0], inp["weight_emb"][0, :, :, 0, 0],
            inp["depths"].reshape(V, -1), inp["extrinsics"][0], inp["intrinsics"][0], tuple(inp["depths"].shape[-2:]))


def gpu_ms(fn, n=5, warm=2):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def ops_section(dev, peaks):
    """The other operators of the path at BASELINE config 3 / 4 sizes: device time of the product kernels, their roofline line
    (SURVEY 8d byte / flop counts), and CPU baselines: the oracle port timed here on bounded samples (stated), and the
    reference's OWN PyTorch code timed on a full unit of work in the build container (profiles/r2_reference_cpu_timings.json:
    /root/reference does not exist on this box)."""
    import numpy as np
    import torch
    from freesplat_b200 import decoder, ptf, synth
    from freesplat_b200.cost_volume import AVGFeatureVolumeManager
    from oracle import cost_volume as ocv, ptf as optf
    out = {}
    torch.set_num_threads(host_threads())
    hbm, tf = float(peaks["hbm_gbs"]), float(peaks.get("bf16_tflops", 1590.0))
    try:
        ref_cpu = json.load(open(os.path.join(ROOT, "profiles", "r2_reference_cpu_timings.json")))
    except Exception:
        ref_cpu = {}

    # ---- cost volume: config 3 = 3 reference views, K = 2 ; config 4 = 10 views, K = 8 ; 48 x 120 x 160, D = 128 ----
    Hf, Wf, D, C = 120, 160, 128, 48
    for tag, V, K in (("cfg3", 3, 2), ("cfg4", 10, 8)):
        inp = synth.cost_volume_inputs(0, V, K, C, Hf, Wf)
        mlp = synth.cost_volume_mlp(0)
        m = AVGFeatureVolumeManager(Hf, Wf, num_depth_bins=D, matching_dim_size=C).to(dev)
        with torch.no_grad():
            for p_, w_ in zip([m.mlp.net[0].weight, m.mlp.net[0].bias, m.mlp.net[2].weight, m.mlp.net[2].bias, m.mlp.net[4].weight,
                               m.mlp.net[4].bias], mlp):
                p_.copy_(w_)
        ginp = {k: v.to(dev) for k, v in inp.items()}
        with torch.no_grad():
            f_ms = gpu_ms(lambda: m(**ginp))
        rows = V * D * Hf * Wf
        flops = rows * (2 * 2624 + K * (2 * 4 * C + 2 * C))
        tflops_tensor = rows * 2 * 2624 * 3 / (f_ms * 1e-3) / 1e12      # 3xTF32: three tensor-core products per useful one
        nbytes = V * ((1 + K) * C * Hf * Wf * 4 + D * Hf * Wf * 4)
        e = {"gpu_ms": f_ms, "views": V, "K": K, "useful_tflops": flops / (f_ms * 1e-3) / 1e12,
             "roofline": {"bound": "tensor", "achieved": tflops_tensor, "peak": tf / 2, "unit": "TFLOP/s",
                          "frac": tflops_tensor / (tf / 2), "note": "issued 3xTF32 MLP flops vs the tf32 dense peak (= half the measured bf16 "
                          "peak); the gather, not the tensor pipe, bounds this kernel (DESIGN 4)"},
             "hbm_gbs": nbytes / (f_ms * 1e-3) / 1e9}
        if tag == "cfg3":
            cur = ginp["cur_feats"].clone().requires_grad_(True); src = ginp["src_feats"].clone().requires_grad_(True)
            gw = torch.randn((V, D, Hf, Wf), device=dev)

            def fb():
                cur.grad = None; src.grad = None
                m(**{**ginp, "cur_feats": cur, "src_feats": src}).backward(gw)
            e["fwd_bwd_gpu_ms"] = gpu_ms(fb, n=3)
            Dsub = 32                                                  # CPU sample: 1 reference view, 32 of the 128 planes
            sub = {k: v[:1] for k, v in inp.items() if k not in ("min_depth", "max_depth")}
            t0 = time.perf_counter()
            ocv.forward(sub["cur_feats"], sub["src_feats"], sub["src_extrinsics"], sub["src_Ks"], sub["cur_invK"], inp["min_depth"],
                        inp["max_depth"], mlp, Dsub, plane_chunk=8)
            c_s = time.perf_counter() - t0
            e["cpu_baseline"] = {"kind": "port", "ms_per_view": c_s * 1e3 * (D / Dsub), "cores": host_threads(),
                                 "sample": f"1 reference view, {Dsub} of {D} planes ({c_s:.1f} s), scaled by {D // Dsub}"}
            if "cost_volume_fwd_one_view_K2_D128_120x160_s" in ref_cpu:
                e["reference_cpu"] = {"kind": "reference (its own PyTorch code, pre-timed in the build container)",
                                      "ms_per_view": ref_cpu["cost_volume_fwd_one_view_K2_D128_120x160_s"] * 1e3,
                                      "fwd_bwd_ms_per_view": ref_cpu.get("cost_volume_fwd_bwd_one_view_s", 0) * 1e3,
                                      "cores": ref_cpu.get("host_threads"), "sample": "one full reference view, all 128 planes, unscaled"}
        out[f"cost_volume_{tag}_fwd"] = e
        del m, ginp

    # ---- PTF: 3 views (config 3) and 10 views (config 4) of 640 x 480 ----
    state = synth.gru_state(0)
    gru = PlainGRU(state, dev)
    for tag, V in (("cfg3_3views", 3), ("cfg4_10views", 10)):
        feats, coords, dens, wemb, depths, ext, Kn, hw = _flat_ptf(synth.ptf_inputs(0, V, H, W))
        gargs = [x.to(dev).contiguous() for x in (feats, coords, dens, wemb, depths, ext, Kn)]
        with torch.no_grad():
            res = ptf.fuse_views(gru, *gargs, hw)
            g_ms = gpu_ms(lambda: ptf.fuse_views(gru, *gargs, hw), n=3)
        N_out = int(res[0].shape[0])
        HW = H * W
        # SURVEY 8d per merge step: 16 N + 8 HW + 344 N_out + 280 HW; N grows from HW to N_out: bounded below by the inputs read
        # once and the final state written once
        nbytes = V * HW * 280 + N_out * 344
        e = {"gpu_ms": g_ms, "views": V, "N_out": N_out,
             "roofline": {"bound": "hbm", "achieved": nbytes / (g_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                          "frac": nbytes / (g_ms * 1e-3) / 1e9 / hbm,
                          "note": "algorithmic bytes = V*HW*280 (candidates read once) + 344*N_out (state written once); the GRU of the "
                                  "matched pairs (tensor cores) is inside the time"}}
        if V == 3:
            np_args = [x.numpy() for x in (feats, coords, dens, wemb, depths, ext, Kn)]
            t0 = time.perf_counter()
            optf.fuse(*np_args, hw, optf.torch_gru_fn(state))
            c_s = time.perf_counter() - t0
            e["cpu_baseline"] = {"kind": "port", "ms": c_s * 1e3, "cores": host_threads(), "sample": f"the full 3-view fold ({c_s:.1f} s)"}
            if "ptf_3views_640x480_s" in ref_cpu:
                e["reference_cpu"] = {"kind": "reference (its own PyTorch code, pre-timed in the build container)",
                                      "ms": ref_cpu["ptf_3views_640x480_s"] * 1e3, "cores": ref_cpu.get("host_threads"),
                                      "sample": "the full 3-view fold, unscaled"}
        out[f"ptf_{tag}"] = e
        del gargs, res

    # ---- BASELINE config 3 as ONE chained training step: cost volume -> depth-head tail -> back-projection -> PTF fold ->
    #      Gaussian head -> raster forward, MSE loss, backward through every operator of the path.  The conv stacks between the
    #      operators (CVEncoder / DepthDecoder / Gaussian MLP: cuDNN, out of scope) are replaced by fixed cheap stand-ins. ----
    try:
        out["config3_train_step"] = config3_step(dev, gru)
    except Exception as exc:
        out["config3_train_step"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    # ---- raster forward + backward, config 3: P = 460 800, 4 target views ----
    sc = synth.pixel_aligned_scene(seed=3, h=H, w=W, n_context=3, n_target=4, keep=460800).to(dev)
    bg = torch.zeros((4, 3), device=dev)
    leaves = [t.clone().requires_grad_(True) for t in (sc.means, sc.covariances, sc.harmonics, sc.opacities)]
    gC = torch.randn((4, 3, H, W), device=dev)

    def train_step():
        for t in leaves:
            t.grad = None
        c, _ = decoder.render_views(sc.extrinsics, sc.intrinsics, sc.near, sc.far, (H, W), bg, *leaves)
        c.backward(gC)
    with torch.no_grad():
        f_ms = gpu_ms(lambda: decoder.render_views(sc.extrinsics, sc.intrinsics, sc.near, sc.far, (H, W), bg, sc.means, sc.covariances,
                                                   sc.harmonics, sc.opacities), n=5)
    out["raster_cfg3_P460800_4views"] = {"fwd_gpu_ms": f_ms, "fwd_bwd_gpu_ms": gpu_ms(train_step, n=5)}
    return out


def config3_step(dev, gru):
    import torch
    import torch.nn.functional as Fn
    from freesplat_b200 import adapter, decoder, depth_head, ptf, synth
    from freesplat_b200.cost_volume import AVGFeatureVolumeManager
    V, K, T, Hf, Wf, D, C = 3, 2, 4, 120, 160, 128, 48
    HW = H * W
    inp = {k: v.to(dev) for k, v in synth.cost_volume_inputs(0, V, K, C, Hf, Wf).items()}
    m = AVGFeatureVolumeManager(Hf, Wf, num_depth_bins=D, matching_dim_size=C).to(dev)
    with torch.no_grad():
        for p_, w_ in zip([m.mlp.net[0].weight, m.mlp.net[0].bias, m.mlp.net[2].weight, m.mlp.net[2].bias, m.mlp.net[4].weight,
                           m.mlp.net[4].bias], synth.cost_volume_mlp(0)):
            p_.copy_(w_)
    g = torch.Generator().manual_seed(0)
    cur = inp["cur_feats"].clone().requires_grad_(True); src = inp["src_feats"].clone().requires_grad_(True)
    latents = (0.5 * torch.randn((V, HW, 64), generator=g)).to(dev).requires_grad_(True)
    candi = torch.linspace(float(torch.log(torch.tensor(synth.NEAR))), float(torch.log(torch.tensor(4.0))), D, device=dev)
    ctx_ext = synth.camera_path(V, spacing=0.25).to(dev)
    Kn = synth.intrinsics(V).to(dev)
    tgt_ext = synth.camera_path(T, spacing=0.08, t0=0.2).to(dev)
    tgt_K = synth.intrinsics(T).to(dev)
    near = torch.full((T,), synth.NEAR, device=dev); far = torch.full((T,), synth.FAR, device=dev)
    bg = torch.zeros((T, 3), device=dev)
    target = torch.rand((T, 3, H, W), generator=g).to(dev)
    raw_bias = torch.zeros(34, device=dev); raw_bias[:3] = -3.0
    for p_ in gru.parameters():
        p_.requires_grad_(True)
    info = {}

    def forward():
        vol = m(**{**inp, "cur_feats": cur, "src_feats": src})                          # [3,128,120,160]   fs_cost_volume_forward
        logits = Fn.interpolate(vol * 8.0, scale_factor=2, mode="nearest")                # stand-in: CVEncoder + DepthDecoder convs
        o = depth_head.depth_regression(logits, candi, True, upsample=True)               # fs_depth_head (TMA)
        depth = o["depth_up"].reshape(V, HW); dens = o["weights_up"].reshape(V, HW)
        coords = adapter.backproject_depth(depth, Kn[0], ctx_ext, (H, W))                 # fs_backproject
        F_, X_, E_, Z_ = ptf.fuse_views(gru, latents, coords, dens, dens, depth, ctx_ext, Kn, (H, W))   # fs_ptf_*
        gs = adapter.gaussian_head(F_[:, :34] + raw_bias, Z_, torch.sigmoid(F_[:, 34]), X_, E_, Kn[0], (H, W))   # fs_gaussian_head
        color, _ = decoder.render_views(tgt_ext, tgt_K, near, far, (H, W), bg, gs.means, gs.covariances, gs.harmonics, gs.opacities)
        info["fused_gaussians"] = int(F_.shape[0])
        return ((color - target) ** 2).mean()

    def train():
        for t_ in (cur, src, latents, *gru.parameters(), *m.parameters()):
            t_.grad = None
        forward().backward()

    def infer():
        with torch.no_grad():
            forward()
    res = {"fwd_bwd_gpu_ms": gpu_ms(train, n=3, warm=2), "fwd_only_gpu_ms": gpu_ms(infer, n=3, warm=1)}
    for p_ in gru.parameters():
        p_.requires_grad_(False)
    res.update(info)
    res["note"] = ("3 context views (cost volume K=2, D=128), PTF fold, Gaussian head, 4 target views, MSE; every operator of the path "
                   "forward and backward in libfreesplat_b200.so (the GRU backward on the tcgen05 data / weight-gradient kernels); "
                   "the fold reads its step counters on the host once per view in training")
    return res


def config5_section(dev, rank, world, steps=3):
    """BASELINE config 5: 10 context views (owned round-robin by the ranks) -> cross-view exchange of the per-view candidates
    (70 floats per pixel) straight into the [V,HW,...] buffers the PTF kernels read (parallel.ViewExchange: one broadcast per
    view and tensor on a side stream; fold step i waits only for view i) -> PTF fold, replicated -> 18 target views sharded over
    the ranks.  Device-timed per stage, max over ranks."""
    import torch
    import torch.distributed as dist
    from freesplat_b200 import adapter, decoder, parallel, ptf, synth
    V, T = 10, 18
    feats, coords, dens, wemb, depths, ext, Kn, hw = _flat_ptf(synth.ptf_inputs(0, V, H, W))
    mine = parallel.shard_views(V, rank, world)
    local = [x[mine].to(dev).contiguous() for x in (feats, coords, dens, wemb, depths)]
    ext_d, K_d = ext.to(dev), Kn.to(dev)
    gru = PlainGRU(synth.gru_state(0), dev)
    tgt_ext = synth.camera_path(T, spacing=0.05, t0=0.1).to(dev)
    tgt_K = synth.intrinsics(T).to(dev)
    near = torch.full((T,), synth.NEAR, device=dev); far = torch.full((T,), synth.FAR, device=dev)
    my_t = parallel.shard_views(T, rank, world)
    sel = torch.tensor(my_t, dtype=torch.long, device=dev)
    bg = torch.zeros((len(my_t), 3), device=dev)
    ex = parallel.ViewExchange(V, H * W, 64, dev)
    raw_gain = torch.ones(34, device=dev); raw_bias = torch.zeros(34, device=dev); raw_bias[:3] = -3.0
    rows = []
    N = 0
    for it in range(steps + 1):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        with torch.no_grad():
            ev[0].record()
            full = ex.exchange(*local)                                 # enqueues the broadcasts on the side stream
            ev[1].record()
            F_, X_, E_, Z_ = ptf.fuse_views(gru, *full, ext_d, K_d, hw, view_ready=ex.ready_events)
            ev[2].record()
            N = int(F_.shape[0])
            # Gaussian head (fs_gaussian_head) on the fused state; the first 34 latent channels stand in for the output of the
            # (out-of-scope) decoder MLP, shifted so that the scales land where trained FreeSplat's do (sigma ~ 0.5-3 px)
            raw = F_[:, :34] * raw_gain + raw_bias
            gs = adapter.gaussian_head(raw, Z_, torch.full((N,), 0.5, device=dev), X_, E_, K_d[0], (H, W))
            if my_t:
                decoder.render_views(tgt_ext[sel], tgt_K[sel], near[sel], far[sel], (H, W), bg, gs.means, gs.covariances,
                                     gs.harmonics, gs.opacities)
            ev[3].record()
        torch.cuda.synchronize()
        if it:
            rows.append([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]), ev[0].elapsed_time(ev[3])])
    t = torch.tensor(rows, dtype=torch.float64, device=dev).median(0).values
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gather_bytes = (V - len(mine)) * H * W * 70 * 4
    return {"workload": f"replica_10views_{W}x{H}_ptf_then_{T}_targets_sharded_x{world}", "fused_gaussians": N,
            "exchange_enqueue_ms": float(t[0]), "ptf_fold_ms": float(t[1]), "render_share_ms": float(t[2]), "total_ms": float(t[3]),
            "exchange_bytes_received_per_rank": gather_bytes, "targets_per_rank": len(my_t),
            "note": "exchange overlaps the fold (step i waits for view i only): its exposed time is inside ptf_fold_ms; "
                    "max over ranks, median of %d steps" % steps}


def cost_volume_sharded_section(dev, rank, world, steps=3):
    """SURVEY 8e row 2 at BASELINE config-4/5 size: 10 context views (owned round-robin), ONE all-gather of the stride-4 matching
    features (3.7 MB per view) and every rank builds the cost volumes (K = 8 nearest sources, D = 128) of ITS reference views
    (parallel.cost_volume_sharded).  Device-timed, max over ranks; `single_rank_ms` is the same 10 volumes built by one GPU."""
    import torch
    import torch.distributed as dist
    from freesplat_b200 import parallel, synth
    from freesplat_b200.cost_volume import AVGFeatureVolumeManager
    V, K, Hf, Wf, D, C = 10, 8, 120, 160, 128, 48
    inp = synth.cost_volume_inputs(0, V, K, C, Hf, Wf)
    feats = inp["cur_feats"]                                        # [V,C,Hf,Wf]
    src_idx = torch.tensor([sorted([j for j in range(V) if j != b], key=lambda j: abs(j - b))[:K] for b in range(V)])
    ext = synth.camera_path(V, spacing=0.25).to(dev)
    Kf = synth.intrinsics(V).clone(); Kf[:, 0] *= Wf; Kf[:, 1] *= Hf
    Kf = Kf.to(dev)
    m = AVGFeatureVolumeManager(Hf, Wf, num_depth_bins=D, matching_dim_size=C).to(dev)
    mine = parallel.shard_views(V, rank, world)
    local = feats[mine].to(dev).contiguous()
    near, far = inp["min_depth"].to(dev), inp["max_depth"].to(dev)
    rows = []
    for it in range(steps + 1):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        dist.barrier()
        torch.cuda.synchronize()
        with torch.no_grad():
            ev[0].record()
            vol, ids = parallel.cost_volume_sharded(m, local, ext, Kf, near, far, src_indices=src_idx)
            ev[1].record()
        torch.cuda.synchronize()
        if it:
            rows.append(ev[0].elapsed_time(ev[1]))
    t = torch.tensor([statistics.median(rows)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    single = None
    if rank == 0:                                                  # the same 10 volumes on one GPU (no exchange)
        allf = feats.to(dev)
        with torch.no_grad():
            single = gpu_ms(lambda: m(cur_feats=allf, src_feats=allf[src_idx.to(dev)], src_extrinsics=inp["src_extrinsics"].to(dev),
                                      src_poses=inp["src_poses"].to(dev), src_Ks=inp["src_Ks"].to(dev), cur_invK=inp["cur_invK"].to(dev),
                                      min_depth=near, max_depth=far), n=3, warm=1)
    return {"workload": f"fvt_10views_K8_D128_{Wf}x{Hf}_context_views_sharded_x{world}", "sharded_ms": float(t[0]),
            "single_rank_ms": single, "views_per_rank": len(mine), "all_gather_bytes_per_rank": (V - len(mine)) * C * Hf * Wf * 4,
            "note": "feature all-gather (NCCL) + cost volumes of the rank's own reference views; max over ranks, median of %d" % steps}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sc = make_scene(0)
    from oracle import raster as oracle
    inps = [oracle_view_inputs(sc, v) for v in range(T_VIEWS)]
    _pick_oracle_threads(oracle, inps[0])
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
        "config": base_config(args.gpus),
        "reference_note": "CPU restatement of the reference's (CUDA-only, un-vendored) rasterizer: oracle/raster_oracle.c, OpenMP; "
                          "the reference has no CPU implementation of this path",
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
    ap.add_argument("--no-ops", action="store_true", help="skip the `ops` (N = 1) / `config5` (N > 1) sections")
    ap.add_argument("--no-graph", action="store_true", help="time the eager launch sequence instead of the CUDA graph")
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
    from freesplat_b200.pipeline import HostRenderPipeline, bind_to_gpu_numa
    numa = bind_to_gpu_numa(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    sc_cpu = make_scene(rank)
    sc = sc_cpu.to(dev)
    V = T_VIEWS
    bg = torch.zeros((V, 3), device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    cov9 = sc.covariances.reshape(-1, 9)
    # the step on a static workspace, recorded once as a CUDA graph (camera records + memset + preprocess + tile scan +
    # scatter + sort/render): one launch per step
    plan = rasterizer.RasterPlan(sc.means, sc.opacities, H, W, shs=sc.harmonics, cov3D_precomp=cov9,
                                 cameras=(sc.extrinsics, sc.intrinsics, sc.near, sc.far, bg), sh_degree=2, sh_layout=1, cov_stride=9,
                                 graph=not args.no_graph)
    R = plan.run_checked()
    for _ in range(args.warmup):
        plan.run()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- (a) device-resident: per-step CUDA events around the graph launch, L2 flushed between steps ----
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        evs[k][0].record()
        plan.run()
        evs[k][1].record()
    barrier()
    assert not plan.check()[1]
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = sum(step_ms)

    # ---- (b) the same work, eager, with events BETWEEN the three stages of the ABI (roofline of the render kernel) ----
    views = decoder.camera_records_fused(sc.extrinsics, sc.intrinsics, sc.near, sc.far, bg, True)
    n_stage = min(args.steps, 100)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(n_stage)]
    st = None
    for k in range(n_stage):
        flush.fill_(k & 0xFF)
        st = rasterizer.raster_forward_raw(sc.means, sc.opacities, views, H, W, shs=sc.harmonics, cov3D_precomp=cov9, sh_degree=2,
                                           sh_layout=1, cov_stride=9, check_overflow="deferred", stage_events=ev[k], reuse_scratch=True)
    barrier()
    render_ms = [e[2].elapsed_time(e[3]) for e in ev]
    pre_ms = [e[0].elapsed_time(e[1]) for e in ev]
    bin_ms = [e[1].elapsed_time(e[2]) for e in ev]

    # ---- (c) the public device-resident call, default (deferred) overflow check: K calls between two events ----
    with torch.no_grad():
        for _ in range(3):
            decoder.render_views(sc.extrinsics, sc.intrinsics, sc.near, sc.far, (H, W), bg, sc.means, sc.covariances, sc.harmonics,
                                 sc.opacities)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_host0 = time.perf_counter()
        a0.record()
        for k in range(args.steps):
            decoder.render_views(sc.extrinsics, sc.intrinsics, sc.near, sc.far, (H, W), bg, sc.means, sc.covariances, sc.harmonics,
                                 sc.opacities)
        a1.record()
        t_host = time.perf_counter() - t_host0
        barrier()
        rasterizer.poll_deferred(block=True)
    api_ms = a0.elapsed_time(a1)

    # ---- (d) end to end from pinned host memory through the public host-to-host API ----
    pin = lambda t: t.contiguous().pin_memory()
    host = dict(extrinsics=pin(sc_cpu.extrinsics), intrinsics=pin(sc_cpu.intrinsics), near=pin(sc_cpu.near), far=pin(sc_cpu.far),
                means=pin(sc_cpu.means), covariances=pin(sc_cpu.covariances), harmonics=pin(sc_cpu.harmonics),
                opacities=pin(sc_cpu.opacities))
    h2d_full = sum(t.numel() * t.element_size() for t in host.values())
    d2h = (V * 3 * H * W + V * H * W) * 4
    pipe = HostRenderPipeline(dev, (H, W), V, depth=2, shard_group=None if world > 1 else "none")
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
    clocks = sampler.stop()          # sampled across the timed regions (device-resident, per-stage, API, host-to-host)
    out_c, out_d = pipe.wait(last)
    assert torch.isfinite(out_c).all() and pipe.reruns == 0
    h2d_rank = pipe.h2d_bytes

    # ---- max over ranks; every rank's own step statistics ----
    t = torch.tensor([total_ms, api_ms] + e2e_runs, dtype=torch.float64, device=dev)
    mine = torch.tensor([min(step_ms), statistics.median(step_ms), max(step_ms), sum(step_ms) / len(step_ms)], dtype=torch.float64, device=dev)
    per_rank = [mine.clone() for _ in range(world)]
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_gather(per_rank, mine)
    total_ms, api_ms = float(t[0]), float(t[1])
    e2e_all = sorted(float(x) for x in t[2:])
    e2e_ms = e2e_all[1]
    cfg5 = None
    cvs = None
    if world > 1 and not args.no_ops:
        try:
            cfg5 = config5_section(dev, rank, world)
        except Exception as exc:             # keep the headline line even if the side section fails
            cfg5 = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        try:
            cvs = cost_volume_sharded_section(dev, rank, world)
        except Exception as exc:
            cvs = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    if rank == 0:
        peaks, peak_src = _peaks()
        peak = float(peaks["hbm_gbs"])
        HW = H * W
        # render fwd = 44 R + 24 HW per view; the kernel also depth-sorts its tile first (8 B key read,
        # 8 B key + 4 B index written per instance): + 20 R
        alg_bytes = 64.0 * R + 24.0 * HW * V
        rd = sum(render_ms) / len(render_ms)
        achieved = alg_bytes / (rd * 1e-3) / 1e9
        traffic, traffic_src = _ncu_traffic("render_fwd_kernel")
        launches = plan.launches_per_run()
        line = {
            "metric": METRIC, "value": world * V * args.steps / (total_ms * 1e-3), "unit": "views/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": base_config(world),
            "tile_instances_R": R,
            "launch": "one CUDA graph launch per step (fs_graph_launch)" if not args.no_graph else "eager",
            "per_rank_ms": [{"rank": r, "min": float(x[0]), "median": float(x[1]), "max": float(x[2]), "mean": float(x[3])}
                            for r, x in enumerate(per_rank)],
            "api": {"value": world * V * args.steps / (api_ms * 1e-3), "unit": "views/s", "ms_per_step": api_ms / args.steps,
                    "call": "freesplat_b200.decoder.render_views (device-resident inputs, deferred overflow check, no host sync)",
                    "host_ms_per_call_rank0": 1e3 * t_host / args.steps},
            "e2e": {"value": world * V * args.steps / (e2e_ms * 1e-3), "unit": "views/s", "h2d_bytes_per_step": h2d_rank,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps,
                    "api": "freesplat_b200.pipeline.HostRenderPipeline (3 streams, depth 2, one CUDA graph per slot)",
                    "upload": ("each rank uploads 1/%d of the Gaussian set (%d of %d bytes) + its cameras; in-place NCCL all-gathers "
                               "replicate it over NVLink" % (world, h2d_rank, h2d_full)) if world > 1 else "full scene per step",
                    "protocol": "median of 3 timings of K steps", "ms_per_step_all": [x / args.steps for x in e2e_all],
                    "h2d_gbs_per_rank": h2d_rank / (e2e_ms / args.steps * 1e-3) / 1e9,
                    "d2h_gbs_per_rank": d2h / (e2e_ms / args.steps * 1e-3) / 1e9, "numa": numa},
            "gpu_launches": launches * args.steps,      # camera records, preprocess, tile scan (+ fallback scatter), sort+render (graph nodes)
            "stage_ms": {"preprocess": sum(pre_ms) / len(pre_ms), "binning": sum(bin_ms) / len(bin_ms), "render": rd},
            "roofline": {"kernel": "render_fwd_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "strict_8d_frac": (44.0 * R + 24.0 * HW * V) / (rd * 1e-3) / 1e9 / peak,
                         "note": "per-tile depth sort + alpha blend in one kernel (64 R + 24 HW V bytes; SURVEY 8d's render-only 44 R + "
                                 "24 HW V in strict_8d_frac); FP32/SFU-issue bound at this size (SURVEY §7), reported against HBM as "
                                 "BASELINE asks"},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            reps = 25                  # ~10 s of host work at ~8 views/s
            v, cores, dt = cpu_reference_views_per_s(sc_cpu, n_views=T_VIEWS, repeats=reps)
            line["cpu_baseline"] = {"value": v, "unit": "views/s", "cores": cores, "kind": "port",
                                    "sample": f"{reps} x {T_VIEWS} views of the full workload ({dt:.1f} s)"}
        if world == 1 and not args.no_ops:
            try:
                line["ops"] = ops_section(dev, peaks)
            except Exception as exc:
                line["ops"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        if cfg5 is not None:
            line["configs"] = {"config5": cfg5, "cost_volume_sharded": cvs}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
