"""Times the cost volume (fwd, fwd+bwd) and PTF at BASELINE config-3/4 sizes on cuda:0; JSON to gpurun_out/."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from freesplat_b200 import ptf, synth  # noqa: E402
from freesplat_b200.cost_volume import AVGFeatureVolumeManager  # noqa: E402

dev = "cuda:0"
res = {}


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for V, K in ((3, 2), (10, 8)):
    Hf, Wf, D = 120, 160, 128
    inp = {k: v.to(dev) for k, v in synth.cost_volume_inputs(0, V, K, 48, Hf, Wf).items()}
    m = AVGFeatureVolumeManager(Hf, Wf, num_depth_bins=D, matching_dim_size=48).to(dev)
    with torch.no_grad():
        f = timeit(lambda: m(**inp))
    cur = inp["cur_feats"].clone().requires_grad_(True); src = inp["src_feats"].clone().requires_grad_(True)
    inp2 = dict(inp, cur_feats=cur, src_feats=src)

    def fb():
        out = m(**inp2)
        out.sum().backward()
    fbt = timeit(fb, n=3, warm=1)
    rows = V * D * Hf * Wf
    res[f"cost_volume_V{V}_K{K}"] = {"fwd_ms": f, "fwd_bwd_ms": fbt, "fwd_tflops": rows * (2 * 2624 + K * 480) / (f * 1e-3) / 1e12,
                                     "alg_bytes_fwd": ((1 + K) * 48 * Hf * Wf * 4 + D * Hf * Wf * 4) * V}

from test_ptf_gpu import GRU  # noqa: E402
for V in (3, 10):
    h, w = 480, 640
    inp = synth.ptf_inputs(0, V, h, w)
    from ptf_helpers import flat_inputs  # noqa: E402
    feats, coords, dens, wemb, depths, ext, K, hw = flat_inputs(inp)
    t = lambda a: torch.from_numpy(a).to(dev)
    g = GRU(); g.load_state_dict(synth.gru_state(0)); g = g.to(dev)
    args = [t(x) for x in (feats, coords, dens, wemb, depths, ext, K)]
    with torch.no_grad():
        out = ptf.fuse_views(g, *args, hw)
        ms = timeit(lambda: ptf.fuse_views(g, *args, hw), n=3, warm=1)
    N = out[0].shape[0]
    tm = []
    with torch.no_grad():
        ptf.fuse_views(g, *args, hw, timings=tm)
    res[f"ptf_V{V}_640x480"] = {"ms": ms, "N_out": N, "ratio": N / (V * h * w),
                                "alg_bytes": 280 * h * w * V + 344 * N,
                                "per_step_ms": {k: round(sum(t[k] for t in tm) / len(tm), 3) for k in ("match_ms", "gru_ms", "merge_ms")},
                                "matched_mean": sum(t["matched"] for t in tm) / len(tm)}
# ---- raster fwd + bwd (BASELINE config 3: 4 target views, MSE loss on colour) ----
from freesplat_b200 import decoder  # noqa: E402
sc = synth.pixel_aligned_scene(seed=0, h=480, w=640, n_context=3, n_target=4, keep=460800).to(dev)
bg = torch.zeros((4, 3), device=dev)
means = sc.means.clone().requires_grad_(True); cov = sc.covariances.clone().requires_grad_(True)
sh = sc.harmonics.clone().requires_grad_(True); op = sc.opacities.clone().requires_grad_(True)
target = torch.rand((4, 3, 480, 640), device=dev)


def raster_fb():
    c, d = decoder.render_views(sc.extrinsics, sc.intrinsics, sc.near, sc.far, (480, 640), bg, means, cov, sh, op)
    ((c - target) ** 2).mean().backward()


def raster_f():
    with torch.no_grad():
        decoder.render_views(sc.extrinsics, sc.intrinsics, sc.near, sc.far, (480, 640), bg, means, cov, sh, op)


res["raster_cfg3_P460800_T4"] = {"fwd_ms": timeit(raster_f, n=10, warm=3), "fwd_bwd_ms": timeit(raster_fb, n=10, warm=3)}
# ---- depth-head tail: 3 views at 640x480 (scale 0 = 240x320, D = 128) ----
from freesplat_b200.depth_head import depth_regression  # noqa: E402
import torch.nn.functional as F  # noqa: E402
Vd, Dd, hd, wd = 3, 128, 240, 320
lg = (torch.randn((Vd, Dd, hd, wd), device=dev) * 5)
cand = (torch.log(torch.tensor(0.5)) + torch.linspace(0, 1, Dd) * torch.log(torch.tensor(30.0))).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit_cold(fn, n=10, warm=3):     # L2 flushed before every timed call (the logits of 3 views fit in the 126 MB L2)
    for _ in range(warm):
        fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n


def torch_tail():      # the reference's op sequence (networks.py:130-152) on the GPU
    planes = F.softmax(lg, dim=1)
    coarse = (cand.view(1, -1, 1, 1) * planes).sum(dim=1, keepdim=True)
    fine = F.interpolate(coarse, scale_factor=2, mode="bilinear", align_corners=True)
    return torch.exp(coarse), torch.exp(fine), F.interpolate(planes, scale_factor=2, mode="bilinear", align_corners=True).max(dim=1, keepdim=True)[0]


with torch.no_grad():
    t_tma = timeit_cold(lambda: depth_regression(lg, cand, True, upsample=True, tile_mode=0))
    t_ldg = timeit_cold(lambda: depth_regression(lg, cand, True, upsample=True, tile_mode=1))
    t_exp = timeit_cold(lambda: depth_regression(lg, cand, True, upsample=False))
    t_ref = timeit_cold(torch_tail, n=5, warm=2)
alg = Vd * (Dd * hd * wd * 4 + 10 * hd * wd * 4)
res["depth_head_V3_640x480"] = {"fused_tma_ms": t_tma, "fused_ldg_ms": t_ldg, "expect_only_ms": t_exp, "torch_ops_ms": t_ref,
                                "alg_bytes": alg, "gbs_tma": alg / (t_tma * 1e-3) / 1e9, "gbs_expect_only": Vd * (Dd + 2) * hd * wd * 4 / (t_exp * 1e-3) / 1e9}
print(json.dumps(res, indent=1))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bench_ops.json"), "w"), indent=1)
