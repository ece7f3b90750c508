"""Diagnostic run on the GPU box: dumps every parity metric as JSON under gpurun_out/ so that a
failing case can be read offline (pytest only shows the first assertion)."""
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from freesplat_b200 import synth  # noqa: E402
from tests import raster_compare as rc  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
res = {"device": torch.cuda.get_device_name(0), "nproc": os.cpu_count()}


def section(name, fn):
    t0 = time.time()
    try:
        res[name] = fn()
    except Exception:
        res[name] = {"error": traceback.format_exc()}
    res[name + "_s"] = round(time.time() - t0, 2)
    json.dump(res, open(os.path.join(OUT, "gpu_check.json"), "w"), indent=1, default=str)


def fwd(scene, bg=(0.0, 0.0, 0.0)):
    st, views = rc.run_cuda(scene, bg=bg)
    m = rc.compare_forward(scene, st, bg=bg)
    m["fails"] = rc.forward_ok(m, n_pixels=st.H * st.W)
    return m


def bwd(scene, depth=False):
    st, views = rc.run_cuda(scene)
    g = torch.Generator().manual_seed(1)
    dC = torch.randn((st.V, 3, st.H, st.W), generator=g)
    dD = torch.randn((st.V, st.H, st.W), generator=g) * 0.2 if depth else None
    m = rc.compare_backward(scene, st, views, dC, dD)
    m["fails"] = rc.backward_ok(m)
    return m


section("fwd_random_256", lambda: fwd(synth.random_scene(seed=0, h=256, w=256, P=10000), (0.2, 0.4, 0.6)))
section("fwd_pixel_small_3v", lambda: fwd(synth.pixel_aligned_scene(seed=1, h=120, w=160, n_context=2, n_target=3, keep=None)))
section("fwd_crowded", lambda: fwd(synth.random_scene(seed=4, h=64, w=64, P=30000, sigma_px=(0.5, 2.0))))
section("bwd_pixel_small", lambda: bwd(synth.pixel_aligned_scene(seed=0, h=96, w=128, n_context=2, n_target=2, keep=None)))
section("bwd_pixel_small_depth", lambda: bwd(synth.pixel_aligned_scene(seed=1, h=96, w=128, n_context=2, n_target=2, keep=None), True))
section("bwd_random", lambda: bwd(synth.random_scene(seed=2, h=128, w=128, P=4000)))
section("fwd_full_config2", lambda: fwd(synth.pixel_aligned_scene(seed=0, h=480, w=640, n_context=2, n_target=3, keep=307200)))
print(json.dumps({k: (v.get("fails", v) if isinstance(v, dict) else v) for k, v in res.items()}, indent=1, default=str)[:6000])
