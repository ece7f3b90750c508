"""Direct binning (FsRasterFwdArgs.bins) vs the count + scatter path: forward time, heaviest tile and fallback flag for the
BASELINE config-2 (P = 307 200, 3 views) and config-3 (P = 460 800, 4 views) scenes.   python tools/bench_bins.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from freesplat_b200 import decoder, rasterizer, synth  # noqa: E402
from tests import raster_compare as rc  # noqa: E402

dev = "cuda:0"
res = {}
for name, kw in (("config2", dict(n_context=2, n_target=3, keep=307200)), ("config3", dict(n_context=3, n_target=4, keep=460800))):
    scene = synth.pixel_aligned_scene(seed=0, h=480, w=640, **kw)
    sc = scene.to(dev)
    V = scene.extrinsics.shape[0]
    bg = torch.zeros((V, 3), device=dev)
    row = {}
    for cap in (0, 2048, 4096, 0, 2048, 4096):
        rasterizer.BIN_CAP = cap
        rasterizer._scratch_cache.clear()
        st, _ = rc.run_cuda(scene)
        nt = st.ranges.shape[0]
        heaviest = int((st.ranges[:, 1] - st.ranges[:, 0]).max())
        flag = int(st.tile_buf[2 * nt]) if cap else 0
        with torch.no_grad():
            for _ in range(5):
                decoder.render_views(sc.extrinsics, sc.intrinsics, sc.near, sc.far, (480, 640), bg, sc.means, sc.covariances, sc.harmonics, sc.opacities)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                decoder.render_views(sc.extrinsics, sc.intrinsics, sc.near, sc.far, (480, 640), bg, sc.means, sc.covariances, sc.harmonics, sc.opacities)
            e1.record(); torch.cuda.synchronize()
        row.setdefault(f"cap{cap}_ms", []).append(round(e0.elapsed_time(e1) / 50, 4))
        row["heaviest_tile"] = heaviest; row["R"] = st.num_rendered()
        row[f"cap{cap}_fallback"] = flag
    res[name] = row
print(json.dumps(res))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bench_bins.json"), "w"), indent=1)
