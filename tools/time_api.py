"""Where does the public-API time go?  Host wall clock per call (with a final synchronise) for the device-resident entry
points at BASELINE config 2, next to the CUDA-graph replay of the same step.  cuda:0.
    python tools/time_api.py > gpurun_out/time_api.json"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from freesplat_b200 import decoder, rasterizer, synth  # noqa: E402

dev = "cuda:0"
sc = synth.pixel_aligned_scene(seed=0, h=480, w=640, n_context=2, n_target=3, keep=307200).to(dev)
bg = torch.zeros((3, 3), device=dev)


def wall(fn, n=100):
    for _ in range(10):
        fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    return {"ms_per_call": (time.perf_counter() - t0) / n * 1e3, "host_enqueue_ms_per_call": t_enq / n * 1e3}


res = {}
cov9 = sc.covariances.reshape(-1, 9)
views = decoder.camera_records_fused(sc.extrinsics, sc.intrinsics, sc.near, sc.far, bg, True)
with torch.no_grad():
    res["camera_records_fused"] = wall(lambda: decoder.camera_records_fused(sc.extrinsics, sc.intrinsics, sc.near, sc.far, bg, True))
    kw = dict(shs=sc.harmonics, cov3D_precomp=cov9, sh_degree=2, sh_layout=1, cov_stride=9)
    res["raster_forward_raw_sync"] = wall(lambda: rasterizer.raster_forward_raw(sc.means, sc.opacities, views, 480, 640, check_overflow="sync", **kw))
    res["raster_forward_raw_deferred"] = wall(lambda: rasterizer.raster_forward_raw(sc.means, sc.opacities, views, 480, 640, check_overflow="deferred", **kw))
    res["raster_forward_raw_deferred_reuse_scratch"] = wall(lambda: rasterizer.raster_forward_raw(sc.means, sc.opacities, views, 480, 640, check_overflow="deferred", reuse_scratch=True, **kw))
    rv = lambda **k: decoder.render_views(sc.extrinsics, sc.intrinsics, sc.near, sc.far, (480, 640), bg, sc.means, sc.covariances, sc.harmonics, sc.opacities, **k)
    res["render_views_sync"] = wall(lambda: rv(check_overflow="sync"))
    res["render_views_default_deferred"] = wall(lambda: rv())
    plan = rasterizer.RasterPlan(sc.means, sc.opacities, 480, 640, shs=sc.harmonics, cov3D_precomp=cov9,
                                 cameras=(sc.extrinsics, sc.intrinsics, sc.near, sc.far, bg), sh_degree=2, sh_layout=1, cov_stride=9)
    plan.run_checked()
    res["raster_plan_graph_launch"] = wall(lambda: plan.run())
    rasterizer.poll_deferred(block=True)
print(json.dumps(res, indent=1))
