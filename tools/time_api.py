"""Where does the public-API time go?  (host wall clock with synchronisation, cuda:0)"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from freesplat_b200 import decoder, rasterizer, synth

dev = "cuda:0"
sc = synth.pixel_aligned_scene(seed=0, h=480, w=640, n_context=2, n_target=3, keep=307200).to(dev)
bg = torch.zeros((3, 3), device=dev)


def wall(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


res = {}
res["camera_records_ms"] = wall(lambda: decoder.camera_records(sc.extrinsics, sc.intrinsics, sc.near, sc.far, bg, True))
views, _ = decoder.camera_records(sc.extrinsics, sc.intrinsics, sc.near, sc.far, bg, True)
with torch.no_grad():
    res["raster_forward_raw_sync_ms"] = wall(lambda: rasterizer.raster_forward_raw(sc.means, sc.opacities, views, 480, 640, shs=sc.harmonics, cov3D_precomp=sc.covariances.reshape(-1, 9), sh_degree=2, sh_layout=1, cov_stride=9))
    res["raster_forward_raw_deferred_ms"] = wall(lambda: rasterizer.raster_forward_raw(sc.means, sc.opacities, views, 480, 640, shs=sc.harmonics, cov3D_precomp=sc.covariances.reshape(-1, 9), sh_degree=2, sh_layout=1, cov_stride=9, check_overflow="deferred"))
    res["render_views_sync_ms"] = wall(lambda: decoder.render_views(sc.extrinsics, sc.intrinsics, sc.near, sc.far, (480, 640), bg, sc.means, sc.covariances, sc.harmonics, sc.opacities))
    res["render_views_deferred_ms"] = wall(lambda: decoder.render_views(sc.extrinsics, sc.intrinsics, sc.near, sc.far, (480, 640), bg, sc.means, sc.covariances, sc.harmonics, sc.opacities, check_overflow="deferred"))
print(json.dumps(res, indent=1))
