"""Upstream-STRUCTURED proxy (baseline/upstream_proxy.cu) vs this repo on the bench workload, same B200.
The proxy is NOT the reference's extension (un-vendored, not installable here); it reproduces its kernel
structure so that a '2023-style CUDA recompiled for sm_100' number stands next to ours.  JSON to gpurun_out/."""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from freesplat_b200 import decoder, rasterizer, synth  # noqa: E402

so = os.path.join(ROOT, "baseline", "libupstream_proxy.so")
if not os.path.exists(so):
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler",
                           "-fPIC", "-shared", "-ccbin", "/usr/bin/g++", "-o", so, os.path.join(ROOT, "baseline", "upstream_proxy.cu")])
L = C.CDLL(so)
L.proxy_render_view.restype = C.c_longlong
dev = "cuda:0"
H, W, P, V = 480, 640, 307200, 3
sc = synth.pixel_aligned_scene(seed=0, h=H, w=W, n_context=2, n_target=V, keep=P).to(dev)
bg = torch.zeros((V, 3), device=dev)
views, _ = decoder.camera_records(sc.extrinsics, sc.intrinsics, sc.near, sc.far, bg, True)
row, col = torch.triu_indices(3, 3)
cov6 = sc.covariances[:, row, col].contiguous()
shs = sc.harmonics.transpose(1, 2).contiguous()
assert L.proxy_setup(C.c_int(P), C.c_int(H), C.c_int(W), C.c_size_t(4 * P)) == 0
oc = torch.empty((V, 3, H, W), device=dev); od = torch.empty((V, H, W), device=dev)
stream = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def proxy_step():
    for v in range(V):                      # the reference renders view by view
        r = L.proxy_render_view(C.c_int(9), C.c_int(2), C.c_void_p(sc.means.data_ptr()), C.c_void_p(cov6.data_ptr()),
                                C.c_void_p(sc.opacities.data_ptr()), C.c_void_p(shs.data_ptr()), C.c_void_p(views[v].data_ptr()),
                                C.c_void_p(oc[v].data_ptr()), C.c_void_p(od[v].data_ptr()), C.c_void_p(stream))
        assert r >= 0


def ours_step():
    return rasterizer.raster_forward_raw(sc.means, sc.opacities, views, H, W, shs=shs, cov3D_precomp=cov6, sh_degree=2,
                                         check_overflow="deferred")


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for k in range(n):
        flush.fill_(k & 255)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n


st = ours_step(); proxy_step(); torch.cuda.synchronize()
same = torch.isclose(st.color, oc, rtol=1e-4, atol=1e-5).float().mean().item()
p_ms, o_ms = timeit(proxy_step), timeit(ours_step)
res = {"workload": "scannet_2views_640x480_P307200_targets3_raster_fwd", "proxy_ms_per_step": p_ms, "ours_ms_per_step": o_ms,
       "proxy_views_per_s": V / (p_ms * 1e-3), "ours_views_per_s": V / (o_ms * 1e-3), "ratio": p_ms / o_ms,
       "images_agree_frac": same,
       "note": "proxy = upstream-structured kernels written for this comparison (cub scan + D2H sync + cub 64-bit radix sort + "
               "uncull 16x16 render), NOT the reference's own extension"}
print(json.dumps(res, indent=1))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bench_proxy.json"), "w"), indent=1)
