"""Raster fwd + bwd at BASELINE config-3 size (P = 460 800, 4 target views, MSE loss): ncu target / timing."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from freesplat_b200 import decoder, synth  # noqa: E402

dev = "cuda:0"
sc = synth.pixel_aligned_scene(seed=0, h=480, w=640, n_context=3, n_target=4, keep=460800).to(dev)
bg = torch.zeros((4, 3), device=dev)
means = sc.means.clone().requires_grad_(True); cov = sc.covariances.clone().requires_grad_(True)
sh = sc.harmonics.clone().requires_grad_(True); op = sc.opacities.clone().requires_grad_(True)
target = torch.rand((4, 3, 480, 640), device=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for it in range(n + 2):
    if it == 2:
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record()
    c, d = decoder.render_views(sc.extrinsics, sc.intrinsics, sc.near, sc.far, (480, 640), bg, means, cov, sh, op)
    ((c - target) ** 2).mean().backward()
e1.record(); torch.cuda.synchronize()
print(f"raster fwd+bwd {e0.elapsed_time(e1) / n:.3f} ms")
