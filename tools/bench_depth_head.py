"""Depth-head tail only (ncu target / timing): 3 views at 640x480 (scale 0 = 240x320, D = 128)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from freesplat_b200.depth_head import depth_regression  # noqa: E402

dev = "cuda:0"
V, D, h, w = 3, 128, 240, 320
lg = torch.randn((V, D, h, w), device=dev) * 5
cand = (torch.log(torch.tensor(0.5)) + torch.linspace(0, 1, D) * torch.log(torch.tensor(30.0))).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
with torch.no_grad():
    for mode in (0, 1):
        tot = 0.0
        for it in range(n + 3):
            flush.zero_(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); depth_regression(lg, cand, True, upsample=True, tile_mode=mode); e1.record(); torch.cuda.synchronize()
            if it >= 3:
                tot += e0.elapsed_time(e1)
        alg = V * (D * h * w * 4 + 10 * h * w * 4)
        print(f"tile_mode {mode}: {tot / n * 1e3:.1f} us  {alg / (tot / n * 1e-3) / 1e9:.0f} GB/s algorithmic")
