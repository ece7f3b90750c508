"""Times the fused cost volume at BASELINE config-3 size on cuda:0 (CUDA events) and dumps JSON."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from freesplat_b200 import synth  # noqa: E402
from freesplat_b200.cost_volume import AVGFeatureVolumeManager  # noqa: E402

V, K, Hf, Wf, D = int(os.environ.get("CV_V", 3)), int(os.environ.get("CV_K", 2)), 120, 160, 128
dev = "cuda:0"
inp = {k: v.to(dev) for k, v in synth.cost_volume_inputs(0, V, K, 48, Hf, Wf).items()}
m = AVGFeatureVolumeManager(Hf, Wf, num_depth_bins=D, matching_dim_size=48).to(dev)
with torch.no_grad():
    for _ in range(3):
        out = m(**inp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        out = m(**inp)
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
rows = V * D * Hf * Wf
flops = rows * (2 * 2624 + K * (2 * 4 * 48 + 2 * 48))
res = {"views": V, "K": K, "ms_per_forward": ms, "ms_per_ref_view": ms / V, "tflops": flops / (ms * 1e-3) / 1e12,
       "alg_bytes": ((1 + K) * 48 * Hf * Wf * 4 + D * Hf * Wf * 4) * V}
print(json.dumps(res))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bench_cv.json"), "w"))
