"""Times the cost-volume backward (config-3 size) on cuda:0; FS_CV_BWD_DEBUG experiments."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from freesplat_b200 import synth
from freesplat_b200.cost_volume import AVGFeatureVolumeManager
V, K, Hf, Wf, D = int(os.environ.get("CV_V", 3)), int(os.environ.get("CV_K", 2)), 120, 160, 128
dev = "cuda:0"
inp = {k: v.to(dev) for k, v in synth.cost_volume_inputs(0, V, K, 48, Hf, Wf).items()}
inp["cur_feats"].requires_grad_(True); inp["src_feats"].requires_grad_(True)
m = AVGFeatureVolumeManager(Hf, Wf, num_depth_bins=D, matching_dim_size=48).to(dev)
out = m(**inp)
g = torch.randn_like(out)
n = int(os.environ.get("CV_N", 5))
for _ in range(2):
    out.backward(g, retain_graph=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    out.backward(g, retain_graph=True)
e1.record(); torch.cuda.synchronize()
print(json.dumps({"dbg": os.environ.get("FS_CV_BWD_DEBUG", "0"), "V": V, "K": K, "ms_bwd": e0.elapsed_time(e1) / n}))
