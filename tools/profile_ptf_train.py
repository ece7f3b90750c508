"""torch.profiler kernel table of the 3-view PTF training fold (forward + backward) at 640x480: where the time goes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from freesplat_b200 import ptf, synth
dev = torch.device("cuda", 0)
V, H, W = 3, 480, 640
feats, coords, dens, wemb, depths, e3, K3, hw = bench._flat_ptf(synth.ptf_inputs(0, V, H, W))
gru = bench.PlainGRU(synth.gru_state(0), dev)
for p_ in gru.parameters():
    p_.requires_grad_(True)
leaves = [x.to(dev).contiguous().requires_grad_(True) for x in (feats, coords, dens, wemb, depths)]
e3, K3 = e3.to(dev), K3.to(dev)
gF = None
def step():
    global gF
    F_, X_, E_, Z_ = ptf.fuse_views(gru, *leaves, e3, K3, hw)
    if gF is None:
        gF = torch.randn_like(F_)
    F_.backward(gF)
for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=35, max_name_column_width=70))
