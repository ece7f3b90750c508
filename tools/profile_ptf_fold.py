"""torch.profiler kernel table of the PTF inference fold (10 views of 640x480): pool vs compacting."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from freesplat_b200 import ptf, synth
dev = torch.device("cuda", 0)
V = 10
feats, coords, dens, wemb, depths, ext, K, hw = bench._flat_ptf(synth.ptf_inputs(0, V, 480, 640))
gru = bench.PlainGRU(synth.gru_state(0), dev)
args = [x.to(dev).contiguous() for x in (feats, coords, dens, wemb, depths, ext, K)]
for pool in (True, False):
    ptf.POOL = pool
    with torch.no_grad():
        for _ in range(3):
            ptf.fuse_views(gru, *args, hw)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            ptf.fuse_views(gru, *args, hw)
            torch.cuda.synchronize()
    print("POOL =", pool)
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=60))
