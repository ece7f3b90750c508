"""CUDA-event time of the PTF TRAINING fold (forward + backward, 3 views of 640x480 = BASELINE config 3) with the GRU backward
on the tensor cores (fs_ptf_gru_bwd_data / _weights) and with the earlier recompute + fp32 cuBLAS path.
    python tools/bench_gru_bwd.py [V]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from freesplat_b200 import ptf, synth
dev = torch.device("cuda", 0)
V = int(sys.argv[1]) if len(sys.argv) > 1 else 3
feats, coords, dens, wemb, depths, ext, K, hw = bench._flat_ptf(synth.ptf_inputs(0, V, 480, 640))
gru = bench.PlainGRU(synth.gru_state(0), dev)
base = [x.to(dev).contiguous() for x in (feats, coords, dens, wemb, depths)]
ext, K = ext.to(dev), K.to(dev)
res = {"views": V}


def step():
    a = [x.detach().requires_grad_(True) for x in base]
    F_, X_, E_, Z_ = ptf.fuse_views(gru, *a, ext, K, hw)
    (F_.sum() + X_.sum() + Z_.sum()).backward()


for mode in ("tc", "cublas", "tc", "cublas"):
    ptf.GRU_BWD = mode
    res.setdefault(mode + "_fwd_bwd_ms", []).append(round(bench.gpu_ms(step, n=3, warm=2), 3))
with torch.no_grad():
    res["inference_ms"] = round(bench.gpu_ms(lambda: ptf.fuse_views(gru, *base, ext, K, hw), n=5, warm=2), 3)
print(json.dumps(res))
