"""CUDA-event time of the PTF inference fold (V views of 640x480, default 10) with the append-only pool and with the compacting
fold.   python tools/bench_ptf_fold.py [V]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from freesplat_b200 import ptf, synth
dev = torch.device("cuda", 0)
V = int(sys.argv[1]) if len(sys.argv) > 1 else 10
feats, coords, dens, wemb, depths, ext, K, hw = bench._flat_ptf(synth.ptf_inputs(0, V, 480, 640))
gru = bench.PlainGRU(synth.gru_state(0), dev)
args = [x.to(dev).contiguous() for x in (feats, coords, dens, wemb, depths, ext, K)]
res = {"views": V}
with torch.no_grad():
    for pool in (True, False, True, False):
        ptf.POOL = pool
        ms = bench.gpu_ms(lambda: ptf.fuse_views(gru, *args, hw), n=5, warm=2)
        res.setdefault("pool_ms" if pool else "compacting_ms", []).append(round(ms, 4))
    res["N_out"] = int(ptf.fuse_views(gru, *args, hw)[0].shape[0])
print(json.dumps(res))
