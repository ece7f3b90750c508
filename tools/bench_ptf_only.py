"""PTF at config-4 size (10 views, 640x480) -- used as the ncu target of tools/profile_ops.sh."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from freesplat_b200 import ptf, synth
from ptf_helpers import flat_inputs
from test_ptf_gpu import GRU
dev = "cuda:0"
V = int(os.environ.get("PTF_V", 6))
inp = synth.ptf_inputs(0, V, 480, 640)
feats, coords, dens, wemb, depths, ext, K, hw = flat_inputs(inp)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
g = GRU(); g.load_state_dict(synth.gru_state(0)); g = g.to(dev)
with torch.no_grad():
    out = ptf.fuse_views(g, *[t(x) for x in (feats, coords, dens, wemb, depths, ext, K)], hw)
torch.cuda.synchronize()
print(out[0].shape)
