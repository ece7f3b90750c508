"""Debug aid: gradients of the PTF training fold with the hand-derived GRU backward (_GruTrain) vs torch autograd through the
GRU module, same inputs (mid-size golden case).  Prints the worst elements."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from freesplat_b200 import ptf, synth
from tests.ptf_helpers import flat_inputs
from tests.test_ptf_gpu import _gru
dev = "cuda:0"
seed, V, h, w = 5, 4, 120, 160
inp = synth.ptf_inputs(seed, V, h, w)
feats, coords, dens, wemb, depths, ext, K, hw = flat_inputs(inp)
gen = torch.Generator().manual_seed(500 + seed)
res = {}
for mode in ("tc", "cublas"):
    ptf.GRU_MODE = mode
    t = lambda a, g=True: torch.from_numpy(np.ascontiguousarray(a)).to(dev).requires_grad_(g)
    tf, tx, td, tw, tz = t(feats), t(coords), t(dens), t(wemb), t(depths)
    gru = _gru(seed, dev)
    F_, X_, E_, Z_ = ptf.fuse_views(gru, tf, tx, td, tw, tz, t(ext, False), t(K, False), hw)
    if mode == "tc":
        wF = torch.randn(F_.shape, generator=torch.Generator().manual_seed(1)).to(dev)
    (F_ * wF).sum().backward()
    res[mode] = dict(F=F_.detach(), feats=tf.grad, dens=td.grad, wemb=tw.grad, **{n: p.grad for n, p in gru.named_parameters()})
print("forward max diff", float((res["tc"]["F"] - res["cublas"]["F"]).abs().max()))
for k in res["tc"]:
    if k == "F":
        continue
    a, b = res["tc"][k], res["cublas"][k]
    d = (a - b).abs()
    i = int(d.argmax())
    print(k, "max abs diff", float(d.max()), "of max", float(d.max() / b.abs().max()), "at", np.unravel_index(i, tuple(a.shape)), float(a.flatten()[i]), float(b.flatten()[i]))
td_ = res["tc"]["dens"] - res["cublas"]["dens"]
idx = torch.nonzero(td_.abs() > 1e-3 * res["cublas"]["dens"].abs().max())
print("dens outliers (view, pixel):", idx[:10].tolist(), "values", [(float(res['tc']['dens'][tuple(i)]), float(res['cublas']['dens'][tuple(i)])) for i in idx[:10]])
print("dens at those:", [float(torch.from_numpy(dens)[tuple(i)]) for i in idx[:10].cpu()])
