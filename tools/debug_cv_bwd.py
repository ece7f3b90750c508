"""Per-gradient comparison of the tensor-core backward (mode 0) with the fp32 backward (mode 1)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from freesplat_b200 import synth, cost_volume as cvm
from freesplat_b200.cost_volume import AVGFeatureVolumeManager
dev = "cuda:0"


def run(V, K, Hf, Wf, D, wscale=2.0, detail=False):
    inp = {k: v.to(dev) for k, v in synth.cost_volume_inputs(11, V, K, 48, Hf, Wf).items()}
    mlp = [w * wscale for w in synth.cost_volume_mlp(11)]
    wts = torch.randn((V, D, Hf, Wf), generator=torch.Generator().manual_seed(3)).to(dev)
    grads = {}
    for mode in (0, 1):
        cvm.MLP_MODE = mode
        m = AVGFeatureVolumeManager(Hf, Wf, num_depth_bins=D, mlp_channels=[49, 32, 32, 1], matching_dim_size=48).to(dev)
        with torch.no_grad():
            for p, w in zip([m.mlp.net[0].weight, m.mlp.net[0].bias, m.mlp.net[2].weight, m.mlp.net[2].bias, m.mlp.net[4].weight, m.mlp.net[4].bias], mlp):
                p.copy_(w)
        cur = inp["cur_feats"].clone().requires_grad_(True); src = inp["src_feats"].clone().requires_grad_(True)
        out = m(**{**inp, "cur_feats": cur, "src_feats": src})
        (out * wts).sum().backward()
        grads[mode] = dict(cur=cur.grad, src=src.grad, W0=m.mlp.net[0].weight.grad, b0=m.mlp.net[0].bias.grad, W1=m.mlp.net[2].weight.grad,
                           b1=m.mlp.net[2].bias.grad, W2=m.mlp.net[4].weight.grad, b2=m.mlp.net[4].bias.grad)
    line = f"V{V} K{K} {Hf}x{Wf} D{D}: "
    for n in grads[0]:
        a, b = grads[0][n].double().cpu().numpy(), grads[1][n].double().cpu().numpy()
        line += f"{n}={np.abs(a-b).max()/(np.abs(b).max()+1e-30):.1e} "
    print(line)
    if detail:
        a, b = grads[0]["cur"].double().cpu().numpy(), grads[1]["cur"].double().cpu().numpy()
        e = np.abs(a - b) / np.abs(b).max()
        bad = np.argwhere(e > 1e-4)
        print("  cur: bad elements", len(bad), "of", e.size)
        if len(bad):
            print("  bad (v,c,h,w) sample:", bad[:12].tolist())
            print("  bad h hist:", np.bincount(bad[:, 2], minlength=Hf).tolist())
            print("  bad w hist:", np.bincount(bad[:, 3], minlength=Wf).tolist())
            print("  bad c hist:", np.bincount(bad[:, 1], minlength=48).tolist())
        a, b = grads[0]["b1"].double().cpu().numpy(), grads[1]["b1"].double().cpu().numpy()
        print("  b1 tc", a[:6], "\n  b1 fp", b[:6])


run(2, 1, 8, 64, 16, detail=True)
run(2, 1, 8, 64, 32)
run(2, 1, 40, 52, 16)
run(4, 3, 40, 52, 40, detail=True)
run(4, 3, 40, 52, 40, wscale=1.0)
