"""Forward+backward CUDA-event timings of every operator of the BASELINE config-3 training step, one by one (the chained
step is in bench.py `ops.config3_train_step`).  cuda:0.   python tools/bench_train_ops.py > gpurun_out/train_ops.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as Fn  # noqa: E402

import bench  # noqa: E402
from freesplat_b200 import adapter, decoder, depth_head, ptf, synth  # noqa: E402

dev = torch.device("cuda", 0)
H, W, V, D = 480, 640, 3, 128
HW = H * W
g = torch.Generator().manual_seed(0)
res = {}


def fb(make, n=3):
    """make() -> (loss-like tensor to call .backward(grad) on, grad)."""
    def run():
        out, go = make()
        out.backward(go)
    return bench.gpu_ms(run, n=n, warm=2)


def fwd(make, n=3):
    def run():
        with torch.no_grad():
            make()
    return bench.gpu_ms(run, n=n, warm=1)


# depth-head tail, scale 0: logits [3,128,240,320]
logits = (4 * torch.randn((V, D, 240, 320), generator=g)).to(dev).requires_grad_(True)
candi = torch.linspace(-0.7, 1.4, D, device=dev)
gu = torch.randn((V, 1, H, W), generator=g).to(dev)


def dh():
    o = depth_head.depth_regression(logits, candi, True, upsample=True)
    return o["depth_up"] + o["weights_up"], gu
res["depth_head_s0"] = {"fwd_ms": fwd(lambda: depth_head.depth_regression(logits, candi, True, upsample=True)), "fwd_bwd_ms": fb(dh)}
vol = torch.randn((V, D, 120, 160), generator=g).to(dev).requires_grad_(True)
res["nearest_x2_standin"] = {"fwd_bwd_ms": fb(lambda: (Fn.interpolate(vol * 8.0, scale_factor=2, mode="nearest"), logits.detach()))}
del logits, vol

# back-projection
depth = (0.5 + 3 * torch.rand((V, HW), generator=g)).to(dev).requires_grad_(True)
Kn = synth.intrinsics(V).to(dev); ext = synth.camera_path(V, spacing=0.25).to(dev)
gm = torch.randn((V, HW, 3), generator=g).to(dev)
res["backproject"] = {"fwd_bwd_ms": fb(lambda: (adapter.backproject_depth(depth, Kn[0], ext, (H, W)), gm))}

# PTF fold, training path (merge fwd/bwd native, GRU through torch autograd) and inference path
feats, coords, dens, wemb, depths, e3, K3, hw = bench._flat_ptf(synth.ptf_inputs(0, V, H, W))
gru = bench.PlainGRU(synth.gru_state(0), dev)
args = [x.to(dev).contiguous() for x in (feats, coords, dens, wemb, depths)]
leaves = [a.clone().requires_grad_(True) for a in args]
with torch.no_grad():
    N = int(ptf.fuse_views(gru, *args, e3.to(dev), K3.to(dev), hw)[0].shape[0])
gF = torch.randn((N, 64), generator=g).to(dev)
timings = []


def ptf_train():
    F_, X_, E_, Z_ = ptf.fuse_views(gru, *leaves, e3.to(dev), K3.to(dev), hw)
    return F_, gF
for p_ in gru.parameters():
    p_.requires_grad_(True)
res["ptf_3views"] = {"fwd_inference_ms": fwd(lambda: ptf.fuse_views(gru, *args, e3.to(dev), K3.to(dev), hw)), "fwd_bwd_training_ms": fb(ptf_train)}
with torch.no_grad():
    ptf.fuse_views(gru, *args, e3.to(dev), K3.to(dev), hw, timings=timings)
res["ptf_3views"]["inference_steps"] = timings
for p_ in gru.parameters():
    p_.requires_grad_(False)

# Gaussian head
raw = torch.randn((N, 34), generator=g).to(dev).requires_grad_(True)
dpt = (0.5 + 3 * torch.rand((N,), generator=g)).to(dev).requires_grad_(True)
op = torch.rand((N,), generator=g).to(dev).requires_grad_(True)
xyz = torch.randn((N, 3), generator=g).to(dev).requires_grad_(True)
E = ext[0][None].expand(N, 4, 4).contiguous().requires_grad_(True)
gc = torch.randn((N, 3, 3), generator=g).to(dev)


def head():
    gs = adapter.gaussian_head(raw, dpt, op, xyz, E, Kn[0], (H, W))
    return gs.covariances, gc
res["gaussian_head"] = {"N": N, "fwd_bwd_ms": fb(head)}
print(json.dumps(res, indent=1))
