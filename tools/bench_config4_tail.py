"""BASELINE config 4 ("ScanNet FVT 10-views with Pixel-wise Triplet Fusion merge, 1xB200"), everything downstream of the
convolutions on the B200 kernels, chained through the public operator mirrors:

  matching features --cost volume (tcgen05)--> [V,128,120,160] volumes            (freesplat_b200.cost_volume)
  plane logits      --depth-head tail (TMA)--> depths / weights [V,480,640]       (freesplat_b200.depth_head)
  depths            --back-projection-------> pixel-aligned coordinates           (freesplat_b200.adapter.backproject_depth)
  latents + coords  --PTF (match, tcgen05 GRU, merge)--> fused Gaussian set       (freesplat_b200.ptf.fuse_views)
  fused latents     --Gaussian head--------->  means / covariances / SH / opacity (freesplat_b200.adapter.gaussian_head)
  Gaussians         --rasterizer (9 target views)--> colour + depth               (freesplat_b200.decoder.render_views)

The convolutional backbone / cost-volume encoder / depth decoder between these stages are the reference's (out of scope,
SURVEY §8): their outputs are replaced by seeded synthetic tensors of the right shapes (smooth depth-consistent logits so
that PTF sees real cross-view overlap).  Prints per-stage CUDA-event timings; JSON to gpurun_out/config4_tail.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from freesplat_b200 import decoder, ptf, synth  # noqa: E402
from freesplat_b200.adapter import backproject_depth, gaussian_head  # noqa: E402
from freesplat_b200.cost_volume import AVGFeatureVolumeManager  # noqa: E402
from freesplat_b200.depth_head import depth_regression  # noqa: E402
from ptf_helpers import flat_inputs  # noqa: E402
from test_ptf_gpu import GRU  # noqa: E402

dev = torch.device("cuda", 0)
V, K, T, h, w, D, F = 10, 8, 9, 480, 640, 128, 64
g = torch.Generator().manual_seed(0)
# ---- synthetic stand-ins for the network outputs -------------------------------------------------------------------
cv_in = {k: v.to(dev) for k, v in synth.cost_volume_inputs(0, V, K, 48, h // 4, w // 4).items()}
cvm = AVGFeatureVolumeManager(h // 4, w // 4, num_depth_bins=D, matching_dim_size=48).to(dev)
pin = synth.ptf_inputs(0, V, h, w)                                  # depth-consistent views of one synthetic room
feats, coords_ref, dens, wemb, depths_ref, ext, Kn, hw = flat_inputs(pin)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
feats, dens, wemb, ext, Kn = t(feats), t(dens), t(wemb), t(ext), t(Kn)
candi = (torch.log(torch.tensor(0.5)) + torch.linspace(0, 1, D) * torch.log(torch.tensor(30.0))).to(dev)
# plane logits at scale 0 (240x320) peaked at the synthetic scene's log depth (what a trained depth decoder produces)
d_half = torch.nn.functional.avg_pool2d(t(depths_ref).reshape(V, 1, h, w), 2).log()
logits = -((candi.view(1, D, 1, 1) - d_half) ** 2) * 60.0 + 0.3 * torch.randn((V, D, h // 2, w // 2), generator=g).to(dev)
gru = GRU(); gru.load_state_dict(synth.gru_state(0)); gru = gru.to(dev)
to_gaussians = torch.nn.Linear(F, 34).to(dev)                       # stand-in for encoder.to_gaussians (encoder_freesplat.py:180-186)
cams = synth.camera_path(T, spacing=0.1, t0=0.3).to(dev)
tK = synth.intrinsics(T).to(dev); near = torch.full((T,), 0.5, device=dev); far = torch.full((T,), 15.0, device=dev)
bg = torch.zeros((T, 3), device=dev)


def ev():
    return torch.cuda.Event(enable_timing=True)


def run():
    e = [ev() for _ in range(7)]
    with torch.no_grad():
        e[0].record()
        vol = cvm(**cv_in)                                                        # [V,128,120,160]
        e[1].record()
        dh = depth_regression(logits, candi, True, upsample=True)                 # depth_up / weights_up [V,1,480,640]
        e[2].record()
        coords = backproject_depth(dh["depth_up"].reshape(V, h, w), Kn[0], ext, (h, w))
        e[3].record()
        F_, X_, E_, Z_ = ptf.fuse_views(gru, feats, coords, dens, dh["weights_up"].reshape(V, -1), dh["depth_up"].reshape(V, -1),
                                        ext, Kn, (h, w))
        e[4].record()
        raw = to_gaussians(F_)
        gs = gaussian_head(raw[:, :34].contiguous(), Z_, torch.sigmoid(raw[:, 0]), X_, E_, Kn[0], (h, w))
        e[5].record()
        col, dep = decoder.render_views(cams, tK, near, far, (h, w), bg, gs.means, gs.covariances, gs.harmonics, gs.opacities)
        e[6].record()
    torch.cuda.synchronize()
    return [e[i].elapsed_time(e[i + 1]) for i in range(6)], F_.shape[0], vol, col


for _ in range(2):
    ms, N, vol, col = run()
acc = np.zeros(6)
n = 5
for _ in range(n):
    ms, N, vol, col = run()
    acc += np.array(ms)
acc /= n
names = ["cost_volume_ms", "depth_head_ms", "backproject_ms", "ptf_ms", "gaussian_head_ms", "render_ms"]
res = {"context_views": V, "src_views_per_ref": K, "target_views": T, "fused_gaussians": int(N), "gs_ratio": N / (V * h * w),
       **{k: float(v) for k, v in zip(names, acc)}, "total_ms": float(acc.sum()),
       "finite": bool(torch.isfinite(col).all() and torch.isfinite(vol).all()), "coverage": float((col.sum(1) > 0).float().mean())}
print(json.dumps(res, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "config4_tail.json"), "w"), indent=1)
