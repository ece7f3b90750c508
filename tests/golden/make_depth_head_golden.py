"""Generates tests/golden/depth_head_*.npz by running the REFERENCE's own DepthDecoder.forward
(/root/reference/src/model/encoder/modules/networks.py:110-152) on seeded features and recording, per scale, the
plane logits that enter its tail (forward hooks on conv_depth[i]) together with the tail's outputs
(depth_pred_s{i}, log_depth_pred_s{i}, depth_pred_s-1, depth_weights).  Run in the build container."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.golden import ref_loader  # noqa: E402


def main():
    ref_loader.load_cost_volume_module()
    nw = sys.modules["refmods.networks"]
    cases = {  # name: (seed, batch, planes, h0, w0, log_planes, near, far, logit gain)
        "depth_head_log_d16": (0, 2, 16, 32, 48, True, 0.5, 15.0, 24.0),
        "depth_head_log_d128": (1, 1, 128, 16, 32, True, 0.5, 15.0, 40.0),
        "depth_head_inv_d64": (2, 1, 64, 16, 48, False, 1.0, 100.0, 40.0),
    }
    for name, (seed, B, D, h0, w0, logp, near, far, gain) in cases.items():
        torch.manual_seed(seed)
        dd = nw.DepthDecoder([8, 8, 8, 8, 8], num_output_channels=1 + 4, near=near, far=far, num_samples=D, log_planes=logp).eval()
        with torch.no_grad():   # sharper plane distributions than a fresh init gives (peaked softmax, as after training)
            for i in range(4):
                dd.conv_depth[f"{i}"][1].weight.mul_(gain)
        feats = [torch.randn(B, 8, h0 // (2 ** i), w0 // (2 ** i)) for i in range(5)]
        cap = {}
        live = {}

        def hook(m, inp, out, i=None):
            cap[i] = out.detach().clone()
            out.retain_grad()                       # gradient of the tail's loss w.r.t. the plane logits (reference autograd)
            live[i] = out
        for i in range(4):
            dd.conv_depth[f"{i}"].register_forward_hook(lambda m, inp, out, i=i: hook(m, inp, out, i))
        out = dd(feats)
        gen = torch.Generator().manual_seed(300 + seed)
        wts = {}
        loss = 0.0
        for i in range(4):
            for key in (f"depth_pred_s{i}_b1hw", f"log_depth_pred_s{i}_b1hw"):
                wts[key] = torch.randn(out[key].shape, generator=gen)
                loss = loss + (out[key] * wts[key]).sum()
        for key in ("depth_pred_s-1_b1hw", "depth_weights"):
            wts[key] = torch.randn(out[key].shape, generator=gen)
            loss = loss + (out[key] * wts[key]).sum()
        loss.backward()
        out = {k: v.detach() for k, v in out.items()}
        z = dict(meta=np.array([seed, B, D, h0, w0, int(logp)]), near_far=np.array([near, far], np.float32),
                 candi=dd.depth_candi_curr.reshape(-1).numpy().astype(np.float32))
        for i in range(4):
            z[f"logits_s{i}"] = cap[i].numpy()
            z[f"depth_s{i}"] = out[f"depth_pred_s{i}_b1hw"].numpy()
            z[f"log_depth_s{i}"] = out[f"log_depth_pred_s{i}_b1hw"].numpy()
            z[f"g_logits_s{i}"] = live[i].grad.numpy()
            z[f"w_depth_s{i}"] = wts[f"depth_pred_s{i}_b1hw"].numpy()
            z[f"w_log_depth_s{i}"] = wts[f"log_depth_pred_s{i}_b1hw"].numpy()
        z["w_depth_up"] = wts["depth_pred_s-1_b1hw"].numpy()
        z["w_weights_up"] = wts["depth_weights"].numpy()
        z["depth_up"] = out["depth_pred_s-1_b1hw"].numpy()
        z["weights_up"] = out["depth_weights"].numpy()
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **z)
        print(name, {k: v.shape for k, v in z.items() if k.startswith(("logits", "depth_up"))},
              "weights range", float(z["weights_up"].min()), float(z["weights_up"].max()))


if __name__ == "__main__":
    main()
