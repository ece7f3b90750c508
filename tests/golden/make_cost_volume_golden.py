"""Generates tests/golden/cost_volume_*.npz by running the REFERENCE's own code
(/root/reference/src/model/encoder/modules/cost_volume.py, unmodified, loaded by ref_loader) on
seeded inputs.  Run in the build container:  python tests/golden/make_cost_volume_golden.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from freesplat_b200 import synth  # noqa: E402
from tests.golden import ref_loader  # noqa: E402

CASES = {
    # name: (seed, views, K, C, Hf, Wf, D)
    "cost_volume_v3k2": (0, 3, 2, 48, 12, 16, 8),
    "cost_volume_v2k1": (1, 2, 1, 48, 10, 14, 6),
    "cost_volume_v4k3_wide": (2, 4, 3, 48, 9, 12, 5),
}


def main():
    cvmod = ref_loader.load_cost_volume_module()
    for name, (seed, V, K, C, Hf, Wf, D) in CASES.items():
        torch.manual_seed(seed)
        spacing = 0.6 if "wide" in name else 0.25     # wide baseline: samples leave the source image
        inp = synth.cost_volume_inputs(seed, V, K, C, Hf, Wf, spacing=spacing)
        mlp = synth.cost_volume_mlp(seed, C)
        m = cvmod.AVGFeatureVolumeManager(Hf, Wf, num_depth_bins=D, mlp_channels=[C + 1, 32, 32, 1], matching_dim_size=C)
        with torch.no_grad():
            for p, w in zip([m.mlp.net[0].weight, m.mlp.net[0].bias, m.mlp.net[2].weight, m.mlp.net[2].bias,
                             m.mlp.net[4].weight, m.mlp.net[4].bias], mlp):
                p.copy_(w)
        cur = inp["cur_feats"].clone().requires_grad_(True)
        src = inp["src_feats"].clone().requires_grad_(True)
        out, planes, _ = m.build_cost_volume(cur_feats=cur, src_feats=src, src_extrinsics=inp["src_extrinsics"],
                                             src_poses=inp["src_poses"], src_Ks=inp["src_Ks"], cur_invK=inp["cur_invK"],
                                             min_depth=inp["min_depth"], max_depth=inp["max_depth"])
        g = torch.Generator().manual_seed(77 + seed)
        wts = torch.randn(out.shape, generator=g)
        (out * wts).sum().backward()
        np.savez_compressed(
            os.path.join(ROOT, "tests", "golden", name + ".npz"),
            meta=np.array([seed, V, K, C, Hf, Wf, D]),
            **{k: v.numpy() for k, v in inp.items()}, **{f"mlp{i}": w.numpy() for i, w in enumerate(mlp)},
            out=out.detach().numpy(), planes=planes[:, :, 0, 0].detach().numpy(), wts=wts.numpy(),
            g_cur=cur.grad.numpy(), g_src=src.grad.numpy(),
            **{f"g_mlp{i}": p.grad.numpy() for i, p in enumerate([m.mlp.net[0].weight, m.mlp.net[0].bias, m.mlp.net[2].weight,
                                                                    m.mlp.net[2].bias, m.mlp.net[4].weight, m.mlp.net[4].bias])})
        print(name, tuple(out.shape), float(out.abs().mean()))


if __name__ == "__main__":
    main()
