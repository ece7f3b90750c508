"""Generates tests/golden/ptf_*.npz by running the REFERENCE's own fuse_gaussians
(/root/reference/src/model/encoder/encoder_freesplat.py:431-522, extracted unmodified by ref_loader,
with the reference's GRU from modules/networks.py:188-214) on seeded inputs.
Run in the build container:  python tests/golden/make_ptf_golden.py"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from freesplat_b200 import synth  # noqa: E402
from tests.golden import ref_loader  # noqa: E402

CASES = {"ptf_v3": (0, 3, 16, 24), "ptf_v4": (1, 4, 12, 20), "ptf_v2_far": (2, 2, 16, 16), "ptf_v3_ties": (3, 3, 12, 16)}


def main():
    fuse, pe, GRU = ref_loader.load_fuse_gaussians()
    for name, (seed, V, h, w) in CASES.items():
        inp = synth.ptf_inputs(seed, V, h, w, spacing=0.9 if "far" in name else 0.2)
        if "ties" in name:
            # exact z-buffer ties: every 5th pixel of view 0 takes the coordinates of its left neighbour, so both project to the
            # same pixel of the later views with IDENTICAL depth -> several winners per pixel (encoder_freesplat.py:470-482)
            c = inp["coords"][0]
            idx = torch.arange(1, h * w, 5)
            c[0, 0, idx] = c[0, 0, idx - 1]
        gru = GRU()
        gru.load_state_dict(synth.gru_state(seed))
        self = SimpleNamespace(gru=gru)
        # inputs that receive gradients in training (encoder outputs) + the GRU parameters
        gi = {k: inp[k].clone().requires_grad_(True) for k in ("densities", "weight_emb", "depths")}
        g_feats = inp["gaussians"][0].clone().requires_grad_(True)
        g_coords = inp["coords"][0].clone().requires_grad_(True)
        feats, coords, extr, depths = fuse(self, [g_feats], [g_coords], gi["densities"], gi["weight_emb"],
                                           gi["depths"], inp["extrinsics"], inp["intrinsics"], inp["image_shape"])
        gen = torch.Generator().manual_seed(500 + seed)
        wF, wX = torch.randn(feats.shape, generator=gen), torch.randn(coords.shape, generator=gen)
        wE, wZ = torch.randn(extr.shape, generator=gen), torch.randn(depths.shape, generator=gen)
        ((feats * wF).sum() + (coords * wX).sum() + (extr * wE).sum() + (depths * wZ).sum()).backward()
        grads = dict(g_in_feats=g_feats.grad, g_in_coords=g_coords.grad, g_in_dens=gi["densities"].grad,
                     g_in_wemb=gi["weight_emb"].grad, g_in_depths=gi["depths"].grad)
        grads.update({"g_gru." + k: v.grad for k, v in gru.named_parameters()})
        feats, coords, extr, depths = feats.detach(), coords.detach(), extr.detach(), depths.detach()
        np.savez_compressed(
            os.path.join(ROOT, "tests", "golden", name + ".npz"), meta=np.array([seed, V, h, w]),
            in_feats=inp["gaussians"][0].numpy(), in_coords=inp["coords"][0].numpy(), in_dens=inp["densities"].numpy(),
            in_wemb=inp["weight_emb"].numpy(), in_depths=inp["depths"].numpy(), in_ext=inp["extrinsics"].numpy(),
            in_K=inp["intrinsics"].numpy(), out_feats=feats.numpy(), out_coords=coords.numpy(), out_ext=extr.numpy(),
            out_depths=depths.numpy(), wF=wF.numpy(), wX=wX.numpy(), wE=wE.numpy(), wZ=wZ.numpy(),
            **{k: v.numpy() for k, v in grads.items()})
        print(name, "N_out", feats.shape[1], "of", V * h * w)


if __name__ == "__main__":
    main()
