"""Mid-size reference goldens (cost volume 60x80xD32, K = 2; PTF 120x160, V = 4) produced by the REFERENCE's own code
(loaded unmodified by ref_loader), in a compact form: the inputs are re-generated from their seeds by freesplat_b200.synth
(a float64 checksum of every input tensor is stored and verified by the tests), integer / order-defining outputs are
stored in full, the large float outputs and gradients as strided samples (the stride is stored).
Run in the build container:  python tests/golden/make_mid_golden.py"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from freesplat_b200 import synth  # noqa: E402
from tests.golden import ref_loader  # noqa: E402

CV_CASE = ("mid_cost_volume_v3k2", 5, 3, 2, 48, 60, 80, 32)        # name, seed, V, K, C, Hf, Wf, D
PTF_CASE = ("mid_ptf_v4", 5, 4, 120, 160)                           # name, seed, V, h, w
SRC_CH_STRIDE, ROW_STRIDE = 4, 8
KINK_EPS = 2e-4


def checksum(t) -> float:
    return float(torch.as_tensor(t).double().sum())


def cost_volume():
    name, seed, V, K, C, Hf, Wf, D = CV_CASE
    cvmod = ref_loader.load_cost_volume_module()
    inp = synth.cost_volume_inputs(seed, V, K, C, Hf, Wf)
    mlp = synth.cost_volume_mlp(seed, C)
    m = cvmod.AVGFeatureVolumeManager(Hf, Wf, num_depth_bins=D, mlp_channels=[C + 1, 32, 32, 1], matching_dim_size=C)
    params = [m.mlp.net[0].weight, m.mlp.net[0].bias, m.mlp.net[2].weight, m.mlp.net[2].bias, m.mlp.net[4].weight, m.mlp.net[4].bias]
    with torch.no_grad():
        for p, w in zip(params, mlp):
            p.copy_(w)
    # voxel rows next to a LeakyReLU kink (|pre-activation| < KINK_EPS in the REFERENCE's own evaluation) get zero loss weight:
    # the derivative jumps by a factor of 100 there, so two fp32 evaluations may legitimately differ; the mask is stored
    pre = {0: [], 2: []}
    hooks = [m.mlp.net[i].register_forward_hook(lambda mod, a, o, i=i: pre[i].append(o.detach().abs().amin(-1) < KINK_EPS)) for i in (0, 2)]
    with torch.no_grad():
        m.build_cost_volume(cur_feats=inp["cur_feats"], src_feats=inp["src_feats"], src_extrinsics=inp["src_extrinsics"],
                            src_poses=inp["src_poses"], src_Ks=inp["src_Ks"], cur_invK=inp["cur_invK"], min_depth=inp["min_depth"],
                            max_depth=inp["max_depth"])
    for hk in hooks:
        hk.remove()
    kink = torch.stack(pre[0], 1) | torch.stack(pre[2], 1)           # [V,D,Hf,Wf]
    cur = inp["cur_feats"].clone().requires_grad_(True)
    src = inp["src_feats"].clone().requires_grad_(True)
    out, planes, _ = m.build_cost_volume(cur_feats=cur, src_feats=src, src_extrinsics=inp["src_extrinsics"], src_poses=inp["src_poses"],
                                         src_Ks=inp["src_Ks"], cur_invK=inp["cur_invK"], min_depth=inp["min_depth"],
                                         max_depth=inp["max_depth"])
    wts = torch.randn(out.shape, generator=torch.Generator().manual_seed(77 + seed))
    (out * (wts * (~kink))).sum().backward()
    np.savez_compressed(
        os.path.join(ROOT, "tests", "golden", name + ".npz"), meta=np.array([seed, V, K, C, Hf, Wf, D, SRC_CH_STRIDE]),
        kink_packed=np.packbits(kink.numpy().reshape(-1)), kink_rows=np.array(int(kink.sum())),
        checksums=np.array([checksum(inp[k]) for k in ("cur_feats", "src_feats", "src_extrinsics", "src_Ks", "cur_invK")] +
                           [checksum(w) for w in mlp] + [checksum(wts)]),
        out=out.detach().numpy(), g_cur_sub=cur.grad[:, ::2].contiguous().numpy(), g_src_sub=src.grad[:, :, ::SRC_CH_STRIDE].contiguous().numpy(),
        **{f"g_mlp{i}": p.grad.numpy() for i, p in enumerate(params)})
    print(name, tuple(out.shape), "kink rows excluded:", int(kink.sum()), "of", kink.numel())


def ptf():
    name, seed, V, h, w = PTF_CASE
    fuse, pe, GRU = ref_loader.load_fuse_gaussians()
    inp = synth.ptf_inputs(seed, V, h, w)
    gru = GRU(); gru.load_state_dict(synth.gru_state(seed))
    gi = {k: inp[k].clone().requires_grad_(True) for k in ("densities", "weight_emb", "depths")}
    g_feats = inp["gaussians"][0].clone().requires_grad_(True)
    g_coords = inp["coords"][0].clone().requires_grad_(True)
    feats, coords, extr, depths = fuse(SimpleNamespace(gru=gru), [g_feats], [g_coords], gi["densities"], gi["weight_emb"], gi["depths"],
                                       inp["extrinsics"], inp["intrinsics"], inp["image_shape"])
    gen = torch.Generator().manual_seed(500 + seed)
    wF, wX = torch.randn(feats.shape, generator=gen), torch.randn(coords.shape, generator=gen)
    wE, wZ = torch.randn(extr.shape, generator=gen), torch.randn(depths.shape, generator=gen)
    ((feats * wF).sum() + (coords * wX).sum() + (extr * wE).sum() + (depths * wZ).sum()).backward()
    E_inv = torch.linalg.inv(inp["extrinsics"][0])            # extrinsic.inverse() of this container's LAPACK (:454)
    S = ROW_STRIDE
    np.savez_compressed(
        os.path.join(ROOT, "tests", "golden", name + ".npz"), meta=np.array([seed, V, h, w, S, feats.shape[1]]),
        checksums=np.array([checksum(inp["gaussians"][0]), checksum(inp["coords"][0]), checksum(inp["densities"]),
                            checksum(inp["weight_emb"]), checksum(inp["depths"]), checksum(inp["extrinsics"]), checksum(wF)]),
        E_inv=E_inv.numpy(), out_coords=coords[0].detach().numpy(), out_depths=depths[0].detach().numpy(),
        out_ext_sub=extr[0, ::S].detach().numpy(), out_ext_sum=np.array(checksum(extr.detach())),
        out_feats_sub=feats[0, ::S].detach().numpy(),
        g_in_feats_sub=g_feats.grad[0, :, ::S].contiguous().numpy(), g_in_coords=g_coords.grad[0, :, :, 0, 0, :].numpy(),
        g_in_dens=gi["densities"].grad[0, :, :, 0, 0].numpy(), g_in_wemb=gi["weight_emb"].grad[0, :, :, 0, 0].numpy(),
        g_in_depths=gi["depths"].grad.reshape(V, -1).numpy(), **{"g_gru." + k: v.grad.numpy() for k, v in gru.named_parameters()})
    print(name, "N_out", feats.shape[1], "of", V * h * w)


if __name__ == "__main__":
    cost_volume()
    ptf()
