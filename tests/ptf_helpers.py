"""Shared by the PTF tests: unpack the encoder-shaped inputs into flat per-view arrays."""
import numpy as np
import torch


def flat_inputs(inp):
    """dict from synth.ptf_inputs / golden npz -> (feats [V,HW,F], coords [V,HW,3], dens [V,HW], wemb [V,HW],
    depths [V,HW], ext [V,4,4], K [V,3,3], (h,w))."""
    g = lambda k: inp[k].numpy() if isinstance(inp[k], torch.Tensor) else np.asarray(inp[k])
    feats = (inp["gaussians"][0] if "gaussians" in inp else inp["in_feats"])
    feats = feats.numpy() if isinstance(feats, torch.Tensor) else np.asarray(feats)
    coords = (inp["coords"][0] if "coords" in inp else inp["in_coords"])
    coords = coords.numpy() if isinstance(coords, torch.Tensor) else np.asarray(coords)
    dens = g("densities") if "densities" in inp else g("in_dens")
    wemb = g("weight_emb") if "weight_emb" in inp else g("in_wemb")
    depths = g("depths") if "depths" in inp else g("in_depths")
    ext = g("extrinsics") if "extrinsics" in inp else g("in_ext")
    K = g("intrinsics") if "intrinsics" in inp else g("in_K")
    V = feats.shape[1]
    h, w = depths.shape[-2:]
    return (feats[0], coords[0, :, :, 0, 0, :], dens[0, :, :, 0, 0], wemb[0, :, :, 0, 0], depths.reshape(V, -1),
            ext[0], K[0], (h, w))


def torch_inverses(ext):
    """extrinsic.inverse() exactly as the reference computes it (torch fp32 on the CPU)."""
    return torch.linalg.inv(torch.from_numpy(np.ascontiguousarray(ext)).float()).numpy()
