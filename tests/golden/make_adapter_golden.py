"""Generates tests/golden/adapter_*.npz by running the REFERENCE's own GaussianAdapter.forward (fusion=False, coords given:
the call at encoder_freesplat.py:376-386) on seeded inputs.  Run in the build container."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from freesplat_b200 import synth  # noqa: E402
from tests.golden import ref_loader  # noqa: E402


def main():
    GA, Cfg = ref_loader.load_gaussian_adapter()
    for name, (seed, N, h, w) in {"adapter_a": (0, 700, 480, 640), "adapter_b": (1, 333, 384, 512)}.items():
        g = torch.Generator().manual_seed(seed)
        ad = GA(Cfg(gaussian_scale_min=0.5, gaussian_scale_max=15.0, sh_degree=2))
        raw = torch.randn((N, 34), generator=g)
        depths = 0.5 + 4 * torch.rand((N,), generator=g)
        opac = torch.rand((N,), generator=g)
        coords = torch.randn((N, 3), generator=g) * 2
        # per-Gaussian (averaged, hence not exactly rigid) camera-to-world matrices, as PTF produces them
        ext = synth.camera_path(4)[torch.randint(0, 4, (N,), generator=g)].clone()
        ext[:, :3, :] += 0.01 * torch.randn((N, 3, 4), generator=g)
        K = synth.intrinsics(1)[0]
        leaves = [t.requires_grad_(True) for t in (raw, depths, opac, coords, ext)]      # gradients by the reference's own autograd
        out = ad.forward(ext[None, None, :, None, None], K[None, None, None, None, None].expand(1, 1, N, 1, 1, 3, 3),
                         torch.zeros(1, 1, N, 1, 1, 2), depths[None, None, :, None, None], opac[None, None, :, None, None],
                         raw[None, None, :, None, None, :], (h, w), coords=coords[None, None, :, None, None, :])
        names = ("means", "covariances", "harmonics", "opacities", "scales", "rotations")
        wts = {k: torch.randn(getattr(out, k)[0, 0, :, 0, 0].shape, generator=g) for k in names}
        sum((getattr(out, k)[0, 0, :, 0, 0] * wts[k]).sum() for k in names).backward()
        grads = dict(g_raw=raw.grad.numpy(), g_depths=depths.grad.numpy(), g_opac=opac.grad.numpy(), g_coords=coords.grad.numpy(),
                     g_ext=ext.grad.numpy(), **{"w_" + k: v.numpy() for k, v in wts.items()})
        raw, depths, opac, coords, ext = (t.detach() for t in leaves)
        sq = lambda t: t[0, 0, :, 0, 0].detach().numpy()
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), meta=np.array([seed, N, h, w]), raw=raw.numpy(),
                            depths=depths.numpy(), opac=opac.numpy(), coords=coords.numpy(), ext=ext.numpy(), K=K.numpy(),
                            means=sq(out.means), covariances=sq(out.covariances), harmonics=sq(out.harmonics),
                            opacities=sq(out.opacities), scales=sq(out.scales), rotations=sq(out.rotations), **grads)
        print(name, N)


if __name__ == "__main__":
    main()
