"""Generates tests/golden/backproject_*.npz by running the REFERENCE's own Create_from_depth_map.project
(gaussian_adapter.py:48-68, driven exactly as GaussianAdapter.forward(fusion=True) does at :175-189) on seeded depth
maps.  Run in the build container."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from freesplat_b200 import synth  # noqa: E402
from tests.golden import ref_loader  # noqa: E402


def main():
    ref_loader.load_gaussian_adapter()
    ga = sys.modules["refpkg.src.model.encoder.common.gaussian_adapter"]
    for name, (seed, V, h, w) in {"backproject_a": (0, 3, 48, 64), "backproject_b": (1, 2, 30, 52)}.items():
        g = torch.Generator().manual_seed(seed)
        K = synth.intrinsics(1)[0].clone()
        K[0, 2] += 0.013 * seed; K[1, 1] *= 1.0 + 0.07 * seed
        c2w = synth.camera_path(V + 1)[1:].clone()
        depths = 0.5 + 6 * torch.rand((V, h, w), generator=g)
        intrinsic = K.clone().view(3, 3)
        intrinsic[:1, :] *= w
        intrinsic[1:2, :] *= h
        conv = ga.Create_from_depth_map(intrinsic, height=h, width=w, depth_trunc=15)
        means = torch.stack([conv.project(depths[j].view(h, w), c2w[j].view(4, 4)) for j in range(V)])
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), meta=np.array([seed, V, h, w]), K=K.numpy(),
                            c2w=c2w.numpy(), depths=depths.numpy(), means=means.numpy())
        print(name, means.shape)


if __name__ == "__main__":
    main()
