"""Generates tests/golden/ply_*.npz by running the REFERENCE's own export_ply (src/model/ply_export.py:26-92) on seeded
Gaussians; the vertex table it passes to plyfile (absent here: stubbed, see ref_loader.load_export_ply) is recorded.
Run in the build container (scipy 1.18)."""
import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from freesplat_b200 import synth  # noqa: E402
from tests.golden import ref_loader  # noqa: E402


def main():
    export_ply, cap = ref_loader.load_export_ply()
    for name, (seed, N) in {"ply_a": (0, 900), "ply_b": (1, 257)}.items():
        g = torch.Generator().manual_seed(seed)
        ext = synth.camera_path(3)[seed + 1].clone()
        means = torch.randn((N, 3), generator=g) * torch.tensor([2.0, 0.7, 3.0]) + torch.tensor([0.3, -0.2, 2.5])
        scales = 0.005 + 0.2 * torch.rand((N, 3), generator=g)
        rot = torch.randn((N, 4), generator=g)
        rot = rot / rot.norm(dim=-1, keepdim=True)
        sh = torch.randn((N, 3, 9), generator=g)
        op = torch.rand((N,), generator=g)
        with tempfile.TemporaryDirectory() as td:
            export_ply(ext, means, scales, rot, sh, op, Path(td) / "x" / "a.ply")
        el = cap["elements"]
        table = np.stack([el[k] for k in el.dtype.names], axis=1).astype(np.float32)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), meta=np.array([seed, N]), ext=ext.numpy(),
                            means=means.numpy(), scales=scales.numpy(), rotations=rot.numpy(), harmonics=sh.numpy(),
                            opacities=op.numpy(), table=table, names=np.array(el.dtype.names))
        print(name, table.shape)


if __name__ == "__main__":
    main()
