"""Generates tests/golden/image_u8.npz by running the REFERENCE's own prep_image (/root/reference/src/misc/image_io.py:36-53,
function extracted unmodified by AST: the module itself imports matplotlib / torchvision) on seeded images.
Run in the build container:  python tests/golden/make_image_golden.py"""
import ast
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.golden import ref_loader  # noqa: E402


def load_prep_image():
    from einops import rearrange, repeat
    path = os.path.join(ref_loader.REF, "src", "misc", "image_io.py")
    tree = ast.parse(open(path).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "prep_image"]
    fn[0].returns = None
    for a in fn[0].args.args:
        a.annotation = None
    ns = {"torch": torch, "np": np, "rearrange": rearrange, "repeat": repeat}
    exec(compile(ast.Module(body=fn, type_ignores=[]), path, "exec"), ns)
    return ns["prep_image"]


def main():
    prep = load_prep_image()
    g = torch.Generator().manual_seed(0)
    cases = {"rgb": torch.rand((3, 37, 53), generator=g) * 1.4 - 0.2,          # values outside [0,1] get clipped
             "batch": torch.rand((3, 3, 20, 31), generator=g) * 1.2 - 0.1,
             "gray": torch.rand((24, 40), generator=g), "rgba": torch.rand((4, 16, 16), generator=g)}
    # exact quantisation edges: k/255 and its fp32 neighbours
    k = torch.arange(0, 256, dtype=torch.float32) / 255
    edges = torch.stack([k, torch.nextafter(k, torch.tensor(2.0)), torch.nextafter(k, torch.tensor(-1.0))]).reshape(1, 3, 256)
    cases["edges"] = edges.expand(3, 3, 256).contiguous()
    out = {}
    for name, img in cases.items():
        out["in_" + name] = img.numpy()
        out["out_" + name] = prep(img)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "image_u8.npz"), **out)
    print({k: v.shape for k, v in out.items() if k.startswith("out_")})


if __name__ == "__main__":
    main()
