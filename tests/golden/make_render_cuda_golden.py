"""Generates tests/golden/render_cuda_*.npz by running the REFERENCE's own raster adapter and decoder
(/root/reference/src/model/decoder/cuda_splatting.py:47-132 `render_cuda`, decoder_splatting_cuda.py:35-75
`DecoderSplattingCUDA.forward`, loaded unmodified by ref_loader.load_decoder) on seeded scenes.  The third-party extension
they import is absent, so a stand-in module `diff_gaussian_rasterization_depth` backed by the CPU oracle
(oracle/raster_oracle.c) takes its place and RECORDS every call the adapter makes: settings + tensors exactly as the
reference hands them over.  The fixture therefore pins
  * what the reference adapter passes to the op per view (the drop-in GaussianRasterizer is fed these on the GPU),
  * the images / depths `render_cuda` and `DecoderSplattingCUDA.forward` return for them (incl. depth / 2),
  * the gradients w.r.t. the Gaussians that flow back through the adapter's own rescale / rearrange / triu ops.
Run in the build container:  python tests/golden/make_render_cuda_golden.py"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from freesplat_b200 import synth  # noqa: E402
from oracle import raster as oracle  # noqa: E402
from tests.golden import ref_loader  # noqa: E402

CASES = {
    # name: (seed, b, v, h, w, keep)
    "render_cuda_a": (0, 1, 3, 48, 64, 2500),
    "render_cuda_b2": (1, 2, 2, 40, 56, 1500),
}
CALLS = []


class _OracleOp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, shs, opacities, cov3D, kw):
        st = oracle.forward(means3D=means3D.detach().numpy(), opacities=opacities.detach().numpy(), shs=shs.detach().numpy(),
                            cov3D_precomp=cov3D.detach().numpy(), **kw)
        ctx.st, ctx.kw = st, kw
        ctx.save_for_backward(means3D, shs)
        t = torch.from_numpy
        return t(st.color), t(st.radii), t(st.depth), t(1.0 - st.final_T)

    @staticmethod
    def backward(ctx, g_color, g_radii, g_depth, g_alpha):
        means3D, shs = ctx.saved_tensors
        kw = ctx.kw
        g = oracle.backward(ctx.st, tanfovx=kw["tanfovx"], tanfovy=kw["tanfovy"], bg=kw["bg"], viewmatrix=kw["viewmatrix"],
                            projmatrix=kw["projmatrix"], campos=kw["campos"], means3D=means3D.detach().numpy(),
                            dL_dcolor=g_color.numpy(), dL_ddepth=None if g_depth is None else g_depth.numpy(),
                            shs=shs.detach().numpy(), sh_degree=kw["sh_degree"])
        t = torch.from_numpy
        return t(g["means3D"]), t(g["means2D"]), t(g["shs"]), t(g["opacities"]), t(g["cov3D"]), None


def _stand_in():
    from freesplat_b200.rasterizer import GaussianRasterizationSettings      # the NamedTuple of the operator surface

    class GaussianRasterizer(torch.nn.Module):
        def __init__(self, raster_settings):
            super().__init__()
            self.raster_settings = raster_settings

        def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                    cov3D_precomp=None):
            rs = self.raster_settings
            assert colors_precomp is None and scales is None and rotations is None
            kw = dict(H=int(rs.image_height), W=int(rs.image_width), tanfovx=float(rs.tanfovx), tanfovy=float(rs.tanfovy),
                      bg=rs.bg.detach().numpy().copy(), viewmatrix=rs.viewmatrix.detach().numpy().reshape(16).copy(),
                      projmatrix=rs.projmatrix.detach().numpy().reshape(16).copy(), campos=rs.campos.detach().numpy().copy(),
                      sh_degree=int(rs.sh_degree), scale_modifier=float(rs.scale_modifier))
            CALLS.append(dict(kw, means3D=means3D.detach().numpy().copy(), opacities=opacities.detach().numpy().copy(),
                              shs=shs.detach().numpy().copy(), cov3D_precomp=cov3D_precomp.detach().numpy().copy(),
                              means2D_shape=tuple(means2D.shape), prefiltered=bool(rs.prefiltered), debug=bool(rs.debug)))
            return _OracleOp.apply(means3D, means2D, shs, opacities, cov3D_precomp, kw)
    m = types.ModuleType("diff_gaussian_rasterization_depth")
    m.GaussianRasterizationSettings = GaussianRasterizationSettings
    m.GaussianRasterizer = GaussianRasterizer
    return m


def main():
    render_cuda, Decoder, Gaussians, DatasetCfg = ref_loader.load_decoder(_stand_in())
    for name, (seed, b, v, h, w, keep) in CASES.items():
        scenes = [synth.pixel_aligned_scene(seed=seed + 10 * i, h=h, w=w, n_context=2, n_target=v, keep=keep) for i in range(b)]
        G = min(s.means.shape[0] for s in scenes)
        stk = lambda f: torch.stack([f(s)[:G] for s in scenes])
        means, cov, sh, op = (stk(lambda s: s.means), stk(lambda s: s.covariances), stk(lambda s: s.harmonics),
                              stk(lambda s: s.opacities))
        ext = torch.stack([s.extrinsics for s in scenes]); K = torch.stack([s.intrinsics for s in scenes])
        near = torch.stack([s.near for s in scenes]); far = torch.stack([s.far for s in scenes])
        bgc = (0.1, 0.3, 0.2)
        leaves = [t.clone().requires_grad_(True) for t in (means, cov, sh, op)]
        dec = Decoder(None, DatasetCfg(list(bgc)))
        CALLS.clear()
        out = dec.forward(Gaussians(*leaves), ext, K, near, far, (h, w), depth_mode="depth")
        gen = torch.Generator().manual_seed(900 + seed)
        wC = torch.randn(out.color.shape, generator=gen)
        (out.color * wC).sum().backward()
        calls = {}
        for i, c in enumerate(CALLS):
            for k, val in c.items():
                calls[f"call{i}_{k}"] = np.asarray(val)
        np.savez_compressed(
            os.path.join(ROOT, "tests", "golden", name + ".npz"), meta=np.array([seed, b, v, h, w, G]), bg=np.array(bgc, np.float32),
            means=means.numpy(), covariances=cov.numpy(), harmonics=sh.numpy(), opacities=op.numpy(), extrinsics=ext.numpy(),
            intrinsics=K.numpy(), near=near.numpy(), far=far.numpy(), color=out.color.detach().numpy(),
            depth=out.depth.detach().numpy(), wC=wC.numpy(), n_calls=np.array(len(CALLS)),
            g_means=leaves[0].grad.numpy(), g_cov=leaves[1].grad.numpy(), g_sh=leaves[2].grad.numpy(), g_op=leaves[3].grad.numpy(),
            **calls)
        print(name, "calls", len(CALLS), "G", G, tuple(out.color.shape), float(out.color.detach().mean()))


if __name__ == "__main__":
    main()
