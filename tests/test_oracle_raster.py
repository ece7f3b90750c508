"""CPU: oracle/raster_oracle.c against the independent dense PyTorch renderer + autograd.

The reference holds no golden vector for the rasterizer (third-party, absent: PARITY UNPINNED);
this is the strongest pin available here: two independent restatements of SURVEY Appendix A
(one tiled C with hand-derived gradients, one dense torch with autograd) must agree."""
import numpy as np
import pytest
import torch

from freesplat_b200 import synth
from oracle import raster as oracle
from tests import dense_torch_raster as dense
from tests.helpers import view_inputs


def _small_scene(seed, P=160, h=48, w=64, sigma=(0.7, 6.0)):
    return synth.random_scene(seed=seed, h=h, w=w, P=P, n_target=1, sh_degree=2, sigma_px=sigma)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_forward_matches_dense(seed):
    sc = _small_scene(seed)
    inp, _ = view_inputs(sc, 0, bg=(0.1, 0.2, 0.3))
    st = oracle.forward(**inp)
    kw = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in inp.items()}
    cov6 = kw.pop("cov3D_precomp")
    color, depth, T, radii = dense.render(cov6=cov6, **kw)
    assert st.R > 0
    np.testing.assert_array_equal(st.radii, radii.numpy().astype(np.int32))
    # a handful of pixels may flip a threshold test between fp32 and fp64
    ok = np.isclose(st.color, color.numpy(), rtol=1e-3, atol=2e-4)
    assert ok.mean() > 0.999, ok.mean()
    assert np.isclose(st.depth, depth.numpy(), rtol=1e-3, atol=1e-3).mean() > 0.999
    assert np.isclose(st.final_T, T.numpy(), rtol=1e-3, atol=2e-4).mean() > 0.999


def test_integer_buffers_consistent():
    sc = _small_scene(3, P=400, h=64, w=80)
    inp, _ = view_inputs(sc, 0)
    st = oracle.forward(**inp)
    assert st.R == int(st.tiles_touched.sum()) == int(st.offsets[-1])
    assert np.all(np.diff(st.keys.astype(np.uint64)) >= 0)              # sorted
    tiles = (st.keys >> np.uint64(32)).astype(np.int64)
    gx, gy = (80 + 15) // 16, (64 + 15) // 16
    for t in range(gx * gy):
        a, b = st.ranges[t]
        assert np.all(tiles[a:b] == t)
        # inside a tile: ascending depth, ties by Gaussian index (stable sort of an index-ordered emission)
        d = st.depths[st.point_list[a:b]]
        assert np.all(np.diff(d) >= 0)
        same = np.diff(d) == 0
        assert np.all(np.diff(st.point_list[a:b].astype(np.int64))[same] > 0)
    assert int((st.ranges[:, 1] - st.ranges[:, 0]).sum()) == st.R


@pytest.mark.parametrize("seed,use_sr", [(0, False), (1, False), (2, True)])
def test_backward_matches_autograd(seed, use_sr):
    sc = _small_scene(seed, P=120, h=40, w=48, sigma=(0.8, 5.0))
    inp, _ = view_inputs(sc, 0, bg=(0.3, 0.1, 0.2))
    g = torch.Generator().manual_seed(100 + seed)
    H, W = inp["H"], inp["W"]
    dL_dcolor = torch.randn((3, H, W), generator=g, dtype=torch.float64)
    dL_ddepth = torch.randn((H, W), generator=g, dtype=torch.float64) * 0.3
    dL_dalpha = torch.randn((H, W), generator=g, dtype=torch.float64) * 0.5
    scale = float(1.0 / sc.near[0])
    if use_sr:
        inp = dict(inp); inp.pop("cov3D_precomp")
        inp["scales"] = (sc.scales * scale).numpy(); inp["rotations"] = sc.rotations.numpy()
    st = oracle.forward(**inp)
    for with_depth in (False, True):
        gr = oracle.backward(st, tanfovx=inp["tanfovx"], tanfovy=inp["tanfovy"], bg=inp["bg"],
                             viewmatrix=inp["viewmatrix"], projmatrix=inp["projmatrix"], campos=inp["campos"],
                             means3D=inp["means3D"], dL_dcolor=dL_dcolor.numpy(),
                             dL_ddepth=dL_ddepth.numpy() if with_depth else None,
                             dL_dalpha=dL_dalpha.numpy() if with_depth else None, shs=inp["shs"],
                             scales=inp.get("scales"), rotations=inp.get("rotations"), sh_degree=inp["sh_degree"])
        t = {k: torch.tensor(inp[k], dtype=torch.float64, requires_grad=True)
             for k in ("means3D", "opacities", "shs") }
        if use_sr:
            t["scales"] = torch.tensor(inp["scales"], dtype=torch.float64, requires_grad=True)
            t["rotations"] = torch.tensor(inp["rotations"], dtype=torch.float64, requires_grad=True)
            extra = dict(scales=t["scales"], rotations=t["rotations"])
        else:
            t["cov6"] = torch.tensor(inp["cov3D_precomp"], dtype=torch.float64, requires_grad=True)
            extra = dict(cov6=t["cov6"])
        color, depth, fT, _ = dense.render(H=H, W=W, tanfovx=inp["tanfovx"], tanfovy=inp["tanfovy"], bg=inp["bg"],
                                          viewmatrix=inp["viewmatrix"], projmatrix=inp["projmatrix"],
                                          campos=inp["campos"], means3D=t["means3D"], opacities=t["opacities"],
                                          shs=t["shs"], sh_degree=inp["sh_degree"], **extra)
        loss = (color * dL_dcolor).sum()
        if with_depth:
            loss = loss + (depth * dL_ddepth).sum() + ((1.0 - fT) * dL_dalpha).sum()   # 4th output of the op: 1 - final_T
        loss.backward()
        pairs = [("means3D", gr["means3D"], t["means3D"].grad), ("opacities", gr["opacities"][:, 0], t["opacities"].grad),
                 ("shs", gr["shs"], t["shs"].grad)]
        if use_sr:
            pairs += [("scales", gr["scales"], t["scales"].grad), ("rotations", gr["rotations"], t["rotations"].grad)]
        else:
            pairs += [("cov3D", gr["cov3D"], t["cov6"].grad)]
        for name, got, want in pairs:
            want = want.numpy()
            scale_ = np.abs(want).max() + 1e-12
            err = np.abs(got - want) / scale_
            # fp32 oracle vs fp64 autograd; a few Gaussians see a flipped threshold pixel
            assert np.quantile(err, 0.98) < 2e-3, (name, with_depth, np.quantile(err, 0.98), err.max())
