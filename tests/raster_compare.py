"""Comparison of the CUDA rasterizer state with the CPU oracle (used by the -m gpu tests and by
tools/gpu_check.py, which dumps the same numbers as JSON for offline reading)."""
from __future__ import annotations

import numpy as np
import torch

from freesplat_b200 import rasterizer
from oracle import raster as oracle
from tests.helpers import view_inputs


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def run_cuda(scene, bg=(0.0, 0.0, 0.0), device="cuda:0", scale_invariant=True, capacity=None):
    from freesplat_b200 import decoder
    sc = scene.to(device)
    V = scene.extrinsics.shape[0]
    # camera records are computed on the CPU and copied, so that the oracle and the kernels see
    # bit-identical inputs (torch's CPU and CUDA inverse()/matmul differ in the last bits)
    bgc = torch.tensor(bg, dtype=torch.float32)[None].expand(V, 3)
    views, _ = decoder.camera_records(scene.extrinsics, scene.intrinsics, scene.near, scene.far, bgc, scale_invariant)
    views = views.to(device)
    row, col = torch.triu_indices(3, 3)
    d_sh = sc.harmonics.shape[-1]
    st = rasterizer.raster_forward_raw(
        sc.means, sc.opacities, views, sc.image_shape[0], sc.image_shape[1],
        shs=sc.harmonics.transpose(1, 2).contiguous(), cov3D_precomp=sc.covariances[:, row, col].contiguous(),
        sh_degree=int(round(d_sh ** 0.5)) - 1, capacity=capacity, debug_buffers=True)
    return st, views


def compare_forward(scene, st, bg=(0.0, 0.0, 0.0), scale_invariant=True) -> dict:
    """Per-view comparison; returns a dict of metrics (all 'exact_*' must be True, counts 0)."""
    out = {}
    V, P, H, W = st.V, st.P, st.H, st.W
    gx, gy = (W + 15) // 16, (H + 15) // 16
    T = gx * gy
    rec = st.rec.cpu().numpy(); radii = st.radii.cpu().numpy(); tt = st.tiles_touched.cpu().numpy().astype(np.uint32)
    ranges = st.ranges.cpu().numpy().astype(np.uint32).reshape(V, T, 2)
    pl = st.point_list.cpu().numpy().astype(np.uint32)
    keys = st.keybuf.cpu().numpy().view(np.uint64) if st.keybuf is not None else None
    color = st.color.cpu().numpy(); depth = st.depth.cpu().numpy(); fT = st.final_T.cpu().numpy()
    nc = st.n_contrib.cpu().numpy().astype(np.uint32)
    clamped = st.clamped.cpu().numpy()
    Rtot = st.num_rendered()
    out["R_total"] = Rtot
    out["overflow"] = st.overflowed()
    per_view = []
    base = 0
    for v in range(V):
        inp, _ = view_inputs(scene, v, scale_invariant=scale_invariant, bg=bg)
        o = oracle.forward(**inp)
        vis = o.radii > 0
        m = {"view": v, "R": o.R, "visible": int(vis.sum())}
        m["exact_radii"] = bool(np.array_equal(radii[v], o.radii))
        m["exact_tiles_touched"] = bool(np.array_equal(tt[v], o.tiles_touched))
        m["exact_xy"] = bool(np.array_equal(_bits(rec[v][vis][:, 0:2]), _bits(o.xy[vis])))
        m["exact_conic"] = bool(np.array_equal(_bits(rec[v][vis][:, 2:5]), _bits(o.conic_opacity[vis][:, 0:3])))
        m["exact_depths"] = bool(np.array_equal(_bits(rec[v][vis][:, 9]), _bits(o.depths[vis])))
        m["exact_rgb"] = bool(np.array_equal(_bits(rec[v][vis][:, 6:9]), _bits(o.rgb[vis])))
        cm = (o.clamped[:, 0] | (o.clamped[:, 1] << 1) | (o.clamped[:, 2] << 2)).astype(np.uint8)
        m["exact_clamped"] = bool(np.array_equal(clamped[v][vis], cm[vis]))
        r_rel = ranges[v].astype(np.int64) - base
        m["exact_ranges"] = bool(np.array_equal(np.where((o.ranges[:, 1] > o.ranges[:, 0])[:, None], r_rel, 0),
                                                o.ranges.astype(np.int64)))
        m["exact_point_list"] = bool(np.array_equal(pl[base:base + o.R], o.point_list))
        if keys is not None:
            # reference key = tile<<32 | depth_bits ; ours = depth_bits<<32 | gaussian, inside the tile's range
            tile_of = np.repeat(np.arange(T, dtype=np.uint64), (o.ranges[:, 1] - o.ranges[:, 0]).astype(np.int64))
            ref_keys = (tile_of << np.uint64(32)) | (keys[base:base + o.R] >> np.uint64(32))
            m["exact_keys"] = bool(np.array_equal(ref_keys, o.keys))
        ok_c = np.isclose(color[v], o.color, rtol=1e-4, atol=1e-5)
        ok_d = np.isclose(depth[v], o.depth, rtol=1e-4, atol=1e-5)
        ok_t = np.isclose(fT[v], o.final_T, rtol=1e-4, atol=1e-6)
        m["color_bad_px"] = int((~ok_c).any(axis=0).sum()); m["color_max_abs"] = float(np.abs(color[v] - o.color).max())
        m["depth_bad_px"] = int((~ok_d).sum()); m["depth_max_abs"] = float(np.abs(depth[v] - o.depth).max())
        m["finalT_bad_px"] = int((~ok_t).sum())
        m["n_contrib_mismatch_px"] = int((nc[v] != o.n_contrib).sum())
        mse = float(((color[v].astype(np.float64) - o.color) ** 2).mean())
        m["psnr_vs_oracle_db"] = float(10 * np.log10(1.0 / max(mse, 1e-20)))
        per_view.append(m)
        base += o.R
    out["views"] = per_view
    out["R_matches"] = bool(base == Rtot)
    return out


def forward_ok(metrics: dict, max_bad_frac=2e-5, n_pixels=None) -> list:
    """List of human-readable failures (empty = parity green)."""
    fails = []
    if metrics["overflow"] or not metrics["R_matches"]:
        fails.append(f"R mismatch/overflow: {metrics['R_total']}")
    for m in metrics["views"]:
        for k, val in m.items():
            if k.startswith("exact_") and not val:
                fails.append(f"view {m['view']}: {k} is not bit-exact")
        lim = max(2, int(max_bad_frac * (n_pixels or 1)))
        for k in ("color_bad_px", "depth_bad_px", "finalT_bad_px", "n_contrib_mismatch_px"):
            if m[k] > lim:
                fails.append(f"view {m['view']}: {k}={m[k]} > {lim}")
    return fails


def compare_backward(scene, st, views, dL_dcolor, dL_ddepth=None, bg=(0.0, 0.0, 0.0), scale_invariant=True, dL_dalpha=None) -> dict:
    """CUDA backward (summed over views) vs the sum of per-view oracle backwards, mapped back through
    the scene rescale (means*s, cov*s^2) the adapter applies."""
    dev = st.rec.device
    sc = scene.to(dev)
    row, col = torch.triu_indices(3, 3)
    shs = sc.harmonics.transpose(1, 2).contiguous()
    g = rasterizer.raster_backward_raw(st, sc.means, sc.opacities, dL_dcolor.to(dev), shs=shs,
                                       cov3D_precomp=sc.covariances[:, row, col].contiguous(),
                                       dL_ddepth=None if dL_ddepth is None else dL_ddepth.to(dev),
                                       dL_dalpha=None if dL_dalpha is None else dL_dalpha.to(dev))
    P = st.P
    want = dict(means3D=np.zeros((P, 3)), cov3D=np.zeros((P, 6)), shs=np.zeros(tuple(shs.shape)), opacities=np.zeros((P, 1)))
    want2d = []
    for v in range(st.V):
        inp, vrec = view_inputs(scene, v, scale_invariant=scale_invariant, bg=bg)
        s = float(vrec[v][40])
        o = oracle.forward(**inp)
        go = oracle.backward(o, tanfovx=inp["tanfovx"], tanfovy=inp["tanfovy"], bg=inp["bg"], viewmatrix=inp["viewmatrix"],
                             projmatrix=inp["projmatrix"], campos=inp["campos"], means3D=inp["means3D"],
                             dL_dcolor=dL_dcolor[v].cpu().numpy(),
                             dL_ddepth=None if dL_ddepth is None else dL_ddepth[v].cpu().numpy(),
                             dL_dalpha=None if dL_dalpha is None else dL_dalpha[v].cpu().numpy(),
                             shs=inp["shs"], sh_degree=inp["sh_degree"])
        want["means3D"] += go["means3D"].astype(np.float64) * s
        want["cov3D"] += go["cov3D"].astype(np.float64) * s * s
        want["shs"] += go["shs"]; want["opacities"] += go["opacities"]
        want2d.append(go["means2D"])
    from tests.helpers import grad_report
    out = {}
    got = dict(means3D=g["means3D"], cov3D=g["cov3D"], shs=g["shs"], opacities=g["opacities"])
    for k in want:
        out[k] = grad_report(got[k].cpu().numpy(), want[k])
    out["means2D"] = grad_report(g["means2D"].cpu().numpy(), np.stack(want2d))
    return out


def backward_ok(metrics: dict) -> list:
    """Element-wise criterion of tests.helpers.grad_report: |got - want| <= 1e-4 |want| + 1e-5 max|want|, at most 5e-4 of the
    elements outside (counted in the report: a pixel whose alpha >= 1/255 or T < 1e-4 decision flips between MUFU.EX2 and libm
    changes the gradient of the Gaussians that cover it), nothing further than 2e-4 of the maximum."""
    return [f"grad {k}: {m}" for k, m in metrics.items() if not m["ok"]]
