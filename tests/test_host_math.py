"""CPU: the product's per-Gaussian math (raster_math.cuh compiled for the host) must make the same
integer decisions, bit for bit, as the oracle -- caught here before any GPU time is spent."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from freesplat_b200 import synth
from oracle import raster as oracle
from tests.helpers import view_inputs

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def harness():
    so = os.path.join(HERE, "host_harness", "libfs_host_harness.so")
    src = os.path.join(HERE, "host_harness", "harness.cpp")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([gxx, "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                           "-o", so, src])
    return C.CDLL(so)


@pytest.mark.parametrize("seed,kind", [(0, "random"), (1, "random"), (2, "pixel")])
def test_projection_bit_exact(harness, seed, kind):
    if kind == "random":
        sc = synth.random_scene(seed=seed, h=256, w=256, P=10000)
    else:
        sc = synth.pixel_aligned_scene(seed=seed, h=96, w=128, n_context=2, n_target=2, keep=None)
    inp, _ = view_inputs(sc, 0)
    st = oracle.forward(render=False, **inp)
    P = st.P
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    radii = np.zeros(P, np.int32); rect = np.zeros((P, 4), np.int32); xy = np.zeros((P, 2), np.float32)
    conic = np.zeros((P, 3), np.float32); depth = np.zeros(P, np.float32); rgb = np.zeros((P, 3), np.float32)
    clamp = np.zeros(P, np.int32); ext = np.zeros((P, 2), np.float32)
    means = np.ascontiguousarray(inp["means3D"], np.float32); cov = np.ascontiguousarray(inp["cov3D_precomp"], np.float32)
    shs = np.ascontiguousarray(inp["shs"], np.float32); op = np.ascontiguousarray(inp["opacities"], np.float32)
    vm = np.ascontiguousarray(inp["viewmatrix"], np.float32); pm = np.ascontiguousarray(inp["projmatrix"], np.float32)
    cp = np.ascontiguousarray(inp["campos"], np.float32)
    harness.fsh_project(C.c_int(P), C.c_int(inp["H"]), C.c_int(inp["W"]), C.c_float(inp["tanfovx"]),
                        C.c_float(inp["tanfovy"]), p(means), p(cov), p(vm), p(pm), p(op), C.c_int(inp["sh_degree"]),
                        C.c_int(shs.shape[1]), p(shs), p(cp), p(radii), p(rect), p(xy), p(conic), p(depth), p(rgb),
                        p(clamp), p(ext))
    vis = st.radii > 0
    assert vis.sum() > 100
    np.testing.assert_array_equal(radii, st.radii)
    np.testing.assert_array_equal(((rect[:, 2] - rect[:, 0]) * (rect[:, 3] - rect[:, 1]))[vis], st.tiles_touched[vis])
    bits = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)
    np.testing.assert_array_equal(bits(xy[vis]), bits(st.xy[vis]))
    np.testing.assert_array_equal(bits(depth[vis]), bits(st.depths[vis]))
    np.testing.assert_array_equal(bits(conic[vis]), bits(st.conic_opacity[vis, :3]))
    np.testing.assert_array_equal(bits(rgb[vis]), bits(st.rgb[vis]))
    cm = st.clamped[:, 0] | (st.clamped[:, 1] << 1) | (st.clamped[:, 2] << 2)
    np.testing.assert_array_equal(clamp[vis], cm[vis])
    # the culling extent must contain every pixel offset whose alpha can reach 1/255
    co = st.conic_opacity[vis]; e = ext[vis]
    tau = np.log(255.0 * np.maximum(co[:, 3].astype(np.float64), 1e-30))
    ok = tau >= 0
    det = co[:, 0].astype(np.float64) * co[:, 2] - co[:, 1].astype(np.float64) ** 2
    hx_true = np.sqrt(np.maximum(2 * tau * co[:, 2] / det, 0)); hy_true = np.sqrt(np.maximum(2 * tau * co[:, 0] / det, 0))
    assert np.all(e[ok, 0] >= hx_true[ok]) and np.all(e[ok, 1] >= hy_true[ok])
