"""GPU parity tests of the rasterizer: CUDA path (through the C ABI) vs the CPU oracle on identical
seeded inputs.  Integer buffers bit-exact; floats within 1e-4 relative (tolerance stated in
raster_compare.py: rtol 1e-4, atol 1e-5; a pixel whose T<1e-4 / alpha<1/255 test flips because
MUFU.EX2 and libm expf differ in the last bits may exceed it -- bounded to 2e-5 of the pixels)."""
import numpy as np
import pytest
import torch

from freesplat_b200 import synth
from tests import raster_compare as rc

pytestmark = pytest.mark.gpu


def _check(scene, bg=(0.0, 0.0, 0.0)):
    st, views = rc.run_cuda(scene, bg=bg)
    m = rc.compare_forward(scene, st, bg=bg)
    fails = rc.forward_ok(m, n_pixels=st.H * st.W)
    assert not fails, (fails, m)
    return st, views, m


def test_config1_random_256():
    """BASELINE config 1: 256x256, 10k random Gaussians."""
    _check(synth.random_scene(seed=0, h=256, w=256, P=10000), bg=(0.2, 0.4, 0.6))


def test_pixel_aligned_batched_views():
    _check(synth.pixel_aligned_scene(seed=1, h=120, w=160, n_context=2, n_target=3, keep=None))


def test_ragged_image_and_empty():
    # H, W not multiples of 16; then a scene with every Gaussian behind the camera (R = 0)
    sc = synth.random_scene(seed=3, h=100, w=77, P=3000)
    _check(sc)
    sc.means[:, :] = sc.means[:, :] * 0 + torch.tensor([0.0, 0.0, -5.0])
    st, _, m = _check(sc, bg=(0.5, 0.25, 0.125))
    assert m["R_total"] == 0
    assert torch.allclose(st.color[0, 0], torch.full_like(st.color[0, 0], 0.5))


def test_many_views_tile_scan_chunks():
    """18 views x 300 tiles = 5400 tile counters: more than one 4096-counter pass of the scan / tile-order kernel."""
    sc = synth.pixel_aligned_scene(seed=2, h=240, w=320, n_context=2, n_target=18, keep=None)
    st, _, m = _check(sc)
    assert st.V == 18 and len(m["views"]) == 18


def test_crowded_tile_global_sort_path():
    # > 4096 instances in a tile: exercises the global-memory sort path and multi-batch rendering
    sc = synth.random_scene(seed=4, h=64, w=64, P=30000, sigma_px=(0.5, 2.0))
    _check(sc)


def test_packed_two_pixel_render_kernel_is_bit_identical(monkeypatch):
    """FS_STAGE_RENDER_PACKED (render_fwd2_kernel: two pixels per lane, FFMA2 / FMUL2, half-warp-independent instance lists)
    must reproduce the default kernel bit for bit: images, final_T, n_contrib, sorted keys; incl. a crowded tile (> 2048 keys:
    its global-memory sort path) and a ragged image."""
    from freesplat_b200 import rasterizer
    for sc in (synth.pixel_aligned_scene(seed=1, h=120, w=160, n_context=2, n_target=3, keep=None),
               synth.random_scene(seed=4, h=64, w=64, P=30000, sigma_px=(0.5, 2.0)), synth.random_scene(seed=3, h=100, w=77, P=3000)):
        monkeypatch.setattr(rasterizer, "RENDER_PACKED", False)
        a, _ = rc.run_cuda(sc, bg=(0.1, 0.2, 0.3))
        monkeypatch.setattr(rasterizer, "RENDER_PACKED", True)
        b, _ = rc.run_cuda(sc, bg=(0.1, 0.2, 0.3))
        R = a.num_rendered()
        assert R == b.num_rendered() and R > 0
        for k in ("color", "depth", "final_T", "n_contrib", "ranges"):
            assert torch.equal(getattr(a, k), getattr(b, k)), k
        assert torch.equal(a.point_list[:R], b.point_list[:R]) and torch.equal(a.keybuf[:R], b.keybuf[:R])


def test_direct_binning_matches_scatter_path_and_falls_back(monkeypatch):
    """FsRasterFwdArgs.bins (preprocess appends the instance keys to per-tile bins, no scatter pass) against the counted +
    scattered path (BIN_CAP = 0): same ranges, sorted keys, lists and images bit for bit.  BIN_CAP = 64 is smaller than the
    heaviest tile of every scene: the device-side fallback (flag word -> scan + scatter inside the same launch sequence) must
    give the same result again; the crowded scene also has a tile above the 4096-key shared-memory sort."""
    from freesplat_b200 import rasterizer
    for sc in (synth.pixel_aligned_scene(seed=1, h=120, w=160, n_context=2, n_target=3, keep=None),
               synth.random_scene(seed=4, h=64, w=64, P=30000, sigma_px=(0.5, 2.0)), synth.random_scene(seed=3, h=100, w=77, P=3000)):
        out = []
        for cap in (0, 2048, 64):
            monkeypatch.setattr(rasterizer, "BIN_CAP", cap)
            rasterizer.reset_capacity_hints()                  # (forgets that this shape fell back before)
            st, _ = rc.run_cuda(sc, bg=(0.1, 0.2, 0.3))
            assert (st.bins is not None) == (cap > 0)
            heavy = int((st.ranges[:, 1] - st.ranges[:, 0]).max())
            assert int(st.status.cpu()[3]) == int(0 < cap < heavy)   # status[3]: the call took the device-side fallback
            R = st.num_rendered()
            heaviest = int((st.ranges[:, 1] - st.ranges[:, 0]).max())
            out.append((R, heaviest, st.ranges.clone(), st.point_list[:R].clone(), st.keybuf[:R].clone(), st.color.clone(), st.depth.clone(),
                        st.final_T.clone(), st.n_contrib.clone()))
        assert out[0][0] == out[1][0] == out[2][0] > 0 and out[0][1] > 64
        # the host noticed the fallback (sync check): the next state of this shape is built without bins
        st2, _ = rc.run_cuda(sc, bg=(0.1, 0.2, 0.3))
        assert st2.bins is None and torch.equal(st2.color, out[0][5])
        rasterizer.reset_capacity_hints()
        for other in out[1:]:
            for a, b in zip(out[0][2:], other[2:]):
                assert torch.equal(a, b)


def test_scan_fused_into_preprocess_is_bit_identical(monkeypatch):
    """[counters | cursors | status] in one buffer -> the tile scan runs in the last preprocess CTA (ticket in status[3]);
    a separate status buffer -> the stand-alone scan kernel.  Same ranges, lists and images; 18 views x 300 tiles covers the
    multi-pass case of the 256-thread scan."""
    from freesplat_b200 import rasterizer
    for sc in (synth.pixel_aligned_scene(seed=1, h=120, w=160, n_context=2, n_target=3, keep=None),
               synth.pixel_aligned_scene(seed=2, h=240, w=320, n_context=2, n_target=18, keep=20000)):
        out = []
        for fused in (False, True):
            monkeypatch.setattr(rasterizer, "FUSED_SCAN", fused)
            st, _ = rc.run_cuda(sc)
            R = st.num_rendered()
            out.append((R, st.ranges.clone(), st.point_list[:R].clone(), st.color.clone(), st.n_contrib.clone()))
        assert out[0][0] == out[1][0] > 0
        for a, b in zip(out[0][1:], out[1][1:]):
            assert torch.equal(a, b)


def test_capacity_overflow_retry():
    sc = synth.random_scene(seed=5, h=128, w=128, P=5000)
    st, _ = rc.run_cuda(sc, capacity=100)       # far too small: must re-run with the reported R
    assert not st.overflowed() and st.num_rendered() > 100
    m = rc.compare_forward(sc, st)
    assert not rc.forward_ok(m, n_pixels=128 * 128), m


def test_full_size_config2():
    """BASELINE config 2 at full size: 640x480, 307 200 pixel-aligned Gaussians, 3 target views."""
    sc = synth.pixel_aligned_scene(seed=0, h=480, w=640, n_context=2, n_target=3, keep=307200)
    st, _, m = _check(sc)
    for v in m["views"]:
        assert v["psnr_vs_oracle_db"] > 80.0     # PSNR parity: << 0.01 dB


@pytest.mark.parametrize("seed,with_depth", [(0, False), (1, True)])
def test_backward_vs_oracle(seed, with_depth):
    sc = synth.pixel_aligned_scene(seed=seed, h=96, w=128, n_context=2, n_target=2, keep=None)
    st, views = rc.run_cuda(sc, bg=(0.1, 0.2, 0.3))
    g = torch.Generator().manual_seed(7 + seed)
    dC = torch.randn((st.V, 3, st.H, st.W), generator=g)
    dD = torch.randn((st.V, st.H, st.W), generator=g) * 0.2 if with_depth else None
    dA = torch.randn((st.V, st.H, st.W), generator=g) * 0.5 if with_depth else None       # 4th output: 1 - final_T
    m = rc.compare_backward(sc, st, views, dC, dD, bg=(0.1, 0.2, 0.3), dL_dalpha=dA)
    assert not rc.backward_ok(m), m


def test_backward_random_scene():
    sc = synth.random_scene(seed=2, h=128, w=128, P=4000)
    st, views = rc.run_cuda(sc)
    g = torch.Generator().manual_seed(11)
    dC = torch.randn((st.V, 3, st.H, st.W), generator=g)
    m = rc.compare_backward(sc, st, views, dC)
    assert not rc.backward_ok(m), m


def test_dropin_module_and_autograd():
    """The reference's call site, verbatim: settings + rasterizer(...) 4-tuple, gradients via autograd."""
    from diff_gaussian_rasterization_depth import GaussianRasterizationSettings, GaussianRasterizer
    from freesplat_b200 import decoder
    from tests.helpers import view_inputs
    from oracle import raster as oracle
    sc = synth.random_scene(seed=6, h=96, w=96, P=2000)
    inp, vrec = view_inputs(sc, 0, bg=(0.0, 0.0, 0.0))
    dev = "cuda:0"
    t = lambda a: torch.tensor(a, device=dev)
    means = t(inp["means3D"]).requires_grad_(True); cov = t(inp["cov3D_precomp"]).requires_grad_(True)
    shs = t(inp["shs"]).requires_grad_(True); op = t(inp["opacities"])[:, None].requires_grad_(True)
    means2D = torch.zeros_like(means, requires_grad=True)
    settings = GaussianRasterizationSettings(
        image_height=inp["H"], image_width=inp["W"], tanfovx=inp["tanfovx"], tanfovy=inp["tanfovy"], bg=t(inp["bg"]),
        scale_modifier=1.0, viewmatrix=t(inp["viewmatrix"]).reshape(4, 4), projmatrix=t(inp["projmatrix"]).reshape(4, 4),
        sh_degree=inp["sh_degree"], campos=t(inp["campos"]), prefiltered=False, debug=False)
    image, radii, depth, alpha = GaussianRasterizer(settings)(means3D=means, means2D=means2D, shs=shs, colors_precomp=None,
                                                               opacities=op, cov3D_precomp=cov)
    assert image.shape == (3, 96, 96) and depth.shape == (96, 96) and radii.shape == (2000,) and alpha.shape == (96, 96)
    o = oracle.forward(**inp)
    assert np.array_equal(radii.cpu().numpy(), o.radii)
    assert np.isclose(image.detach().cpu().numpy(), o.color, rtol=1e-4, atol=1e-5).mean() > 0.9999
    from tests.helpers import grad_report
    g = torch.Generator().manual_seed(3)
    dC = torch.randn((3, 96, 96), generator=g)
    dD, dA = torch.randn((96, 96), generator=g) * 0.2, torch.randn((96, 96), generator=g) * 0.5
    # depth and the accumulated alpha are differentiable outputs of the module this replaces: a loss on all three
    ((image * dC.to(dev)).sum() + (depth * dD.to(dev)).sum() + (alpha * dA.to(dev)).sum()).backward()
    go = oracle.backward(o, tanfovx=inp["tanfovx"], tanfovy=inp["tanfovy"], bg=inp["bg"], viewmatrix=inp["viewmatrix"],
                         projmatrix=inp["projmatrix"], campos=inp["campos"], means3D=inp["means3D"], dL_dcolor=dC.numpy(),
                         dL_ddepth=dD.numpy(), dL_dalpha=dA.numpy(), shs=inp["shs"], sh_degree=inp["sh_degree"])
    for name, got, want in [("means3D", means.grad, go["means3D"]), ("cov", cov.grad, go["cov3D"]), ("shs", shs.grad, go["shs"]),
                            ("op", op.grad, go["opacities"]), ("means2D", means2D.grad, go["means2D"])]:
        rep = grad_report(got.cpu().numpy(), want)
        assert rep["ok"], (name, rep)
    vis = GaussianRasterizer(settings).markVisible(means.detach())
    assert np.array_equal(vis.cpu().numpy(), o.depths > 0.2) or np.array_equal(vis.cpu().numpy()[o.radii > 0], np.ones((o.radii > 0).sum(), bool))


def test_per_view_dropin_calls_match_batched():
    """V separate calls of the drop-in op in the upstream layouts == render_views (one batched launch sequence, in-place
    layouts); the reference's own render_cuda / DecoderSplattingCUDA are covered by tests/test_reference_adapter_gpu.py."""
    from freesplat_b200 import decoder
    from tests.helpers import render_per_view_upstream_layout
    sc = synth.pixel_aligned_scene(seed=2, h=96, w=128, n_context=2, n_target=3, keep=None).to("cuda:0")
    V = 3
    bg = torch.zeros((V, 3), device="cuda:0")
    with torch.no_grad():
        c1, d1 = decoder.render_views(sc.extrinsics, sc.intrinsics, sc.near, sc.far, sc.image_shape, bg, sc.means,
                                      sc.covariances, sc.harmonics, sc.opacities)
        c2, d2 = render_per_view_upstream_layout(sc.extrinsics, sc.intrinsics, sc.near, sc.far, sc.image_shape, bg, sc.means,
                                                 sc.covariances, sc.harmonics, sc.opacities)
    # render_views builds its camera records with the fused fp64 kernel, the per-view driver with torch fp32 ops: the
    # matrices agree to ~1e-7, so the images agree except where a Gaussian's integer radius / a pixel threshold flips
    ok = torch.isclose(c1, c2, rtol=1e-3, atol=1e-4)
    assert ok.float().mean() > 0.9995 and torch.isclose(d1, d2, rtol=1e-3, atol=1e-3).float().mean() > 0.9995
    va = decoder.camera_records_fused(sc.extrinsics, sc.intrinsics, sc.near, sc.far, bg)
    vb, _ = decoder.camera_records(sc.extrinsics, sc.intrinsics, sc.near, sc.far, bg)
    assert torch.allclose(va, vb, rtol=2e-5, atol=2e-6), (va - vb).abs().max()


def test_native_layout_gradients_match_upstream_layout():
    """render_views reads [G,3,3] covariances and [G,3,d_sh] harmonics in place (sh_layout=1, cov_stride=9); its
    gradients must equal those of the reference-style path (transpose + triu gather + per-view op) through autograd."""
    from freesplat_b200 import decoder
    dev = "cuda:0"
    sc = synth.pixel_aligned_scene(seed=3, h=64, w=96, n_context=2, n_target=2, keep=None).to(dev)
    V = 2
    bg = torch.tensor([[0.1, 0.2, 0.3]], device=dev).expand(V, 3).contiguous()
    g = torch.Generator().manual_seed(5)
    dC = torch.randn((V, 3, 64, 96), generator=g).to(dev)
    grads = []
    for mode in ("native", "upstream"):
        m = sc.means.clone().requires_grad_(True); c = sc.covariances.clone().requires_grad_(True)
        s = sc.harmonics.clone().requires_grad_(True); o = sc.opacities.clone().requires_grad_(True)
        if mode == "native":
            col, dep = decoder.render_views(sc.extrinsics, sc.intrinsics, sc.near, sc.far, sc.image_shape, bg, m, c, s, o)
        else:
            from tests.helpers import render_per_view_upstream_layout
            col, dep = render_per_view_upstream_layout(sc.extrinsics, sc.intrinsics, sc.near, sc.far, sc.image_shape, bg, m, c, s, o)
        (col * dC).sum().backward()
        grads.append((col.detach(), m.grad, c.grad, s.grad, o.grad))
    assert torch.isclose(grads[0][0], grads[1][0], rtol=1e-3, atol=1e-4).float().mean() > 0.9995
    for a, b, name in zip(grads[0][1:], grads[1][1:], ("means", "cov", "sh", "opacity")):
        scale = b.abs().max() + 1e-20
        err = (a - b).abs() / scale
        # camera records differ by ~1e-7 between the two adapters (fused fp64 kernel vs torch fp32): a Gaussian whose
        # radius or a pixel whose threshold flips changes a few entries; everything else agrees to atomics noise
        assert torch.quantile(err.flatten()[:4_000_000], 0.999) < 2e-4, name
    assert torch.equal(grads[0][2][:, 1, 0], torch.zeros_like(grads[0][2][:, 1, 0]))   # lower triangle: no gradient


def test_host_pipeline_matches_direct_call():
    """HostRenderPipeline (3 streams, double-buffered, deferred overflow check) returns what render_views returns."""
    from freesplat_b200 import decoder
    from freesplat_b200.pipeline import HostRenderPipeline
    dev = "cuda:0"
    scenes = [synth.pixel_aligned_scene(seed=s, h=96, w=128, n_context=2, n_target=2, keep=None) for s in (0, 1, 2, 3, 4)]
    pipe = HostRenderPipeline(dev, (96, 128), 2, depth=2)
    pin = lambda t: t.contiguous().pin_memory()
    got = []
    for sc in scenes:
        host = dict(extrinsics=pin(sc.extrinsics), intrinsics=pin(sc.intrinsics), near=pin(sc.near), far=pin(sc.far),
                    means=pin(sc.means), covariances=pin(sc.covariances), harmonics=pin(sc.harmonics), opacities=pin(sc.opacities))
        slot = pipe.submit(host)
        c, d = pipe.wait(slot)
        got.append((c.clone(), d.clone()))
    for sc, (c, d) in zip(scenes, got):
        g = sc.to(dev)
        with torch.no_grad():
            c2, d2 = decoder.render_views(g.extrinsics, g.intrinsics, g.near, g.far, g.image_shape, torch.zeros((2, 3), device=dev),
                                          g.means, g.covariances, g.harmonics, g.opacities)
        assert torch.equal(c, c2.cpu()) and torch.equal(d, d2.cpu())


def test_host_pipeline_recovers_from_deferred_overflow():
    """A scene that needs more tile instances than the workspace holds, submitted AFTER the pipeline is in its sync-free
    steady state: the overflow is noticed when the slot is next touched and the step is re-run on a grown workspace."""
    from freesplat_b200 import decoder, rasterizer
    from freesplat_b200.pipeline import HostRenderPipeline
    dev = "cuda:0"
    h, w, V = 96, 128, 2
    small = synth.pixel_aligned_scene(seed=0, h=h, w=w, n_context=2, n_target=V, keep=None)
    big = synth.pixel_aligned_scene(seed=1, h=h, w=w, n_context=2, n_target=V, keep=None)
    big.covariances = big.covariances * 400.0            # 20x larger footprints: many more (Gaussian, tile) instances
    pin = lambda t: t.contiguous().pin_memory()
    host = lambda sc: dict(extrinsics=pin(sc.extrinsics), intrinsics=pin(sc.intrinsics), near=pin(sc.near), far=pin(sc.far),
                           means=pin(sc.means), covariances=pin(sc.covariances), harmonics=pin(sc.harmonics), opacities=pin(sc.opacities))
    key = (torch.device(dev).index, small.means.shape[0], V, h, w)
    rasterizer._capacity_hint[key] = 1 << 16             # a workspace sized for the small scene only
    try:
        pipe = HostRenderPipeline(dev, (h, w), V, depth=2)
        for _ in range(3):
            pipe.submit(host(small))
        slot = pipe.submit(host(big))
        pipe.submit(host(small))
        pipe.submit(host(small))                         # reuses the big scene's slot: resolves it first
        c, d = pipe.wait(slot)                           # (already overwritten by now: only checks that nothing raises)
        slot = pipe.submit(host(big))
        c, d = pipe.wait(slot)
        c, d = c.clone(), d.clone()
        pipe.drain()
        assert pipe.reruns >= 1
        g = big.to(dev)
        with torch.no_grad():
            c2, d2 = decoder.render_views(g.extrinsics, g.intrinsics, g.near, g.far, (h, w), torch.zeros((V, 3), device=dev),
                                          g.means, g.covariances, g.harmonics, g.opacities, check_overflow="sync")
        assert torch.equal(c, c2.cpu()) and torch.equal(d, d2.cpu())
    finally:
        rasterizer.reset_capacity_hints()


def test_deferred_check_reports_overflow_at_the_next_call():
    from freesplat_b200 import _lib, decoder, rasterizer
    dev = "cuda:0"
    h, w, V = 96, 128, 2
    sc = synth.pixel_aligned_scene(seed=1, h=h, w=w, n_context=2, n_target=V, keep=None).to(dev)
    bg = torch.zeros((V, 3), device=dev)
    key = (torch.device(dev).index, sc.means.shape[0], V, h, w)
    rasterizer.reset_capacity_hints()
    rasterizer._capacity_hint[key] = 1024
    try:
        with torch.no_grad():
            decoder.render_views(sc.extrinsics, sc.intrinsics, sc.near, sc.far, (h, w), bg, sc.means, sc.covariances, sc.harmonics,
                                 sc.opacities, check_overflow="deferred")
            torch.cuda.synchronize()
            with pytest.raises(_lib.FreeSplatB200Error, match="tile instances"):
                rasterizer.poll_deferred(block=True)
            assert rasterizer._capacity_hint[key] > 1024          # grown: the re-run fits
            c, d = decoder.render_views(sc.extrinsics, sc.intrinsics, sc.near, sc.far, (h, w), bg, sc.means, sc.covariances,
                                        sc.harmonics, sc.opacities, check_overflow="deferred")
            rasterizer.poll_deferred(block=True)
            c2, d2 = decoder.render_views(sc.extrinsics, sc.intrinsics, sc.near, sc.far, (h, w), bg, sc.means, sc.covariances,
                                          sc.harmonics, sc.opacities, check_overflow="sync")
        assert torch.equal(c, c2) and torch.equal(d, d2)
    finally:
        rasterizer.reset_capacity_hints()


def test_raster_plan_graph_equals_eager():
    """RasterPlan (static workspace, one CUDA-graph launch per step) returns bit-identical images to the eager call sequence,
    step after step, also when the scene in its input buffers changes."""
    from freesplat_b200 import decoder, rasterizer
    dev = "cuda:0"
    h, w, V = 96, 128, 3
    scs = [synth.pixel_aligned_scene(seed=s, h=h, w=w, n_context=2, n_target=V, keep=None).to(dev) for s in (0, 1)]
    bg = torch.zeros((V, 3), device=dev)
    buf = {k: getattr(scs[0], k).clone() for k in ("means", "opacities", "harmonics", "covariances", "extrinsics", "intrinsics", "near", "far")}
    plan = rasterizer.RasterPlan(buf["means"], buf["opacities"], h, w, shs=buf["harmonics"], cov3D_precomp=buf["covariances"].reshape(-1, 9),
                                 cameras=(buf["extrinsics"], buf["intrinsics"], buf["near"], buf["far"], bg), sh_degree=2, sh_layout=1,
                                 cov_stride=9)
    assert plan.graph is not None
    for it in range(4):
        sc = scs[it % 2]
        for k in buf:
            buf[k].copy_(getattr(sc, k))
        R = plan.run_checked()
        with torch.no_grad():
            c2, d2 = decoder.render_views(sc.extrinsics, sc.intrinsics, sc.near, sc.far, (h, w), bg, sc.means, sc.covariances,
                                          sc.harmonics, sc.opacities, check_overflow="sync")
        assert R > 0 and torch.equal(plan.color, c2) and torch.equal(plan.depth, d2)


def test_training_loop_reuses_its_workspace():
    """Steady-state training steps must not allocate device memory: the autograd node may not keep its own outputs alive
    (a reference cycle would park every step's workspace until Python's cyclic GC runs)."""
    import gc
    from freesplat_b200 import decoder
    dev = "cuda:0"
    sc = synth.pixel_aligned_scene(seed=0, h=120, w=160, n_context=2, n_target=3, keep=None).to(dev)
    params = [x.detach().clone().requires_grad_(True) for x in (sc.means, sc.covariances, sc.harmonics, sc.opacities)]
    bg = torch.zeros((3, 3), device=dev)
    gc.collect(); gc.disable()
    try:
        allocs = []
        for it in range(8):
            for p in params:
                p.grad = None
            col, dep = decoder.render_views(sc.extrinsics, sc.intrinsics, sc.near, sc.far, (120, 160), bg, *params)
            (col ** 2).mean().backward()
            torch.cuda.synchronize()
            allocs.append(torch.cuda.memory_stats()["num_device_alloc"])
    finally:
        gc.enable()
    assert allocs[-1] == allocs[3], allocs
