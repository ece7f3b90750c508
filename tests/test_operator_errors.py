"""CPU: the operator mirrors keep the reference's argument rules and error behaviour, and never fall back to the CPU:
wrong argument combinations raise the upstream wrapper's messages BEFORE any CUDA work; CPU tensors raise
FreeSplatB200Error from every operator."""
import pytest
import torch

from freesplat_b200 import _lib


def _settings():
    from diff_gaussian_rasterization_depth import GaussianRasterizationSettings
    eye = torch.eye(4)
    return GaussianRasterizationSettings(image_height=16, image_width=16, tanfovx=0.5, tanfovy=0.5, bg=torch.zeros(3),
                                         scale_modifier=1.0, viewmatrix=eye, projmatrix=eye, sh_degree=0, campos=torch.zeros(3),
                                         prefiltered=False, debug=False)


def test_dropin_module_exports_the_reference_names():
    import diff_gaussian_rasterization_depth as m
    assert set(m.__all__) == {"GaussianRasterizationSettings", "GaussianRasterizer"}
    assert m.GaussianRasterizationSettings._fields == ("image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier",
                                                       "viewmatrix", "projmatrix", "sh_degree", "campos", "prefiltered", "debug")


def test_rasterizer_argument_rules_match_upstream_messages():
    from diff_gaussian_rasterization_depth import GaussianRasterizer
    r = GaussianRasterizer(_settings())
    P = 4
    m3, m2, op = torch.zeros(P, 3), torch.zeros(P, 3), torch.ones(P, 1)
    sh, col, cov = torch.zeros(P, 1, 3), torch.zeros(P, 3), torch.zeros(P, 6)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(m3, m2, op, shs=sh, colors_precomp=col, cov3D_precomp=cov)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(m3, m2, op, cov3D_precomp=cov)
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        r(m3, m2, op, shs=sh)
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        r(m3, m2, op, shs=sh, scales=torch.ones(P, 3), rotations=torch.zeros(P, 4), cov3D_precomp=cov)


def test_cpu_tensors_raise_everywhere():
    from diff_gaussian_rasterization_depth import GaussianRasterizer
    from freesplat_b200 import adapter, depth_head, ply_export, ptf
    from freesplat_b200.cost_volume import AVGFeatureVolumeManager
    E = _lib.FreeSplatB200Error
    P = 4
    with pytest.raises(E):
        GaussianRasterizer(_settings())(torch.zeros(P, 3), torch.zeros(P, 3), torch.ones(P, 1), shs=torch.zeros(P, 1, 3),
                                        cov3D_precomp=torch.zeros(P, 6))
    with pytest.raises(E):
        depth_head.depth_regression(torch.zeros(1, 4, 3, 3), torch.zeros(4))
    with pytest.raises(E):
        adapter.backproject_depth(torch.ones(1, 4, 4), torch.eye(3), torch.eye(4)[None], (4, 4))
    with pytest.raises(E):
        adapter.gaussian_head(torch.zeros(2, 34), torch.ones(2), torch.ones(2), torch.zeros(2, 3), torch.eye(4).repeat(2, 1, 1),
                              torch.eye(3), (4, 4))
    with pytest.raises(E):
        ply_export.vertex_table(torch.eye(4), torch.zeros(2, 3), torch.ones(2, 3), torch.ones(2, 4), torch.zeros(2, 3, 9), torch.ones(2))
    with pytest.raises(E):
        ptf.fuse_views(torch.nn.Identity(), torch.zeros(2, 4, 64), torch.zeros(2, 4, 3), torch.zeros(2, 4), torch.zeros(2, 4),
                       torch.ones(2, 4), torch.eye(4).repeat(2, 1, 1), torch.eye(3).repeat(2, 1, 1), (2, 2))
    cv = AVGFeatureVolumeManager(4, 4, num_depth_bins=4, matching_dim_size=48)
    with pytest.raises(E):
        cv(cur_feats=torch.zeros(1, 48, 4, 4), src_feats=torch.zeros(1, 1, 48, 4, 4), src_extrinsics=torch.eye(4)[None, None],
           src_poses=torch.eye(4)[None, None], src_Ks=torch.eye(4)[None, None], cur_invK=torch.eye(4)[None],
           min_depth=torch.tensor(0.5), max_depth=torch.tensor(15.0))


def test_numa_binding_helper_never_raises():
    """No NVML in the build container: the helper reports why it did nothing and leaves the affinity alone."""
    import os
    from freesplat_b200.pipeline import bind_to_gpu_numa
    before = os.sched_getaffinity(0)
    info = bind_to_gpu_numa(0)
    assert isinstance(info, dict) and "bound" in info
    if not info["bound"]:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)
