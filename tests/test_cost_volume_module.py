"""CPU: the module mirror keeps the reference's parameter / buffer names and shapes (checkpoints load,
SURVEY §8b) and refuses to run without the CUDA library (no fallback)."""
import pytest
import torch

from freesplat_b200 import _lib
from freesplat_b200.cost_volume import AVGFeatureVolumeManager, pack_mlp, unpack_mlp


def test_state_dict_names_match_reference():
    m = AVGFeatureVolumeManager(120, 160, num_depth_bins=128, matching_dim_size=48)
    sd = m.state_dict()
    want = {
        "linear_ramp_1d11": (1, 128, 1, 1), "backprojector.pix_coords_13N": (1, 3, 19200), "projector.eps": (1, 1, 1),
        "mlp.net.0.weight": (32, 49), "mlp.net.0.bias": (32,), "mlp.net.2.weight": (32, 32), "mlp.net.2.bias": (32,),
        "mlp.net.4.weight": (1, 32), "mlp.net.4.bias": (1,),
    }
    assert {k: tuple(v.shape) for k, v in sd.items()} == want
    flat = pack_mlp([sd[f"mlp.net.{i}.{n}"] for i in (0, 2, 4) for n in ("weight", "bias")])
    assert flat.numel() == 32 * 49 + 32 + 32 * 32 + 32 + 32 + 1
    assert all(torch.equal(a, sd[f"mlp.net.{i}.{n}"]) for a, (i, n) in
               zip(unpack_mlp(flat), [(i, n) for i in (0, 2, 4) for n in ("weight", "bias")]))


def test_reference_names_if_available():
    from tests.golden import ref_loader
    if not ref_loader.available():
        pytest.skip("/root/reference not present (GPU box)")
    cv = ref_loader.load_cost_volume_module()
    ref = cv.AVGFeatureVolumeManager(24, 32, num_depth_bins=16, mlp_channels=[49, 32, 32, 1], matching_dim_size=48)
    ours = AVGFeatureVolumeManager(24, 32, num_depth_bins=16, mlp_channels=[49, 32, 32, 1], matching_dim_size=48)
    a = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    b = {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    assert a == b
    ours.load_state_dict(ref.state_dict())
    assert torch.equal(ours.backprojector.pix_coords_13N, ref.backprojector.pix_coords_13N)


def test_cpu_tensors_raise():
    from freesplat_b200 import synth
    inp = synth.cost_volume_inputs(0, 2, 1, 48, 8, 8)
    m = AVGFeatureVolumeManager(8, 8, num_depth_bins=4, matching_dim_size=48)
    with pytest.raises(_lib.FreeSplatB200Error):
        m(**inp)
