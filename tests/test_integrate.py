"""CPU: freesplat_b200.integrate.patch() rebinds exactly the reference's hot-path names (checked on a stand-in package
with the reference's module layout; the real checkout needs timm / jaxtyping / lightning, absent here) and unpatch()
restores them."""
import sys
import types

import pytest


def _fake_src(root="fakesrc"):
    mods = {}
    for name in (root, f"{root}.model", f"{root}.model.encoder", f"{root}.model.encoder.encoder_freesplat", f"{root}.model.decoder",
                 f"{root}.model.decoder.decoder_splatting_cuda"):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
        mods[name] = m
    enc = mods[f"{root}.model.encoder.encoder_freesplat"]

    class AVGFeatureVolumeManager:            # stands for `from .modules.cost_volume import AVGFeatureVolumeManager` (:19)
        pass

    class EncoderFreeSplat:
        def fuse_gaussians(self, *a, **k):
            return "reference"
    enc.AVGFeatureVolumeManager, enc.EncoderFreeSplat = AVGFeatureVolumeManager, EncoderFreeSplat
    dec = mods[f"{root}.model.decoder.decoder_splatting_cuda"]

    class DecoderOutput:
        def __init__(self, color, depth):
            self.color, self.depth = color, depth

    class DecoderSplattingCUDA:
        def forward(self, *a, **k):
            return "reference"
    dec.DecoderOutput, dec.DecoderSplattingCUDA = DecoderOutput, DecoderSplattingCUDA
    return enc, dec, AVGFeatureVolumeManager, EncoderFreeSplat.fuse_gaussians, DecoderSplattingCUDA.forward


def test_patch_and_unpatch():
    from freesplat_b200 import cost_volume, integrate, ptf
    enc, dec, ref_cv, ref_fuse, ref_fwd = _fake_src()
    try:
        done = integrate.patch("fakesrc", decoder=True)
        assert enc.AVGFeatureVolumeManager is cost_volume.AVGFeatureVolumeManager
        assert enc.EncoderFreeSplat.fuse_gaussians is ptf.fuse_gaussians
        assert dec.DecoderSplattingCUDA.forward is not ref_fwd
        assert len(done) == 3 and all(v.startswith("freesplat_b200.") for v in done.values())
        out = dec.DecoderSplattingCUDA.forward(object(), None, None, None, None, None, (4, 4), no_color=True)
        assert isinstance(out, dec.DecoderOutput) and out.color is None and out.depth is None
    finally:
        integrate.unpatch()
    assert enc.AVGFeatureVolumeManager is ref_cv and enc.EncoderFreeSplat.fuse_gaussians is ref_fuse
    assert dec.DecoderSplattingCUDA.forward is ref_fwd


def test_patch_fails_loudly_without_the_package():
    from freesplat_b200 import integrate
    with pytest.raises(ModuleNotFoundError):
        integrate.patch("no_such_freesplat_checkout")
