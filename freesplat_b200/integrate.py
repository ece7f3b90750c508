"""One call that swaps the B200 operators into an (unmodified) FreeSplat checkout -- INTEGRATION.md as code.

    import freesplat_b200.integrate as fsi
    fsi.patch()            # before the model is built:  python -m src.main +experiment=...

What it touches, by the reference's own names (nothing under `src/` is edited):
  * `diff_gaussian_rasterization_depth`             -- nothing to patch: the top-level package of this repo is found first
                                                      on PYTHONPATH and `src/model/decoder/cuda_splatting.py:5-8` imports it;
  * `src.model.encoder.encoder_freesplat.AVGFeatureVolumeManager`
                                                   -- rebound to `freesplat_b200.cost_volume.AVGFeatureVolumeManager` (same
                                                      constructor, parameter and buffer names: checkpoints load), so the
                                                      construction at `encoder_freesplat.py:157-160` builds the fused operator;
  * `EncoderFreeSplat.fuse_gaussians`              -- rebound to `freesplat_b200.ptf.fuse_gaussians` (same signature,
                                                      `encoder_freesplat.py:431-432`);
  * optionally (`decoder=True`) `DecoderSplattingCUDA.forward` -> the batched all-views renderer
    (`freesplat_b200.decoder.DecoderSplattingB200.forward`), which produces the same `DecoderOutput(color, depth)`.
`unpatch()` restores the originals.  A missing `libfreesplat_b200.so` fails here, loudly, not at the first training step."""
from __future__ import annotations

import importlib
from typing import Any, Dict, Tuple

from . import _lib

_saved: Dict[Tuple[Any, str], Any] = {}


def _swap(obj, name, new):
    key = (obj, name)
    if key not in _saved:
        _saved[key] = getattr(obj, name)
    setattr(obj, name, new)


def patch(src_package: str = "src", decoder: bool = False) -> Dict[str, str]:
    """Rebinds the reference's hot-path classes / methods to the B200 operators.  Returns {patched name: replacement}."""
    _lib.lib()                                   # raises FreeSplatB200Error when the CUDA library is missing / stale
    from . import cost_volume, ptf
    done = {}
    enc = importlib.import_module(f"{src_package}.model.encoder.encoder_freesplat")
    _swap(enc, "AVGFeatureVolumeManager", cost_volume.AVGFeatureVolumeManager)
    done[f"{enc.__name__}.AVGFeatureVolumeManager"] = "freesplat_b200.cost_volume.AVGFeatureVolumeManager"
    _swap(enc.EncoderFreeSplat, "fuse_gaussians", ptf.fuse_gaussians)
    done[f"{enc.__name__}.EncoderFreeSplat.fuse_gaussians"] = "freesplat_b200.ptf.fuse_gaussians"
    if decoder:
        from . import decoder as dec
        mod = importlib.import_module(f"{src_package}.model.decoder.decoder_splatting_cuda")
        b200 = dec.DecoderSplattingB200

        def forward(self, gaussians, extrinsics, intrinsics, near, far, image_shape, depth_mode=None, no_color=False):
            # decoder_splatting_cuda.py:35-75: same arguments, DecoderOutput(color, depth | None)
            ref_out = getattr(mod, "DecoderOutput", None)
            color, depth = (None, None) if no_color else b200.forward(self, gaussians, extrinsics, intrinsics, near, far,
                                                                      image_shape, depth_mode)
            return (color, depth) if ref_out is None else ref_out(color, depth)
        _swap(mod.DecoderSplattingCUDA, "forward", forward)
        done[f"{mod.__name__}.DecoderSplattingCUDA.forward"] = "freesplat_b200.decoder.DecoderSplattingB200.forward"
    return done


def unpatch() -> None:
    for (obj, name), old in list(_saved.items()):
        setattr(obj, name, old)
    _saved.clear()
