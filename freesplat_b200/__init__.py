"""freesplat_b200 -- B200-native (sm_100a) implementation of FreeSplat's data-parallel hot path:
Gaussian rasterization fwd/bwd, plane-sweep cost volume, Pixel-wise Triplet Fusion.

Importing the package does not load the CUDA library; the first operator call does, and fails
loudly when libfreesplat_b200.so is absent (no CPU fallback)."""
__version__ = "0.1.0"
