"""Drop-in for the third-party module FreeSplat imports at
/root/reference/src/model/decoder/cuda_splatting.py:5-8
(`from diff_gaussian_rasterization_depth import GaussianRasterizationSettings, GaussianRasterizer`).

With /root/repo on PYTHONPATH the reference's src/model runs unchanged on the B200 kernels.
"""
from freesplat_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer  # noqa: F401

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer"]
