"""ex4dgs_b200 - B200-native (sm_100a) differentiable 4D-Gaussian rasterizer, drop-in for the
`diff_gaussian_rasterization_df` extension of juno181/Ex4DGS (hot path only)."""
from .rasterizer import (GaussianRasterizationSettings, GaussianRasterizer, _RasterizeGaussians,
                         rasterize_gaussians, set_default_flags, get_default_flags, last_inexact_thresholds, SegmentedSH,
                         set_host_wait, get_host_wait, set_capacity_hint, frame_overflowed)

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "_RasterizeGaussians",
           "rasterize_gaussians", "set_default_flags", "get_default_flags", "last_inexact_thresholds", "SegmentedSH",
           "set_host_wait", "get_host_wait", "set_capacity_hint", "frame_overflowed"]
