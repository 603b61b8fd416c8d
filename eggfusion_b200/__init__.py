"""eggfusion_b200 -- B200-native (sm_100a) differentiable Gaussian-surfel rasterizer: a drop-in for the
render/optimise hot path of EGG-Fusion (the `diff_gaussian_rasterization` extension of the reference).

    from eggfusion_b200 import GaussianRasterizationSettings, GaussianRasterizer

or put `eggfusion_b200/dropin` on sys.path and keep `from diff_gaussian_rasterization import ...` unchanged.
"""
from ._lib import build, load  # noqa: F401
from .rasterizer import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    rasterize_gaussians,
    cpu_deep_copy_tuple,
)

from .fusion import preprocess_surfels, project_surfels_to_frame  # noqa: F401,E402

__all__ = ["preprocess_surfels", "project_surfels_to_frame", "build", "load", "GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians",
           "cpu_deep_copy_tuple"]
