"""Drop-in package: same import name as the reference extension
(/root/reference/submodules/diff-gaussian-surfels/diff_gaussian_rasterization/__init__.py).

Put the parent directory (`eggfusion_b200/dropin`) first on sys.path -- or install it -- and EGG-Fusion's
`from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer`
(src/core/render.py:8-11) resolves to the sm_100a implementation without touching src/.
"""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _root not in sys.path:
    sys.path.insert(0, _root)

from eggfusion_b200.rasterizer import (  # noqa: E402,F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    _RasterizeGaussians,
    cpu_deep_copy_tuple,
    rasterize_gaussians,
)
from eggfusion_b200.fusion import preprocess_surfels, project_surfels_to_frame  # noqa: E402,F401
