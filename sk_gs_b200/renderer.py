"""`render_gs_offical` - same signature, same returned dict as the reference adapter
/root/reference/networks/renderer/gaussian_render_origin.py:11-68 (the entry point `SkeletonGaussianSplatting.render`
calls at networks/sk_gs.py:1228 and `GaussianSplatting.render` at networks/gaussian_splatting.py:309)."""
from __future__ import annotations

import torch
from torch import Tensor

from .diff_gaussian_rasterization import GaussianRasterizationSettings, rasterize_gaussians


def render_gs_offical(points: Tensor, opacity: Tensor, raster_settings: GaussianRasterizationSettings,
                      scales: Tensor = None, rotations: Tensor = None, covariance: Tensor = None,
                      sh_features: Tensor = None, colors=None, extras=None, **kwargs):
    """Render the scene.  `rotations` are (x, y, z, w) as everywhere inside SK_GS; the reference permutes them to
    (w, x, y, z) for the upstream module (gaussian_render_origin.py:41-42) - here the kernels read xyzw directly, which
    saves that gather and its backward scatter."""
    # zero tensor that receives the gradient of the 2D (screen-space) means (gaussian_render_origin.py:30-34)
    screenspace_points = torch.zeros_like(points, requires_grad=True) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:  # noqa
        pass
    if (sh_features is None) == (colors is None):
        raise Exception('Please provide excatly one of either SHs or precomputed colors!')
    rendered_image, radii, depth, alpha = rasterize_gaussians(
        points, screenspace_points, sh_features, colors, opacity, scales, rotations, covariance, raster_settings,
        quat_wxyz=False)
    return {
        'images': rendered_image,
        'viewspace_points': screenspace_points,
        'visibility_filter': radii > 0,
        'radii': radii,
        'depths': depth,
        'alpha': alpha,
    }
