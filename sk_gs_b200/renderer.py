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


# ------------------------------------------------------------------------------------------------------------------
# Boundary B2: the in-tree extension entry points `get_C_function('rasterize_gaussians')` /
# `('rasterize_gaussians_backward')` (reference networks/renderer/gaussian_render.py:69-93,152-174; C++
# my_ext/_C/src/nerf/gaussian_rasterizer_forward.cu:260-315, gaussian_rasterizer_backwrad.cu:200-261).
# Same positional arguments and returned tuples; rotations are (x,y,z,w); no background inside (the reference blends
# it in Python, networks/sk_gs.py:1230-1231).  `extras` (per-Gaussian extra channels) are not on the SK_GS hot path and
# are rejected.  The three "buffers" of the reference become one opaque RasterState carried in the geomBuffer slot.
# ------------------------------------------------------------------------------------------------------------------
def rasterize_gaussians_b2(image_height, image_width, tanfovx, tanfovy, sh_degree, scale_modifier, prefiltered, debug,
                           colmap, viewmatrix, projmatrix, campos, means3D, opacities, sh, scales, rotations, extras,
                           colors_precomp, cov3Ds_precomp):
    from . import diff_gaussian_rasterization as DGR
    if extras is not None and extras.numel() > 0:
        raise RuntimeError('extras are not supported by sk_gs_b200 (not on the SK_GS hot path)')
    if not colmap:
        raise RuntimeError('only the colmap/upstream convention is implemented (every shipped config uses it)')
    rs = GaussianRasterizationSettings(image_height, image_width, tanfovx, tanfovy, None, scale_modifier, viewmatrix,
                                       projmatrix, sh_degree, campos, prefiltered, debug)
    color, depth, alpha, radii, state = DGR.rasterize_forward(rs, means3D, opacities, sh, colors_precomp, scales,
                                                              rotations, cov3Ds_precomp, quat_wxyz=False)
    num_rendered = state.num_rendered
    return num_rendered, color, alpha[0], radii, state, state.binning, state.img, None


def rasterize_gaussians_backward_b2(scale_modifier, tanfovx, tanfovy, sh_degree, debug, colmap, viewmatrix, projmatrix,
                                    campos, means3D, colors, extras, scales, rotations, cov3D, sh, R, radii,
                                    out_opacity, dL_dcolor, dL_dopacity, dL_dextra, grad_means2D, grad_conic,
                                    grad_opacity, geomBuffer, binningBuffer, imgBuffer):
    """-> (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations, dL_dextras)
    like the reference; `geomBuffer` is the RasterState returned by rasterize_gaussians_b2."""
    from . import diff_gaussian_rasterization as DGR
    state = geomBuffer
    g = DGR.rasterize_backward(state, dL_dcolor, None, None if dL_dopacity is None else dL_dopacity.reshape(1, *dL_dopacity.shape[-2:]))
    return (g['means2D'], g['colors_precomp'], g['opacities'], g['means3D'], g['cov3D_precomp'], g['shs'], g['scales'],
            g['rotations'], None)
