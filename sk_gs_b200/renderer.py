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
# it in Python, networks/sk_gs.py:1230-1231).  Like the reference, the state between forward and backward travels in
# three uint8 tensors (geomBuffer, binningBuffer, imgBuffer = this library's geom / binning / img arenas) plus the
# integer `num_rendered`; the binning arena is sized for exactly num_rendered entries, so backward can rebuild the arena
# layout from its arguments alone - nothing Python-side is attached to the tensors.
# Deviations: `extras` (per-Gaussian extra channels) and `colmap=False` (the in-tree row-major
# camera convention of gaussian_preprocess.cu) are not on the SK_GS hot path and are rejected loudly.
# ------------------------------------------------------------------------------------------------------------------
def rasterize_gaussians_b2(image_height, image_width, tanfovx, tanfovy, sh_degree, scale_modifier, prefiltered, debug,
                           colmap, viewmatrix, projmatrix, campos, means3D, opacities, sh, scales, rotations, extras,
                           colors_precomp, cov3Ds_precomp):
    """-> (num_rendered, color[3,H,W], opacity[H,W], radii[P], geomBuffer, binningBuffer, imgBuffer, out_extras)."""
    import ctypes as C

    from . import _lib
    from . import diff_gaussian_rasterization as DGR
    if extras is not None and extras.numel() > 0:
        raise RuntimeError('extras are not supported by sk_gs_b200 (not on the SK_GS hot path)')
    if not colmap:
        raise RuntimeError('only the colmap/upstream convention is implemented (every shipped config uses it)')
    if not means3D.is_cuda:
        raise RuntimeError('means3D must be a CUDA tensor (sk_gs_b200 has no CPU path)')
    if means3D.ndim != 2 or means3D.shape[1] != 3:
        raise RuntimeError('means3D must have dimensions (num_points, 3)')
    L = _lib.lib()
    dev = means3D.device
    f = DGR._f32c
    absent = DGR._absent
    means3D, opacities = f(means3D), f(opacities)
    sh = None if absent(sh) else f(sh)
    colors_precomp = None if absent(colors_precomp) else f(colors_precomp)
    scales = None if absent(scales) else f(scales)
    rotations = None if absent(rotations) else f(rotations)
    cov3Ds_precomp = None if absent(cov3Ds_precomp) else f(cov3Ds_precomp)
    P, M = means3D.shape[0], (0 if sh is None else int(sh.shape[1]))
    H, W = int(image_height), int(image_width)
    rs = GaussianRasterizationSettings(H, W, tanfovx, tanfovy, None, scale_modifier, viewmatrix, projmatrix, sh_degree,
                                       campos, prefiltered, debug)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev)
        st = stream.cuda_stream
        s, keep = DGR._make_settings(rs, dev, quat_wxyz=False)
        lay0 = DGR.layout_query(P, W, H, 0)
        geom = torch.empty(lay0.geom_bytes, dtype=torch.uint8, device=dev)
        img = torch.empty(lay0.img_bytes, dtype=torch.uint8, device=dev)
        radii = torch.empty(P, dtype=torch.int32, device=dev)
        color = torch.empty(3, H, W, device=dev)
        depth = torch.empty(H, W, device=dev)
        alpha = torch.empty(H, W, device=dev)
        words = DGR._pinned_words(dev)
        _lib.check(L.skgs_raster_forward_geometry(
            C.byref(s), P, M, _lib.ptr(means3D), _lib.ptr(sh), _lib.ptr(colors_precomp), _lib.ptr(opacities),
            _lib.ptr(scales), _lib.ptr(rotations), _lib.ptr(cov3Ds_precomp), geom.data_ptr(), radii.data_ptr(), None, 0,
            img.data_ptr(), words.data_ptr(), st), 'skgs_raster_forward_geometry')
        stream.synchronize()  # the reference returns num_rendered as a host integer as well (blocking cudaMemcpy)
        num_rendered = int(words[0].item()) & 0xffffffff
        R_cap = max(num_rendered, 1)
        lay = DGR.layout_query(P, W, H, R_cap)
        binning = torch.empty(lay.binning_bytes, dtype=torch.uint8, device=dev)
        _lib.check(L.skgs_raster_forward_render(
            C.byref(s), P, geom.data_ptr(), binning.data_ptr(), R_cap, num_rendered, img.data_ptr(), radii.data_ptr(),
            0, color.data_ptr(), depth.data_ptr(), alpha.data_ptr(), None, st), 'skgs_raster_forward_render')
    return num_rendered, color, alpha, radii, geom, binning, img, None


def rasterize_gaussians_backward_b2(scale_modifier, tanfovx, tanfovy, sh_degree, debug, colmap, viewmatrix, projmatrix,
                                    campos, means3D, colors, extras, scales, rotations, cov3D, sh, R, radii,
                                    out_opacity, dL_dcolor, dL_dopacity, dL_dextra, grad_means2D, grad_conic,
                                    grad_opacity, geomBuffer, binningBuffer, imgBuffer, _debug_flags: int = 0):
    """-> (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations, dL_dextras)
    like the reference.  The three buffers are the uint8 tensors returned by rasterize_gaussians_b2; everything else
    backward needs is among its arguments (the image size comes from dL_dcolor)."""
    from . import diff_gaussian_rasterization as DGR
    if not colmap:
        raise RuntimeError('only the colmap/upstream convention is implemented (every shipped config uses it)')
    if geomBuffer.dtype != torch.uint8 or binningBuffer.dtype != torch.uint8 or imgBuffer.dtype != torch.uint8:
        raise RuntimeError('geomBuffer / binningBuffer / imgBuffer must be the uint8 tensors of rasterize_gaussians')
    f, absent = DGR._f32c, DGR._absent
    dev = means3D.device
    means3D = f(means3D)
    sh = None if absent(sh) else f(sh)
    colors = None if absent(colors) else f(colors)
    scales = None if absent(scales) else f(scales)
    rotations = None if absent(rotations) else f(rotations)
    cov3D = None if absent(cov3D) else f(cov3D)
    P, M = means3D.shape[0], (0 if sh is None else int(sh.shape[1]))
    H, W = int(dL_dcolor.shape[-2]), int(dL_dcolor.shape[-1])
    rs = GaussianRasterizationSettings(H, W, tanfovx, tanfovy, None, scale_modifier, viewmatrix, projmatrix, sh_degree,
                                       campos, False, debug)
    s, keep = DGR._make_settings(rs, dev, quat_wxyz=False)
    R_cap = max(int(R), 1)
    lay = DGR.layout_query(P, W, H, R_cap)
    if geomBuffer.numel() < lay.geom_bytes or binningBuffer.numel() < lay.binning_bytes or \
            imgBuffer.numel() < lay.img_bytes:
        raise RuntimeError('buffers do not belong to a forward call of this shape (P, image size, num_rendered)')
    state = DGR.RasterState(s, keep + (means3D, sh, colors, None, scales, rotations, cov3D), P, M, R_cap, geomBuffer,
                            binningBuffer, imgBuffer, radii, lay, int(R))
    dA = None if dL_dopacity is None else dL_dopacity.reshape(1, H, W)
    g = DGR.rasterize_backward(state, dL_dcolor, None, dA, debug_flags=_debug_flags)
    return (g['means2D'], g['colors_precomp'], g['opacities'], g['means3D'], g['cov3D_precomp'], g['shs'], g['scales'],
            g['rotations'], None)
