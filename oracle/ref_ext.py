"""oracle/ref_ext.py - thin harness around the REFERENCE's own rasterizer extension (oracle/_ref/_ref_raster*.so,
compiled by oracle/build_ref.sh from /root/reference/my_ext/_C/src/nerf/gaussian_*.cu, unmodified).

TEST / BENCH INFRASTRUCTURE ONLY.  Used (a) as a second, GPU-side parity reference (tests/test_gpu_reference_ext.py)
and (b) as the "reference extension" speed baseline of BASELINE.md section 2 (bench.py --impl reference).
The call sequence mirrors /root/reference/networks/renderer/gaussian_render.py:51-188 (`_RasterizeGaussians`) with
`colmap=True`, plus the Python-side background blend of networks/sk_gs.py:1230-1231.  FK + LBS for the reference arm
is the torch-op formulation in oracle/fk_lbs.py run on the GPU (lietorch / pytorch3d are not installable here).
"""
from __future__ import annotations

import glob
import importlib.util
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_mod = None


def available() -> bool:
    return bool(glob.glob(os.path.join(_HERE, '_ref', '_ref_raster*.so')))


def module():
    global _mod
    if _mod is None:
        paths = glob.glob(os.path.join(_HERE, '_ref', '_ref_raster*.so'))
        if not paths:
            raise RuntimeError('oracle/_ref/_ref_raster*.so not built (run oracle/build_ref.sh where /root/reference exists)')
        spec = importlib.util.spec_from_file_location('_ref_raster', paths[0])
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod


class _RefRasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, opacities, scales, rotations, H, W, tanfovx, tanfovy, sh_degree,
                scale_modifier, viewmatrix, projmatrix, campos):
        m = module()
        empty = torch.Tensor([])
        args = (H, W, tanfovx, tanfovy, sh_degree, scale_modifier, False, False, True, viewmatrix, projmatrix, campos,
                means3D, opacities, sh, scales, rotations, None, empty, empty)
        num_rendered, color, opacity, radii, geomBuffer, binningBuffer, imgBuffer, _ = m.rasterize_gaussians(*args)
        ctx.cfg = (tanfovx, tanfovy, sh_degree, scale_modifier, num_rendered)
        ctx.save_for_backward(viewmatrix, projmatrix, campos, means3D, scales, rotations, sh, geomBuffer,
                              binningBuffer, imgBuffer, radii, opacity)
        ctx.mark_non_differentiable(radii)
        return color, opacity, radii

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_color, g_opacity, _g_radii):
        m = module()
        tanfovx, tanfovy, sh_degree, scale_modifier, R = ctx.cfg
        (viewmatrix, projmatrix, campos, means3D, scales, rotations, sh, geomBuffer, binningBuffer, imgBuffer, radii,
         opacity) = ctx.saved_tensors
        empty = torch.Tensor([])
        if g_opacity is None:
            g_opacity = torch.zeros_like(opacity)
        out = m.rasterize_gaussians_backward(scale_modifier, tanfovx, tanfovy, sh_degree, False, True, viewmatrix,
                                             projmatrix, campos, means3D, empty, None, scales, rotations, empty, sh, R,
                                             radii, opacity, g_color.contiguous(), g_opacity.contiguous(), None, None,
                                             None, None, geomBuffer, binningBuffer, imgBuffer)
        dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations, _ = out
        return (dL_dmeans3D, dL_dmeans2D, dL_dsh, dL_dopacity, dL_dscales, dL_drotations) + (None,) * 9


def render(points, opacity, scales, rotations, sh_features, cam, sh_degree=3, scale_modifier=1.0):
    """Returns dict(images[3,H,W] with background, opacity[H,W], radii, viewspace_points) - reference semantics."""
    dev = points.device
    screenspace = torch.zeros_like(points, requires_grad=True) + 0
    try:
        screenspace.retain_grad()
    except Exception:  # noqa
        pass
    color, op, radii = _RefRasterize.apply(points, screenspace, sh_features, opacity, scales, rotations, cam.H, cam.W,
                                           cam.tanfovx, cam.tanfovy, sh_degree, scale_modifier,
                                           cam.viewmatrix.to(dev), cam.projmatrix.to(dev), cam.campos.to(dev))
    images = color + (1 - op[None]) * cam.bg.to(dev)[:, None, None]  # networks/sk_gs.py:1230-1231
    return dict(images=images, opacity=op, radii=radii, viewspace_points=screenspace)
