"""oracle/ref_ext.py - thin harness around the REFERENCE's own rasterizer extension (oracle/_ref/_ref_raster*.so,
compiled by oracle/build_ref.sh from /root/reference/my_ext/_C/src/nerf/gaussian_*.cu, unmodified).

TEST / BENCH INFRASTRUCTURE ONLY.  Used (a) as the GPU-side parity reference (tests/test_gpu_reference_ext.py) and
(b) as the "reference extension" speed baseline of BASELINE.md section 2 (bench.py --impl reference).
The call sequence mirrors /root/reference/networks/renderer/gaussian_render.py:51-188 (`_RasterizeGaussians`) with
`colmap=True`, plus the Python-side background blend of networks/sk_gs.py:1230-1231.  FK + LBS for the reference arm
is the torch-op formulation in oracle/fk_lbs.py run on the GPU (lietorch / pytorch3d are not installable here).

Two builds of the same unmodified sources exist (oracle/build_ref.sh): the default one (FMA contraction on, what a user
runs: speed baseline, tolerance-level parity) and `nofma=True` (-fmad=false: the arithmetic the source text states; the
bit-exact target for radii, sorted keys, point_list and tile ranges).
"""
from __future__ import annotations

import glob
import importlib.util
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_mods = {}


def _name(nofma: bool) -> str:
    return '_ref_raster_nofma' if nofma else '_ref_raster'


def available(nofma: bool = False) -> bool:
    return bool(glob.glob(os.path.join(_HERE, '_ref', _name(nofma) + '.*so')))


def module(nofma: bool = False):
    name = _name(nofma)
    if name not in _mods:
        paths = glob.glob(os.path.join(_HERE, '_ref', name + '.*so'))
        if not paths:
            raise RuntimeError(f'oracle/_ref/{name}*.so not built (run oracle/build_ref.sh where /root/reference exists)')
        spec = importlib.util.spec_from_file_location(name, paths[0])
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _mods[name] = mod
    return _mods[name]


class _RefRasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, opacities, scales, rotations, H, W, tanfovx, tanfovy, sh_degree,
                scale_modifier, viewmatrix, projmatrix, campos, nofma, keep):
        m = module(nofma)
        empty = torch.Tensor([])
        args = (H, W, tanfovx, tanfovy, sh_degree, scale_modifier, False, False, True, viewmatrix, projmatrix, campos,
                means3D, opacities, sh, scales, rotations, None, empty, empty)
        num_rendered, color, opacity, radii, geomBuffer, binningBuffer, imgBuffer, _ = m.rasterize_gaussians(*args)
        ctx.cfg = (tanfovx, tanfovy, sh_degree, scale_modifier, num_rendered, nofma)
        ctx.save_for_backward(viewmatrix, projmatrix, campos, means3D, scales, rotations, sh, geomBuffer,
                              binningBuffer, imgBuffer, radii, opacity)
        ctx.mark_non_differentiable(radii)
        if keep is not None:
            keep.update(num_rendered=num_rendered, geomBuffer=geomBuffer, binningBuffer=binningBuffer,
                        imgBuffer=imgBuffer, P=means3D.shape[0], H=H, W=W)
        return color, opacity, radii

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_color, g_opacity, _g_radii):
        tanfovx, tanfovy, sh_degree, scale_modifier, R, nofma = ctx.cfg
        m = module(nofma)
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
        return (dL_dmeans3D, dL_dmeans2D, dL_dsh, dL_dopacity, dL_dscales, dL_drotations) + (None,) * 11


def render(points, opacity, scales, rotations, sh_features, cam, sh_degree=3, scale_modifier=1.0, nofma=False,
           keep_buffers=False):
    """Returns dict(images[3,H,W] with background, opacity[H,W], radii, viewspace_points) - reference semantics.
    With `keep_buffers` the dict also carries 'buffers' = the reference's num_rendered / geomBuffer / binningBuffer /
    imgBuffer (parse with `parse_buffers`)."""
    dev = points.device
    screenspace = torch.zeros_like(points, requires_grad=True) + 0
    try:
        screenspace.retain_grad()
    except Exception:  # noqa
        pass
    keep = {} if keep_buffers else None
    color, op, radii = _RefRasterize.apply(points, screenspace, sh_features, opacity, scales, rotations, cam.H, cam.W,
                                           cam.tanfovx, cam.tanfovy, sh_degree, scale_modifier,
                                           cam.viewmatrix.to(dev), cam.projmatrix.to(dev), cam.campos.to(dev), nofma, keep)
    images = color + (1 - op[None]) * cam.bg.to(dev)[:, None, None]  # networks/sk_gs.py:1230-1231
    out = dict(images=images, opacity=op, radii=radii, viewspace_points=screenspace, color_nobg=color)
    if keep_buffers:
        out['buffers'] = keep
    return out


# --------------------------------------------------------------------------------------------------------------------
# The reference's internal state, read back for bit-exact comparison.  Layout = the `obtain()` sequences of
# GeometryState / BinningState / ImageState::fromChunk (my_ext/_C/src/nerf/gaussian_rasterizer_imp.cu:39-73,
# include/gaussian_render.h:111-117): consecutive arrays, each start rounded up to 128 bytes of the ABSOLUTE address.
# --------------------------------------------------------------------------------------------------------------------
def _carve(buf: torch.Tensor, specs):
    """specs: [(name, numpy dtype, element count)] in obtain() order -> {name: numpy array (host copy)}."""
    base = buf.data_ptr()
    host = buf.cpu().numpy()
    out = {}
    addr = base
    for name, dtype, count in specs:
        addr = (addr + 127) & ~127
        nbytes = int(count) * np.dtype(dtype).itemsize
        if name is not None:
            out[name] = host[addr - base: addr - base + nbytes].view(dtype).copy()
        addr += nbytes
    return out


def parse_buffers(b) -> dict:
    """{'depths','clamped','means2D','cov3D','conic_opacity','rgb','tiles_touched',   (geometry, per Gaussian)
        'point_list','point_list_unsorted','keys','keys_unsorted',                     (binning, R entries)
        'n_contrib','ranges'}                                                          (image; ranges [tiles, 2])"""
    P, R, H, W = b['P'], int(b['num_rendered']), b['H'], b['W']
    out = _carve(b['geomBuffer'], [('depths', np.float32, P), ('clamped', np.uint8, 3 * P), (None, np.int32, P),
                                   ('means2D', np.float32, 2 * P), ('cov3D', np.float32, 6 * P),
                                   ('conic_opacity', np.float32, 4 * P), ('rgb', np.float32, 3 * P),
                                   ('tiles_touched', np.uint32, P)])
    out['means2D'] = out['means2D'].reshape(P, 2)
    out['cov3D'] = out['cov3D'].reshape(P, 6)
    out['conic_opacity'] = out['conic_opacity'].reshape(P, 4)
    out['rgb'] = out['rgb'].reshape(P, 3)
    out['clamped'] = out['clamped'].reshape(P, 3)
    if R > 0:
        out.update(_carve(b['binningBuffer'], [('point_list', np.uint32, R), ('point_list_unsorted', np.uint32, R),
                                               ('keys', np.uint64, R), ('keys_unsorted', np.uint64, R)]))
    else:
        out.update(point_list=np.zeros(0, np.uint32), point_list_unsorted=np.zeros(0, np.uint32),
                   keys=np.zeros(0, np.uint64), keys_unsorted=np.zeros(0, np.uint64))
    N = H * W
    img = _carve(b['imgBuffer'], [('n_contrib', np.uint32, N), ('ranges', np.uint32, 2 * N)])
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    out['n_contrib'] = img['n_contrib'].reshape(H, W)
    out['ranges'] = img['ranges'].reshape(N, 2)[:tiles]
    return out
