"""Drop-in for the un-vendored `diff_gaussian_rasterization` module the reference imports
(/root/reference/networks/renderer/gaussian_render_origin.py:7, gui.py:535; boundary B1 of SURVEY.md 8b).

Same names and call shapes: `GaussianRasterizationSettings(image_height, image_width, tanfovx, tanfovy, bg,
scale_modifier, viewmatrix, projmatrix, sh_degree, campos, prefiltered, debug)` as built at
networks/gaussian_splatting.py:271-284, and `GaussianRasterizer(raster_settings)(means3D, means2D, shs, colors_precomp,
opacities, scales, rotations (w,x,y,z), cov3D_precomp)` returning the 4-tuple `(image[3,H,W], radii int32[P],
depth[1,H,W], alpha[1,H,W])` accepted at gaussian_render_origin.py:53-54.  The gradient of the screen-space means is
delivered as the gradient of the `means2D` input ([P,3], z = 0), exactly how densification consumes it
(networks/gaussian_splatting.py:503-513).

Everything runs in the hand-written sm_100a kernels of libskgs_b200.so on torch's CURRENT stream; there is no
PyTorch/CPU fallback.  Host synchronisation: none in the kernels; the wrapper waits only on an event recorded right
after the preprocess+scan stage (to learn R and validate the binning capacity) while the rest of the forward is
already queued - the GPU never idles, unlike the reference's blocking cudaMemcpy
(my_ext/_C/src/nerf/gaussian_rasterizer_forward.cu:208-209).
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple, Optional

import torch
from torch import Tensor, nn

from . import _lib


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: Tensor
    scale_modifier: float
    viewmatrix: Tensor
    projmatrix: Tensor
    sh_degree: int
    campos: Tensor
    prefiltered: bool = False
    debug: bool = False


# ------------------------------------------------------------------------------------------------- capacity policy
class _Capacity:
    """Binning-arena capacity estimate per (device, P, W, H): last observed R with 25 % head room."""

    def __init__(self):
        self.last = {}
        self.growth = 1.25

    def get(self, key):
        return self.last.get(key)

    def put(self, key, R):
        self.last[key] = int(R)


_capacity = _Capacity()
_pinned = {}


def new_header_words() -> Tensor:
    """A pinned host int32[4] that a fixed-capacity forward (`rasterize_forward(..., fixed_capacity=, header_words=)`)
    fills asynchronously with {R, num_visible, -, overflow}.  Every captured CUDA graph owns one, so that the overflow
    flag of one graph is never confused with another's (allocate it BEFORE stream capture starts)."""
    return torch.zeros(4, dtype=torch.int32).pin_memory()


def capacity_known(device, P: int, W: int, H: int) -> bool:
    """Has a forward of this shape run on `device` before (so that the binning arena can be sized up front)?"""
    device = torch.device(device)
    return _capacity.get((device.index, P, W, H)) is not None


def last_header_words(device) -> Tensor:
    """Pinned host int32[4] = {R, num_visible, -, overflow} of the most recent EAGER forward on `device` (valid after
    that forward has completed on the device).  Fixed-capacity (graph) calls write to their own buffer instead."""
    return _pinned_words(torch.device(device))


def _pinned_words(device) -> Tensor:
    key = device.index if device.index is not None else torch.cuda.current_device()
    buf = _pinned.get(key)  # one per device: allocating pinned memory is illegal during stream capture
    if buf is None:
        buf = torch.zeros(4, dtype=torch.int32).pin_memory()
        _pinned[key] = buf
    return buf


def _f32c(t: Optional[Tensor]) -> Optional[Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _absent(t: Optional[Tensor]) -> bool:
    """The reference encodes 'absent' both as None and as an empty tensor (networks/renderer/gaussian_render.py:257-267)."""
    return t is None or t.numel() == 0


class RasterState(NamedTuple):
    """Opaque per-call state kept for backward (the analogue of geomBuffer/binningBuffer/imgBuffer + num_rendered,
    networks/renderer/gaussian_render.py:19-31,112-117)."""
    settings: object
    keep: tuple
    P: int
    M: int
    R_cap: int
    geom: Tensor
    binning: Tensor
    img: Tensor
    radii: Tensor
    layout: object
    num_rendered: int

    def header(self) -> '_lib.RasterHeader':
        """Host copy of the device header (synchronises the device)."""
        raw = bytes(self.geom[int(self.layout.header):int(self.layout.header) + C.sizeof(_lib.RasterHeader)].cpu().numpy())
        return _lib.RasterHeader.from_buffer_copy(raw)

    def sorted_lists(self):
        """(keys uint64-as-int64 [R], point_list int32 [R]) sorted by (tile, depth), as the per-tile sort left them in
        layout.keys / layout.vals (synchronises the device).  R = min(num_rendered, R_cap)."""
        h = self.header()
        lay = self.layout
        R = min(int(h.num_rendered), int(self.R_cap))
        ko, vo = lay.keys, lay.vals
        keys = self.binning[ko:ko + 8 * R].view(torch.int64)
        vals = self.binning[vo:vo + 4 * R].view(torch.int32)
        return keys, vals

    @property
    def overflow_ptr(self) -> int:
        """Device address of header.overflow (uint32) of this call: non-zero <=> R exceeded the binning capacity."""
        return self.geom.data_ptr() + int(self.layout.header) + 12


def _make_settings(rs: GaussianRasterizationSettings, device, quat_wxyz: bool, debug_flags: int = 0):
    view, proj, campos = _f32c(rs.viewmatrix.to(device)), _f32c(rs.projmatrix.to(device)), _f32c(rs.campos.to(device))
    bg = None if rs.bg is None else _f32c(rs.bg.to(device))
    s = _lib.RasterSettings(int(rs.image_height), int(rs.image_width), float(rs.tanfovx), float(rs.tanfovy),
                            float(rs.scale_modifier), int(rs.sh_degree), int(quat_wxyz), int(bool(rs.prefiltered)),
                            int(bool(rs.debug)) | debug_flags, view.data_ptr(), proj.data_ptr(), campos.data_ptr(),
                            None if bg is None else bg.data_ptr())
    return s, (view, proj, campos, bg)


def layout_query(P: int, W: int, H: int, R_cap: int):
    lay = _lib.RasterLayout()
    _lib.check(_lib.lib().skgs_raster_layout_query(P, W, H, R_cap, C.byref(lay)), 'skgs_raster_layout_query')
    return lay


def rasterize_forward(rs: GaussianRasterizationSettings, means3D, opacities, shs=None, colors_precomp=None,
                      scales=None, rotations=None, cov3D_precomp=None, quat_wxyz: bool = True, debug_flags: int = 0,
                      fixed_capacity: Optional[int] = None, header_words: Optional[Tensor] = None,
                      deform: Optional[dict] = None):
    """Non-autograd forward.  Returns (color, depth, alpha, radii, RasterState).

    `deform` (sk_gs_b200.pipeline.HotPath): run the FUSED per-Gaussian forward - skinning + assembly + preprocess + key
    emission in one kernel (skgs_deform_forward_geometry).  means3D / opacities / scales / rotations are then OUTPUT
    buffers the kernel fills from deform['xyz' | 'scaling' | 'rotation' | 'opacity'] and the skeleton deform['sk'];
    needs a known capacity (`fixed_capacity`, or a previous call of the same shape: `capacity_known`).

    `fixed_capacity`: size the binning arena for exactly that many (Gaussian, tile) pairs and never look at R on the host
    (no wait at all: CUDA-graph capturable).  It is a property of THIS call - nothing process-wide changes.  If R exceeds
    it the image of the call is INVALID: `header.overflow` = 1 on the device (RasterState.overflow_ptr, which
    skgs_adam_step can be told to honour) and in `header_words[3]` once the call has completed.  The caller owns the
    check: see HotPath.overflowed() / TrainLoop.replay()."""
    L = _lib.lib()
    if not means3D.is_cuda:
        raise RuntimeError('means3D must be a CUDA tensor (sk_gs_b200 has no CPU path)')
    device = means3D.device
    means3D = _f32c(means3D)
    if means3D.ndim != 2 or means3D.shape[1] != 3:
        raise RuntimeError('means3D must have dimensions (num_points, 3)')  # gaussian_rasterizer_forward.cu:271-273
    P = means3D.shape[0]
    shs = None if _absent(shs) else _f32c(shs)
    colors_precomp = None if _absent(colors_precomp) else _f32c(colors_precomp)
    scales = None if _absent(scales) else _f32c(scales)
    rotations = None if _absent(rotations) else _f32c(rotations)
    cov3D_precomp = None if _absent(cov3D_precomp) else _f32c(cov3D_precomp)
    opacities = _f32c(opacities)
    M = 0 if shs is None else int(shs.shape[1])
    W, H = int(rs.image_width), int(rs.image_height)
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream(device)
        st = stream.cuda_stream
        s, keep = _make_settings(rs, device, quat_wxyz, debug_flags)
        lay0 = layout_query(P, W, H, 0)
        geom = torch.empty(lay0.geom_bytes, dtype=torch.uint8, device=device)
        img = torch.empty(lay0.img_bytes, dtype=torch.uint8, device=device)
        radii = torch.empty(P, dtype=torch.int32, device=device)
        color = torch.empty(3, H, W, dtype=torch.float32, device=device)
        depth = torch.empty(1, H, W, dtype=torch.float32, device=device)
        alpha = torch.empty(1, H, W, dtype=torch.float32, device=device)
        fixed = None if fixed_capacity is None else int(fixed_capacity)
        if fixed is not None and header_words is None:
            raise RuntimeError('fixed_capacity needs header_words=new_header_words(): somebody has to own the overflow flag')
        # {R, num_visible, -, overflow}; in fixed mode it is written but never waited on
        words = header_words if fixed is not None else _pinned_words(device)
        key = (device.index, P, W, H)
        est = _capacity.get(key)
        R_known = None

        def geometry(binning, R_cap):
            if deform is not None:
                if binning is None:
                    raise RuntimeError('the fused forward needs a known binning capacity (capacity_known / fixed_capacity)')
                d = deform
                _lib.check(L.skgs_deform_forward_geometry(
                    C.byref(d['sk']), C.byref(s), P, M, d['xyz'].data_ptr(), d['scaling'].data_ptr(),
                    d['rotation'].data_ptr(), d['opacity'].data_ptr(), _lib.ptr(shs), means3D.data_ptr(),
                    scales.data_ptr(), rotations.data_ptr(), opacities.data_ptr(), d['d_rot'].data_ptr(),
                    d['weights'].data_ptr(), d['indices'].data_ptr(), d['sk_T'].data_ptr(), d['workspace'].data_ptr(),
                    geom.data_ptr(), radii.data_ptr(), binning.data_ptr(), R_cap, img.data_ptr(), words.data_ptr(), st),
                    'skgs_deform_forward_geometry')
                return
            _lib.check(L.skgs_raster_forward_geometry(
                C.byref(s), P, M, _lib.ptr(means3D), _lib.ptr(shs), _lib.ptr(colors_precomp), _lib.ptr(opacities),
                _lib.ptr(scales), _lib.ptr(rotations), _lib.ptr(cov3D_precomp), geom.data_ptr(), radii.data_ptr(),
                _lib.ptr(binning), R_cap, img.data_ptr(), words.data_ptr(), st), 'skgs_raster_forward_geometry')

        def render(binning, R_cap, hint, keys_emitted):
            _lib.check(L.skgs_raster_forward_render(
                C.byref(s), P, geom.data_ptr(), binning.data_ptr(), R_cap, int(hint), img.data_ptr(), radii.data_ptr(),
                int(keys_emitted), color.data_ptr(), depth.data_ptr(), alpha.data_ptr(),
                words.data_ptr() if fixed is not None else None, st), 'skgs_raster_forward_render')

        if fixed is not None or est is not None:
            # capacity known up front (a captured graph's fixed one, or the estimate from the previous call of this
            # shape): preprocess, scan and key emission are ONE kernel
            R_cap = fixed if fixed is not None else max(int(est * _capacity.growth) + 4096, 4096)
            lay = layout_query(P, W, H, R_cap)
            binning = torch.empty(lay.binning_bytes, dtype=torch.uint8, device=device)
            geometry(binning, R_cap)
            if fixed is None:
                ev = torch.cuda.Event()
                ev.record(stream)  # the header words are on their way; everything below is queued behind them
            render(binning, R_cap, est if est is not None else R_cap, True)
            if fixed is not None:
                R_known = -1
            else:
                ev.synchronize()  # preprocess finished long ago; the GPU is busy with the rest of the forward
                R_known = int(words[0].item()) & 0xffffffff
                _capacity.put(key, R_known)
        else:  # first call for this shape: one blocking read of R, like the reference does on every call
            geometry(None, 0)
            stream.synchronize()
            R_known = int(words[0].item()) & 0xffffffff
            _capacity.put(key, R_known)
            R_cap = -1  # forces the sizing pass below
        while fixed is None and R_known > R_cap:  # first call, or under-estimated: size the arena from R, emit + render
            R_cap = max(int(R_known * _capacity.growth) + 4096, 4096)
            lay = layout_query(P, W, H, R_cap)
            binning = torch.empty(lay.binning_bytes, dtype=torch.uint8, device=device)
            render(binning, R_cap, R_known, False)
    state = RasterState(s, keep + (means3D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp), P, M,
                        R_cap, geom, binning, img, radii, lay, R_known)
    return color, depth, alpha, radii, state


def rasterize_backward(state: RasterState, dL_dcolor, dL_ddepth=None, dL_dalpha=None, out=None, debug_flags: int = 0):
    """Returns dict of gradients (None where the input was absent).  `out` may supply preallocated, contiguous fp32
    tensors for any of the keys (e.g. views of a flat all-reduce arena); every output is fully overwritten.
    `debug_flags` are OR-ed into settings.debug for this call (bit 2: the reference extension's T_final recovery)."""
    L = _lib.lib()
    (view, proj, campos, bg, means3D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp) = state.keep
    device = means3D.device
    P, M = state.P, state.M
    dL_dcolor = _f32c(dL_dcolor)
    dL_ddepth = None if dL_ddepth is None else _f32c(dL_ddepth)
    dL_dalpha = None if dL_dalpha is None else _f32c(dL_dalpha)

    def new(*shape):
        return torch.empty(*shape, dtype=torch.float32, device=device)

    out = out or {}

    def pick(name, *shape):
        t = out.get(name)
        if t is None:
            return new(*shape)
        assert t.is_contiguous() and t.dtype == torch.float32 and t.numel() == torch.Size(shape).numel(), name
        return t

    g = {
        'means3D': pick('means3D', P, 3), 'means2D': pick('means2D', P, 3), 'opacities': pick('opacities', P, 1),
        'shs': None if shs is None else pick('shs', P, M, 3),
        'colors_precomp': None if colors_precomp is None else pick('colors_precomp', P, 3),
        'scales': None if scales is None else pick('scales', P, 3),
        'rotations': None if rotations is None else pick('rotations', P, 4),
        'cov3D_precomp': None if cov3D_precomp is None else pick('cov3D_precomp', P, 6),
    }
    settings = state.settings
    if debug_flags:
        settings = _lib.RasterSettings.from_buffer_copy(bytes(state.settings))
        settings.debug |= int(debug_flags)
    with torch.cuda.device(device):
        st = torch.cuda.current_stream(device).cuda_stream
        _lib.check(L.skgs_raster_backward(
            C.byref(settings), P, M, _lib.ptr(means3D), _lib.ptr(shs), _lib.ptr(colors_precomp), _lib.ptr(scales),
            _lib.ptr(rotations), _lib.ptr(cov3D_precomp), state.radii.data_ptr(), state.geom.data_ptr(),
            state.binning.data_ptr(), state.R_cap, state.img.data_ptr(), dL_dcolor.data_ptr(), _lib.ptr(dL_ddepth),
            _lib.ptr(dL_dalpha), _lib.ptr(g['means3D']), _lib.ptr(g['means2D']), _lib.ptr(g['shs']),
            _lib.ptr(g['colors_precomp']), _lib.ptr(g['opacities']), _lib.ptr(g['scales']), _lib.ptr(g['rotations']),
            _lib.ptr(g['cov3D_precomp']), st), 'skgs_raster_backward')
    return g


def rasterize_assemble_backward(state: RasterState, scaling, rotation, opacity_logit, d_rot, dL_dcolor, dL_ddepth=None,
                                dL_dalpha=None, out=None, debug_flags: int = 0, private_lbs_inputs: bool = False):
    """Rasterizer backward with the assembly backward fused into its per-Gaussian kernel
    (skgs_raster_assemble_backward).  Returns {'xyz' (= dL/dpoints = dL/d_xyz), 'means2D', 'shs', 'scaling', 'rotation'
    (= dL/d_rot as well), 'opacity', 'dd_scale', 'dd_xyz', 'dd_rot'}.  `out` may supply preallocated tensors (e.g. arena
    views).  'dd_xyz' / 'dd_rot' are what the LBS backward reads: the 'xyz' / 'rotation' tensors themselves, or - with
    `private_lbs_inputs` - separate copies the kernel writes as well (needed when 'xyz' / 'rotation' live in an arena
    whose all-reduce runs concurrently with the LBS backward)."""
    L = _lib.lib()
    (view, proj, campos, bg, means3D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp) = state.keep
    if shs is None or scales is None or rotations is None:
        raise RuntimeError('the fused backward needs SH colours and scale / rotation inputs')
    device = means3D.device
    P, M = state.P, state.M
    dL_dcolor = _f32c(dL_dcolor)
    dL_ddepth = None if dL_ddepth is None else _f32c(dL_ddepth)
    dL_dalpha = None if dL_dalpha is None else _f32c(dL_dalpha)
    out = out or {}

    def pick(name, *shape):
        t = out.get(name)
        if t is None:
            return torch.empty(*shape, dtype=torch.float32, device=device)
        assert t.is_contiguous() and t.dtype == torch.float32 and t.numel() == torch.Size(shape).numel(), name
        return t

    g = {'xyz': pick('xyz', P, 3), 'means2D': pick('means2D', P, 3), 'shs': pick('shs', P, M, 3),
         'scaling': pick('scaling', P, 3), 'rotation': pick('rotation', P, 4), 'opacity': pick('opacity', P, 1),
         'dd_scale': pick('dd_scale', P, 3)}
    g['dd_xyz'] = pick('dd_xyz', P, 3) if private_lbs_inputs else g['xyz']
    g['dd_rot'] = pick('dd_rot', P, 4) if private_lbs_inputs else g['rotation']
    settings = state.settings
    if debug_flags:
        settings = _lib.RasterSettings.from_buffer_copy(bytes(state.settings))
        settings.debug |= int(debug_flags)
    with torch.cuda.device(device):
        st = torch.cuda.current_stream(device).cuda_stream
        _lib.check(L.skgs_raster_assemble_backward(
            C.byref(settings), P, M, _lib.ptr(means3D), _lib.ptr(shs), _lib.ptr(scales), _lib.ptr(rotations),
            state.radii.data_ptr(), state.geom.data_ptr(), state.binning.data_ptr(), state.R_cap, state.img.data_ptr(),
            dL_dcolor.data_ptr(), _lib.ptr(dL_ddepth), _lib.ptr(dL_dalpha), scaling.data_ptr(), rotation.data_ptr(),
            opacity_logit.data_ptr(), _lib.ptr(d_rot), g['xyz'].data_ptr(), g['means2D'].data_ptr(),
            g['shs'].data_ptr(), g['scaling'].data_ptr(), g['rotation'].data_ptr(), g['opacity'].data_ptr(),
            g['dd_scale'].data_ptr(), g['dd_xyz'].data_ptr() if private_lbs_inputs else None,
            g['dd_rot'].data_ptr() if private_lbs_inputs else None, st), 'skgs_raster_assemble_backward')
    return g


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                raster_settings, quat_wxyz):
        color, depth, alpha, radii, state = rasterize_forward(raster_settings, means3D, opacities, shs, colors_precomp,
                                                              scales, rotations, cov3D_precomp, quat_wxyz)
        ctx.state = state
        ctx.mark_non_differentiable(radii)
        return color, radii, depth, alpha

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_color, g_radii, g_depth, g_alpha):
        state = ctx.state
        if state.P == 0:
            return (None,) * 10
        if g_color is None:
            g_color = torch.zeros(3, state.settings.image_height, state.settings.image_width, device=state.radii.device)
        g = rasterize_backward(state, g_color, g_depth, g_alpha)
        ctx.state = None
        return (g['means3D'], g['means2D'], g['shs'], g['colors_precomp'], g['opacities'], g['scales'],
                g['rotations'], g['cov3D_precomp'], None, None)


def rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings, quat_wxyz: bool = True):
    return _RasterizeGaussians.apply(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings, quat_wxyz)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions: Tensor) -> Tensor:
        """Frustum test of the upstream module: p_view.z > 0.2 (gaussian_preprocess_colmap.cu:63-82)."""
        with torch.no_grad():
            V = self.raster_settings.viewmatrix.to(positions)
            z = positions @ V[:3, 2] + V[3, 2]
            return z > 0.2

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                   self.raster_settings, True)
