"""oracle/raster.py - numpy/ctypes front end of the C restatement in raster_oracle.c.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference leg.  The product package sk_gs_b200/ never imports it.

The API mirrors the stages of the reference rasterizer (my_ext/_C/src/nerf/gaussian_rasterizer_forward.cu:157-250,
gaussian_rasterizer_backwrad.cu:148-198): preprocess -> scan -> duplicate/sort/ranges -> composite, and back.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, 'liboracle_raster.so')
_lib = None

f32p = np.ctypeslib.ndpointer(np.float32, flags='C_CONTIGUOUS')
i32p = np.ctypeslib.ndpointer(np.int32, flags='C_CONTIGUOUS')
u32p = np.ctypeslib.ndpointer(np.uint32, flags='C_CONTIGUOUS')
u64p = np.ctypeslib.ndpointer(np.uint64, flags='C_CONTIGUOUS')
u8p = np.ctypeslib.ndpointer(np.uint8, flags='C_CONTIGUOUS')


def build(force: bool = False) -> str:
    """Compile raster_oracle.c with the flags in oracle/Makefile (gcc, -ffp-contract=off -mfma -fopenmp)."""
    src = os.path.join(_HERE, 'raster_oracle.c')
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(['make', '-C', _HERE, '-B', 'liboracle_raster.so'], check=True, capture_output=True)
    return _LIB_PATH


def _opt(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def lib():
    global _lib
    if _lib is None:
        with open('/proc/cpuinfo') as f:
            if ' fma' not in f.read():
                raise RuntimeError('oracle needs an FMA-capable x86-64 host (compiled with -mfma)')
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_exp_scalar.restype = C.c_float
        L.orc_exp_scalar.argtypes = [C.c_float]
        L.orc_higher_msb.restype = C.c_uint32
        L.orc_higher_msb.argtypes = [C.c_uint32]
        L.orc_scan.restype = C.c_uint64
        L.orc_scan.argtypes = [C.c_int, u32p, u32p]
        vp = C.c_void_p
        L.orc_preprocess_fwd.restype = None
        L.orc_preprocess_fwd.argtypes = [C.c_int, C.c_int, C.c_int, f32p, vp, C.c_float, vp, C.c_int, f32p, vp, vp, vp,
                                         f32p, f32p, f32p, C.c_int, C.c_int, C.c_float, C.c_float, i32p, f32p, f32p,
                                         f32p, f32p, f32p, u8p, u32p]
        L.orc_binning.restype = None
        L.orc_binning.argtypes = [C.c_int, C.c_uint64, f32p, f32p, i32p, u32p, C.c_int, C.c_int, u64p, u32p, u64p,
                                  u32p, u32p]
        L.orc_composite_fwd.restype = None
        L.orc_composite_fwd.argtypes = [C.c_int, C.c_int, u32p, u32p, f32p, f32p, f32p, f32p, f32p, f32p, f32p, f32p,
                                        u32p, f32p]
        L.orc_composite_bwd.restype = None
        L.orc_composite_bwd.argtypes = [C.c_int, C.c_int, C.c_int, u32p, u32p, f32p, f32p, f32p, f32p, f32p, u32p,
                                        f32p, f32p, vp, vp, f32p, f32p, f32p, f32p, f32p]
        L.orc_preprocess_bwd.restype = None
        L.orc_preprocess_bwd.argtypes = [C.c_int, C.c_int, C.c_int, f32p, i32p, vp, u8p, vp, vp, C.c_int, C.c_float,
                                         f32p, f32p, f32p, f32p, C.c_int, C.c_int, C.c_float, C.c_float, f32p, f32p,
                                         f32p, vp, f32p, f32p, vp, vp, vp]
        L.orc_sh_to_rgb.restype = None
        L.orc_sh_to_rgb.argtypes = [C.c_int, C.c_int, C.c_int, f32p, f32p, f32p, u8p]
        L.orc_cov3d.restype = None
        L.orc_cov3d.argtypes = [C.c_int, f32p, C.c_float, f32p, C.c_int, f32p]
        L.orc_cov2d.restype = None
        L.orc_cov2d.argtypes = [C.c_int, f32p, f32p, f32p, C.c_float, C.c_float, C.c_float, C.c_float, f32p]
        _lib = L
    return _lib


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(n: int):
    lib().orc_set_num_threads(int(n))


def orc_exp(x: np.ndarray) -> np.ndarray:
    L = lib()
    x = np.asarray(x, np.float32)
    return np.array([L.orc_exp_scalar(float(v)) for v in x.ravel()], np.float32).reshape(x.shape)


def sh_to_rgb(deg, shs, dirs):
    """shs [n, M, 3], dirs [n, 3] (not necessarily unit) -> (rgb [n,3] = max(SH + 0.5, 0), clamped [n,3])."""
    shs, dirs = f32(shs), f32(dirs)
    n, M = shs.shape[0], shs.shape[1]
    rgb, cl = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.uint8)
    lib().orc_sh_to_rgb(n, int(deg), M, shs, dirs, rgb, cl)
    return rgb, cl


def cov3d(scales, rots, mod=1.0, quat_wxyz=False):
    scales, rots = f32(scales), f32(rots)
    out = np.zeros((scales.shape[0], 6), np.float32)
    lib().orc_cov3d(scales.shape[0], scales, float(mod), rots, int(quat_wxyz), out)
    return out


def cov2d(means, cov6, viewmatrix, fx, fy, tanfovx, tanfovy):
    means, cov6 = f32(means), f32(cov6)
    out = np.zeros((means.shape[0], 3), np.float32)
    lib().orc_cov2d(means.shape[0], means, cov6, f32(viewmatrix).reshape(-1), float(fx), float(fy), float(tanfovx),
                    float(tanfovy), out)
    return out


def f32(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float32))


@dataclass
class Settings:
    """Same fields as diff_gaussian_rasterization.GaussianRasterizationSettings
    (networks/gaussian_splatting.py:271-284), numpy-typed."""
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: np.ndarray
    scale_modifier: float
    viewmatrix: np.ndarray  # [4,4] = Tw2v.T (row-major storage, indexed column-major by the kernels)
    projmatrix: np.ndarray  # [4,4] = (Tv2c @ Tw2v).T
    sh_degree: int
    campos: np.ndarray
    quat_wxyz: bool = True  # upstream convention at boundary B1


@dataclass
class Geom:
    radii: np.ndarray
    means2D: np.ndarray
    depths: np.ndarray
    cov3D: np.ndarray
    conic_opacity: np.ndarray
    rgb: np.ndarray
    clamped: np.ndarray
    tiles_touched: np.ndarray


@dataclass
class Bins:
    R: int
    offsets: np.ndarray
    keys_unsorted: np.ndarray
    vals_unsorted: np.ndarray
    keys: np.ndarray
    point_list: np.ndarray
    ranges: np.ndarray  # [tiles, 2]


def preprocess_fwd(s: Settings, means3D, opacities, scales=None, rotations=None, shs=None, colors_precomp=None,
                   cov3D_precomp=None) -> Geom:
    means3D = f32(means3D)
    P = means3D.shape[0]
    opac = f32(opacities).reshape(-1)
    scales, rotations, shs = f32(scales), f32(rotations), f32(shs)
    colors_precomp, cov3D_precomp = f32(colors_precomp), f32(cov3D_precomp)
    M = 0 if shs is None else shs.shape[1]
    g = Geom(np.zeros(P, np.int32), np.zeros((P, 2), np.float32), np.zeros(P, np.float32),
             np.zeros((P, 6), np.float32), np.zeros((P, 4), np.float32), np.zeros((P, 3), np.float32),
             np.zeros((P, 3), np.uint8), np.zeros(P, np.uint32))
    if colors_precomp is not None:
        g.rgb[...] = colors_precomp
    if cov3D_precomp is not None:
        g.cov3D[...] = cov3D_precomp
    lib().orc_preprocess_fwd(P, int(s.sh_degree), M, means3D, _opt(scales), float(s.scale_modifier), _opt(rotations),
                             int(s.quat_wxyz), opac, _opt(shs), _opt(cov3D_precomp), _opt(colors_precomp),
                             f32(s.viewmatrix).reshape(-1), f32(s.projmatrix).reshape(-1), f32(s.campos),
                             int(s.image_width), int(s.image_height), float(s.tanfovx), float(s.tanfovy), g.radii,
                             g.means2D, g.depths, g.cov3D, g.conic_opacity, g.rgb, g.clamped, g.tiles_touched)
    return g


def binning(s: Settings, g: Geom) -> Bins:
    P = g.radii.shape[0]
    W, H = int(s.image_width), int(s.image_height)
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    offsets = np.zeros(P, np.uint32)
    R = int(lib().orc_scan(P, g.tiles_touched, offsets)) if P > 0 else 0
    n = max(R, 1)
    b = Bins(R, offsets, np.zeros(n, np.uint64), np.zeros(n, np.uint32), np.zeros(n, np.uint64),
             np.zeros(n, np.uint32), np.zeros((tiles, 2), np.uint32))
    lib().orc_binning(P, R, g.means2D, g.depths, g.radii, offsets, W, H, b.keys_unsorted, b.vals_unsorted, b.keys,
                      b.point_list, b.ranges)
    for name in ('keys_unsorted', 'vals_unsorted', 'keys', 'point_list'):
        setattr(b, name, getattr(b, name)[:R])
    return b


@dataclass
class Image:
    color: np.ndarray  # [3,H,W]
    depth: np.ndarray  # [H,W]
    alpha: np.ndarray  # [H,W]
    n_contrib: np.ndarray
    final_T: np.ndarray


def composite_fwd(s: Settings, g: Geom, b: Bins) -> Image:
    W, H = int(s.image_width), int(s.image_height)
    img = Image(np.zeros((3, H, W), np.float32), np.zeros((H, W), np.float32), np.zeros((H, W), np.float32),
                np.zeros((H, W), np.uint32), np.zeros((H, W), np.float32))
    pl = b.point_list if b.R > 0 else np.zeros(1, np.uint32)
    lib().orc_composite_fwd(W, H, b.ranges, pl, g.means2D, g.conic_opacity, g.rgb, g.depths, f32(s.bg), img.color,
                            img.depth, img.alpha, img.n_contrib, img.final_T)
    return img


@dataclass
class GeomGrads:
    dL_dmean2D: np.ndarray  # [P,3]
    dL_dconic: np.ndarray  # [P,4] (x, y, -, w)
    dL_dopacity: np.ndarray  # [P]
    dL_dcolors: np.ndarray  # [P,3]
    dL_dz: np.ndarray  # [P]


def composite_bwd(s: Settings, g: Geom, b: Bins, img: Image, dL_dcolor, dL_ddepth=None, dL_dalpha=None) -> GeomGrads:
    W, H = int(s.image_width), int(s.image_height)
    P = g.radii.shape[0]
    gg = GeomGrads(np.zeros((P, 3), np.float32), np.zeros((P, 4), np.float32), np.zeros(P, np.float32),
                   np.zeros((P, 3), np.float32), np.zeros(P, np.float32))
    pl = b.point_list if b.R > 0 else np.zeros(1, np.uint32)
    dD, dA = f32(dL_ddepth), f32(dL_dalpha)
    lib().orc_composite_bwd(P, W, H, b.ranges, pl, g.means2D, g.conic_opacity, g.rgb, g.depths, f32(s.bg),
                            img.n_contrib, img.final_T, f32(dL_dcolor), _opt(dD), _opt(dA), gg.dL_dmean2D, gg.dL_dconic,
                            gg.dL_dopacity, gg.dL_dcolors, gg.dL_dz)
    return gg


@dataclass
class InputGrads:
    dL_dmeans3D: np.ndarray
    dL_dcov3D: np.ndarray
    dL_dsh: Optional[np.ndarray]
    dL_dscales: Optional[np.ndarray]
    dL_drotations: Optional[np.ndarray]
    dL_dopacity: np.ndarray
    dL_dcolors: np.ndarray
    dL_dmeans2D: np.ndarray


def preprocess_bwd(s: Settings, g: Geom, gg: GeomGrads, means3D, scales=None, rotations=None, shs=None,
                   cov3D_precomp=None) -> InputGrads:
    means3D = f32(means3D)
    P = means3D.shape[0]
    scales, rotations, shs = f32(scales), f32(rotations), f32(shs)
    M = 0 if shs is None else shs.shape[1]
    dmeans = np.zeros((P, 3), np.float32)
    dcov = np.zeros((P, 6), np.float32)
    dsh = None if shs is None else np.zeros((P, M, 3), np.float32)
    dsc = None if scales is None else np.zeros((P, 3), np.float32)
    drot = None if rotations is None else np.zeros((P, 4), np.float32)
    lib().orc_preprocess_bwd(P, int(s.sh_degree), M, means3D, g.radii, _opt(shs), g.clamped, _opt(scales),
                             _opt(rotations), int(s.quat_wxyz), float(s.scale_modifier), g.cov3D,
                             f32(s.viewmatrix).reshape(-1), f32(s.projmatrix).reshape(-1), f32(s.campos),
                             int(s.image_width), int(s.image_height), float(s.tanfovx), float(s.tanfovy),
                             gg.dL_dmean2D, gg.dL_dconic, gg.dL_dcolors, gg.dL_dz.ctypes.data_as(C.c_void_p), dmeans,
                             dcov, _opt(dsh), _opt(dsc), _opt(drot))
    return InputGrads(dmeans, dcov, dsh, dsc, drot, gg.dL_dopacity.copy(), gg.dL_dcolors.copy(), gg.dL_dmean2D.copy())


def render_forward(s: Settings, means3D, opacities, scales=None, rotations=None, shs=None, colors_precomp=None,
                   cov3D_precomp=None):
    """Whole forward: returns (Image, Geom, Bins) - the (image, radii, depth, alpha) 4-tuple of the upstream contract
    (networks/renderer/gaussian_render_origin.py:53-54) is (img.color, geom.radii, img.depth, img.alpha)."""
    g = preprocess_fwd(s, means3D, opacities, scales, rotations, shs, colors_precomp, cov3D_precomp)
    b = binning(s, g)
    img = composite_fwd(s, g, b)
    return img, g, b


def render_backward(s: Settings, g: Geom, b: Bins, img: Image, dL_dcolor, means3D, scales=None, rotations=None,
                    shs=None, cov3D_precomp=None, dL_ddepth=None, dL_dalpha=None) -> InputGrads:
    gg = composite_bwd(s, g, b, img, dL_dcolor, dL_ddepth, dL_dalpha)
    return preprocess_bwd(s, g, gg, means3D, scales, rotations, shs, cov3D_precomp)
