"""ctypes binding of libskgs_b200.so (the C ABI declared in include/skgs_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, a RuntimeError is raised
(the reference degrades to Python fallbacks on ImportError, my_ext/_C/__init__.py:100-126; this package must not).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('SKGS_LIB', os.path.join(_HERE, 'libskgs_b200.so'))  # SKGS_LIB: tuning builds (tools/)

c_f32p = C.c_void_p  # all device pointers travel as integers (tensor.data_ptr())


class RasterSettings(C.Structure):
    _fields_ = [
        ('image_height', C.c_int32), ('image_width', C.c_int32), ('tanfovx', C.c_float), ('tanfovy', C.c_float),
        ('scale_modifier', C.c_float), ('sh_degree', C.c_int32), ('quat_wxyz', C.c_int32),
        ('prefiltered', C.c_int32), ('debug', C.c_int32), ('viewmatrix', C.c_void_p), ('projmatrix', C.c_void_p),
        ('campos', C.c_void_p), ('bg', C.c_void_p),
    ]


_LAYOUT_FIELDS = ['geom_bytes', 'binning_bytes', 'img_bytes', 'header', 'means2D', 'depths', 'cov3D', 'conic_opacity',
                  'rgbd', 'cull', 'clamped', 'tiles_touched', 'point_offsets', 'scan_state', 'geom_grads',
                  'keys', 'vals', 'tile_pairs', 'tile_grid', 'tile_cursors', 'ranges', 'n_contrib',
                  'final_T', 'tile_order', 'work_counters']


class RasterLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in _LAYOUT_FIELDS]


class RasterHeader(C.Structure):
    """struct skgs_raster_header (lives at geom + layout.header)."""
    _fields_ = [('num_rendered', C.c_uint32), ('num_visible', C.c_uint32), ('scan_ticket', C.c_uint32),
                ('overflow', C.c_uint32), ('reserved', C.c_uint32 * 28)]


class Skeleton(C.Structure):
    _fields_ = [
        ('M', C.c_int32), ('L', C.c_int32), ('root', C.c_int32), ('K', C.c_int32), ('mode', C.c_int32),
        ('temperature', C.c_float), ('joints', C.c_void_p), ('sk_r', C.c_void_p), ('sk_r_delta', C.c_void_p),
        ('sk_r_delta_dim', C.c_int32), ('sk_d_rot', C.c_void_p), ('sk_d_scale', C.c_void_p), ('g_tr', C.c_void_p),
        ('parents', C.c_void_p), ('sp_W', C.c_void_p), ('sp_radius', C.c_void_p), ('sp_weight', C.c_void_p),
    ]


class Superpoints(C.Structure):
    """struct skgs_superpoints (sp-stage LBS)."""
    _fields_ = [
        ('M', C.c_int32), ('K', C.c_int32), ('mode', C.c_int32), ('method', C.c_int32), ('temperature', C.c_float),
        ('sp_points', C.c_void_p), ('sp_t', C.c_void_p), ('sp_r', C.c_void_p), ('sp_rot', C.c_void_p),
        ('sp_scale', C.c_void_p), ('sp_W', C.c_void_p), ('sp_radius', C.c_void_p), ('sp_weight', C.c_void_p),
    ]


class DensifyConfig(C.Structure):
    _fields_ = [('do_densify', C.c_int32), ('do_prune', C.c_int32), ('grad_threshold', C.c_float),
                ('densify_extent', C.c_float), ('min_opacity', C.c_float), ('max_screen_size', C.c_float),
                ('prune_extent', C.c_float)]


class DensifyTensor(C.Structure):
    _fields_ = [('in_', C.c_void_p), ('out', C.c_void_p), ('m_in', C.c_void_p), ('m_out', C.c_void_p),
                ('v_in', C.c_void_p), ('v_out', C.c_void_p), ('width', C.c_int32), ('role', C.c_int32)]


class AdamTensor(C.Structure):
    _fields_ = [
        ('param', C.c_void_p), ('grad', C.c_void_p), ('exp_avg', C.c_void_p), ('exp_avg_sq', C.c_void_p),
        ('numel', C.c_int64), ('lr', C.c_double), ('lr2', C.c_double), ('period', C.c_int32), ('split', C.c_int32),
        ('cols', C.c_int32), ('K', C.c_int32), ('knn_indices', C.c_void_p),
    ]


class JointMlp(C.Structure):
    _fields_ = [
        ('M', C.c_int32), ('degree_p', C.c_int32), ('degree_t', C.c_int32), ('width', C.c_int32), ('depth', C.c_int32),
        ('skip_mask', C.c_int32), ('n_out', C.c_int32), ('rotation_head', C.c_int32), ('theta', C.c_void_p),
    ]


ABI_VERSION = 4  # must equal skgs_abi_version() of the loaded library
ADAM_MAX_TENSORS = 16
LBS_MODES = {'W': 0, 'kernel': 1, 'weighted_kernel': 2, 'dist': 3}
WARP_METHODS = {'LBS': 0, 'LBS_c': 1, 'largest': 2}

# every symbol include/skgs_b200.h declares: (restype, argtypes)
_vp, _i32, _i64 = C.c_void_p, C.c_int32, C.c_int64
_SIGNATURES = {
    'skgs_last_error': (C.c_char_p, []),
    'skgs_abi_version': (C.c_int, []),
    'skgs_built_for_sm': (C.c_int, []),
    'skgs_launch_count': (C.c_uint64, []),
    'skgs_profile_enable': (None, [C.c_int]),
    'skgs_profile_collect': (C.c_int, [C.c_char_p, C.c_size_t]),
    'skgs_raster_layout_query': (C.c_int, [_i32, _i32, _i32, _i64, C.POINTER(RasterLayout)]),
    'skgs_raster_forward': (C.c_int, [C.POINTER(RasterSettings), _i32, _i32] + [_vp] * 8 + [_vp, _i64, _vp] +
                            [_vp] * 6),
    'skgs_raster_forward_geometry': (C.c_int, [C.POINTER(RasterSettings), _i32, _i32] + [_vp] * 7 +
                                     [_vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    'skgs_raster_forward_render': (C.c_int, [C.POINTER(RasterSettings), _i32, _vp, _vp, _i64, _i64, _vp, _vp, _i32] +
                                   [_vp] * 5),
    'skgs_deform_forward_geometry': (C.c_int, [C.POINTER(Skeleton), C.POINTER(RasterSettings), _i32, _i32] + [_vp] * 5 +
                                     [_vp] * 7 + [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    'skgs_raster_backward': (C.c_int, [C.POINTER(RasterSettings), _i32, _i32] + [_vp] * 7 + [_vp, _vp, _i64, _vp] +
                             [_vp] * 12),
    'skgs_raster_assemble_backward': (C.c_int, [C.POINTER(RasterSettings), _i32, _i32] + [_vp] * 5 +
                                      [_vp, _vp, _i64, _vp] + [_vp] * 3 + [_vp] * 4 + [_vp] * 9 + [_vp]),
    'skgs_fk_lbs_forward': (C.c_int, [C.POINTER(Skeleton), _i32] + [_vp] * 9),
    'skgs_fk_lbs_workspace_bytes': (C.c_size_t, [_i32]),
    'skgs_fk_lbs_backward': (C.c_int, [C.POINTER(Skeleton), _i32] + [_vp] * 20),
    'skgs_sp_lbs_workspace_bytes': (C.c_size_t, [_i32]),
    'skgs_sp_lbs_forward': (C.c_int, [C.POINTER(Superpoints), _i32] + [_vp] * 9),
    'skgs_sp_lbs_backward': (C.c_int, [C.POINTER(Superpoints), _i32] + [_vp] * 20),
    'skgs_assemble_forward': (C.c_int, [_i32] + [_vp] * 12),
    'skgs_assemble_backward': (C.c_int, [_i32] + [_vp] * 16),
    'skgs_image_loss_workspace_bytes': (C.c_size_t, [_i32, _i32]),
    'skgs_image_loss': (C.c_int, [_i32, _i32, _vp, _vp, _i32, _i32, C.c_float, C.c_float, C.c_float, _vp, _vp, _vp,
                                  _vp]),
    'skgs_adam_step': (C.c_int, [C.POINTER(AdamTensor), _i32, _i32, C.c_double, C.c_double, C.c_double, C.c_float,
                                 _vp, _vp, _vp]),
    'skgs_densify_stats': (C.c_int, [_i32, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    'skgs_densify_workspace_bytes': (C.c_size_t, [_i32]),
    'skgs_densify_plan': (C.c_int, [C.POINTER(DensifyConfig), _i32] + [_vp] * 11),
    'skgs_densify_apply': (C.c_int, [C.POINTER(DensifyTensor), _i32, _i32] + [_vp] * 7),
    'skgs_opacity_reset': (C.c_int, [_i32, _vp, _vp, _vp, C.c_float, _vp]),
    'skgs_joint_mlp_layout': (C.c_int, [C.POINTER(JointMlp), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                        C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    'skgs_joint_mlp_workspace_bytes': (C.c_size_t, [C.POINTER(JointMlp)]),
    'skgs_joint_mlp_forward': (C.c_int, [C.POINTER(JointMlp)] + [_vp] * 7),
    'skgs_joint_mlp_backward': (C.c_int, [C.POINTER(JointMlp)] + [_vp] * 7),
    'skgs_multimem_allreduce': (C.c_int, [_vp, _i64, _i32, _i32, _vp]),
    'skgs_accumulate_f32': (C.c_int, [_vp, _vp, _i64, _vp]),
    'skgs_max_i32': (C.c_int, [_vp, _vp, _i64, _vp]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f'{LIB_PATH} is missing: run `python __graft_entry__.py` (nvcc, sm_100a) first. '
                               f'sk_gs_b200 has no CPU / PyTorch fallback.')
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the export is missing
            fn.restype = res
            fn.argtypes = args
        if L.skgs_abi_version() != ABI_VERSION:
            raise RuntimeError(f'{LIB_PATH} has ABI version {L.skgs_abi_version()}, this package needs {ABI_VERSION}: '
                               f'rebuild it (python __graft_entry__.py)')
        _lib = L
    return _lib


def check_exports():
    """Load the library and verify every declared symbol is exported (no compute call; works without a GPU)."""
    L = lib()
    assert L.skgs_abi_version() == ABI_VERSION
    assert L.skgs_built_for_sm() == 100
    return sorted(_SIGNATURES)


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f'{what} failed ({rc}): {lib().skgs_last_error().decode()}')


def launch_count() -> int:
    return int(lib().skgs_launch_count())


def profile_enable(on: bool):
    lib().skgs_profile_enable(int(bool(on)))


def profile_collect():
    """{kernel name: (launches, total microseconds)} recorded since profile_enable(True)."""
    buf = C.create_string_buffer(1 << 16)
    lib().skgs_profile_collect(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, n, us = line.split()
        out[name] = (int(n), float(us))
    return out


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
