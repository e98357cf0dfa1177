"""C ABI of libskgs_b200.so without a GPU: the library loads, exports every symbol include/skgs_b200.h declares, the
host-only entry points work and argument errors are reported the way the header promises."""
import ctypes as C
import os
import re

import pytest

from sk_gs_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    with open(os.path.join(ROOT, 'include', 'skgs_b200.h')) as f:
        src = f.read()
    return sorted(set(re.findall(r'SKGS_API\s+[\w\s\*]+?\b(skgs_\w+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    declared = _declared_symbols()
    assert len(declared) >= 14
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), f'{name} declared in include/skgs_b200.h but not exported'
    assert sorted(_lib.check_exports()) == declared  # the ctypes binding covers exactly the header
    assert L.skgs_abi_version() == _lib.ABI_VERSION and L.skgs_built_for_sm() == 100


def test_built_for_sm100a_only():
    import subprocess
    out = subprocess.run(['cuobjdump', '-lelf', _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r'sm_(\d+a?)', out))
    assert archs == {'100a'}, archs


def test_layout_query_is_host_only_and_consistent():
    L = _lib.lib()
    lay = _lib.RasterLayout()
    assert L.skgs_raster_layout_query(100000, 800, 800, 1000000, C.byref(lay)) == 0
    offs = {n: getattr(lay, n) for n in _lib._LAYOUT_FIELDS}
    assert all(v % 256 == 0 for k, v in offs.items())
    assert offs['header'] == 0 and offs['scan_state'] > 0 and offs['means2D'] > offs['scan_state']
    assert lay.geom_bytes >= 100000 * (8 + 4 + 24 + 16 + 16 + 16 + 1 + 4 + 4 + 48)
    assert lay.binning_bytes >= 1000000 * 20
    assert lay.img_bytes >= 2500 * 8 + 2 * 800 * 800 * 4
    # keys / values, the per-tile segments of the scatter pass, then the tile grid + cursors one memset resets
    assert len({offs['keys'], offs['vals'], offs['tile_pairs']}) == 3
    assert offs['tile_grid'] > offs['tile_pairs'] and offs['tile_cursors'] > offs['tile_grid']
    assert C.sizeof(_lib.RasterHeader) == 128 and _lib.RasterHeader.overflow.offset == 12


def test_errors_are_reported_not_swallowed():
    L = _lib.lib()
    lay = _lib.RasterLayout()
    assert L.skgs_raster_layout_query(-1, 800, 800, 0, C.byref(lay)) == -1
    assert b'bad sizes' in L.skgs_last_error()
    assert L.skgs_raster_layout_query(10, 800, 800, 1 << 31, C.byref(lay)) == -1
    assert b'2^31' in L.skgs_last_error()
    # NULL settings / skeleton are rejected before any CUDA call
    assert L.skgs_raster_forward_geometry(None, 0, 0, *([None] * 10), 0, *([None] * 3)) == -1
    assert b'settings is NULL' in L.skgs_last_error()
    assert L.skgs_fk_lbs_forward(None, 0, *([None] * 9)) == -1
    sk = _lib.Skeleton(M=2000, L=1, root=0, K=5, mode=0, temperature=1.0)
    assert L.skgs_fk_lbs_forward(C.byref(sk), 0, *([None] * 9)) == -1
    assert b'M=2000' in L.skgs_last_error()
    with pytest.raises(RuntimeError):
        _lib.check(-1, 'demo')


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under sk_gs_b200/ may import, link or call it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'sk_gs_b200')):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                with open(os.path.join(dirpath, fn)) as f:
                    txt = f.read()
                if re.search(r'^\s*(from|import)\s+oracle\b', txt, re.M) or 'liboracle' in txt or '#include "../../oracle' in txt:
                    bad.append(fn)
    assert not bad, bad


def test_cpu_tensors_fail_loudly():
    import torch
    from sk_gs_b200 import diff_gaussian_rasterization as DGR
    from sk_gs_b200.fk_lbs import fk_lbs
    from sk_gs_b200.pipeline import raster_settings_for
    from sk_gs_b200 import scene as S
    sc = S.make_scene('c1', P=8)
    rs = raster_settings_for(sc.cameras[0], 'cpu')
    with pytest.raises(RuntimeError, match='no CPU path'):
        DGR.rasterize_forward(rs, sc.xyz, torch.sigmoid(sc.opacity), shs=torch.cat((sc.f_dc, sc.f_rest), 1),
                              scales=sc.scaling.exp(), rotations=sc.rotation)
    with pytest.raises(RuntimeError, match='no CPU path'):
        fk_lbs(sc.xyz, sc.joints, sc.sk_r, sc.sk_d_rot, sc.sk_d_scale, sc.g_tr, sc.parents, sc.root, sp_W=sc.sp_W)


def test_widening_entry_points_validate_arguments_before_any_device_work():
    """skgs_image_loss / skgs_adam_step / skgs_joint_mlp_*: bad arguments return SKGS_ERR_INVALID_ARG with a message."""
    L = _lib.lib()
    assert L.skgs_image_loss_workspace_bytes(800, 800) >= 9 * 800 * 800 * 4
    assert L.skgs_image_loss(0, 10, *([None] * 2), 0, 0, 1.0, 1.0, 1.0, *([None] * 4)) == -1
    assert b'empty image' in L.skgs_last_error()
    assert L.skgs_image_loss(8, 8, *([None] * 2), 0, 0, 1.0, 1.0, 1.0, *([None] * 4)) == -1
    assert b'null pointer' in L.skgs_last_error()
    assert L.skgs_image_loss(8, 8, 1, 1, 5, 0, 1.0, 1.0, 1.0, 1, 1, None, None) == -1
    assert b'target_pixel_stride' in L.skgs_last_error()
    assert L.skgs_image_loss(8, 8, 1, 1, 3, 7, 1.0, 1.0, 1.0, 1, 1, None, None) == -1
    assert b'method' in L.skgs_last_error()

    table = (_lib.AdamTensor * 1)()
    assert L.skgs_adam_step(table, 17, 1, 0.9, 0.999, 1e-15, 1.0, None, None, None) == -1
    assert b'count 17' in L.skgs_last_error()
    assert L.skgs_adam_step(table, 1, 0, 0.9, 0.999, 1e-15, 1.0, None, None, None) == -1
    assert b'step must be >= 1' in L.skgs_last_error()
    assert L.skgs_adam_step(table, 1, 1, 1.5, 0.999, 1e-15, 1.0, None, None, None) == -1
    assert b'hyper-parameters' in L.skgs_last_error()
    assert L.skgs_adam_step(table, 1, 1, 0.9, 0.999, 1e-15, 1.0, None, None, None) == 0  # numel 0: nothing to do
    table[0] = _lib.AdamTensor(None, None, None, None, 8, 1e-3, 1e-3, 0, 0, 0, 0, None)
    assert L.skgs_adam_step(table, 1, 1, 0.9, 0.999, 1e-15, 1.0, None, None, None) == -1
    assert b'null pointer' in L.skgs_last_error()
    table[0] = _lib.AdamTensor(1, 1, 1, 1, 8, 1e-3, 1e-3, 4, 4, 0, 0, None)
    assert L.skgs_adam_step(table, 1, 1, 0.9, 0.999, 1e-15, 1.0, None, None, None) == -1
    assert b'split < period' in L.skgs_last_error()
    table[0] = _lib.AdamTensor(1, 1, 1, 1, 8, 1e-3, 1e-3, 0, 0, 4, 5, 1)
    assert L.skgs_adam_step(table, 1, 1, 0.9, 0.999, 1e-15, 1.0, None, None, None) == -1
    assert b'compact gradient' in L.skgs_last_error()

    net = _lib.JointMlp(32, 10, 6, 256, 8, 1 << 4, 11, 1, None)
    tot = C.c_int64()
    assert L.skgs_joint_mlp_layout(C.byref(net), None, None, None, C.byref(tot)) == 0 and tot.value == 502539
    assert L.skgs_joint_mlp_workspace_bytes(C.byref(net)) > 8 * 32 * 256 * 4
    assert L.skgs_joint_mlp_forward(C.byref(net), *([None] * 7)) == -1  # theta == NULL
    assert b'null pointer' in L.skgs_last_error()
    bad = _lib.JointMlp(32, 10, 6, 256, 8, 1 << 4, 12, 1, None)
    assert L.skgs_joint_mlp_layout(C.byref(bad), None, None, None, C.byref(tot)) == -1
    assert b'11 outputs' in L.skgs_last_error()
    assert L.skgs_joint_mlp_workspace_bytes(C.byref(bad)) == 0
    assert L.skgs_joint_mlp_layout(None, None, None, None, None) == -1


def test_sass_carries_the_instructions_the_design_relies_on():
    """Evidence that the shipped binary is the design DESIGN.md describes (no GPU needed, cuobjdump on the in-tree .so):
    NVLS in-switch reduction (multimem.ld_reduce -> LDGMC...ADD.F32x4, multimem.st -> STG.E.128.STRONG.SYS), one 16-byte
    vector RED per gradient group in the compositing backward, match.any ranking in the per-tile sort, and nothing
    compiled for an architecture other than sm_100a."""
    import subprocess
    sass = subprocess.run(['cuobjdump', '-sass', _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert len(re.findall(r'LDGMC\.E\.ADD\.F32x4', sass)) >= 4            # allreduce_mm.cu
    assert 'STG.E.128.STRONG.SYS' in sass                                 # multimem.st of the reduced slice
    assert len(re.findall(r'REDG\.E\.ADD\.F32x4', sass)) >= 3             # composite.cu flush_slots: 3 x red.v4.f32
    assert 'MATCH.ANY' in sass                                            # tile_sort.cu radix ranking
    assert 'REDG.E.ADD.F64' in sass                                       # image_loss.cu loss sums
    funcs = set(re.findall(r'Function : (\S+)', sass))
    for name in ('composite_fwd_kernel', 'composite_bwd_kernel', 'tile_plan_kernel', 'tile_scatter_kernel', 'tile_sort_kernel',
                 'deform_preprocess_kernel', 'preprocess_scan_kernel', 'densify_apply_kernel', 'sp_table_kernel',
                 'lbs_fwd_kernel', 'fk_table_kernel', 'lbs_bwd_jm_kernel', 'multimem_allreduce_kernel', 'ssim_stats_kernel',
                 'ssim_grad_kernel', 'adam_kernel', 'small_gemm_kernel'):
        assert any(name in f for f in funcs), name
