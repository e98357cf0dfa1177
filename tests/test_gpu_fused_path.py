"""The fused per-Gaussian forward (fk_table_kernel + deform_preprocess_kernel: skinning + assembly + preprocess + prefix
sum + key emission in one kernel) is an execution plan of the same operators, built from the same device functions with
the same flags: every tensor it produces must equal the operator-by-operator path (skgs_fk_lbs_forward ->
skgs_assemble_forward -> skgs_raster_forward_*) BIT FOR BIT, and the backward contexts it builds must give the same
gradients."""
import numpy as np
import pytest
import torch

from sk_gs_b200 import scene as S
from sk_gs_b200.pipeline import HotPath

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.parametrize('name,P,K', [('c1', 3000, 5), ('c1', None, 3), ('c4', 40000, 8), ('c2', None, 5)])
def test_fused_forward_is_bit_identical_to_the_operator_path(name, P, K):
    sc = S.make_scene(name, P=P, views=1)
    sc.K = K
    hp = HotPath(sc, DEV, merged_sh=True, requires_grad=False)
    out_u, ctx_u = hp.forward_raw(0, fused=False)          # first call of the shape: also primes the capacity estimate
    out_f, ctx_f = hp.forward_raw(0, fused=True)
    torch.cuda.synchronize()
    assert out_f.get('_fused') and not out_u.get('_fused')
    for k in ('images', 'depths', 'alpha', 'radii'):
        assert torch.equal(out_u[k], out_f[k]), k
    su, sf = out_u['_raster_state'], out_f['_raster_state']
    # assembled Gaussians (inputs of the rasterizer): keep = (view, proj, campos, bg, means3D, shs, colors, opacities,
    # scales, rotations, cov3D)
    for idx, what in ((4, 'points'), (7, 'opacity'), (8, 'scales'), (9, 'rotations')):
        assert torch.equal(su.keep[idx], sf.keep[idx]), what
    for a, b, what in zip(out_u['_sk'][3:], out_f['_sk'][3:], ('sk_T', 'sk_d_rot', 'sk_d_scale', 'g_tr', 'weights',
                                                               'indices')):
        assert torch.equal(a, b), what
    assert torch.equal(out_u['_sk'][1], out_f['_sk'][1])   # d_rot
    ku, vu = su.sorted_lists()
    kf, vf = sf.sorted_lists()
    assert torch.equal(ku, kf) and torch.equal(vu, vf)
    H, W = sc.cfg.H, sc.cfg.W
    dL = (torch.randn(3, H, W, generator=torch.Generator().manual_seed(3)) / (3 * H * W)).to(DEV)
    gu, _ = hp.backward_raw(ctx_u, dL)
    gf, _ = hp.backward_raw(ctx_f, dL)
    g3, _ = hp.backward_raw(ctx_u, dL, fused=False)   # assembly backward as a kernel of its own
    for n in ('xyz', 'scaling', 'rotation', 'opacity', 'shs', 'sp_W', 'joints', 'sk_r', 'g_tr'):
        scale = float(g3[n].abs().max())
        assert float((gu[n] - g3[n]).abs().max()) <= 1e-4 * scale, n   # two runs of the atomics + fma contraction
    for n in ('xyz', 'scaling', 'rotation', 'opacity', 'shs', 'sp_W', 'joints', 'sk_r', 'sk_d_rot', 'sk_d_scale',
              'g_tr', 'viewspace_points'):
        a, b = gu[n], gf[n]
        scale = float(a.abs().max())
        assert scale > 0 and float((a - b).abs().max()) <= 1e-4 * scale, n   # same kernels; only the atomics' order differs


def test_fused_forward_in_a_captured_graph_and_train_loop():
    """capture_step uses the fused forward (its capacity is fixed): a replay must reproduce the eager operator path."""
    sc = S.make_scene('c1', views=1)
    hp = HotPath(sc, DEV, merged_sh=True, requires_grad=False)
    H, W = sc.cfg.H, sc.cfg.W
    dL = (torch.randn(3, H, W, generator=torch.Generator().manual_seed(4)) / (3 * H * W)).to(DEV)
    out_u, ctx_u = hp.forward_raw(0, fused=False)
    gu, _ = hp.backward_raw(ctx_u, dL)
    graph, out, grads = hp.capture_step(0, dL, headroom=1.5)
    assert out.get('_fused')
    graph.replay()
    torch.cuda.synchronize()
    assert not hp.overflowed()
    assert torch.equal(out['images'], out_u['images']) and torch.equal(out['radii'], out_u['radii'])
    for n in ('xyz', 'shs', 'sp_W', 'joints', 'sk_r', 'g_tr'):
        scale = float(gu[n].abs().max())
        assert float((gu[n] - grads[n]).abs().max()) <= 2e-5 * scale, n
