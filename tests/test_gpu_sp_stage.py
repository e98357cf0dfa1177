"""GPU parity of the sp-stage LBS (`calc_LBS_weight` + `warp`, networks/sk_gs.py:751-828) against
(a) the golden vectors the REFERENCE's own functions produced (tests/golden/sp_stage.npz, fp64) and (b) the torch oracle
at the reference's default size (P = 100 000 Gaussians, M = 512 superpoints, K = 5; exps/default.yaml:25).

Tolerances (fp32): forward 2e-6 abs on O(1) quantities (5e-6 at full size), gradients 1e-4 relative to each tensor's
largest magnitude.  KNN indices must agree exactly."""
import os

import numpy as np
import pytest
import torch

from oracle import fk_lbs as OF
from sk_gs_b200.sp_lbs import sp_warp, sp_warp_backward_raw, sp_warp_forward_raw
from skgs_test_util import rel_err
from test_oracle_golden import _sp_case

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
GRAD_RTOL = 1e-4


def test_golden_cases_of_the_reference_functions():
    """All 24 (mode x method x sep_rot) cases: outputs and gradients of the CUDA path vs the reference's fp64 results."""
    d = np.load(os.path.join(G, 'sp_stage.npz'))
    dev = torch.device('cuda:0')
    for i in range(int(d['n'])):
        points, leaves, sp_r, sp_rot, kw = _sp_case(d, i, dtype=torch.float32, device=dev)
        outs = sp_warp(points, leaves['sp_points'], leaves['sp_t'], sp_r, sp_rot, leaves['sp_scale'], **kw)
        d_points, d_rotation, d_scales, spT, w, idx = outs
        assert np.array_equal(idx.cpu().numpy(), d[f'c{i}_idx']), i
        for name, got in (('w', w), ('d_points', d_points), ('d_rotation', d_rotation), ('d_scales', d_scales),
                          ('spT', spT)):
            assert np.abs(got.detach().cpu().double().numpy() - d[f'c{i}_{name}']).max() <= 2e-6, (i, name)
        loss = sum((o * torch.from_numpy(d[f'c{i}_cot{j}']).float().to(dev)).sum() for j, o in
                   enumerate((d_points, d_rotation, d_scales, spT)))
        names = [k[len(f'c{i}_grad_'):] for k in d.files if k.startswith(f'c{i}_grad_')]
        grads = torch.autograd.grad(loss, [leaves[n] for n in names], allow_unused=True)
        for n, g in zip(names, grads):
            ref = d[f'c{i}_grad_{n}']
            got = np.zeros_like(ref) if g is None else g.cpu().numpy()
            assert rel_err(got, ref) <= GRAD_RTOL or np.abs(ref).max() == 0 and np.abs(got).max() == 0, \
                (i, str(d[f'c{i}_cfg']), n, rel_err(got, ref))


def _scene(P, M, seed, dev, dtype):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).to(device=dev, dtype=dtype)  # noqa: E731
    points, sp_points = r(P, 3, scale=0.5), r(M, 3, scale=0.5)
    leaves = dict(sp_points=sp_points, sp_t=r(M, 3, scale=0.05), raw_r=r(M, 4, scale=0.2), raw_g=r(M, 4, scale=0.2),
                  sp_scale=r(M, 3, scale=0.01), sp_W=r(P, M), sp_radius=r(M, scale=0.2) - 1.5, sp_weight=r(M))
    return points, leaves


@pytest.mark.parametrize('mode,method,sep_rot', [('W', 'LBS', True), ('W', 'largest', False),
                                                 ('weighted_kernel', 'LBS_c', True), ('dist', 'LBS', False),
                                                 ('kernel', 'LBS_c', False)])
def test_default_size_against_oracle(mode, method, sep_rot):
    """P = 100K, M = 512, K = 5 (the sp-stage of exps/default.yaml / d_nerf_sc_gs.yaml / d_nerf_sp_gs.yaml): the fp64
    oracle on the GPU, the CUDA path in fp32."""
    dev = torch.device('cuda:0')
    P, M, K = 100_000, 512, 5
    points, base = _scene(P, M, 77, dev, torch.float64)
    bias64 = torch.tensor([0, 0, 0, 1.0], dtype=torch.float64, device=dev)

    def run(dtype, fn):
        lv = {k: v.to(dtype).clone().requires_grad_() for k, v in base.items()}
        bias = bias64.to(dtype)
        sp_r = torch.nn.functional.normalize(lv['raw_r'] + bias, dim=-1)
        sp_rot = torch.nn.functional.normalize(lv['raw_g'] + bias, dim=-1) if sep_rot else None
        kw = dict(K=K, mode=mode, method=method, sp_W=lv['sp_W'] if mode == 'W' else None,
                  sp_radius=lv['sp_radius'] if 'kernel' in mode else None,
                  sp_weight=lv['sp_weight'] if mode == 'weighted_kernel' else None, temperature=0.05)
        return lv, fn(points.to(dtype), lv['sp_points'], lv['sp_t'], sp_r, sp_rot, lv['sp_scale'], **kw)

    lo, o = run(torch.float64, OF.sp_stage)
    lg, g = run(torch.float32, sp_warp)
    assert torch.equal(g[5], o[5])
    for j in (0, 1, 2, 3, 4):
        assert (g[j].double() - o[j]).abs().max().item() <= 5e-6, j
    gen = torch.Generator().manual_seed(3)
    cot = [torch.randn(t.shape, generator=gen, dtype=torch.float64).to(dev) for t in o[:5]]
    names = ['sp_points', 'sp_t', 'raw_r', 'sp_scale'] + (['raw_g'] if sep_rot else []) + \
        (['sp_W'] if mode == 'W' else []) + (['sp_radius'] if 'kernel' in mode else []) + \
        (['sp_weight'] if mode == 'weighted_kernel' else [])
    go = torch.autograd.grad(sum((t * c).sum() for t, c in zip(o[:5], cot)), [lo[n] for n in names], allow_unused=True)
    gg = torch.autograd.grad(sum((t * c.float()).sum() for t, c in zip(g[:5], cot)), [lg[n] for n in names],
                             allow_unused=True)
    for n, a, b in zip(names, gg, go):
        if b is None or float(b.abs().max()) == 0:
            assert a is None or float(a.abs().max()) == 0, n
            continue
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) <= GRAD_RTOL, (n, rel_err(a.cpu().numpy(), b.cpu().numpy()))


def test_compact_sp_W_gradient_and_small_M_kernel():
    """M <= 256 takes the joint-major backward kernel; the compact [P, K] sp_W gradient is the dense one gathered."""
    dev = torch.device('cuda:0')
    P, M, K = 20_000, 64, 5
    points, lv = _scene(P, M, 78, dev, torch.float32)
    sp_r = torch.nn.functional.normalize(lv['raw_r'] + torch.tensor([0, 0, 0, 1.0], device=dev), dim=-1)
    gen = torch.Generator().manual_seed(4)
    cots = [torch.randn(P, k, generator=gen).to(dev) for k in (3, 4, 3)]
    res = {}
    for compact in (False, True):
        out, ctx = sp_warp_forward_raw(points, lv['sp_points'], lv['sp_t'], sp_r, None, lv['sp_scale'], K=K, mode='W',
                                       sp_W=lv['sp_W'], method='LBS')
        res[compact] = (out, sp_warp_backward_raw(ctx, *cots, compact_sp_W=compact))
    (out, dense), (_, comp) = res[False], res[True]
    idx = out[5]
    assert torch.equal(torch.gather(dense[5], 1, idx), comp[5])
    assert float(dense[5].abs().sum()) == pytest.approx(float(comp[5].abs().sum()), rel=1e-6)
    for a, b in zip(dense[:5], comp[:5]):
        # same kernels twice: only the order of the floating-point atomics differs
        assert (a is None and b is None) or float((a - b).abs().max()) <= 1e-5 * float(a.abs().max())
    # against the oracle (fp64)
    l64 = {k: v.double().clone().requires_grad_() for k, v in lv.items()}
    sp_r64 = torch.nn.functional.normalize(l64['raw_r'] + torch.tensor([0, 0, 0, 1.0], dtype=torch.float64, device=dev),
                                           dim=-1)
    sp_r64_leaf = sp_r64.detach().clone().requires_grad_()
    o = OF.sp_stage(points.double(), l64['sp_points'], l64['sp_t'], sp_r64_leaf, None, l64['sp_scale'], K=K, mode='W',
                    sp_W=l64['sp_W'], method='LBS')
    go = torch.autograd.grad(sum((t * c.double()).sum() for t, c in zip(o[:3], cots)),
                             [l64['sp_t'], sp_r64_leaf, l64['sp_scale'], l64['sp_W']])
    assert rel_err(dense[1].cpu().numpy(), go[0].cpu().numpy()) <= GRAD_RTOL
    assert rel_err(dense[2].cpu().numpy(), go[1].cpu().numpy()) <= GRAD_RTOL   # tangent + blend gradient of sp_r
    assert rel_err(dense[4].cpu().numpy(), go[2].cpu().numpy()) <= GRAD_RTOL
    assert rel_err(dense[5].cpu().numpy(), go[3].cpu().numpy()) <= GRAD_RTOL


def test_argument_errors_are_loud():
    dev = torch.device('cuda:0')
    points, lv = _scene(100, 8, 79, dev, torch.float32)
    with pytest.raises(RuntimeError):
        sp_warp(points, lv['sp_points'], lv['sp_t'], lv['raw_r'], K=5, mode='W', sp_W=None)
    with pytest.raises(RuntimeError):
        sp_warp(points, lv['sp_points'], lv['sp_t'], lv['raw_r'], K=9, mode='dist')
    with pytest.raises(RuntimeError):
        sp_warp(points.cpu(), lv['sp_points'].cpu(), lv['sp_t'].cpu(), lv['raw_r'].cpu(), K=2, mode='dist')


def test_edge_cases_single_superpoint_and_empty_point_set():
    """M = K = 1 (every Gaussian follows the one transform with weight 1) and P = 0."""
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(5)
    pts = torch.randn(50, 3, generator=g).to(dev)
    c, t = torch.randn(1, 3, generator=g).to(dev), torch.randn(1, 3, generator=g).to(dev)
    q = torch.nn.functional.normalize(torch.randn(1, 4, generator=g), dim=-1).to(dev)
    for method in ('LBS', 'LBS_c', 'largest'):
        d_points, d_rot, d_scales, spT, w, idx = sp_warp(pts, c, t, q, None, None, K=1, mode='dist', method=method)
        ref = OF.sp_stage(pts.double(), c.double(), t.double(), q.double(), None, None, K=1, mode='dist', method=method)
        assert d_scales is None and torch.all(idx == 0) and torch.all(w == 1)
        assert float((d_points.double() - ref[0]).abs().max()) <= 2e-6
        assert float((d_rot - q.expand(50, 4)).abs().max()) <= 1e-7
    out = sp_warp(pts[:0], c, t, q, None, None, K=1, mode='dist')
    assert out[0].shape == (0, 3) and out[4].shape == (0, 1) and out[3].shape == (1, 7)
