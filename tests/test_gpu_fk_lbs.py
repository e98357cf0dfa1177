"""GPU parity of the fused FK + LBS kernels (and the assembly kernels) against the torch oracle.

Tolerances (fp32): forward 2e-6 abs on O(1) quantities; gradients 1e-4 relative to each tensor's largest magnitude.
KNN indices must agree exactly (synthetic joints are generic: no exact distance ties)."""
import pytest
import torch

from oracle import fk_lbs as OF
from sk_gs_b200 import scene as S
from sk_gs_b200.fk_lbs import assemble, fk_lbs
from skgs_test_util import rel_err

pytestmark = pytest.mark.gpu
FWD_ATOL = 2e-6
GRAD_RTOL = 1e-4


def _leaf(t, dev=None, dtype=torch.float32):
    t = t.to(dtype).clone()
    if dev is not None:
        t = t.to(dev)
    return t.requires_grad_(True)


def _run_both(sc, mode, delta=None, P=None):
    dev = torch.device('cuda:0')
    names = ['joints', 'sk_r', 'sk_d_rot', 'sk_d_scale', 'g_tr', 'sp_W', 'sp_radius', 'sp_weight']
    cpu = {n: _leaf(getattr(sc, n), dtype=torch.float64) for n in names}
    gpu = {n: _leaf(getattr(sc, n), dev) for n in names}
    kw = dict(K=sc.K, mode=mode, temperature=0.7)
    o = OF.sk_stage(sc.xyz.double(), cpu['joints'], cpu['sk_r'], cpu['sk_d_rot'], cpu['sk_d_scale'], cpu['g_tr'],
                    sc.parents.long(), sc.root, sp_W=cpu['sp_W'], sp_radius=cpu['sp_radius'],
                    sp_weight=cpu['sp_weight'], sk_r_delta=None if delta is None else delta.double(), **kw)
    g = fk_lbs(sc.xyz.to(dev), gpu['joints'], gpu['sk_r'], gpu['sk_d_rot'], gpu['sk_d_scale'], gpu['g_tr'],
               sc.parents.to(dev), sc.root, sp_W=gpu['sp_W'] if mode == 'W' else None,
               sp_radius=gpu['sp_radius'] if mode in ('kernel', 'weighted_kernel') else None,
               sp_weight=gpu['sp_weight'] if mode == 'weighted_kernel' else None,
               sk_r_delta=None if delta is None else delta.to(dev), **kw)
    return cpu, gpu, o, g


@pytest.mark.parametrize('mode', ['W', 'kernel', 'weighted_kernel', 'dist'])
@pytest.mark.parametrize('name,P', [('c1', 5000), ('c4', 3000)])
def test_forward_and_backward(mode, name, P):
    sc = S.make_scene(name, P=P, seed=31)
    cpu, gpu, o, g = _run_both(sc, mode)
    d_xyz, d_rot, d_scale, sk_T, _, _, _, w, idx = g
    assert torch.equal(idx.cpu(), o[8])
    for got, ref in [(d_xyz, o[0]), (d_rot, o[1]), (d_scale, o[2]), (sk_T, o[3]), (w, o[7])]:
        assert (got.cpu().double() - ref).abs().max().item() <= FWD_ATOL
    # backward with random cotangents on every differentiable output
    gen = torch.Generator().manual_seed(5)
    cot = [torch.randn(t.shape, generator=gen, dtype=torch.float64) for t in (o[0], o[1], o[2], o[3], o[7])]
    loss_c = sum((t * c).sum() for t, c in zip((o[0], o[1], o[2], o[3], o[7]), cot))
    loss_c.backward()
    dev = d_xyz.device
    loss_g = sum((t * c.float().to(dev)).sum() for t, c in zip((d_xyz, d_rot, d_scale, sk_T, w), cot))
    loss_g.backward()
    names = ['joints', 'sk_r', 'sk_d_rot', 'sk_d_scale', 'g_tr']
    if mode == 'W':
        names.append('sp_W')
    if mode in ('kernel', 'weighted_kernel'):
        names.append('sp_radius')
    if mode == 'weighted_kernel':
        names.append('sp_weight')
    for n in names:
        ref, got = cpu[n].grad, gpu[n].grad
        assert got is not None and ref is not None, n
        assert rel_err(got.cpu().numpy(), ref.numpy()) <= GRAD_RTOL, (n, rel_err(got.cpu().numpy(), ref.numpy()))


def test_repose_delta_forward():
    """gui.py reposing path: sk_r <- Exp(delta) * sk_r (networks/sk_gs.py:1087-1088), axis-angle and quaternion forms."""
    sc = S.make_scene('c1', P=2000, seed=32)
    gen = torch.Generator().manual_seed(6)
    for delta in (0.3 * torch.randn(sc.cfg.M, 3, generator=gen),
                  torch.nn.functional.normalize(torch.randn(sc.cfg.M, 4, generator=gen), dim=-1)):
        cpu, gpu, o, g = _run_both(sc, 'W', delta=delta)
        assert (g[3].cpu().double() - o[3]).abs().max().item() <= FWD_ATOL
        assert (g[0].cpu().double() - o[0]).abs().max().item() <= 5e-6


def test_chain_and_star_trees():
    """Deep chain (L = 5 for 33 joints) and radius-1 star (the reference's unsupported L == 0 case)."""
    for parent in ([-1] + list(range(32)), [-1] + [0] * 7):
        sc = S.make_scene('c1', P=1000, seed=33)
        M = len(parent)
        g = torch.Generator().manual_seed(9)
        sc.joints = torch.randn(M, 3, generator=g) * 0.5
        sc.parents, sc.joint_depth, sc.root = S.find_root_table(parent)
        sc.sk_r = torch.nn.functional.normalize(torch.randn(M, 4, generator=g), dim=-1)
        sc.sk_d_rot, sc.sk_d_scale = 0.01 * torch.randn(M, 4, generator=g), 0.001 * torch.randn(M, 3, generator=g)
        sc.sp_W = torch.randn(1000, M, generator=g)
        sc.sp_radius, sc.sp_weight = torch.zeros(M), torch.zeros(M)
        cpu, gpu, o, gg = _run_both(sc, 'W')
        assert (gg[3].cpu().double() - o[3]).abs().max().item() <= 5e-6
        # the serial recurrence is the same function (oracle cross-check)
        ser = OF.skeleton_warp_serial(OF.local_transforms(sc.joints.double(), sc.sk_r.double()), sc.g_tr.double(),
                                      sc.parents[:, 0].long(), sc.root)
        assert (ser - o[3]).abs().max().item() <= 1e-12


def test_assemble_forward_backward():
    sc = S.make_scene('c1', P=4000, seed=34)
    dev = torch.device('cuda:0')
    gen = torch.Generator().manual_seed(2)
    d = [0.1 * torch.randn(4000, k, generator=gen) for k in (3, 4, 3)]
    cpu = [_leaf(t, dtype=torch.float64) for t in (sc.xyz, sc.scaling, sc.rotation, sc.opacity, *d)]
    gpu = [_leaf(t, dev) for t in (sc.xyz, sc.scaling, sc.rotation, sc.opacity, *d)]
    ref = OF.assemble(cpu[0], cpu[1], cpu[2], cpu[3], sc.f_dc.double(), sc.f_rest.double(), cpu[4], cpu[5], cpu[6])[:4]
    got = assemble(*gpu)
    cot = [torch.randn(t.shape, generator=gen, dtype=torch.float64) for t in ref]
    sum((t * c).sum() for t, c in zip(ref, cot)).backward()
    sum((t * c.float().to(dev)).sum() for t, c in zip(got, cot)).backward()
    for a, b in zip(got, ref):
        assert (a.cpu().double() - b).abs().max().item() <= 2e-6
    for a, b in zip(gpu, cpu):
        assert rel_err(a.grad.cpu().numpy(), b.grad.numpy()) <= 1e-5


def test_many_joints_uses_atomic_fallback():
    """M > 256 takes the shared-memory-atomics LBS backward instead of the joint-major kernel; same results."""
    M, P = 300, 3000
    g = torch.Generator().manual_seed(41)
    sc = S.make_scene('c1', P=P, seed=41)
    parent = [-1] + [int(torch.randint(0, j, (1,), generator=g)) for j in range(1, M)]
    sc.joints = torch.randn(M, 3, generator=g) * 0.6
    sc.parents, sc.joint_depth, sc.root = S.find_root_table(parent)
    sc.sk_r = torch.nn.functional.normalize(torch.randn(M, 4, generator=g), dim=-1)
    sc.sk_d_rot, sc.sk_d_scale = 0.01 * torch.randn(M, 4, generator=g), 0.001 * torch.randn(M, 3, generator=g)
    sc.sp_W = torch.randn(P, M, generator=g)
    sc.sp_radius, sc.sp_weight = torch.full((M,), -1.5), torch.zeros(M)
    for mode in ('W', 'weighted_kernel'):
        cpu, gpu, o, gg = _run_both(sc, mode)
        assert torch.equal(gg[8].cpu(), o[8])
        cot = [torch.randn(t.shape, generator=g, dtype=torch.float64) for t in (o[0], o[1], o[2])]
        sum((t * c).sum() for t, c in zip(o[:3], cot)).backward()
        sum((t * c.float().to(t.device)).sum() for t, c in zip(gg[:3], cot)).backward()
        for n in ['joints', 'sk_r', 'sk_d_rot', 'sk_d_scale', 'g_tr'] + (['sp_W'] if mode == 'W' else ['sp_radius', 'sp_weight']):
            assert rel_err(gpu[n].grad.cpu().numpy(), cpu[n].grad.numpy()) <= GRAD_RTOL, (mode, n)
