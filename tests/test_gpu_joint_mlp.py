"""GPU parity of the joint-rotation network through the C ABI: skgs_joint_mlp_forward / _backward vs
oracle/deform_net.py (float64 autograd; its structure is pinned on the reference's own SimpleDeformationNetwork by
tests/test_oracle_deform_net.py), over several depths / widths / skip patterns / joint counts, then the module mirror
and the HotPath / TrainLoop plumbing.  Tolerances: outputs 1e-5 of the tensor's max, gradients 1e-4 (north_star)."""
import os

import numpy as np
import pytest
import torch

from oracle import deform_net as OD
from sk_gs_b200 import scene as S
from sk_gs_b200.deform_net import (NetConfig, SimpleDeformationNetwork, joint_mlp_backward_raw,
                                   joint_mlp_forward_raw)
from sk_gs_b200.pipeline import HotPath
from sk_gs_b200.train import TrainLoop

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _flatten(cfg: NetConfig, ws, bs):
    theta = torch.zeros(cfg.param_count, dtype=ws[0].dtype)
    views = cfg.views(theta)
    with torch.no_grad():
        for i in range(cfg.depth):
            views[i][0].copy_(ws[i].detach())
            views[i][1].copy_(bs[i].detach())
        views[-1][0].copy_(torch.cat([w.detach() for w in ws[cfg.depth:]], 0))
        views[-1][1].copy_(torch.cat([b.detach() for b in bs[cfg.depth:]], 0))
    return theta


def _check(cfg: NetConfig, ws, bs, joints, t, tag):
    """forward + backward of the library against float64 autograd of the oracle, layer by layer."""
    ws64 = [w.detach().clone().double().requires_grad_(True) for w in ws]
    bs64 = [b.detach().clone().double().requires_grad_(True) for b in bs]
    j64 = joints.detach().clone().double().requires_grad_(True)
    outs = OD.forward(ws64, bs64, j64, t.double(), cfg.degree_p, cfg.degree_t, cfg.skips, cuda_formula=True,
                      rotation_head=cfg.rotation_head)
    gen = torch.Generator().manual_seed(5)
    gs = [torch.randn(o.shape, generator=gen, dtype=torch.float64) for o in outs]
    ref = torch.autograd.grad(outs, [j64] + ws64 + bs64, gs)
    theta = _flatten(cfg, ws, bs).float().to(DEV)
    t_dev = t.float().reshape(1).to(DEV)
    got, ctx = joint_mlp_forward_raw(cfg, theta, joints.float().to(DEV), t_dev)
    for o, r, name in zip(got, outs, ('sk_r', 'd_rot', 'd_scale')):
        assert _rel(o, r) <= 1e-5, (tag, 'forward', name, _rel(o, r))
    d_theta, d_joints = joint_mlp_backward_raw(ctx, *[g.float().to(DEV) for g in gs])
    torch.cuda.synchronize()
    n = len(ws)
    ref_w, ref_b = list(ref[1:1 + n]), list(ref[1 + n:])
    ref_w = ref_w[:cfg.depth] + [torch.cat(ref_w[cfg.depth:], 0)]
    ref_b = ref_b[:cfg.depth] + [torch.cat(ref_b[cfg.depth:], 0)]
    views = cfg.views(d_theta)
    scale = max(float(r.abs().max()) for r in ref_w)
    for i in range(cfg.depth, -1, -1):  # in the order the backward computes them
        ew = float((views[i][0].cpu().double() - ref_w[i]).abs().max()) / scale
        eb = float((views[i][1].cpu().double() - ref_b[i]).abs().max()) / scale
        assert ew <= 1e-4 and eb <= 1e-4, (tag, 'backward layer', i, 'dW', ew, 'db', eb)
    assert _rel(d_joints, ref[0]) <= 1e-4, (tag, 'd_joints', _rel(d_joints, ref[0]))
    # joints gradient may be skipped
    d_theta2, none = joint_mlp_backward_raw(ctx, *[g.float().to(DEV) for g in gs], need_joints=False)
    assert none is None and torch.equal(d_theta2, d_theta)


@pytest.mark.parametrize('depth,width,skips,M,rot', [
    (1, 8, (), 5, False), (2, 16, (0,), 7, False), (4, 16, (1, 2), 5, True), (8, 32, (4,), 16, True),
    (8, 256, (4,), 32, True), (8, 256, (4,), 64, True), (8, 256, (4,), 1, True), (8, 256, (4,), 33, False),
    (3, 40, (2,), 70, True)])
def test_joint_mlp_matches_oracle(depth, width, skips, M, rot):
    cfg = NetConfig(10, 6, width, depth, skips, rotation_head=rot)
    ws, bs = OD.init_params(seed=depth * 100 + width, head_std=0.05, width=width, depth=depth, skips=skips)
    gen = torch.Generator().manual_seed(M)
    bs[-1] = torch.randn(3, generator=gen) * 0.05
    joints = torch.randn(M, 3, generator=gen) * 0.4
    _check(cfg, ws, bs, joints, torch.tensor([0.37]), (depth, width, skips, M, rot))


def test_joint_mlp_on_reference_golden_weights():
    """the weights of the reference-generated fixture (two structures, one with two skip layers)"""
    d = np.load(os.path.join(G, 'deform_net.npz'))
    for ci in range(int(d['n'])):
        c = d[f'cfg{ci}']
        width, depth, skips = int(c[1]), int(c[2]), tuple(int(x) for x in c[3:])
        n = depth + 3
        ws = [torch.from_numpy(d[f'w{ci}_{i}']) for i in range(n)]
        bs = [torch.from_numpy(d[f'b{ci}_{i}']) for i in range(n)]
        cfg = NetConfig(10, 6, width, depth, skips, rotation_head=False)
        _check(cfg, ws, bs, torch.from_numpy(d[f'joints{ci}']), torch.from_numpy(d[f't{ci}']), ('golden', ci))
        # against the reference module's own numbers ('freq_torch' encoders: exact cos instead of sin(y + fl32(pi/2)),
        # a 4e-5 difference in the encoded input)
        theta = _flatten(cfg, ws, bs).float().to(DEV)
        got, _ = joint_mlp_forward_raw(cfg, theta, torch.from_numpy(d[f'joints{ci}']).float().to(DEV),
                                       torch.from_numpy(d[f't{ci}']).float().to(DEV))
        for o, name in zip(got, ('o_r', 'o_rot', 'o_s')):
            assert _rel(o, torch.from_numpy(d[f'{name}{ci}'])) <= 2e-3


def test_module_mirror_and_checkpoint_layout():
    net = SimpleDeformationNetwork(pos_enc_p_cfg=dict(degree=10), pos_enc_t_cfg=dict(degree=6), width=64, depth=8,
                                   skips=(4,)).to(DEV)
    sd = net.reference_state_dict()
    assert sd['dynamic_net.net.5.weight'].shape == (64, 64 + 76) and sd['dynamic_net.last.2.weight'].shape == (3, 64)
    joints = (torch.randn(12, 3, device=DEV) * 0.4).requires_grad_(True)
    t = torch.tensor(0.25)
    outs = net(joints, t)
    assert [tuple(o.shape) for o in outs] == [(12, 4), (12, 4), (12, 3)]
    gs = [torch.randn_like(o) for o in outs]
    torch.autograd.backward(outs, gs)
    ws = [sd[f'dynamic_net.net.{i}.weight'].cpu() for i in range(8)] + [sd[f'dynamic_net.last.{j}.weight'].cpu() for j in range(3)]
    bs = [sd[f'dynamic_net.net.{i}.bias'].cpu() for i in range(8)] + [sd[f'dynamic_net.last.{j}.bias'].cpu() for j in range(3)]
    ws64 = [w.double().requires_grad_(True) for w in ws]
    bs64 = [b.double().requires_grad_(True) for b in bs]
    j64 = joints.detach().cpu().double().requires_grad_(True)
    ref = OD.forward(ws64, bs64, j64, t.double(), rotation_head=False)
    for o, r in zip(outs, ref):
        assert _rel(o, r) <= 1e-5
    gr = torch.autograd.grad(ref, [j64] + ws64, [g.cpu().double() for g in gs])
    assert _rel(joints.grad, gr[0]) <= 1e-4
    got_w0 = net.cfg.views(net.theta.grad)[0][0]
    assert _rel(got_w0, gr[1]) <= 1e-4
    other = SimpleDeformationNetwork(pos_enc_p_cfg=dict(degree=10), pos_enc_t_cfg=dict(degree=6), width=64, depth=8,
                                     skips=(4,)).to(DEV)
    other.load_reference_state_dict({k: v.clone() for k, v in sd.items()})
    assert torch.equal(other.theta, net.theta)
    with pytest.raises(RuntimeError):
        net(torch.zeros(3, 3), t)  # CPU tensor: no fallback
    with pytest.raises(NotImplementedError):
        SimpleDeformationNetwork(p_in_channels=5)


def test_hot_path_with_joint_network():
    """raw (graph-capturable) chain == autograd chain, gradients reach theta and joints, an iteration optimises"""
    cfg = S.CONFIGS['c1']
    sc = S.make_scene(cfg, views=1)
    dL = (torch.randn(3, cfg.H, cfg.W, generator=torch.Generator().manual_seed(2)) / (3 * cfg.H * cfg.W)).to(DEV)
    a = HotPath(sc, DEV, merged_sh=True, joint_mlp=True, head_std=0.02)
    out_a = a.step(0, dL)
    b = HotPath(sc, DEV, requires_grad=False, merged_sh=True, joint_mlp=True, head_std=0.02)
    assert torch.equal(a.params['theta'].detach(), b.params['theta'])
    out_b, grads = b.step_grads(0, dL)
    assert torch.equal(out_a['images'].detach(), out_b['images'])
    for n in ('theta', 'joints', 'xyz', 'sp_W', 'g_tr'):
        assert _rel(grads[n], a.params[n].grad) <= 1e-5, n
    assert float(grads['theta'].abs().max()) > 0
    from sk_gs_b200 import diff_gaussian_rasterization as DGR
    target = torch.rand(3, cfg.H, cfg.W, generator=torch.Generator().manual_seed(1)).to(DEV)
    # Adam moves every weight of the network by +-lr per step: with the reference's 1e-3 and a random target d_scale
    # balloons within a few steps (R x2.5 per step); a small lr keeps the fixed-capacity graph inside its headroom
    loop = TrainLoop(b, lrs={'theta': 1e-5})
    assert 'theta' in loop.names and 'sk_r' not in loop.names
    eager, eager_R = [], []
    for _ in range(3):
        eager.append(float(loop.step(0, target)['loss_terms'][2]))
        torch.cuda.synchronize()
        eager_R.append(int(DGR.last_header_words(DEV)[0]))
    assert eager[-1] < eager[0]
    c = HotPath(sc, DEV, requires_grad=False, merged_sh=True, joint_mlp=True, head_std=0.02)
    loop_c = TrainLoop(c, lrs={'theta': 1e-5})
    loop_c.capture(0, target, headroom=4.0)
    cap = loop_c.out['_capacity']
    replayed, replay_words = [], []
    for _ in range(3):
        out = loop_c.replay()
        torch.cuda.synchronize()
        replayed.append(float(out['loss_terms'][2]))
        replay_words.append(out['_header_words'].tolist())
    info = dict(eager=eager, eager_R=eager_R, replayed=replayed, replay_words=replay_words, capacity=cap)
    assert not c.overflowed(), info
    assert np.abs(np.array(replayed) - np.array(eager)).max() <= 1e-5, info
