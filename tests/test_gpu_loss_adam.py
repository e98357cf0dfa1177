"""GPU parity of the two widening rows, through the C ABI:
  skgs_image_loss  vs  golden vectors of the reference's own ImageLoss / SSIM_Loss modules (tests/golden/loss.npz) and
                       vs oracle/losses.py at the benchmark's image size;
  skgs_adam_step   vs  golden torch.optim.Adam trajectories (tests/golden/adam.npz) and vs oracle/losses.adam_step;
  TrainLoop        vs  the same iteration assembled from the autograd path + oracle loss + torch.optim.Adam.
Tolerances (floating point): loss terms 2e-6 absolute, gradients 1e-4 of the tensor's max (north_star), Adam 2e-6."""
import os

import numpy as np
import pytest
import torch

from oracle import losses as OL
from sk_gs_b200 import scene as S
from sk_gs_b200.losses import ImageLoss, SSIM_Loss, image_loss_raw, image_ssim_loss
from sk_gs_b200.optim import Adam, adam_hyper, adam_step_raw
from sk_gs_b200.pipeline import HotPath
from sk_gs_b200.train import DEFAULT_LRS, TrainLoop

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
DEV = 'cuda:0'


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


# ------------------------------------------------------------------------------------------------------------ loss
@pytest.mark.parametrize('layout', ['chw', 'hwc', 'rgba'])
def test_image_loss_matches_reference_golden(layout):
    d = np.load(os.path.join(G, 'loss.npz'))
    for i in range(int(d['n'])):
        img = torch.from_numpy(d[f'img_{i}'])[0].permute(2, 0, 1).contiguous().to(DEV)
        gt_hwc = torch.from_numpy(d[f'gt_{i}'])[0].to(DEV)
        if layout == 'chw':
            tgt = gt_hwc.permute(2, 0, 1).contiguous()
        elif layout == 'hwc':
            tgt = gt_hwc
        else:
            tgt = torch.cat([gt_hwc, torch.rand_like(gt_hwc[..., :1])], -1)[..., :3]  # RGB view of RGBA (sk_gs.py:1525)
            assert not tgt.is_contiguous()
        terms, g = image_loss_raw(img, tgt, 0.8, 0.2, 'l1')
        terms = terms.cpu().double()
        assert abs(float(terms[0]) - float(d[f'l1_f64_{i}'])) <= 2e-6
        assert abs(float(terms[1]) - float(d[f'ssim_f64_{i}'])) <= 2e-6
        assert abs(float(terms[2]) - float(d[f'total_f64_{i}'])) <= 2e-6
        ref = torch.from_numpy(d[f'g_total_f64_{i}'])[0].permute(2, 0, 1)
        assert _rel(g.cpu().double(), ref) <= 1e-4
        # the SSIM part alone (the L1 part is a sign pattern and hides relative errors of the rest)
        _, g_s = image_loss_raw(img, tgt, 0.0, 1.0, 'l1', grad_scale=2.0)
        ref_s = torch.from_numpy(d[f'g_ssim_f64_{i}'])[0].permute(2, 0, 1) * 2.0
        assert _rel(g_s.cpu().double(), ref_s) <= 1e-4
        t_m, g_m = image_loss_raw(img, tgt, 1.0, 0.0, 'mse')
        assert abs(float(t_m[0]) - float(d[f'mse_f64_{i}'])) <= 2e-6
        assert _rel(g_m.cpu().double(), torch.from_numpy(d[f'g_mse_f64_{i}'])[0].permute(2, 0, 1)) <= 1e-5
        t_f, g_f = image_loss_raw(img, tgt, 0.8, 0.2, 'l1', need_grad=False)  # forward only
        assert g_f is None and torch.equal(t_f.cpu().double(), terms)


def test_loss_modules_mirror_reference_interface():
    d = np.load(os.path.join(G, 'loss.npz'))
    img = torch.from_numpy(d['img_1']).to(DEV).requires_grad_(True)  # [1,H,W,3] as at sk_gs.py:1527
    gt = torch.from_numpy(d['gt_1']).to(DEV)
    l1 = ImageLoss(method='l1')(img, gt)
    ss = SSIM_Loss()(img, gt)
    (0.8 * l1 + 0.2 * ss).backward()
    assert abs(float(l1) - float(d['l1_f64_1'])) <= 2e-6 and abs(float(ss) - float(d['ssim_f64_1'])) <= 2e-6
    assert img.grad.shape == img.shape
    assert _rel(img.grad.cpu().double(), torch.from_numpy(d['g_total_f64_1'])) <= 1e-4
    img2 = torch.from_numpy(d['img_1'])[0].permute(2, 0, 1).contiguous().to(DEV).requires_grad_(True)  # [3,H,W]
    tot = image_ssim_loss(img2, gt)
    (3.0 * tot).backward()
    assert abs(float(tot) - float(d['total_f64_1'])) <= 2e-6
    assert _rel(img2.grad.cpu().double(), 3.0 * torch.from_numpy(d['g_total_f64_1'])[0].permute(2, 0, 1)) <= 1e-4
    with pytest.raises(ValueError):
        ImageLoss(method='huber')
    with pytest.raises(RuntimeError):
        image_loss_raw(torch.rand(3, 8, 8), torch.rand(3, 8, 8))  # CPU tensors: no fallback
    with pytest.raises(RuntimeError):
        image_loss_raw(torch.rand(3, 8, 8, device=DEV), torch.rand(3, 8, 9, device=DEV))


@pytest.mark.parametrize('hw', [(800, 800), (1080, 1920), (33, 1), (1, 1)])
def test_image_loss_full_size_against_oracle_and_properties(hw):
    H, W = hw
    g = torch.Generator().manual_seed(H * 7 + W)
    gt = torch.rand(3, H, W, generator=g)
    img = (gt + 0.1 * torch.randn(3, H, W, generator=g)).clamp(0, 1)
    terms, grad = image_loss_raw(img.to(DEV), gt.to(DEV))
    t_ref, g_ref = OL.image_loss(img.double(), gt.double())
    assert float((terms.cpu().double() - t_ref).abs().max()) <= 2e-6
    assert _rel(grad.cpu().double(), g_ref) <= 1e-4
    # identical images: both terms vanish, the SSIM gradient vanishes
    t0, g0 = image_loss_raw(gt.to(DEV), gt.to(DEV))
    assert float(t0.abs().max()) <= 1e-6
    assert float(g0.abs().max()) <= 1e-4 * float(g_ref.abs().max())
    # the gradient is linear in grad_scale, exactly for powers of two
    _, g4 = image_loss_raw(img.to(DEV), gt.to(DEV), grad_scale=4.0)
    assert torch.equal(g4, 4.0 * grad)


# ------------------------------------------------------------------------------------------------------------ Adam
def test_adam_matches_torch_optim_golden():
    d = np.load(os.path.join(G, 'adam.npz'))
    n, lrs = int(d['n']), [float(x) for x in d['lrs']]
    p = [torch.from_numpy(d[f'p0_{i}']).to(DEV) for i in range(n)]
    m = [torch.zeros_like(t) for t in p]
    v = [torch.zeros_like(t) for t in p]
    for t in range(int(d['steps'])):
        gs = [torch.from_numpy(d[f'g{t}_{i}']).to(DEV) for i in range(n)]
        adam_step_raw(p, gs, m, v, lrs, t + 1, 0.9, 0.999, 1e-15)
        for i in range(n):
            for got, name in ((p[i], 'p'), (m[i], 'm'), (v[i], 'v')):
                ref = torch.from_numpy(d[f'{name}{t + 1}_{i}'])
                assert _rel(got.cpu(), ref) <= 2e-6, (name, t, i)


@pytest.mark.parametrize('cols', [24, 30, 5])
def test_adam_variants_agree(cols):
    """vector / scalar / unaligned / compact-gradient (row length a multiple of 4 or not) / interleaved-lr /
    dynamic-table code paths against each other and against the oracle."""
    g = torch.Generator().manual_seed(11)
    rows, K = 3001, 5
    idx = torch.stack([torch.randperm(cols, generator=g)[:K] for _ in range(rows)])
    gk = torch.randn(rows, K, generator=g)
    dense = OL.scatter_knn_grad(gk, idx, cols)
    p0 = torch.randn(rows, cols, generator=g)
    big0, gbig = torch.randn(40000 + 3, generator=g), torch.randn(40000 + 3, generator=g)
    sh0, gsh = torch.randn(500, 16, 3, generator=g), torch.randn(500, 16, 3, generator=g) * 1e-3
    lr, lr2 = 2.5e-3, 1.25e-4

    def run(compact, dynamic):
        p = [p0.to(DEV), big0.to(DEV)[1:], sh0.to(DEV), torch.zeros(0, device=DEV)]  # [1:] -> not 16-byte aligned
        m = [torch.zeros_like(t) for t in p]
        v = [torch.zeros_like(t) for t in p]
        if p[1].data_ptr() % 16 == 0:
            pytest.skip('allocator returned an unexpected alignment')
        grads = [gk.to(DEV) if compact else dense.to(DEV), gbig.to(DEV)[1:].contiguous(), gsh.to(DEV),
                 torch.zeros(0, device=DEV)]
        lrs = [1e-3, 1e-2, (lr, lr2, 48, 3), 1.0]
        for step in (1, 2, 3):
            dyn = torch.tensor(adam_hyper(lrs, step), dtype=torch.float32, device=DEV) if dynamic else None
            adam_step_raw(p, grads, m, v, lrs, 1 if dynamic else step, 0.9, 0.999, 1e-15, grad_scale=0.5,
                          knn_indices=[idx.to(DEV) if compact else None, None, None, None], dynamic_hyper=dyn)
        return [t.cpu() for t in p + m + v]

    base = run(False, False)
    for other in (run(True, False), run(False, True), run(True, True)):
        for a, b in zip(base, other):
            assert a.shape == b.shape and float((a - b).abs().max()) <= 1e-7 * max(1.0, float(b.abs().max())) \
                if a.numel() else True
    # oracle (float32 torch on the CPU)
    ref_p, ref_m, ref_v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    rb, rbm, rbv = big0[1:].clone(), torch.zeros(40002), torch.zeros(40002)
    lr_map = torch.where(torch.arange(48) < 3, lr, lr2).repeat(500 * 48 // 48).view(500, 16, 3)
    rs, rsm, rsv = sh0.clone(), torch.zeros_like(sh0), torch.zeros_like(sh0)
    for step in (1, 2, 3):
        ref_p, ref_m, ref_v = OL.adam_step(ref_p, dense * 0.5, ref_m, ref_v, 1e-3, step)
        rb, rbm, rbv = OL.adam_step(rb, gbig[1:] * 0.5, rbm, rbv, 1e-2, step)
        new, rsm, rsv = OL.adam_step(rs, gsh * 0.5, rsm, rsv, 1.0, step)
        rs = rs + (new - rs) * lr_map
    assert _rel(base[0], ref_p) <= 2e-6 and _rel(base[1], rb) <= 2e-6 and _rel(base[2], rs) <= 2e-6
    assert _rel(base[4], ref_m) <= 2e-6 and _rel(base[8], ref_v) <= 2e-6


def test_adam_class_follows_torch_optimizer():
    g = torch.Generator().manual_seed(3)
    shapes = [(1000, 3), (1000, 1), (17,)] * 7  # 21 tensors -> two launches
    init = [torch.randn(*s, generator=g) for s in shapes]
    ours = [t.to(DEV).requires_grad_(True) for t in init]
    ref = [t.to(DEV).requires_grad_(True) for t in init]
    groups = lambda ps: [{'params': [p], 'lr': 1e-3 * (1 + i % 4), 'name': str(i)} for i, p in enumerate(ps)]  # noqa
    o1 = Adam(groups(ours), lr=0.0, eps=1e-15)
    o2 = torch.optim.Adam(groups(ref), lr=0.0, eps=1e-15)
    for step in range(3):
        for a, b in zip(ours, ref):
            gr = torch.randn(a.shape, generator=g).to(DEV)
            a.grad, b.grad = gr.clone(), gr.clone()
        if step == 1:  # schedulers write group['lr'] (gaussian_splatting.py:466-471)
            o1.param_groups[0]['lr'] = o2.param_groups[0]['lr'] = 5e-2
        o1.step()
        o2.step()
    for a, b in zip(ours, ref):
        assert _rel(a.detach(), b.detach()) <= 2e-6
    o1.zero_grad()
    assert all(p.grad is None for p in ours)
    with pytest.raises(RuntimeError):
        adam_step_raw([torch.zeros(4)], [torch.zeros(4)], [torch.zeros(4)], [torch.zeros(4)], [1e-3], 1)  # CPU
    with pytest.raises(RuntimeError):
        z = torch.zeros(4, device=DEV)
        adam_step_raw([z], [z], [z], [z], [1e-3], 0)  # step < 1


# ------------------------------------------------------------------------------------------------------- iteration
def _reference_iterations(sc, target, n):
    """render (autograd path of this repo, parity-tested separately) -> oracle loss -> torch.optim.Adam."""
    hp = HotPath(sc, DEV)
    names = ['xyz', 'f_dc', 'f_rest', 'opacity', 'scaling', 'rotation', 'sp_W', 'joints', 'sk_r', 'sk_d_rot',
             'sk_d_scale', 'g_tr']
    lr_of = dict(DEFAULT_LRS, f_dc=DEFAULT_LRS['shs'][0], f_rest=DEFAULT_LRS['shs'][1])
    opt = torch.optim.Adam([{'params': [hp.params[n_]], 'lr': lr_of[n_]} for n_ in names], lr=0.0, eps=1e-15)
    losses = []
    for _ in range(n):
        opt.zero_grad()
        img = hp.render(0)['images']
        loss = 0.8 * OL.pixel_loss(img.permute(1, 2, 0)[None], target.permute(1, 2, 0)[None]) + \
            0.2 * OL.ssim_loss(img[None], target[None])
        loss.backward()
        opt.step()
        losses.append(float(loss))
    return hp, losses


@pytest.mark.parametrize('graph', [False, True])
def test_train_iteration_matches_autograd_plus_torch_adam(graph):
    cfg = S.CONFIGS['c1']
    sc = S.make_scene(cfg, views=1)
    target = torch.rand(3, cfg.H, cfg.W, generator=torch.Generator().manual_seed(1)).to(DEV)
    n = 3
    ref, ref_losses = _reference_iterations(sc, target, n)
    hp = HotPath(sc, DEV, requires_grad=False, merged_sh=True)
    loop = TrainLoop(hp)
    before = {k: hp.params[k].detach().clone() for k in loop.names}
    losses = []
    if graph:
        loop.capture(0, target, headroom=4.0)
        for k in loop.names:  # capture leaves parameters and moments untouched
            assert torch.equal(hp.params[k], before[k]) and float(loop.exp_avg[k].abs().max()) == 0.0
        for _ in range(n):
            out = loop.replay()
            torch.cuda.synchronize()
            losses.append(float(out['loss_terms'][2]))
        assert not hp.overflowed()
    else:
        for _ in range(n):
            losses.append(float(loop.step(0, target)['loss_terms'][2]))
    assert np.abs(np.array(losses) - np.array(ref_losses)).max() <= 1e-5
    assert losses[-1] < losses[0]  # it optimises
    ref_sh = torch.cat((ref.params['f_dc'], ref.params['f_rest']), 1).detach()
    lr_sh = torch.where(torch.arange(48, device=DEV) < 3, DEFAULT_LRS['shs'][0], DEFAULT_LRS['shs'][1]).view(1, 16, 3)
    for k in loop.names:
        want = ref_sh if k == 'shs' else ref.params[k].detach()
        lr = lr_sh if k == 'shs' else DEFAULT_LRS[k]
        # Adam normalises the gradient: where it is at noise level its sign - and with it a whole +-lr step - may
        # differ between two correct implementations, so compare in units of lr and bound the fraction of outliers
        err = ((hp.params[k] - want).abs() / lr).flatten()
        moved = ((want - before[k]).abs() / lr).flatten()
        assert float(moved.max()) > 0.5, k
        assert float((err > 0.05).float().mean()) <= 0.02, (k, float(err.max()))
        assert float(err.median()) <= 1e-2, k
