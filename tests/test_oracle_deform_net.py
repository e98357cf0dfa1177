"""oracle/deform_net.py against golden vectors of the REFERENCE's own SimpleDeformationNetwork (networks/sk_gs.py:134-164,
imported unmodified by tests/golden/make_golden.py with the pure-torch encoder variant).  No GPU."""
import os

import numpy as np
import torch

from oracle import deform_net as OD

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _case(d, ci):
    cfg = d[f'cfg{ci}']
    M, width, depth, skips = int(cfg[0]), int(cfg[1]), int(cfg[2]), tuple(int(x) for x in cfg[3:])
    n = depth + 3
    ws = [torch.from_numpy(d[f'w{ci}_{i}']).requires_grad_(True) for i in range(n)]
    bs = [torch.from_numpy(d[f'b{ci}_{i}']).requires_grad_(True) for i in range(n)]
    return M, width, depth, skips, ws, bs


def test_structure_matches_reference_network():
    d = np.load(os.path.join(G, 'deform_net.npz'))
    for ci in range(int(d['n'])):
        M, width, depth, skips, ws, bs = _case(d, ci)
        enc, shapes = OD.layer_shapes(width=width, depth=depth, skips=skips)
        assert enc == 76 and [tuple(w.shape) for w in ws] == shapes
        joints = torch.from_numpy(d[f'joints{ci}']).requires_grad_(True)
        t = torch.from_numpy(d[f't{ci}'])
        outs = OD.forward(ws, bs, joints, t, skips=skips, cuda_formula=False, rotation_head=False)
        for o, name in zip(outs, ('o_r', 'o_rot', 'o_s')):
            assert np.abs(o.detach().numpy() - d[f'{name}{ci}']).max() <= 1e-12
        grads = torch.autograd.grad(outs, [joints] + ws + bs,
                                    [torch.from_numpy(d[f'{g}{ci}']) for g in ('g_r', 'g_rot', 'g_s')])
        n = len(ws)
        assert np.abs(grads[0].numpy() - d[f'd_joints{ci}']).max() <= 1e-9 * np.abs(d[f'd_joints{ci}']).max()
        for i in range(n):
            assert np.abs(grads[1 + i].numpy() - d[f'dw{ci}_{i}']).max() <= 1e-10 * max(1.0, np.abs(d[f'dw{ci}_{i}']).max())
            assert np.abs(grads[1 + n + i].numpy() - d[f'db{ci}_{i}']).max() <= 1e-10 * max(1.0, np.abs(d[f'db{ci}_{i}']).max())


def test_cuda_encoder_formula_is_close_to_the_exact_one():
    x = torch.randn(50, 3, dtype=torch.float64) * 0.5
    a = OD.freq_encode(x, 10, cuda_formula=True)
    b = OD.freq_encode(x, 10, cuda_formula=False)
    assert a.shape == (50, 63) and torch.equal(a[:, :3], x)
    assert float((a - b).abs().max()) <= 4e-5  # fp32 rounding of (2^9 x + pi/2)
    # layout: [x | sin f0 | cos f0 | sin f1 | ...] (freqencoder.cu:21-29)
    assert torch.allclose(b[:, 3:6], torch.sin(x)) and torch.allclose(b[:, 6:9], torch.cos(x))
    assert torch.allclose(b[:, 9:12], torch.sin(2 * x))
    # backward through the stored outputs (freqencoder.cu:52-56)
    xg = x.clone().requires_grad_(True)
    OD.freq_encode(xg, 10, cuda_formula=True).sum().backward()
    xe = x.clone().requires_grad_(True)
    OD.freq_encode(xe, 10, cuda_formula=False).sum().backward()
    assert float((xg.grad - xe.grad).abs().max()) <= 4e-5 * 1024


def test_rotation_head_and_init():
    ws, bs = OD.init_params(seed=1, width=32, depth=8)
    assert float(bs[-1].abs().max()) == 0.0 and float(ws[-1].abs().max()) < 1e-5
    outs = OD.forward(ws, bs, torch.randn(7, 3), torch.tensor([0.5]))
    assert outs[0].shape == (7, 4) and outs[1].shape == (7, 4) and outs[2].shape == (7, 3)
    assert torch.allclose(outs[0].norm(dim=-1), torch.ones(7), atol=1e-6)
    assert float((outs[0] - torch.tensor([0., 0., 0., 1.])).abs().max()) < 1e-4  # near identity at init (sk_gs.py:542-545)
