"""oracle/losses.py against golden vectors of the REFERENCE's own loss modules (networks/losses/ssim.py, image_loss.py,
imported unmodified by tests/golden/make_golden.py) and of torch.optim.Adam (the reference's optimizer).  No GPU."""
import os

import numpy as np
import torch

from oracle import losses as OL

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def test_window_matches_reference_definition():
    w = OL.gaussian_window()
    assert w.dtype == torch.float32 and w.numel() == 11
    assert abs(float(w.sum()) - 1.0) <= 1e-6 and torch.equal(w, w.flip(0))
    assert abs(float(w[5] / w[4]) - np.exp(1.0 / 4.5)) <= 1e-6


def test_losses_match_reference_modules():
    d = np.load(os.path.join(G, 'loss.npz'))
    for i in range(int(d['n'])):
        for tag, dt, tol in (('f64', torch.float64, 1e-12), ('f32', torch.float32, 2e-6)):
            img = torch.from_numpy(d[f'img_{i}']).to(dt)  # [1,H,W,3]
            gt = torch.from_numpy(d[f'gt_{i}']).to(dt)
            chw, tchw = img[0].permute(2, 0, 1).contiguous(), gt[0].permute(2, 0, 1).contiguous()
            terms, g = OL.image_loss(chw, tchw, 0.8, 0.2, 'l1')
            assert abs(float(terms[0]) - float(d[f'l1_{tag}_{i}'])) <= tol
            assert abs(float(terms[1]) - float(d[f'ssim_{tag}_{i}'])) <= tol
            assert abs(float(terms[2]) - float(d[f'total_{tag}_{i}'])) <= tol
            ref = torch.from_numpy(d[f'g_total_{tag}_{i}'])[0].permute(2, 0, 1)
            assert float((g - ref).abs().max()) <= tol * 10 * max(1.0, float(ref.abs().max()) * 1e4)
            t2, g2 = OL.image_loss(chw, tchw, 1.0, 0.0, 'mse')
            assert abs(float(t2[0]) - float(d[f'mse_{tag}_{i}'])) <= tol
            ref2 = torch.from_numpy(d[f'g_mse_{tag}_{i}'])[0].permute(2, 0, 1)
            assert float((g2 - ref2).abs().max()) <= tol
            t3, g3 = OL.image_loss(chw, tchw, 0.0, 1.0, 'l1', grad_scale=3.0)
            ref3 = torch.from_numpy(d[f'g_ssim_{tag}_{i}'])[0].permute(2, 0, 1) * 3.0
            assert float((g3 - ref3).abs().max()) <= tol * 10 * max(1.0, float(ref3.abs().max()) * 1e4)


def test_ssim_of_identical_images_is_one():
    x = torch.rand(1, 3, 40, 33, dtype=torch.float64)
    assert abs(float(OL.ssim_loss(x, x))) <= 1e-12


def test_adam_matches_torch_optim():
    d = np.load(os.path.join(G, 'adam.npz'))
    lrs = d['lrs']
    for i in range(int(d['n'])):
        p = torch.from_numpy(d[f'p0_{i}'])
        m, v = torch.zeros_like(p), torch.zeros_like(p)
        for t in range(int(d['steps'])):
            g = torch.from_numpy(d[f'g{t}_{i}'])
            p, m, v = OL.adam_step(p, g, m, v, float(lrs[i]), t + 1)
            for got, name in ((p, 'p'), (m, 'm'), (v, 'v')):
                ref = torch.from_numpy(d[f'{name}{t + 1}_{i}'])
                assert float((got - ref).abs().max()) <= 2e-6 * max(float(ref.abs().max()), 1e-30), (name, t, i)


def test_scatter_knn_grad():
    idx = torch.tensor([[0, 3, 5], [2, 1, 4]])
    g = torch.tensor([[1., 2., 3.], [4., 5., 6.]])
    out = OL.scatter_knn_grad(g, idx, 6)
    assert out.tolist() == [[1, 0, 0, 2, 0, 3], [0, 5, 4, 0, 6, 0]]
