"""The closed-form SSIM gradient that csrc/image_loss.cu implements (separable window statistics, three partial
derivative maps, adjoint convolution) restated in numpy and checked against autograd of the oracle.  No GPU: this
guards the derivation, the kernel itself is checked on the GPU in tests/test_gpu_loss_adam.py."""
import numpy as np
import torch

from oracle import losses as OL


def _conv_sep(a, w):
    """zero-padded separable 11-tap convolution of a [C,H,W] array."""
    C, H, W = a.shape
    p = np.pad(a, ((0, 0), (5, 5), (5, 5)))
    h = sum(w[k] * p[:, :, k:k + W] for k in range(11))
    return sum(w[k] * h[:, k:k + H, :] for k in range(11))


def closed_form(x, y, w_img, w_ssim, mse=False):
    w = OL.gaussian_window().double().numpy()
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    mu1, mu2 = _conv_sep(x, w), _conv_sep(y, w)
    e11, e22, e12 = _conv_sep(x * x, w), _conv_sep(y * y, w), _conv_sep(x * y, w)
    s1, s2, s12 = e11 - mu1 * mu1, e22 - mu2 * mu2, e12 - mu1 * mu2
    A1, A2 = 2 * mu1 * mu2 + C1, 2 * s12 + C2
    B1, B2 = mu1 * mu1 + mu2 * mu2 + C1, s1 + s2 + C2
    S = A1 * A2 / (B1 * B2)
    Dm = 2 * mu2 * (A2 - A1) / (B1 * B2) - 2 * mu1 * S * (1 / B1 - 1 / B2)
    D11 = -S / B2
    D12 = 2 * A1 / (B1 * B2)
    dS = _conv_sep(Dm, w) + 2 * x * _conv_sep(D11, w) + y * _conv_sep(D12, w)
    n = x.size
    d = x - y
    pix = (d * d).mean() if mse else np.abs(d).mean()
    dpix = 2 * d if mse else np.sign(d)
    terms = np.array([pix, 1 - S.mean(), w_img * pix + w_ssim * (1 - S.mean())])
    return terms, w_img * dpix / n - w_ssim * dS / n


def test_closed_form_equals_autograd():
    g = torch.Generator().manual_seed(5)
    for (H, W), mse in (((23, 31), False), ((40, 37), True), ((7, 5), False)):
        y = torch.rand(3, H, W, generator=g, dtype=torch.float64)
        x = (y + 0.2 * torch.randn(3, H, W, generator=g, dtype=torch.float64)).clamp(0, 1)
        terms, grad = closed_form(x.numpy(), y.numpy(), 0.8, 0.2, mse)
        t_ref, g_ref = OL.image_loss(x, y, 0.8, 0.2, 'mse' if mse else 'l1')
        # the reference convolves with the 2-D window rounded to fp32 (ssim.py:16), the kernel with its two 1-D factors:
        # a 1e-8 relative difference, far below the fp32 tolerances of the GPU test
        assert np.abs(terms - t_ref.numpy()).max() <= 1e-8
        assert np.abs(grad - g_ref.numpy()).max() <= 1e-6 * np.abs(g_ref.numpy()).max()
