"""oracle/losses.py - pure-torch CPU restatement of the photometric loss and of the Adam update.

TEST INFRASTRUCTURE ONLY (see oracle/raster_oracle.c header): only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg import this module.  The product (sk_gs_b200/) never does.

Parity status: PINNED.  `ssim_loss` / `pixel_loss` are checked against the reference's own `SSIM_Loss` and `ImageLoss`
modules (imported unmodified from /root/reference by tests/golden/make_golden.py -> tests/golden/loss.npz, values and
autograd gradients); `adam_step` is checked against torch.optim.Adam itself (the reference's optimizer,
networks/gaussian_splatting.py:445-453, is torch's) on trajectories stored in tests/golden/adam.npz.
"""
from __future__ import annotations

import math
from typing import Tuple

import torch
import torch.nn.functional as F
from torch import Tensor


def gaussian_window(window_size: int = 11, sigma: float = 1.5) -> Tensor:
    """networks/losses/ssim.py:9-11: fp32 samples of exp(-(x - size//2)^2 / (2 sigma^2)), normalised in fp32."""
    g = torch.tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)],
                     dtype=torch.float32)
    return g / g.sum()


def ssim_map(img1: Tensor, img2: Tensor, window_size: int = 11) -> Tensor:
    """networks/losses/ssim.py:14-18 (2-D window = outer product, fp32) and :46-62 (_ssim).  img*: [B,C,H,W]."""
    C = img1.shape[-3]
    w1 = gaussian_window(window_size).unsqueeze(1)
    window = (w1 @ w1.t()).float()[None, None].expand(C, 1, window_size, window_size).to(img1)
    pad = window_size // 2
    mu1 = F.conv2d(img1, window, padding=pad, groups=C)
    mu2 = F.conv2d(img2, window, padding=pad, groups=C)
    mu1_sq, mu2_sq, mu1_mu2 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    sigma1_sq = F.conv2d(img1 * img1, window, padding=pad, groups=C) - mu1_sq
    sigma2_sq = F.conv2d(img2 * img2, window, padding=pad, groups=C) - mu2_sq
    sigma12 = F.conv2d(img1 * img2, window, padding=pad, groups=C) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    return ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))


def ssim_loss(img1: Tensor, img2: Tensor) -> Tensor:
    """SSIM_Loss.forward with reduction='mean' (ssim.py:27-41): 1 - mean of the map.  [B,C,H,W] or [B,H,W,3]."""
    if img1.shape[-1] == 3:
        img1 = img1.permute(0, 3, 1, 2)
    if img2.shape[-1] == 3:
        img2 = img2.permute(0, 3, 1, 2)
    return 1.0 - ssim_map(img1, img2).mean()


def pixel_loss(pred: Tensor, gt: Tensor, method: str = 'l1') -> Tensor:
    """ImageLoss.forward, unmasked (networks/losses/image_loss.py:20-33): mean |.| or mean (.)^2 over [..., :3]."""
    pred, gt = pred[..., :3], gt[..., :3]
    return (pred - gt).abs().mean() if method == 'l1' else ((pred - gt) ** 2).mean()


def image_loss(image_chw: Tensor, target_chw: Tensor, w_image: float = 0.8, w_ssim: float = 0.2, method: str = 'l1',
               grad_scale: float = 1.0) -> Tuple[Tensor, Tensor]:
    """The two image terms of SkeletonGaussianSplatting.loss (networks/sk_gs.py:1524-1529; weights exps/default.yaml:83-84)
    and the gradient w.r.t. the rendered image.  Returns (terms[3] = pixel, ssim, weighted total; dL/dimage [3,H,W])."""
    x = image_chw.detach().clone().requires_grad_(True)
    a = pixel_loss(x.permute(1, 2, 0)[None], target_chw.permute(1, 2, 0)[None], method)
    b = ssim_loss(x[None], target_chw[None])
    total = w_image * a + w_ssim * b
    (g,) = torch.autograd.grad(total * grad_scale, x)
    return torch.stack([a.detach(), b.detach(), total.detach()]), g


def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, lr: float, step: int, beta1: float = 0.9,
              beta2: float = 0.999, eps: float = 1e-15) -> Tuple[Tensor, Tensor, Tensor]:
    """torch.optim.Adam, one parameter, no weight decay / amsgrad / maximize (torch/optim/adam.py `_single_tensor_adam`);
    hyper-parameters of exps/default.yaml:121-125.  `step` is the 1-based count AFTER the increment.  Out of place."""
    m = m + (g - m) * (1 - beta1)
    v = v * beta2 + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


def scatter_knn_grad(grad_knn: Tensor, indices: Tensor, cols: int) -> Tensor:
    """Dense [rows, cols] gradient of sp_W from its compact [rows, K] form (the gather at sk_gs.py:767 transposed)."""
    out = torch.zeros(grad_knn.shape[0], cols, dtype=grad_knn.dtype)
    out.scatter_add_(1, indices.long(), grad_knn)
    return out
