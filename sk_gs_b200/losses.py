"""Photometric loss of the training step: mirrors the reference's `ImageLoss`
(/root/reference/networks/losses/image_loss.py:8-36) and `SSIM_Loss` (networks/losses/ssim.py:21-41) modules, plus the
fused form `image_ssim_loss` that evaluates both terms of `SkeletonGaussianSplatting.loss` (networks/sk_gs.py:1524-1529,
weights exps/default.yaml:83-84) and their gradient in two launches of libskgs_b200.so (csrc/image_loss.cu).

No CPU path: CPU tensors raise.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor, nn

from . import _lib
from .diff_gaussian_rasterization import _f32c

_METHODS = {'l1': 0, 'mse': 1}


def _as_chw(image: Tensor) -> Tensor:
    """[3,H,W] contiguous view of a [3,H,W] / [1,3,H,W] / [H,W,3] / [1,H,W,3] image (a copy only if the memory is
    pixel-major; the rasterizer's output permuted to HWC, sk_gs.py:1229, maps back without a copy)."""
    if image.ndim == 4:
        if image.shape[0] != 1:
            raise RuntimeError('image loss takes one image per call (the reference passes [1,H,W,C], sk_gs.py:1527)')
        image = image[0]
    if image.ndim != 3:
        raise RuntimeError(f'expected an image with 3 or 4 dims, got shape {tuple(image.shape)}')
    if image.shape[0] != 3 and image.shape[-1] == 3:
        image = image.permute(2, 0, 1)
    if image.shape[0] != 3:
        raise RuntimeError(f'expected 3 colour channels, got shape {tuple(image.shape)}')
    return _f32c(image)


def _target(target: Tensor, layout: Optional[str] = None) -> Tuple[Tensor, int]:
    """(tensor, pixel stride): channel-major targets are used as they are (stride 0); pixel-major RGB / RGBA targets
    (datasets hand out [H,W,3|4], sk_gs.py:1525 slices [..., :3]) are read in place with stride 3 / 4.
    `layout` ('chw' | 'hwc') overrides the shape heuristic, which cannot tell [3,H,W] from [H,W,3|4] when H or W is 3
    or 4 (it then prefers pixel-major, what the datasets hand out)."""
    if target.ndim == 4:
        if target.shape[0] != 1:
            raise RuntimeError('image loss takes one target per call, got a batch of %d' % target.shape[0])
        target = target[0]
    if target.ndim != 3:
        raise RuntimeError(f'expected a target with 3 or 4 dims, got shape {tuple(target.shape)}')
    if layout not in (None, 'chw', 'hwc'):
        raise ValueError(f"layout must be 'chw', 'hwc' or None, got {layout!r}")
    if layout == 'chw' or (layout is None and target.shape[0] == 3 and target.shape[-1] not in (3, 4)):
        if target.shape[0] != 3:
            raise RuntimeError(f'channel-major target needs 3 channels, got shape {tuple(target.shape)}')
        return _f32c(target), 0
    if target.shape[-1] not in (3, 4):
        raise RuntimeError(f'cannot interpret target of shape {tuple(target.shape)}')
    H, W, Ct = target.shape
    if Ct == 3 and target.dtype == torch.float32 and target.stride() == (W * 4, 4, 1):
        return target, 4  # RGB view of a contiguous RGBA image: read in place
    return _f32c(target), Ct


def image_loss_raw(image: Tensor, target: Tensor, w_image: float = 0.8, w_ssim: float = 0.2, method: str = 'l1',
                   grad_scale: float = 1.0, need_grad: bool = True, out: Optional[dict] = None,
                   target_layout: Optional[str] = None):
    """Autograd-free call: returns (terms float32[3] on the device = pixel term, 1 - mean SSIM, weighted total;
    dL/dimage [3,H,W] or None).  `out` may carry preallocated 'terms', 'dL_dimage', 'workspace' (CUDA-graph capture)."""
    if not image.is_cuda or not target.is_cuda:
        raise RuntimeError('image loss needs CUDA tensors (sk_gs_b200 has no CPU path)')
    L = _lib.lib()
    img = _as_chw(image.detach())
    tgt, stride = _target(target.detach(), target_layout)
    H, W = img.shape[1:]
    th, tw = (tgt.shape[1:] if stride == 0 else tgt.shape[:2])
    if (th, tw) != (H, W):
        raise RuntimeError(f'image {H}x{W} and target {th}x{tw} differ in size')
    dev = img.device
    out = {} if out is None else out
    nbytes = L.skgs_image_loss_workspace_bytes(H, W)
    ws = out.get('workspace')
    if ws is None or ws.numel() < nbytes:
        ws = out['workspace'] = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    terms = out.get('terms')
    if terms is None:
        terms = out['terms'] = torch.empty(3, device=dev)
    grad = None
    if need_grad:
        grad = out.get('dL_dimage')
        if grad is None:
            grad = out['dL_dimage'] = torch.empty_like(img)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(L.skgs_image_loss(H, W, img.data_ptr(), tgt.data_ptr(), stride, _METHODS[method], float(w_image),
                                     float(w_ssim), float(grad_scale), ws.data_ptr(), terms.data_ptr(),
                                     _lib.ptr(grad), st), 'skgs_image_loss')
    return terms, grad


class _ImageSSIMLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, target, w_image, w_ssim, method, index):
        need = image.requires_grad
        terms, grad = image_loss_raw(image, target, w_image, w_ssim, method, 1.0, need_grad=need)
        ctx.shape = image.shape
        ctx.save_for_backward(grad)
        return terms[2 if index is None else index].clone()

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        if grad is None:
            return (None,) * 6
        grad = grad * g
        shape = ctx.shape  # map [3,H,W] back to the caller's layout
        if len(shape) == 4:
            grad = grad[None] if shape[1] == 3 else grad.permute(1, 2, 0)[None]
        elif shape[0] != 3:
            grad = grad.permute(1, 2, 0)
        return grad, None, None, None, None, None


def image_ssim_loss(image: Tensor, target: Tensor, lambda_image: float = 0.8, lambda_ssim: float = 0.2,
                    method: str = 'l1') -> Tensor:
    """`lambda_image * ImageLoss(method)(image, target) + lambda_ssim * SSIM_Loss()(image, target)` in one pass
    (the two `self.loss_funcs(...)` lines at sk_gs.py:1528-1529).  Differentiable w.r.t. `image`."""
    return _ImageSSIMLoss.apply(image, target, float(lambda_image), float(lambda_ssim), method, None)


class ImageLoss(nn.Module):
    """Drop-in for networks/losses/image_loss.py:8-36 (unmasked; the shipped configs use {method: l1})."""

    def __init__(self, method='mse', masked=False, **kwargs):
        super().__init__()
        if method not in _METHODS:
            raise ValueError(f"method={method} for {self.__class__.__name__} is not supported")
        if masked:
            raise NotImplementedError('masked image loss is outside the hot path (no shipped config enables it)')
        self.method, self.masked = method, masked

    def forward(self, pred_image: Tensor, gt_image: Tensor, mask: Tensor = None):
        return _ImageSSIMLoss.apply(pred_image[..., :3] if pred_image.shape[-1] == 4 else pred_image, gt_image, 1.0,
                                    0.0, self.method, 0)

    def __repr__(self):
        return f"{self.__class__.__name__}(method={self.method}, masked={self.masked})"


class SSIM_Loss(nn.Module):
    """Drop-in for networks/losses/ssim.py:21-41 (window 11, reduction 'mean')."""

    def __init__(self, window_size=11, reduction='mean', **kwargs):
        super().__init__()
        if window_size != 11 or reduction != 'mean':
            raise NotImplementedError('only window_size=11, reduction="mean" (the reference defaults) are built')
        self.window_size, self.reduction = window_size, reduction

    def forward(self, img1: Tensor, img2: Tensor):
        return _ImageSSIMLoss.apply(img1, img2, 0.0, 1.0, 'l1', 1)
