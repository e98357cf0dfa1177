"""One-launch Adam: `Adam` takes the same param groups the reference builds
(/root/reference/networks/gaussian_splatting.py:445-453: one group per Gaussian attribute with its own 'lr' and 'name';
hyper-parameters exps/default.yaml:121-125) and exposes `param_groups` / `state` / `step()` / `zero_grad()` like
torch.optim.Adam, so the lr schedulers that write `group['lr']` (gaussian_splatting.py:466-471) keep working.
All arithmetic runs in libskgs_b200.so (csrc/adam.cu), one kernel per step for up to 16 tensors.

No CPU path: CPU parameters raise.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch
from torch import Tensor

from . import _lib


def adam_step_raw(params: Sequence[Tensor], grads: Sequence[Tensor], exp_avgs: Sequence[Tensor],
                  exp_avg_sqs: Sequence[Tensor], lrs: Sequence[float], step: int, beta1: float = 0.9,
                  beta2: float = 0.999, eps: float = 1e-15, grad_scale: float = 1.0,
                  knn_indices: Optional[Sequence[Optional[Tensor]]] = None, dynamic_hyper: Optional[Tensor] = None):
    """In-place Adam update of `params` (autograd-free; CUDA-graph capturable: nothing is allocated).
    `knn_indices[i]` (int64 [rows, K]) marks tensor i as having a compact [rows, K] gradient (see include/skgs_b200.h).
    `lrs[i]` is a float, or `(lr, lr2, period, split)` for two interleaved param groups in one array (element e uses lr
    if e % period < split else lr2 - the merged SH array, see include/skgs_b200.h).
    `dynamic_hyper` (device float32 [1 + 2n], see `adam_hyper`) overrides step / lr inside a replayed CUDA graph."""
    n = len(params)
    if not (len(grads) == len(exp_avgs) == len(exp_avg_sqs) == len(lrs) == n):
        raise RuntimeError('adam_step_raw: argument lists differ in length')
    if n == 0:
        return
    if dynamic_hyper is not None and (n > _lib.ADAM_MAX_TENSORS or dynamic_hyper.numel() != 1 + 2 * n or
                                      dynamic_hyper.dtype != torch.float32 or not dynamic_hyper.is_cuda):
        raise RuntimeError(f'dynamic_hyper must be a CUDA float32 tensor of {1 + 2 * n} elements (at most '
                           f'{_lib.ADAM_MAX_TENSORS} tensors per call)')
    for t in list(params) + list(grads) + list(exp_avgs) + list(exp_avg_sqs):
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError('adam_step_raw needs contiguous float32 CUDA tensors (no CPU path)')
    L = _lib.lib()
    dev = params[0].device
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        for lo in range(0, n, _lib.ADAM_MAX_TENSORS):
            hi = min(n, lo + _lib.ADAM_MAX_TENSORS)
            table = (_lib.AdamTensor * (hi - lo))()
            for j, i in enumerate(range(lo, hi)):
                p, g, m, v = params[i], grads[i], exp_avgs[i], exp_avg_sqs[i]
                idx = None if knn_indices is None else knn_indices[i]
                if m.shape != p.shape or v.shape != p.shape:
                    raise RuntimeError('adam_step_raw: moment shapes differ from the parameter')
                cols = K = 0
                if idx is None:
                    if g.numel() != p.numel():
                        raise RuntimeError(f'adam_step_raw: gradient {tuple(g.shape)} vs parameter {tuple(p.shape)}')
                else:
                    if p.ndim != 2 or idx.dtype != torch.int64 or not idx.is_contiguous() or idx.shape != g.shape or \
                            idx.shape[0] != p.shape[0]:
                        raise RuntimeError('adam_step_raw: compact gradient needs param [rows, cols], grad and int64 '
                                           'indices [rows, K]')
                    cols, K = int(p.shape[1]), int(idx.shape[1])
                lr, lr2, period, split = lrs[i] if isinstance(lrs[i], (tuple, list)) else (lrs[i], lrs[i], 0, 0)
                table[j] = _lib.AdamTensor(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(),
                                           float(lr), float(lr2), int(period), int(split), cols, K, _lib.ptr(idx))
            _lib.check(L.skgs_adam_step(table, hi - lo, int(step), float(beta1), float(beta2), float(eps),
                                        float(grad_scale), _lib.ptr(dynamic_hyper), st), 'skgs_adam_step')


def adam_hyper(lrs: Sequence[float], step: int, beta1: float = 0.9, beta2: float = 0.999) -> List[float]:
    """Host-side values of the `dynamic_hyper` table for iteration `step` (1-based)."""
    bc1 = 1.0 - beta1 ** step
    out = [(1.0 - beta2 ** step) ** 0.5]
    for lr in lrs:
        lr, lr2 = (lr[0], lr[1]) if isinstance(lr, (tuple, list)) else (lr, lr)
        out += [lr / bc1, lr2 / bc1]
    return out


class Adam:
    """torch.optim.Adam look-alike (no weight decay, amsgrad or maximize - the reference uses none of them)."""

    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        groups = list(params)
        if len(groups) == 0:
            raise ValueError('optimizer got an empty parameter list')
        if not isinstance(groups[0], dict):
            groups = [{'params': groups}]
        self.defaults = dict(lr=lr, betas=tuple(betas), eps=eps)
        self.param_groups: List[dict] = []
        self.state = {}
        for g in groups:
            self.add_param_group(g)

    def add_param_group(self, group: dict):
        group = dict(group)
        ps = group['params']
        group['params'] = [ps] if isinstance(ps, Tensor) else list(ps)
        for k, v in self.defaults.items():
            group.setdefault(k, v)
        self.param_groups.append(group)

    def zero_grad(self, set_to_none: bool = True):
        for group in self.param_groups:
            for p in group['params']:
                if p.grad is not None:
                    if set_to_none:
                        p.grad = None
                    else:
                        p.grad.zero_()

    @torch.no_grad()
    def step(self, knn_grads: Optional[dict] = None):
        """One update of every parameter that has a gradient.  `knn_grads` optionally maps a parameter to
        `(grad [rows, K], indices int64 [rows, K])` - the compact gradient of the skinning table."""
        knn_grads = knn_grads or {}
        buckets = {}
        for group in self.param_groups:
            key = (tuple(group['betas']), float(group['eps']))
            for p in group['params']:
                compact = knn_grads.get(p)
                if p.grad is None and compact is None:
                    continue
                st = self.state.get(p)
                if st is None:
                    st = self.state[p] = {'step': 0, 'exp_avg': torch.zeros_like(p, memory_format=torch.contiguous_format),
                                          'exp_avg_sq': torch.zeros_like(p, memory_format=torch.contiguous_format)}
                st['step'] += 1
                g, idx = (p.grad.contiguous(), None) if compact is None else compact
                buckets.setdefault(key + (st['step'],), []).append((p, g, st, float(group['lr']), idx))
        for (betas, eps, step), items in buckets.items():
            adam_step_raw([i[0].data for i in items], [i[1] for i in items], [i[2]['exp_avg'] for i in items],
                          [i[2]['exp_avg_sq'] for i in items], [i[3] for i in items], step, betas[0], betas[1], eps,
                          knn_indices=[i[4] for i in items])
