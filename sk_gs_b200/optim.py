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
                  knn_indices: Optional[Sequence[Optional[Tensor]]] = None, dynamic_hyper: Optional[Tensor] = None,
                  skip_flag_ptr: Optional[int] = None):
    """In-place Adam update of `params` (autograd-free; CUDA-graph capturable: nothing is allocated).
    `knn_indices[i]` (int64 [rows, K]) marks tensor i as having a compact [rows, K] gradient (see include/skgs_b200.h).
    `lrs[i]` is a float, or `(lr, lr2, period, split)` for two interleaved param groups in one array (element e uses lr
    if e % period < split else lr2 - the merged SH array, see include/skgs_b200.h).
    `dynamic_hyper` (device float32 [1 + 2n], see `adam_hyper`) overrides step / lr inside a replayed CUDA graph.
    `skip_flag_ptr`: device address of a uint32; non-zero at execution time turns the step into a no-op (pass
    RasterState.overflow_ptr so that gradients of an overflowed fixed-capacity render never reach the parameters)."""
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
                                        float(grad_scale), _lib.ptr(dynamic_hyper), skip_flag_ptr, st),
                       'skgs_adam_step')


def adam_hyper(lrs: Sequence[float], step: int, beta1: float = 0.9, beta2: float = 0.999) -> List[float]:
    """Host-side values of the `dynamic_hyper` table for iteration `step` (1-based)."""
    bc1 = 1.0 - beta1 ** step
    out = [(1.0 - beta2 ** step) ** 0.5]
    for lr in lrs:
        lr, lr2 = (lr[0], lr[1]) if isinstance(lr, (tuple, list)) else (lr, lr)
        out += [lr / bc1, lr2 / bc1]
    return out


class Adam(torch.optim.Optimizer):
    """`torch.optim.Adam` with the update done by ONE launch of skgs_adam_step per step (no weight decay, amsgrad or
    maximize - the reference uses none of them).

    It IS a `torch.optim.Optimizer`: `param_groups`, `state` (per parameter 'step' tensor, 'exp_avg', 'exp_avg_sq' - the
    keys and types torch.optim.Adam uses, so checkpoints are interchangeable with it), `state_dict()` /
    `load_state_dict()`, `add_param_group()`, `zero_grad()`, lr schedulers and `GradScaler.step` all work, and so does the
    reference's `GaussianSplatting.change_optimizer` (networks/gaussian_splatting.py:515-563), which tests
    `isinstance(optimizers, torch.optim.Optimizer)` and edits `group['params'][0]` / `optimizer.state[...]` in place when
    densification clones, splits or prunes Gaussians."""

    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1):
            raise ValueError(f'invalid Adam hyper-parameters lr={lr} betas={betas} eps={eps}')
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps))

    @torch.no_grad()
    def step(self, closure=None, knn_grads: Optional[dict] = None):
        """One update of every parameter that has a gradient.  `knn_grads` optionally maps a parameter to
        `(grad [rows, K], indices int64 [rows, K])` - the compact gradient of the skinning table."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        knn_grads = knn_grads or {}
        buckets = {}
        for group in self.param_groups:
            key = (tuple(group['betas']), float(group['eps']))
            for p in group['params']:
                compact = knn_grads.get(p)
                if p.grad is None and compact is None:
                    continue
                st = self.state[p]
                if len(st) == 0:
                    st['step'] = torch.tensor(0.0, dtype=torch.float32)  # host tensor, like torch's default (capturable=False)
                    st['exp_avg'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st['step'] += 1
                g, idx = (p.grad.contiguous(), None) if compact is None else compact
                buckets.setdefault(key + (int(st['step']),), []).append((p, g, st, float(group['lr']), idx))
        for (betas, eps, step), items in buckets.items():
            adam_step_raw([i[0].data for i in items], [i[1] for i in items], [i[2]['exp_avg'] for i in items],
                          [i[2]['exp_avg_sq'] for i in items], [i[3] for i in items], step, betas[0], betas[1], eps,
                          knn_indices=[i[4] for i in items])
        return loss
