"""oracle/densify.py - CPU restatement (torch, fp32 like the reference) of the densification bookkeeping of
/root/reference/networks/gaussian_splatting.py:503-665.

TEST INFRASTRUCTURE ONLY: only tests/ import this module.  The product (sk_gs_b200/) never does.

Parity status: PINNED - tests/test_oracle_golden.py checks it against tests/golden/densify.npz, which
tests/golden/make_golden.py produced by running the reference's own add_densification_stats / densify / prune /
reset_opacity (with change_optimizer on a real torch.optim.Adam) unmodified.

The reference does clone -> split -> drop split originals -> prune as four rounds of cat / mask indexing; this
restatement follows the same rounds literally (it is the oracle, not the product) so that the one-gather CUDA plan can be
checked against both."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

Triple = Tuple[Tensor, Optional[Tensor], Optional[Tensor]]


def add_densification_stats(grad_accum: Tensor, denom: Tensor, max_radii2D: Tensor, radii: Tensor, vs_grad: Tensor):
    """:672-675 + :503-513.  In place."""
    mask = radii > 0
    max_radii2D[mask] = torch.max(max_radii2D[mask], radii[mask].float())
    grad_accum[mask] += torch.norm(vs_grad[mask, :2], dim=-1)
    denom[mask] += 1


def _quat_to_R(q: Tensor) -> Tensor:
    """my_ext/ops_3d/quaternion.py:162-172 (xyzw, normalised)."""
    x, y, z, w = torch.nn.functional.normalize(q, dim=-1).unbind(-1)
    return torch.stack([
        1 - 2 * y * y - 2 * z * z, 2 * x * y - 2 * w * z, 2 * w * y + 2 * x * z,
        2 * x * y + 2 * w * z, 1 - 2 * x * x - 2 * z * z, 2 * y * z - 2 * w * x,
        2 * x * z - 2 * w * y, 2 * w * x + 2 * y * z, 1 - 2 * x * x - 2 * y * y], dim=-1).reshape(-1, 3, 3)


def _cat(t: Dict[str, Triple], new: Dict[str, Tensor]) -> Dict[str, Triple]:
    """change_optimizer op='concat' (:537-538, :548-552)."""
    out = {}
    for n, (p, m, v) in t.items():
        a = new[n]
        out[n] = (torch.cat([p, a]), None if m is None else torch.cat([m, torch.zeros_like(a)]),
                  None if v is None else torch.cat([v, torch.zeros_like(a)]))
    return out


def _keep(t: Dict[str, Triple], keep: Tensor) -> Dict[str, Triple]:
    """change_optimizer op='prune' (:539-541, :553-555)."""
    return {n: (p[keep], None if m is None else m[keep], None if v is None else v[keep]) for n, (p, m, v) in t.items()}


def densify_and_prune(tensors: Dict[str, Triple], grad_accum: Tensor, denom: Tensor, max_radii2D: Tensor,
                      do_densify: bool, do_prune: bool, grad_threshold: float, densify_extent: float,
                      min_opacity: float, max_screen_size: float, prune_extent: float, noise: Tensor):
    """-> (tensors, grad_accum, denom, max_radii2D).  `tensors` needs 'xyz', 'scaling', 'rotation', 'opacity'."""
    t = dict(tensors)
    if do_densify:
        grads = grad_accum / denom  # :641-642
        grads[grads.isnan()] = 0.0
        # clone (:624-638)
        scale = torch.exp(t['scaling'][0])
        sel = (grads >= grad_threshold) & (scale.amax(dim=1) <= densify_extent)
        t = _cat(t, {n: p[sel] for n, (p, _, _) in t.items()})
        # split (:589-622)
        P1 = t['xyz'][0].shape[0]
        padded = torch.zeros(P1, device=grads.device)
        padded[:grads.shape[0]] = grads
        scale = torch.exp(t['scaling'][0])
        sel = (padded >= grad_threshold) & (scale.amax(dim=1) > densify_extent)
        ns = int(sel.sum())
        stds = scale[sel].repeat(2, 1)
        samples = noise[:2 * ns] * stds
        R = _quat_to_R(t['rotation'][0][sel]).repeat(2, 1, 1)
        new = {n: p[sel].repeat(2, *[1] * (p.ndim - 1)) for n, (p, _, _) in t.items()}
        new['xyz'] = torch.bmm(R, samples[..., None]).squeeze(-1) + t['xyz'][0][sel].repeat(2, 1)
        new['scaling'] = torch.log(scale[sel].repeat(2, 1) / (0.8 * 2))
        t = _cat(t, new)
        t = _keep(t, ~torch.cat([sel, sel.new_zeros(2 * ns)]))
        P2 = t['xyz'][0].shape[0]
        grad_accum, denom, max_radii2D = (torch.zeros(P2, device=grads.device) for _ in range(3))  # :583-586
    if do_prune:  # :653-660
        mask = torch.sigmoid(t['opacity'][0]).reshape(-1) < min_opacity
        if max_screen_size and max_screen_size > 0:
            mask = mask | (max_radii2D > max_screen_size) | (torch.exp(t['scaling'][0]).amax(dim=1) > prune_extent)
        t = _keep(t, ~mask)
        grad_accum, denom, max_radii2D = grad_accum[~mask], denom[~mask], max_radii2D[~mask]
    return t, grad_accum, denom, max_radii2D


def reset_opacity(opacity: Tensor, cap: float = 0.01) -> Tensor:
    """:662-665 (moments are zeroed by change_optimizer op='replace')."""
    o = torch.min(torch.sigmoid(opacity), torch.ones_like(opacity) * cap)
    return torch.log(o / (1 - o))
