"""Densification bookkeeping: the host-side mirror of the reference's adaptive control
(/root/reference/networks/gaussian_splatting.py:503-703) over libskgs_b200.so's densify kernels.

  DensifyStats            xyz_gradient_accum / denom / max_radii2D (:97-99, :496-501)
  add_densification_stats :503-513 + the max-radii update of adaptive_control (:669-675), one kernel, graph-capturable
  densify_and_prune       densify (:640-645: clone :624-638, split :589-622) and prune (:653-660) fused: one plan pass
                          + ONE gather over every per-Gaussian tensor and its Adam moments, instead of four rounds of
                          `change_optimizer` (:515-563) cat / mask-index copies
  reset_opacity           :662-665
  AdaptiveControl         the schedule of adaptive_control (:667-703, exps/default.yaml:65-74) driving a TrainLoop:
                          swaps the parameter / moment tensors and re-captures the CUDA graph when P changes

No CPU path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib

ROLES = {'copy': 0, 'xyz': 1, 'scaling': 2}


class DensifyStats:
    """Per-Gaussian statistics the densification decisions read (networks/gaussian_splatting.py:496-501)."""

    def __init__(self, P: int, device):
        self.grad_accum = torch.zeros(P, device=device)
        self.denom = torch.zeros(P, device=device)
        self.max_radii2D = torch.zeros(P, device=device)

    @property
    def P(self) -> int:
        return self.grad_accum.shape[0]


def _check(*ts):
    for t in ts:
        if t is not None and (not t.is_cuda or not t.is_contiguous()):
            raise RuntimeError('densify needs contiguous CUDA tensors (sk_gs_b200 has no CPU path)')


def add_densification_stats(stats: DensifyStats, radii: Tensor, viewspace_grad: Tensor,
                            skip_flag_ptr: Optional[int] = None):
    """One step's statistics on the current stream (capturable).  radii int32 [P] (multi-view: the MAX over views),
    viewspace_grad float32 [P, >=2] (multi-view: the SUM, :509-512)."""
    _check(radii, viewspace_grad)
    P = stats.P
    if radii.dtype != torch.int32 or radii.numel() != P or viewspace_grad.dtype != torch.float32 or \
            viewspace_grad.shape[0] != P or viewspace_grad.ndim != 2:
        raise RuntimeError('add_densification_stats: radii int32 [P], viewspace_grad float32 [P, >=2]')
    st = torch.cuda.current_stream(radii.device).cuda_stream
    _lib.check(_lib.lib().skgs_densify_stats(P, radii.data_ptr(), viewspace_grad.data_ptr(),
                                             int(viewspace_grad.shape[1]), stats.max_radii2D.data_ptr(),
                                             stats.grad_accum.data_ptr(), stats.denom.data_ptr(), skip_flag_ptr, st),
               'skgs_densify_stats')


@dataclass
class DensifyResult:
    tensors: Dict[str, Tuple[Tensor, Optional[Tensor], Optional[Tensor]]]  # name -> (param, exp_avg, exp_avg_sq)
    stats: DensifyStats
    counts: Dict[str, int]   # n_keep, n_clone, n_split (sources), n_selected, n_new
    src: Tensor              # int32 [n_new] source Gaussian of every new slot
    kind: Tensor             # uint8 [n_new] 0 kept, 1 clone, 2 / 3 split sample n = 0 / 1


def densify_and_prune(tensors: Dict[str, Tuple[Tensor, Optional[Tensor], Optional[Tensor]]], stats: DensifyStats,
                      do_densify: bool, do_prune: bool, grad_threshold: float = 0.0002, densify_extent: float = 0.01,
                      min_opacity: float = 0.005, max_screen_size: float = 0.0, prune_extent: float = 0.1,
                      noise: Optional[Tensor] = None, generator: Optional[torch.Generator] = None,
                      roles: Optional[Dict[str, str]] = None) -> DensifyResult:
    """`tensors`: every per-Gaussian tensor [P, ...] with its Adam moments (or None); must contain 'xyz', 'scaling',
    'rotation' (xyzw) and 'opacity'.  Returns NEW tensors [n_new, ...] in the reference's order (kept | clones | split
    samples n=0 | n=1); statistics are zeroed when densification ran (:583-586), gathered when only pruning did
    (:574-576).  `noise` ([2 n_selected, 3] standard normal) overrides the generator (tests)."""
    L = _lib.lib()
    names = list(tensors)
    for need in ('xyz', 'scaling', 'rotation', 'opacity'):
        if need not in tensors:
            raise RuntimeError(f'densify_and_prune: tensors must contain {need!r}')
    if len(names) > 16:
        raise RuntimeError('densify_and_prune: at most 16 tensors per call')
    xyz, scaling, rotation, opacity = (tensors[n][0] for n in ('xyz', 'scaling', 'rotation', 'opacity'))
    P, dev = xyz.shape[0], xyz.device
    for n in names:
        p, m, v = tensors[n]
        _check(p, m, v)
        if p.dtype != torch.float32 or p.shape[0] != P:
            raise RuntimeError(f'densify_and_prune: {n} must be float32 [P, ...]')
    roles = dict(roles or {})
    roles.setdefault('xyz', 'xyz')
    roles.setdefault('scaling', 'scaling')
    cfg = _lib.DensifyConfig(int(do_densify), int(do_prune), float(grad_threshold), float(densify_extent),
                             float(min_opacity), float(max_screen_size or 0.0), float(prune_extent))
    src = torch.empty(2 * P, dtype=torch.int32, device=dev)
    kind = torch.empty(2 * P, dtype=torch.uint8, device=dev)
    noise_row = torch.empty(2 * P, dtype=torch.int32, device=dev)
    counts_dev = torch.zeros(8, dtype=torch.int32, device=dev)
    ws = torch.empty(L.skgs_densify_workspace_bytes(P), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(L.skgs_densify_plan(C.byref(cfg), P, scaling.data_ptr(), opacity.data_ptr(),
                                       stats.grad_accum.data_ptr(), stats.denom.data_ptr(),
                                       stats.max_radii2D.data_ptr(), src.data_ptr(), kind.data_ptr(),
                                       noise_row.data_ptr(), counts_dev.data_ptr(), ws.data_ptr(), st),
                   'skgs_densify_plan')
        c = counts_dev.tolist()  # the one host synchronisation: the new arrays have to be allocated
        counts = dict(n_keep=c[0], n_clone=c[1], n_split=c[2], n_selected=c[3], n_new=c[4])
        n_new, n_sel = counts['n_new'], counts['n_selected']
        if noise is None:
            noise = torch.randn(max(2 * n_sel, 1), 3, device=dev, generator=generator)
        else:
            _check(noise)
            if noise.dtype != torch.float32 or noise.numel() < 6 * n_sel:
                raise RuntimeError(f'densify_and_prune: noise must be float32 [>= {2 * n_sel}, 3]')
        out: Dict[str, Tuple[Tensor, Optional[Tensor], Optional[Tensor]]] = {}
        table = (_lib.DensifyTensor * len(names))()
        for j, n in enumerate(names):
            p, m, v = tensors[n]
            width = p[0].numel() if P > 0 else 1
            shape = (n_new,) + tuple(p.shape[1:])
            po = torch.empty(shape, device=dev)
            mo = None if m is None else torch.empty(shape, device=dev)
            vo = None if v is None else torch.empty(shape, device=dev)
            out[n] = (po, mo, vo)
            table[j] = _lib.DensifyTensor(p.data_ptr(), po.data_ptr(), _lib.ptr(m), _lib.ptr(mo), _lib.ptr(v),
                                          _lib.ptr(vo), int(width), ROLES[roles.get(n, 'copy')])
        _lib.check(L.skgs_densify_apply(table, len(names), n_new, src.data_ptr(), kind.data_ptr(),
                                        noise_row.data_ptr(), scaling.data_ptr(), rotation.data_ptr(),
                                        noise.data_ptr(), st), 'skgs_densify_apply')
    new_stats = DensifyStats(n_new, dev)
    if not do_densify:  # prune only: the statistics of the survivors are kept (:574-576)
        idx = src[:n_new].long()
        new_stats.grad_accum = stats.grad_accum[idx]
        new_stats.denom = stats.denom[idx]
        new_stats.max_radii2D = stats.max_radii2D[idx]
    return DensifyResult(out, new_stats, counts, src[:n_new], kind[:n_new])


def reset_opacity(opacity: Tensor, exp_avg: Optional[Tensor] = None, exp_avg_sq: Optional[Tensor] = None,
                  cap: float = 0.01):
    """In place: opacity logit = logit(min(sigmoid(opacity), cap)), Adam moments zeroed (:662-665, :553-555)."""
    _check(opacity, exp_avg, exp_avg_sq)
    st = torch.cuda.current_stream(opacity.device).cuda_stream
    _lib.check(_lib.lib().skgs_opacity_reset(opacity.numel(), opacity.data_ptr(), _lib.ptr(exp_avg),
                                             _lib.ptr(exp_avg_sq), float(cap), st), 'skgs_opacity_reset')


def check_interval(step: int, interval: int, start: Optional[int] = None, end: Optional[int] = None) -> bool:
    """my_ext/utils/utils.py:126-146 `check_interval_v2(..., close='()')`: step % interval == 0 inside (start, end)."""
    if interval <= 0:
        return False
    if start is not None and start >= 0 and step <= start:
        return False
    if end is not None and end >= 0 and step >= end:
        return False
    return step % interval == 0


# adaptive_control_cfg of exps/default.yaml:65-74
DEFAULT_CONTROL = dict(opacity_reset_interval=(3000, 3000, -1), densify_interval=(100, 500, 25_000),
                       prune_interval=(100, 500, 25_000), densify_grad_threshold=0.0002, densify_percent_dense=0.01,
                       prune_opacity_threshold=0.005, prune_max_screen_size=20, prune_percent_dense=0.1)

PER_GAUSSIAN = ('xyz', 'shs', 'scaling', 'rotation', 'opacity', 'sp_W')


class AdaptiveControl:
    """`adaptive_control` (:667-703) for a `TrainLoop`: per-step statistics inside the captured graph, and - on the
    steps the schedule names - densify / prune / opacity reset followed by a re-capture of the graph for the new P."""

    def __init__(self, loop, cameras_extent: float, cfg: Optional[dict] = None, white_background: bool = False,
                 seed: int = 0):
        self.loop = loop
        self.cfg = dict(DEFAULT_CONTROL)
        self.cfg.update(cfg or {})
        self.extent = float(cameras_extent)
        self.white = white_background
        self.stats = DensifyStats(loop.hp.params['xyz'].shape[0], loop.hp.device)
        self.generator = torch.Generator(device=loop.hp.device).manual_seed(seed)
        self.history = []  # (step, n_before, n_after_densify, n_after_prune)
        loop.after_adam = self._record  # runs inside the captured graph, after the Adam kernel
        loop.extra_state = lambda: [self.stats.grad_accum, self.stats.denom, self.stats.max_radii2D]

    def _record(self, out, grads):
        st = out.get('_raster_state')
        add_densification_stats(self.stats, out['radii'], grads['viewspace_points'],
                                None if st is None else st.overflow_ptr)

    def _names(self):
        return [n for n in PER_GAUSSIAN if n in self.loop.names]

    def after_step(self, step: int) -> bool:
        """Call after iteration `step` (0-based, as the reference passes it); returns True when the Gaussian set (and
        with it the captured graph) changed."""
        cfg, loop = self.cfg, self.loop
        di, pi = cfg['densify_interval'], cfg['prune_interval']
        step = step + 1
        if step >= max(di[2], pi[2]):
            return False
        do_d, do_p = check_interval(step, *di), check_interval(step, *pi)
        changed = False
        if do_d or do_p:
            torch.cuda.synchronize(loop.hp.device)
            size_on = step > cfg['opacity_reset_interval'][0] and cfg['prune_max_screen_size'] > 0
            p = loop.hp.params
            tensors = {n: (p[n].data, loop.exp_avg[n], loop.exp_avg_sq[n]) for n in self._names()}
            n0 = self.stats.P
            res = densify_and_prune(
                tensors, self.stats, do_d, do_p, grad_threshold=cfg['densify_grad_threshold'],
                densify_extent=cfg['densify_percent_dense'] * self.extent,
                min_opacity=cfg['prune_opacity_threshold'],
                max_screen_size=cfg['prune_max_screen_size'] if size_on else 0.0,
                prune_extent=cfg['prune_percent_dense'] * self.extent, generator=self.generator)
            c = res.counts
            self.history.append((step, n0, n0 + c['n_clone'] + c['n_selected'] if do_d else n0, c['n_new']))
            if c['n_new'] != n0 or c['n_keep'] != n0:
                loop.replace_gaussians(res.tensors)
                self.stats = res.stats
                changed = True
            elif do_d:  # nothing selected: densification_postfix still zeroes the statistics (:583-586)
                for t in (self.stats.grad_accum, self.stats.denom, self.stats.max_radii2D):
                    t.zero_()
        oi = cfg['opacity_reset_interval']
        if check_interval(step, *oi) or (self.white and step == di[1]):
            reset_opacity(loop.hp.params['opacity'].data, loop.exp_avg['opacity'], loop.exp_avg_sq['opacity'])
        if changed:
            loop.recapture_after_resize()
        return changed
