"""sp-stage LBS boundary: `sp_warp(...)` does what the reference's `calc_LBS_weight`
(/root/reference/networks/sk_gs.py:751-774) followed by `warp` (:776-828) do inside `sp_stage` (:830-856): K nearest
superpoints, skinning weights (4 modes), blend of the per-superpoint rigid transforms the deformation network predicts.
Same kernels as the skeleton stage (fk_lbs.py), the transform table comes from (sp_t, sp_r) instead of forward
kinematics.  Differentiable w.r.t. everything except `points` (detached at :834).

Deviations from the reference call: sp_t must be a tensor (the `isinstance(sp_t, SE3)` branch of `warp`
is a lietorch object, un-vendored); method 'largest' picks argmax_k w_k on every call (the reference caches `p2sp` from
the last training step, :850-851); hyper-feature KNN (:754-756) is not offered.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
from torch import Tensor

from . import _lib
from .diff_gaussian_rasterization import _f32c


def _opt(t):
    return None if t is None else _f32c(t)


class _SpWarp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, sp_points, sp_t, sp_r, sp_rot, sp_scale, sp_W, sp_radius, sp_weight, K, mode,
                temperature, method):
        if not points.is_cuda:
            raise RuntimeError('sp_warp needs CUDA tensors (sk_gs_b200 has no CPU path)')
        L = _lib.lib()
        device = points.device
        points, sp_points, sp_t, sp_r = _f32c(points.detach()), _f32c(sp_points), _f32c(sp_t), _f32c(sp_r)
        sp_rot, sp_scale, sp_W = _opt(sp_rot), _opt(sp_scale), _opt(sp_W)
        sp_radius, sp_weight = _opt(sp_radius), _opt(sp_weight)
        P, M = points.shape[0], sp_points.shape[0]
        if mode == 'W' and (sp_W is None or tuple(sp_W.shape) != (P, M)):
            raise RuntimeError(f'mode W needs sp_W of shape [{P}, {M}]')
        sp = _lib.Superpoints(M, int(K), _lib.LBS_MODES[mode], _lib.WARP_METHODS[method], float(temperature),
                              sp_points.data_ptr(), sp_t.data_ptr(), sp_r.data_ptr(), _lib.ptr(sp_rot),
                              _lib.ptr(sp_scale), _lib.ptr(sp_W), _lib.ptr(sp_radius), _lib.ptr(sp_weight))
        d_points = torch.empty(P, 3, device=device)
        d_rotation = torch.empty(P, 4, device=device)
        d_scales = torch.empty(P, 3, device=device)
        spT = torch.empty(M, 7, device=device)
        weights = torch.empty(P, K, device=device)
        indices = torch.empty(P, K, dtype=torch.int64, device=device)
        ws = torch.empty(L.skgs_sp_lbs_workspace_bytes(M), dtype=torch.uint8, device=device)
        with torch.cuda.device(device):
            st = torch.cuda.current_stream(device).cuda_stream
            _lib.check(L.skgs_sp_lbs_forward(C.byref(sp), P, points.data_ptr(), d_points.data_ptr(),
                                             d_rotation.data_ptr(), d_scales.data_ptr(), spT.data_ptr(),
                                             weights.data_ptr(), indices.data_ptr(), ws.data_ptr(), st),
                       'skgs_sp_lbs_forward')
        ctx.sp = sp
        ctx.keep = (points, sp_points, sp_t, sp_r, sp_rot, sp_scale, sp_W, sp_radius, sp_weight, spT, weights, indices)
        ctx.mode = mode
        ctx.mark_non_differentiable(indices)
        return d_points, d_rotation, d_scales, spT, weights, indices

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_dp, g_dr, g_ds, g_spT, g_w, _g_idx):
        return _sp_warp_backward_impl(ctx, g_dp, g_dr, g_ds, g_spT, g_w)


def _sp_warp_backward_impl(ctx, g_dp, g_dr, g_ds, g_spT, g_w, out=None):
    out = out or {}
    L = _lib.lib()
    points, sp_points, sp_t, sp_r, sp_rot, sp_scale, sp_W, sp_radius, sp_weight, spT, weights, indices = ctx.keep
    device = points.device
    P, M, K = points.shape[0], sp_points.shape[0], indices.shape[1]
    need = ctx.needs_input_grad

    def pick(name, *shape):
        t = out.get(name)
        return torch.empty(*shape, dtype=torch.float32, device=device) if t is None else t

    d_c, d_t, d_r = pick('sp_points', M, 3), pick('sp_t', M, 3), pick('sp_r', M, 4)
    d_rot = pick('sp_rot', M, 4) if sp_rot is not None else None
    d_scale = pick('sp_scale', M, 3) if sp_scale is not None else None
    compact = bool(getattr(ctx, 'compact_sp_W', False))
    d_W = pick('sp_W', P, M) if (ctx.mode == 'W' and need[6] and not compact) else None
    d_W_knn = pick('sp_W', P, K) if (ctx.mode == 'W' and need[6] and compact) else None
    d_radius = pick('sp_radius', M) if sp_radius is not None and ctx.mode in ('kernel', 'weighted_kernel') else None
    d_weight = pick('sp_weight', M) if sp_weight is not None and ctx.mode == 'weighted_kernel' else None
    ws = torch.empty(L.skgs_sp_lbs_workspace_bytes(M), dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        st = torch.cuda.current_stream(device).cuda_stream
        _lib.check(L.skgs_sp_lbs_backward(
            C.byref(ctx.sp), P, points.data_ptr(), spT.data_ptr(), weights.data_ptr(), indices.data_ptr(),
            _lib.ptr(_opt(g_dp)), _lib.ptr(_opt(g_dr)), _lib.ptr(_opt(g_ds)), _lib.ptr(_opt(g_spT)), _lib.ptr(_opt(g_w)),
            d_c.data_ptr(), d_t.data_ptr(), d_r.data_ptr(), _lib.ptr(d_rot), _lib.ptr(d_scale), _lib.ptr(d_W),
            _lib.ptr(d_W_knn), _lib.ptr(d_radius), _lib.ptr(d_weight), ws.data_ptr(), st), 'skgs_sp_lbs_backward')
    return (None, d_c, d_t, d_r, d_rot, d_scale, d_W if not compact else d_W_knn, d_radius, d_weight, None, None, None,
            None)


def sp_warp(points: Tensor, sp_points: Tensor, sp_t: Tensor, sp_r: Tensor, sp_rot: Optional[Tensor] = None,
            sp_scale: Optional[Tensor] = None, K: int = 5, mode: str = 'W', sp_W: Optional[Tensor] = None,
            sp_radius: Optional[Tensor] = None, sp_weight: Optional[Tensor] = None, temperature: float = 1.0,
            method: str = 'LBS'):
    """-> (d_points [P,3], d_rotation [P,4], d_scales [P,3] | None, spT [M,7], weights [P,K], indices [P,K]): the first
    four are `warp`'s return (networks/sk_gs.py:828), the last two `calc_LBS_weight`'s (:774)."""
    d_points, d_rotation, d_scales, spT, weights, indices = _SpWarp.apply(
        points, sp_points, sp_t, sp_r, sp_rot, sp_scale, sp_W, sp_radius, sp_weight, K, mode, temperature, method)
    return d_points, d_rotation, (d_scales if sp_scale is not None else None), spT, weights, indices


# --------------------------------------------------------------------------------------------- raw (non-autograd) calls
def sp_warp_forward_raw(points, sp_points, sp_t, sp_r, sp_rot=None, sp_scale=None, K=5, mode='W', sp_W=None,
                        sp_radius=None, sp_weight=None, temperature=1.0, method='LBS'):
    from .fk_lbs import _Ctx
    ctx = _Ctx([False, True, True, True, sp_rot is not None, sp_scale is not None, sp_W is not None,
                sp_radius is not None, sp_weight is not None] + [False] * 4)
    out = _SpWarp.forward(ctx, points, sp_points, sp_t, sp_r, sp_rot, sp_scale, sp_W, sp_radius, sp_weight, K, mode,
                          temperature, method)
    return out, ctx


def sp_warp_backward_raw(ctx, g_dp=None, g_dr=None, g_ds=None, g_spT=None, g_w=None, compact_sp_W=False, out=None):
    """-> (d_sp_points, d_sp_t, d_sp_r, d_sp_rot, d_sp_scale, d_sp_W, d_sp_radius, d_sp_weight); with compact_sp_W the
    sp_W gradient is [P, K] in KNN order (the dense [P, M] table has K non-zeros per row)."""
    ctx.compact_sp_W = compact_sp_W
    return _sp_warp_backward_impl(ctx, g_dp, g_dr, g_ds, g_spT, g_w, out)[1:9]
