"""FK + LBS boundary: `fk_lbs(...)` has the return contract of the reference's
`SkeletonGaussianSplatting.sk_stage` (/root/reference/networks/sk_gs.py:1109-1150, minus the joint MLP) and is
differentiable w.r.t. everything except `xyz` (the reference detaches it, :1113).  `assemble(...)` is the output
assembly of `forward` (:1192,1202-1203).  All arithmetic runs in libskgs_b200.so on torch's current stream.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
from torch import Tensor

from . import _lib
from .diff_gaussian_rasterization import _f32c


def _skeleton(joints, sk_r, sk_d_rot, sk_d_scale, g_tr, parents, root, K, mode, sp_W, sp_radius, sp_weight,
              temperature, sk_r_delta):
    M = joints.shape[0]
    L = 0 if parents is None or parents.numel() == 0 else int(parents.shape[1])
    sk = _lib.Skeleton(M, L, int(root), int(K), _lib.LBS_MODES[mode], float(temperature), joints.data_ptr(),
                       sk_r.data_ptr(), _lib.ptr(sk_r_delta), 0 if sk_r_delta is None else int(sk_r_delta.shape[-1]),
                       sk_d_rot.data_ptr(), sk_d_scale.data_ptr(), _lib.ptr(g_tr),
                       None if L == 0 else parents.data_ptr(), _lib.ptr(sp_W), _lib.ptr(sp_radius),
                       _lib.ptr(sp_weight))
    return sk


def global_transform(g_tr: Optional[Tensor]) -> Optional[Tensor]:
    """The global transform in the 7-vector form (t, q xyzw) the kernels take, from any of the forms `kinematic`
    accepts (/root/reference/networks/sk_gs.py:1092-1103): None, a 7-vector, a 6-vector twist (tau, phi) mapped by the
    SE3 exponential (lietorch SE3::Exp = (SO3::Exp(phi), V(phi) tau) with the left Jacobian V,
    my_ext/_C/include/lie.h:323-331,161-176,142-159) or a [4, 4] matrix (ops_3d.rigid.Rt_to_quaternion,
    my_ext/ops_3d/rigid.py:196-205).  A handful of torch operations on seven numbers, differentiable; lietorch `SE3`
    objects are not accepted (pass `.vec()`)."""
    if g_tr is None:
        return None
    if g_tr.shape[-2:] == (4, 4):
        Rt = g_tr.reshape(4, 4)
        w = 0.5 * torch.sqrt((Rt[0, 0] + Rt[1, 1] + Rt[2, 2] + 1).clamp_min(1e-10))
        w_ = 0.25 / w
        q = torch.stack([(Rt[2, 1] - Rt[1, 2]) * w_, (Rt[0, 2] - Rt[2, 0]) * w_, (Rt[1, 0] - Rt[0, 1]) * w_, w])
        return torch.cat([Rt[:3, 3], torch.nn.functional.normalize(q, dim=-1)])
    v = g_tr.reshape(-1)
    if v.numel() == 7:
        return v
    if v.numel() != 6:
        raise ValueError(f'g_tr got shape {tuple(g_tr.shape)}')  # same message as sk_gs.py:1103
    tau, phi = v[:3], v[3:]
    theta2 = (phi * phi).sum()
    theta = torch.sqrt(theta2)
    small = bool(theta < 1e-6)  # lie.h EPS: Taylor branches
    if small:
        imag, real = 0.5 - theta2 / 48.0 + theta2 * theta2 / 3840.0, 1.0 - theta2 / 8.0 + theta2 * theta2 / 384.0
        c1, c2 = 0.5 - theta2 / 24.0, 1.0 / 6.0 - theta2 / 120.0
    else:
        imag, real = torch.sin(0.5 * theta) / theta, torch.cos(0.5 * theta)
        c1, c2 = (1.0 - torch.cos(theta)) / theta2, (theta - torch.sin(theta)) / (theta2 * theta)
    q = torch.nn.functional.normalize(torch.cat([imag * phi, real.reshape(1) if torch.is_tensor(real) else
                                                 phi.new_tensor([real])]), dim=-1)
    pt = torch.linalg.cross(phi, tau)
    t = tau + c1 * pt + c2 * torch.linalg.cross(phi, pt)  # (I + c1 [phi]x + c2 [phi]x^2) tau
    return torch.cat([t, q])


class _FkLbs(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, joints, sk_r, sk_d_rot, sk_d_scale, g_tr, sp_W, sp_radius, sp_weight, parents, root, K,
                mode, temperature, sk_r_delta):
        if not xyz.is_cuda:
            raise RuntimeError('fk_lbs needs CUDA tensors (sk_gs_b200 has no CPU path)')
        L = _lib.lib()
        device = xyz.device
        xyz, joints, sk_r = _f32c(xyz.detach()), _f32c(joints), _f32c(sk_r)
        sk_d_rot, sk_d_scale = _f32c(sk_d_rot), _f32c(sk_d_scale)
        g_tr = None if g_tr is None else _f32c(g_tr.reshape(-1))
        sp_W = None if sp_W is None else _f32c(sp_W)
        sp_radius = None if sp_radius is None else _f32c(sp_radius)
        sp_weight = None if sp_weight is None else _f32c(sp_weight)
        sk_r_delta = None if sk_r_delta is None else _f32c(sk_r_delta)
        parents = None if parents is None else parents.to(device=device, dtype=torch.int32).contiguous()
        P, M = xyz.shape[0], joints.shape[0]
        if mode == 'W' and (sp_W is None or tuple(sp_W.shape) != (P, M)):
            raise RuntimeError(f'mode W needs sp_W of shape [{P}, {M}]')
        sk = _skeleton(joints, sk_r, sk_d_rot, sk_d_scale, g_tr, parents, root, K, mode, sp_W, sp_radius, sp_weight,
                       temperature, sk_r_delta)
        d_xyz = torch.empty(P, 3, device=device)
        d_rot = torch.empty(P, 4, device=device)
        d_scale = torch.empty(P, 3, device=device)
        sk_T = torch.empty(M, 7, device=device)
        weights = torch.empty(P, K, device=device)
        indices = torch.empty(P, K, dtype=torch.int64, device=device)
        ws = torch.empty(L.skgs_fk_lbs_workspace_bytes(M), dtype=torch.uint8, device=device)
        with torch.cuda.device(device):
            st = torch.cuda.current_stream(device).cuda_stream
            _lib.check(L.skgs_fk_lbs_forward(C.byref(sk), P, xyz.data_ptr(), d_xyz.data_ptr(), d_rot.data_ptr(),
                                             d_scale.data_ptr(), sk_T.data_ptr(), weights.data_ptr(),
                                             indices.data_ptr(), ws.data_ptr(), st), 'skgs_fk_lbs_forward')
        ctx.sk = sk
        ctx.keep = (xyz, joints, sk_r, sk_d_rot, sk_d_scale, g_tr, sp_W, sp_radius, sp_weight, parents, sk_r_delta,
                    sk_T, weights, indices)
        ctx.mode = mode
        ctx.mark_non_differentiable(indices)
        return d_xyz, d_rot, d_scale, sk_T, weights, indices

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_dxyz, g_drot, g_dscale, g_skT, g_w, _g_idx):
        return _fk_lbs_backward_impl(ctx, g_dxyz, g_drot, g_dscale, g_skT, g_w)


def _fk_lbs_backward_impl(ctx, g_dxyz, g_drot, g_dscale, g_skT, g_w, out=None):
    if True:
        out = out or {}
        L = _lib.lib()
        (xyz, joints, sk_r, sk_d_rot, sk_d_scale, g_tr, sp_W, sp_radius, sp_weight, parents, sk_r_delta, sk_T, weights,
         indices) = ctx.keep
        device = xyz.device
        P, M = xyz.shape[0], joints.shape[0]
        need = ctx.needs_input_grad

        def new(*shape):
            return torch.empty(*shape, dtype=torch.float32, device=device)

        def pick(name, *shape):
            t = out.get(name)
            return new(*shape) if t is None else t

        d_joints, d_sk_r = pick('joints', M, 3), pick('sk_r', M, 4)
        d_sk_d_rot, d_sk_d_scale = pick('sk_d_rot', M, 4), pick('sk_d_scale', M, 3)
        d_g_tr = None if g_tr is None else pick('g_tr', 7)
        compact = bool(getattr(ctx, 'compact_sp_W', False))
        d_sp_W = pick('sp_W', P, M) if (ctx.mode == 'W' and need[6] and not compact) else None
        d_sp_W_knn = pick('sp_W', P, indices.shape[1]) if (ctx.mode == 'W' and need[6] and compact) else None
        d_sp_radius = new(M) if sp_radius is not None and ctx.mode in ('kernel', 'weighted_kernel') else None
        d_sp_weight = new(M) if sp_weight is not None and ctx.mode == 'weighted_kernel' else None
        ws = torch.empty(L.skgs_fk_lbs_workspace_bytes(M), dtype=torch.uint8, device=device)
        with torch.cuda.device(device):
            st = torch.cuda.current_stream(device).cuda_stream
            _lib.check(L.skgs_fk_lbs_backward(
                C.byref(ctx.sk), P, xyz.data_ptr(), sk_T.data_ptr(), weights.data_ptr(), indices.data_ptr(),
                _lib.ptr(None if g_dxyz is None else _f32c(g_dxyz)), _lib.ptr(None if g_drot is None else _f32c(g_drot)),
                _lib.ptr(None if g_dscale is None else _f32c(g_dscale)),
                _lib.ptr(None if g_skT is None else _f32c(g_skT)), _lib.ptr(None if g_w is None else _f32c(g_w)),
                d_joints.data_ptr(), d_sk_r.data_ptr(), d_sk_d_rot.data_ptr(), d_sk_d_scale.data_ptr(),
                _lib.ptr(d_g_tr), _lib.ptr(d_sp_W), _lib.ptr(d_sp_W_knn), _lib.ptr(d_sp_radius), _lib.ptr(d_sp_weight),
                ws.data_ptr(), st),
                'skgs_fk_lbs_backward')
        return (None, d_joints, d_sk_r, d_sk_d_rot, d_sk_d_scale, d_g_tr, d_sp_W if not compact else d_sp_W_knn,
                d_sp_radius, d_sp_weight, None, None, None, None, None, None)


def fk_lbs(xyz: Tensor, joints: Tensor, sk_r: Tensor, sk_d_rot: Tensor, sk_d_scale: Tensor, g_tr: Optional[Tensor],
           parents: Tensor, root: int, K: int = 5, mode: str = 'W', sp_W: Optional[Tensor] = None,
           sp_radius: Optional[Tensor] = None, sp_weight: Optional[Tensor] = None, temperature: float = 1.0,
           sk_r_delta: Optional[Tensor] = None):
    """Returns the 9-tuple of `sk_stage` (networks/sk_gs.py:1150):
    (d_xyz[P,3], d_rot[P,4], d_scale[P,3], sk_T[M,7], sk_d_rot[M,4], sk_d_scale[M,3], g_tr[7], weights[P,K], indices[P,K])."""
    g_tr = global_transform(g_tr)  # 6-vector twist / 4x4 matrix -> (t, q), as `kinematic` does (:1092-1103)
    d_xyz, d_rot, d_scale, sk_T, weights, indices = _FkLbs.apply(
        xyz, joints, sk_r, sk_d_rot, sk_d_scale, g_tr, sp_W, sp_radius, sp_weight, parents, root, K, mode, temperature,
        sk_r_delta)
    return d_xyz, d_rot, d_scale, sk_T, sk_d_rot, sk_d_scale, g_tr, weights, indices


class _Assemble(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, scaling, rotation, opacity, d_xyz, d_rot, d_scale):
        L = _lib.lib()
        device = xyz.device
        xyz, scaling, rotation, opacity = _f32c(xyz), _f32c(scaling), _f32c(rotation), _f32c(opacity)
        d_xyz = None if d_xyz is None else _f32c(d_xyz)
        d_rot = None if d_rot is None else _f32c(d_rot)
        d_scale = None if d_scale is None else _f32c(d_scale)
        P = xyz.shape[0]
        points, scales = torch.empty_like(xyz), torch.empty_like(scaling)
        rotations, opacities = torch.empty_like(rotation), torch.empty_like(opacity)
        with torch.cuda.device(device):
            st = torch.cuda.current_stream(device).cuda_stream
            _lib.check(L.skgs_assemble_forward(P, xyz.data_ptr(), scaling.data_ptr(), rotation.data_ptr(),
                                               opacity.data_ptr(), _lib.ptr(d_xyz), _lib.ptr(d_rot), _lib.ptr(d_scale),
                                               points.data_ptr(), scales.data_ptr(), rotations.data_ptr(),
                                               opacities.data_ptr(), st), 'skgs_assemble_forward')
        ctx.keep = (scaling, rotation, opacity, d_rot)
        ctx.has = (d_xyz is not None, d_rot is not None, d_scale is not None)
        return points, scales, rotations, opacities

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gp, gs, gr, go):
        return _assemble_backward_impl(ctx, gp, gs, gr, go)


def _assemble_backward_impl(ctx, gp, gs, gr, go, out=None):
    if True:
        out = out or {}
        L = _lib.lib()
        scaling, rotation, opacity, d_rot = ctx.keep
        device = scaling.device
        P = scaling.shape[0]
        need = ctx.needs_input_grad

        def new(like, cond, name=None):
            if not cond:
                return None
            t = out.get(name) if name else None
            return torch.empty_like(like) if t is None else t

        dxyz, dscaling = new(scaling, need[0], 'xyz'), new(scaling, need[1], 'scaling')
        drotation, dopacity = new(rotation, need[2], 'rotation'), new(opacity, need[3], 'opacity')
        dd_xyz = new(scaling, ctx.has[0] and need[4])
        dd_rot = new(rotation, ctx.has[1] and need[5])
        dd_scale = new(scaling, ctx.has[2] and need[6])
        with torch.cuda.device(device):
            st = torch.cuda.current_stream(device).cuda_stream
            _lib.check(L.skgs_assemble_backward(
                P, scaling.data_ptr(), rotation.data_ptr(), opacity.data_ptr(), _lib.ptr(d_rot),
                _lib.ptr(None if gp is None else _f32c(gp)), _lib.ptr(None if gs is None else _f32c(gs)),
                _lib.ptr(None if gr is None else _f32c(gr)), _lib.ptr(None if go is None else _f32c(go)),
                _lib.ptr(dxyz), _lib.ptr(dscaling), _lib.ptr(drotation), _lib.ptr(dopacity), _lib.ptr(dd_xyz),
                _lib.ptr(dd_rot), _lib.ptr(dd_scale), st), 'skgs_assemble_backward')
        return dxyz, dscaling, drotation, dopacity, dd_xyz, dd_rot, dd_scale


def assemble(_xyz, _scaling, _rotation, _opacity, d_xyz=None, d_rot=None, d_scale=None):
    """points = _xyz + d_xyz, scales = exp(_scaling) + d_scale, rotations = normalize(_rotation + d_rot),
    opacity = sigmoid(_opacity)   (networks/sk_gs.py:1162-1163,1192,1202-1203; gaussian_splatting.py:155-160)."""
    return _Assemble.apply(_xyz, _scaling, _rotation, _opacity, d_xyz, d_rot, d_scale)


# --------------------------------------------------------------------------------------------- raw (non-autograd) calls
class _Ctx:
    """Minimal stand-in for an autograd ctx so that the Function bodies can be driven by hand (HotPath.step_manual:
    no autograd engine, no AccumulateGrad nodes -> CUDA-graph capturable and cheaper on the host)."""

    def __init__(self, needs):
        self.needs_input_grad = needs

    def mark_non_differentiable(self, *a):
        pass


def fk_lbs_forward_raw(xyz, joints, sk_r, sk_d_rot, sk_d_scale, g_tr, parents, root, K=5, mode='W', sp_W=None,
                       sp_radius=None, sp_weight=None, temperature=1.0, sk_r_delta=None):
    ctx = _Ctx([False, True, True, True, True, g_tr is not None, sp_W is not None, sp_radius is not None,
                sp_weight is not None] + [False] * 6)
    out = _FkLbs.forward(ctx, xyz, joints, sk_r, sk_d_rot, sk_d_scale, g_tr, sp_W, sp_radius, sp_weight, parents, root,
                         K, mode, temperature, sk_r_delta)
    return out, ctx


def fk_lbs_backward_raw(ctx, g_dxyz=None, g_drot=None, g_dscale=None, g_skT=None, g_w=None, compact_sp_W=False,
                        out=None):
    """-> (d_joints, d_sk_r, d_sk_d_rot, d_sk_d_scale, d_g_tr, d_sp_W, d_sp_radius, d_sp_weight); with compact_sp_W the
    sp_W gradient is returned as [P, K] (values in KNN order, see `scatter_sp_W_grad`) instead of dense [P, M]."""
    ctx.compact_sp_W = compact_sp_W
    r = _fk_lbs_backward_impl(ctx, g_dxyz, g_drot, g_dscale, g_skT, g_w, out)
    return r[1:9]


def assemble_forward_raw(xyz, scaling, rotation, opacity, d_xyz=None, d_rot=None, d_scale=None):
    ctx = _Ctx([True] * 7)
    return _Assemble.forward(ctx, xyz, scaling, rotation, opacity, d_xyz, d_rot, d_scale), ctx


def assemble_backward_raw(ctx, gp, gs, gr, go, out=None, need=None):
    """-> (dxyz, dscaling, drotation, dopacity, dd_xyz, dd_rot, dd_scale); `need` overrides which of the 7 are produced
    (dL/d_xyz equals dL/dpoints, callers that already hold the latter can skip it)."""
    if need is not None:
        ctx.needs_input_grad = list(need)
    return _assemble_backward_impl(ctx, gp, gs, gr, go, out)


def scatter_sp_W_grad(d_knn: Tensor, indices: Tensor, M: int) -> Tensor:
    """Expand the compact [P, K] logit gradients to the dense [P, M] gradient of sp_W."""
    out = torch.zeros(d_knn.shape[0], M, dtype=d_knn.dtype, device=d_knn.device)
    return out.scatter_(1, indices, d_knn)
