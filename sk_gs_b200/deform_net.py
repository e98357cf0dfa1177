"""Joint-rotation network of the `sk` stage: `SimpleDeformationNetwork` mirrors the reference module of
the same name (/root/reference/networks/sk_gs.py:134-164: constructor arguments, `forward(points, t)` returning the list
of head outputs, state-dict key names of `dynamic_net.net.{i}` / `dynamic_net.last.{j}`), with every launch inside
libskgs_b200.so (csrc/joint_mlp.cu).  All parameters live in ONE flat tensor `theta` (a single Adam segment, a single
all-reduce block); the per-layer weights and biases are views of it.

No CPU path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor, nn

from . import _lib
from .diff_gaussian_rasterization import _f32c

HEADS = (4, 4, 3)  # sk_dims (sk_gs.py:519): rotation quaternion, d_rot, d_scale


class NetConfig:
    """Static description of one network (everything but M and theta of struct skgs_joint_mlp)."""

    def __init__(self, degree_p=10, degree_t=6, width=256, depth=8, skips=(4,), rotation_head=False):
        self.degree_p, self.degree_t, self.width, self.depth = int(degree_p), int(degree_t), int(width), int(depth)
        self.skips = tuple(int(s) for s in skips)
        if any(s < 0 or s >= depth for s in self.skips):
            raise ValueError(f'skips {self.skips} outside [0, depth)')
        self.skip_mask = sum(1 << s for s in set(self.skips))
        self.rotation_head = bool(rotation_head)
        desc = self.struct(0, None)
        n = self.depth + 1
        w, b, d, tot = (C.c_int64 * n)(), (C.c_int64 * n)(), (C.c_int32 * n)(), C.c_int64()
        _lib.check(_lib.lib().skgs_joint_mlp_layout(C.byref(desc), w, b, d, C.byref(tot)), 'skgs_joint_mlp_layout')
        self.weight_offsets, self.bias_offsets, self.in_dims = list(w), list(b), list(d)
        self.param_count = int(tot.value)
        self.enc = self.in_dims[0]

    def struct(self, M: int, theta: Optional[Tensor]):
        return _lib.JointMlp(int(M), self.degree_p, self.degree_t, self.width, self.depth, self.skip_mask, sum(HEADS),
                             int(self.rotation_head), _lib.ptr(theta))

    def out_dim(self, i: int) -> int:
        return self.width if i < self.depth else sum(HEADS)

    def views(self, theta: Tensor) -> List[Tuple[Tensor, Tensor]]:
        """[(weight [out, in], bias [out])] for the hidden layers, then the merged heads - views of `theta`."""
        out = []
        for i in range(self.depth + 1):
            o, k = self.out_dim(i), self.in_dims[i]
            w0, b0 = self.weight_offsets[i], self.bias_offsets[i]
            out.append((theta[w0:w0 + o * k].view(o, k), theta[b0:b0 + o]))
        return out


def joint_mlp_forward_raw(cfg: NetConfig, theta: Tensor, joints: Tensor, t_dev: Tensor, out: Optional[dict] = None):
    """Autograd-free forward.  `t_dev`: CUDA float32 tensor with one element.  Returns ((sk_r, d_rot, d_scale), ctx);
    `out` may carry preallocated 'sk_r', 'd_rot', 'd_scale', 'workspace' (CUDA-graph capture)."""
    for x in (theta, joints, t_dev):
        if not x.is_cuda:
            raise RuntimeError('joint_mlp needs CUDA tensors (sk_gs_b200 has no CPU path)')
    if theta.dtype != torch.float32 or not theta.is_contiguous() or theta.numel() != cfg.param_count:
        raise RuntimeError(f'theta must be a contiguous float32 tensor of {cfg.param_count} elements')
    if t_dev.dtype != torch.float32 or t_dev.numel() != 1:
        raise RuntimeError('t must be a float32 CUDA tensor with one element')
    joints = _f32c(joints)
    if joints.ndim != 2 or joints.shape[1] != 3:
        raise RuntimeError(f'joints must be [M, 3], got {tuple(joints.shape)}')
    L = _lib.lib()
    M, dev = joints.shape[0], joints.device
    out = {} if out is None else out
    desc = cfg.struct(M, theta)
    nbytes = L.skgs_joint_mlp_workspace_bytes(C.byref(desc))
    ws = out.get('workspace')
    if ws is None or ws.numel() < nbytes:
        ws = out['workspace'] = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    res = []
    for name, w in (('sk_r', 4), ('d_rot', 4), ('d_scale', 3)):
        t_ = out.get(name)
        if t_ is None:
            t_ = out[name] = torch.empty(M, w, device=dev)
        res.append(t_)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(L.skgs_joint_mlp_forward(C.byref(desc), joints.data_ptr(), t_dev.data_ptr(), res[0].data_ptr(),
                                            res[1].data_ptr(), res[2].data_ptr(), ws.data_ptr(), st),
                   'skgs_joint_mlp_forward')
    return tuple(res), (cfg, theta, joints, ws)


def joint_mlp_backward_raw(ctx, g_sk_r: Optional[Tensor], g_d_rot: Optional[Tensor], g_d_scale: Optional[Tensor],
                           out: Optional[dict] = None, need_joints: bool = True):
    """Backward of the forward that produced `ctx` (its workspace must be untouched).  Returns (dL/dtheta [param_count],
    dL/djoints [M,3] or None); `out` may carry preallocated 'theta' / 'joints'."""
    cfg, theta, joints, ws = ctx
    L = _lib.lib()
    M, dev = joints.shape[0], joints.device
    out = {} if out is None else out
    d_theta = out.get('theta')
    if d_theta is None:
        d_theta = torch.empty(cfg.param_count, device=dev)
    d_joints = None
    if need_joints:
        d_joints = out.get('joints')
        if d_joints is None:
            d_joints = torch.empty(M, 3, device=dev)
    gs = [None if g is None else _f32c(g) for g in (g_sk_r, g_d_rot, g_d_scale)]
    desc = cfg.struct(M, theta)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(L.skgs_joint_mlp_backward(C.byref(desc), _lib.ptr(gs[0]), _lib.ptr(gs[1]), _lib.ptr(gs[2]),
                                             d_theta.data_ptr(), _lib.ptr(d_joints), ws.data_ptr(), st),
                   'skgs_joint_mlp_backward')
    return d_theta, d_joints


class _JointMlp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, theta, joints, t_dev, cfg):
        (a, b, c), raw = joint_mlp_forward_raw(cfg, theta.detach(), joints.detach(), t_dev)
        ctx.raw = raw
        ctx.need_joints = joints.requires_grad
        return a, b, c

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, ga, gb, gc):
        d_theta, d_joints = joint_mlp_backward_raw(ctx.raw, ga, gb, gc, need_joints=ctx.need_joints)
        return d_theta, d_joints, None, None


class SimpleDeformationNetwork(nn.Module):
    """Drop-in for networks/sk_gs.py:134-164 as the `sk` stage configures it (sk_gs.py:520-522, exps/default.yaml:48-55):
    'freq' encoders, 3-d positions, scalar time, heads (4, 4, 3).  `rotation_head=True` additionally folds
    `normalize(out_r + (0,0,0,1))` (sk_gs.py:1075-1076) into the network, which is what HotPath uses."""

    def __init__(self, p_in_channels=3, t_in_channels=1, out_channels=HEADS, width=256, depth=8, skips=(4,),
                 pos_enc_p='freq', pos_enc_p_cfg: dict = None, pos_enc_t='freq', pos_enc_t_cfg: dict = None,
                 rotation_head=False):
        super().__init__()
        if p_in_channels != 3 or t_in_channels != 1 or tuple(out_channels) != HEADS:
            raise NotImplementedError('built for 3-d joints, scalar time and heads (4, 4, 3) (sk_feature_dim = 0, '
                                      'exps/default.yaml:47)')
        if pos_enc_p not in ('freq', 'frequency') or pos_enc_t not in ('freq', 'frequency'):
            raise NotImplementedError("only the 'freq' position encoders are built")
        self.cfg = NetConfig((pos_enc_p_cfg or {}).get('degree', 4), (pos_enc_t_cfg or {}).get('degree', 4), width, depth,
                             skips, rotation_head)
        self.theta = nn.Parameter(torch.empty(self.cfg.param_count))
        self.reset_parameters()

    def reset_parameters(self):
        """nn.Linear's default init for every layer (kaiming_uniform(a=sqrt 5) == U(+-1/sqrt(in)) for weight and bias)."""
        with torch.no_grad():
            for i, (w, b) in enumerate(self.cfg.views(self.theta)):
                bound = 1.0 / math.sqrt(self.cfg.in_dims[i])
                w.uniform_(-bound, bound)
                b.uniform_(-bound, bound)

    def reset_heads(self, std: float = 1e-6):
        """What SkeletonGaussianSplatting.reset_parameters does to the heads (sk_gs.py:542-545)."""
        with torch.no_grad():
            w, b = self.cfg.views(self.theta)[-1]
            w.normal_(0, std)
            b.zero_()

    # ---- interchange with the reference's checkpoint layout
    def reference_state_dict(self) -> Dict[str, Tensor]:
        sd = {}
        views = self.cfg.views(self.theta.detach())
        for i in range(self.cfg.depth):
            sd[f'dynamic_net.net.{i}.weight'], sd[f'dynamic_net.net.{i}.bias'] = views[i]
        w, b = views[-1]
        r = 0
        for j, h in enumerate(HEADS):
            sd[f'dynamic_net.last.{j}.weight'], sd[f'dynamic_net.last.{j}.bias'] = w[r:r + h], b[r:r + h]
            r += h
        return sd

    def load_reference_state_dict(self, sd: Dict[str, Tensor]):
        mine = self.reference_state_dict()
        missing = [k for k in mine if k not in sd]
        if missing:
            raise KeyError(f'missing keys {missing}')
        with torch.no_grad():
            for k, v in mine.items():
                if tuple(sd[k].shape) != tuple(v.shape):
                    raise RuntimeError(f'{k}: shape {tuple(sd[k].shape)} vs {tuple(v.shape)}')
                v.copy_(sd[k])

    def forward(self, points: Tensor, t: Tensor):
        """points [M,3] (the joints), t: tensor holding the time -> [out_r [M,4], d_rot [M,4], d_scale [M,3]]."""
        if not points.is_cuda or not self.theta.is_cuda:
            raise RuntimeError('SimpleDeformationNetwork needs CUDA tensors (sk_gs_b200 has no CPU path)')
        t_dev = t.detach().reshape(-1)[:1].to(device=points.device, dtype=torch.float32).contiguous()
        return list(_JointMlp.apply(self.theta, points, t_dev, self.cfg))
