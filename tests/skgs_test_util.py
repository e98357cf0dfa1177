"""Shared helpers of the test-suite: scene -> oracle inputs / product inputs."""
from __future__ import annotations

import numpy as np
import torch

from oracle import fk_lbs as OF
from oracle import raster as OR
from sk_gs_b200 import scene as S


def oracle_deform(sc: S.Scene, mode='W', dtype=torch.float32, requires_grad=False):
    """FK + LBS + assembly with the torch oracle.  Returns (dict of rasterizer inputs, sk_stage tuple, leaf params)."""
    names = ['xyz', 'scaling', 'rotation', 'opacity', 'f_dc', 'f_rest', 'sp_W', 'joints', 'sk_r', 'sk_d_rot',
             'sk_d_scale', 'g_tr', 'sp_radius', 'sp_weight']
    p = {n: getattr(sc, n).to(dtype).clone().requires_grad_(requires_grad) for n in names}
    out = OF.sk_stage(p['xyz'], p['joints'], p['sk_r'], p['sk_d_rot'], p['sk_d_scale'], p['g_tr'], sc.parents.long(),
                      sc.root, K=sc.K, mode=mode, sp_W=p['sp_W'], sp_radius=p['sp_radius'], sp_weight=p['sp_weight'])
    pts, scl, rot, op, sh = OF.assemble(p['xyz'], p['scaling'], p['rotation'], p['opacity'], p['f_dc'], p['f_rest'],
                                        *out[:3])
    return dict(points=pts, scales=scl, rotations=rot, opacity=op, sh_features=sh), out, p


def oracle_settings(cam: S.Camera, sh_degree=3, quat_wxyz=False, scale_modifier=1.0) -> OR.Settings:
    return OR.Settings(cam.H, cam.W, cam.tanfovx, cam.tanfovy, cam.bg.numpy(), scale_modifier,
                       cam.viewmatrix.numpy(), cam.projmatrix.numpy(), sh_degree, cam.campos.numpy(),
                       quat_wxyz=quat_wxyz)


def np32(t):
    return t.detach().cpu().float().contiguous().numpy()


def arena_view(buf: torch.Tensor, offset: int, dtype, count: int):
    """View `count` elements of `dtype` at byte `offset` of a uint8 arena tensor."""
    nbytes = count * torch.empty((), dtype=dtype).element_size()
    return buf[offset:offset + nbytes].view(dtype)


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    denom = max(np.abs(b).max(), 1e-30)
    return float(np.abs(a - b).max() / denom)
