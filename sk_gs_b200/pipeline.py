"""The hot path end to end on one GPU: parameters -> FK + LBS -> assembly -> rasterize (-> backward).

Mirrors what `SkeletonGaussianSplatting.render` does per view in the `sk` stage
(/root/reference/networks/sk_gs.py:1206-1242 -> forward :1160-1204 -> sk_stage :1109-1150 -> render_gs_offical),
minus the joint MLP, the loss and the optimizer (out of scope for this path)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
from torch import Tensor

from .diff_gaussian_rasterization import GaussianRasterizationSettings
from .fk_lbs import assemble, fk_lbs
from .renderer import render_gs_offical
from .scene import Camera, Scene

PARAM_NAMES = ['xyz', 'scaling', 'rotation', 'opacity', 'f_dc', 'f_rest', 'sp_W', 'joints', 'sk_r', 'sk_d_rot',
               'sk_d_scale', 'g_tr']


def raster_settings_for(cam: Camera, device, sh_degree: int = 3, scale_modifier: float = 1.0):
    """Same construction as networks/gaussian_splatting.py:271-284."""
    return GaussianRasterizationSettings(
        image_height=cam.H, image_width=cam.W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=cam.bg.to(device),
        scale_modifier=scale_modifier, viewmatrix=cam.viewmatrix.to(device), projmatrix=cam.projmatrix.to(device),
        sh_degree=sh_degree, campos=cam.campos.to(device), prefiltered=False, debug=False)


class HotPath:
    """Holds the parameters of one synthetic scene on a device and runs the per-view step."""

    def __init__(self, scene: Scene, device='cuda', mode: str = 'W', requires_grad: bool = True):
        self.scene = scene
        self.device = torch.device(device)
        self.mode = mode
        self.K = scene.K
        self.params: Dict[str, Tensor] = {}
        for n in PARAM_NAMES:
            self.params[n] = getattr(scene, n).to(self.device).clone().requires_grad_(requires_grad)
        self.sp_radius = scene.sp_radius.to(self.device).clone().requires_grad_(requires_grad and mode != 'W')
        self.sp_weight = scene.sp_weight.to(self.device).clone().requires_grad_(requires_grad and mode != 'W')
        self.parents = scene.parents.to(self.device)
        self.root = scene.root
        self.settings = [raster_settings_for(c, self.device, scene.sh_degree) for c in scene.cameras]

    def deform(self):
        p = self.params
        out = fk_lbs(p['xyz'], p['joints'], p['sk_r'], p['sk_d_rot'], p['sk_d_scale'], p['g_tr'], self.parents,
                     self.root, K=self.K, mode=self.mode, sp_W=p['sp_W'] if self.mode == 'W' else None,
                     sp_radius=self.sp_radius if self.mode != 'W' else None,
                     sp_weight=self.sp_weight if self.mode == 'weighted_kernel' else None)
        d_xyz, d_rot, d_scale = out[:3]
        points, scales, rotations, opacity = assemble(p['xyz'], p['scaling'], p['rotation'], p['opacity'], d_xyz,
                                                      d_rot, d_scale)
        sh = torch.cat((p['f_dc'], p['f_rest']), dim=1)
        return dict(points=points, scales=scales, rotations=rotations, opacity=opacity, sh_features=sh), out

    def render(self, view: int = 0):
        net_out, sk_out = self.deform()
        out = render_gs_offical(raster_settings=self.settings[view], **net_out)
        out['_sk'] = sk_out
        return out

    def step(self, view: int = 0, dL_dimage: Optional[Tensor] = None):
        """forward + backward of one view; returns the rendered image.  Gradients land in .grad of the parameters."""
        out = self.render(view)
        img = out['images']
        if dL_dimage is None:
            img.sum().backward()
        else:
            img.backward(dL_dimage)
        return out

    def zero_grad(self):
        for t in list(self.params.values()) + [self.sp_radius, self.sp_weight]:
            t.grad = None

    def grads(self) -> Dict[str, Tensor]:
        return {n: t.grad for n, t in self.params.items() if t.grad is not None}
