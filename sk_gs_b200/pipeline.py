"""The hot path end to end on one GPU: parameters -> FK + LBS -> assembly -> rasterize (-> backward).

Mirrors what `SkeletonGaussianSplatting.render` does per view in the `sk` stage
(/root/reference/networks/sk_gs.py:1206-1242 -> forward :1160-1204 -> sk_stage :1109-1150 -> render_gs_offical),
with the joint rotations as leaf parameters by default.  The step before the path (joint-rotation network, SURVEY.md 8f-1,
`joint_mlp=True`) and the two steps after it - photometric loss (8f-2, networks/sk_gs.py:1524-1529) and Adam (8f-3,
networks/gaussian_splatting.py:445-453) - are optional stages of `step_grads` / `sk_gs_b200.train.TrainLoop`; the
benchmarked metric (BASELINE.json) excludes them."""
from __future__ import annotations

from typing import Dict, Optional

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

    def __init__(self, scene: Scene, device='cuda', mode: str = 'W', requires_grad: bool = True,
                 merged_sh: bool = False, joint_mlp: bool = False, mlp_seed: int = 0, head_std: float = 1e-6):
        """`merged_sh`: keep the SH coefficients as ONE [P,16,3] parameter ('shs'; 'f_dc' / 'f_rest' become views of it)
        instead of the reference's two Parameters that are concatenated on every step (gaussian_splatting.py:155-157).
        `joint_mlp`: the joint rotations come from the joint-rotation network (sk_gs.py:1074-1076) evaluated
        at time `self.t` instead of being leaf parameters; its flat parameter vector is `params['theta']`."""
        self.scene = scene
        self.device = torch.device(device)
        self.mode = mode
        self.K = scene.K
        self.params: Dict[str, Tensor] = {}
        for n in PARAM_NAMES:
            self.params[n] = getattr(scene, n).to(self.device).clone().requires_grad_(requires_grad)
        if merged_sh:
            shs = torch.cat((self.params['f_dc'].detach(), self.params['f_rest'].detach()), dim=1)
            self.params['shs'] = shs.requires_grad_(requires_grad)
            self.params['f_dc'], self.params['f_rest'] = shs.detach()[:, :1], shs.detach()[:, 1:]
        self.sp_radius = scene.sp_radius.to(self.device).clone().requires_grad_(requires_grad and mode != 'W')
        self.sp_weight = scene.sp_weight.to(self.device).clone().requires_grad_(requires_grad and mode != 'W')
        self.mlp = None
        if joint_mlp:
            from .deform_net import SimpleDeformationNetwork
            gen_state = torch.random.get_rng_state()
            torch.manual_seed(mlp_seed)
            self.mlp = SimpleDeformationNetwork(pos_enc_p_cfg=dict(degree=10), pos_enc_t_cfg=dict(degree=6), width=256,
                                                depth=8, skips=(4,), rotation_head=True)  # exps/default.yaml:48-55
            self.mlp.reset_heads(head_std)
            torch.random.set_rng_state(gen_state)
            self.mlp.to(self.device)
            self.mlp.theta.requires_grad_(requires_grad)
            self.params['theta'] = self.mlp.theta
            for n in ('sk_r', 'sk_d_rot', 'sk_d_scale'):  # produced by the network now
                del self.params[n]
            self.t = torch.tensor([0.37], device=self.device)
        self.parents = scene.parents.to(self.device)
        self.root = scene.root
        self.settings = [raster_settings_for(c, self.device, scene.sh_degree) for c in scene.cameras]

    def deform(self):
        p = self.params
        if self.mlp is not None:
            sk_r, sk_d_rot, sk_d_scale = self.mlp(p['joints'], self.t)
        else:
            sk_r, sk_d_rot, sk_d_scale = p['sk_r'], p['sk_d_rot'], p['sk_d_scale']
        out = fk_lbs(p['xyz'], p['joints'], sk_r, sk_d_rot, sk_d_scale, p['g_tr'], self.parents,
                     self.root, K=self.K, mode=self.mode, sp_W=p['sp_W'] if self.mode == 'W' else None,
                     sp_radius=self.sp_radius if self.mode != 'W' else None,
                     sp_weight=self.sp_weight if self.mode == 'weighted_kernel' else None)
        d_xyz, d_rot, d_scale = out[:3]
        points, scales, rotations, opacity = assemble(p['xyz'], p['scaling'], p['rotation'], p['opacity'], d_xyz,
                                                      d_rot, d_scale)
        sh = p['shs'] if 'shs' in p else torch.cat((p['f_dc'], p['f_rest']), dim=1)
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

    # ------------------------------------------------------------------------------------------------ CUDA graph
    def step_grads(self, view: int, dL_dimage: Optional[Tensor], compact_sp_W: bool = False, before_backward=None,
                   arena=None, after_forward=None, mid_backward=None, target: Optional[Tensor] = None,
                   loss: Optional[dict] = None, fixed_capacity: Optional[int] = None,
                   header_words: Optional[Tensor] = None):
        """forward + backward of one view with the operators driven by hand (no autograd engine, no AccumulateGrad
        nodes): FK+LBS -> assembly -> rasterize, then the three backward calls in reverse.  Returns
        (outputs, {parameter name: gradient}); `.grad` is not touched.  Same kernels and same results as step(); this
        form is cheaper on the host and can be captured into a CUDA graph.
        With `target` (and `dL_dimage=None`) the upstream gradient comes from the fused L1 + SSIM loss between the rendered
        image and `target` (`loss` = keyword arguments of losses.image_loss_raw); outputs gain 'loss_terms'."""
        from . import diff_gaussian_rasterization as DGR
        from .fk_lbs import (assemble_backward_raw, assemble_forward_raw, fk_lbs_backward_raw, fk_lbs_forward_raw)
        with torch.no_grad():
            p = self.params
            W = self.mode == 'W'
            cm = None
            if self.mlp is not None:
                from .deform_net import joint_mlp_backward_raw, joint_mlp_forward_raw
                if not hasattr(self, '_mlp_buffers'):
                    self._mlp_buffers = {}
                (sk_r, sk_d_rot, sk_d_scale), cm = joint_mlp_forward_raw(self.mlp.cfg, p['theta'].data, p['joints'],
                                                                         self.t, out=self._mlp_buffers.setdefault(view, {}))
            else:
                sk_r, sk_d_rot, sk_d_scale = p['sk_r'], p['sk_d_rot'], p['sk_d_scale']
            (d_xyz, d_rot, d_scale, sk_T, weights, indices), c1 = fk_lbs_forward_raw(
                p['xyz'], p['joints'], sk_r, sk_d_rot, sk_d_scale, p['g_tr'], self.parents, self.root,
                K=self.K, mode=self.mode, sp_W=p['sp_W'] if W else None, sp_radius=None if W else self.sp_radius,
                sp_weight=self.sp_weight if self.mode == 'weighted_kernel' else None)
            (points, scales, rotations, opacity), c2 = assemble_forward_raw(p['xyz'], p['scaling'], p['rotation'],
                                                                            p['opacity'], d_xyz, d_rot, d_scale)
            sh = p['shs'] if 'shs' in p else torch.cat((p['f_dc'], p['f_rest']), dim=1)
            color, depth, alpha, radii, st = DGR.rasterize_forward(self.settings[view], points, opacity, shs=sh,
                                                                   scales=scales, rotations=rotations, quat_wxyz=False,
                                                                   fixed_capacity=fixed_capacity,
                                                                   header_words=header_words)
            join_after = after_forward(radii) if after_forward is not None else None  # e.g. radii MAX on a side stream
            if before_backward is not None:
                before_backward()  # e.g. join the stream that uploads dL_dimage / the target while the forward runs
            loss_terms = None
            if target is not None:
                from .losses import image_loss_raw
                if not hasattr(self, '_loss_buffers'):
                    self._loss_buffers = {}
                loss_terms, dL_dimage = image_loss_raw(color, target, out=self._loss_buffers.setdefault(view, {}),
                                                       **(loss or {}))
            # with an arena (sk_gs_b200.dist.GradArena) every final gradient is written straight into its slot of the
            # flat all-reduce buffer: no packing copies before the exchange
            A = (lambda n: arena.view(n) if n in arena.offsets else None) if arena is not None else (lambda n: None)
            g = DGR.rasterize_backward(st, dL_dimage, out={'means3D': A('xyz'), 'means2D': A('viewspace_points'),
                                                          'shs': A('shs')})
            _, dscaling, drotation, dopacity, dd_xyz, dd_rot, dd_scale = assemble_backward_raw(
                c2, g['means3D'], g['scales'], g['rotations'], g['opacities'],
                out={'scaling': A('scaling'), 'rotation': A('rotation'), 'opacity': A('opacity')},
                need=[False, True, True, True, True, True, True])
            dxyz = g['means3D']  # d points / d _xyz is the identity
            join_mid = mid_backward() if mid_backward is not None else None  # all rasterizer-side gradients are final
            d_joints, d_sk_r, d_sk_d_rot, d_sk_d_scale, d_g_tr, d_sp_W, d_sp_radius, d_sp_weight = fk_lbs_backward_raw(
                c1, dd_xyz, dd_rot, dd_scale, compact_sp_W=compact_sp_W,
                out={n: A(n) for n in ('joints', 'sk_r', 'sk_d_rot', 'sk_d_scale', 'g_tr', 'sp_W')})
            d_theta = None
            if cm is not None:  # back through the joint-rotation network; joints also feed the network input
                d_theta, d_joints_net = joint_mlp_backward_raw(cm, d_sk_r, d_sk_d_rot, d_sk_d_scale,
                                                               out={'theta': A('theta')})
                d_joints.add_(d_joints_net)
        grads = {'xyz': dxyz, 'scaling': dscaling, 'rotation': drotation, 'opacity': dopacity, 'shs': g['shs'],
                 'f_dc': g['shs'][:, :1], 'f_rest': g['shs'][:, 1:], 'sp_W': d_sp_W, 'joints': d_joints, 'sk_r': d_sk_r,
                 'sk_d_rot': d_sk_d_rot, 'sk_d_scale': d_sk_d_scale, 'g_tr': d_g_tr, 'viewspace_points': g['means2D'],
                 'sp_radius': d_sp_radius, 'sp_weight': d_sp_weight, 'theta': d_theta}
        if join_after is not None:
            join_after()
        if join_mid is not None:
            join_mid()
        out = {'images': color, 'depths': depth, 'alpha': alpha, 'radii': radii, 'visibility_filter': None,
               'loss_terms': loss_terms, '_raster_state': st,
               'viewspace_points': None, '_sk': (d_xyz, d_rot, d_scale, sk_T, sk_d_rot, sk_d_scale, p['g_tr'],
                                                 weights, indices)}
        return out, grads

    def capture_step(self, view: int, dL_dimage: Tensor, headroom: float = 1.3, compact_sp_W: bool = False,
                     uploads=None, dL_host: Optional[Tensor] = None, epilogue=None, arena=None, after_forward=None,
                     mid_backward=None, target: Optional[Tensor] = None, loss: Optional[dict] = None,
                     target_host: Optional[Tensor] = None):
        """Capture forward + backward of one view into a CUDA graph (static shapes, fixed binning capacity = headroom x
        the R observed in an eager warm-up).  Returns (graph, outputs, grads); replay with graph.replay(), results appear
        in the returned tensors.  The capacity belongs to THIS graph (nothing process-wide changes; eager calls keep
        sizing their arena from R).  Every graph owns a pinned header-word buffer (outputs['_header_words']): after a
        replay has finished, `self.overflowed()` tells whether R exceeded the capacity of any graph captured from this
        HotPath, and outputs['_raster_state'].overflow_ptr is the device-side flag (skgs_adam_step honours it)."""
        from . import _lib
        from . import diff_gaussian_rasterization as DGR
        kw = dict(arena=arena, after_forward=after_forward, mid_backward=mid_backward, target=target, loss=loss)
        o_, g_ = self.step_grads(view, dL_dimage, compact_sp_W, **kw)
        torch.cuda.synchronize(self.device)
        R = int(DGR.last_header_words(self.device)[0])
        words = DGR.new_header_words()  # pinned memory must be allocated before the capture starts
        kw.update(fixed_capacity=int(R * headroom) + 4096, header_words=words)
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(3):
                o_, g_ = self.step_grads(view, dL_dimage, compact_sp_W, **kw)
                if epilogue is not None:
                    epilogue(o_, g_)  # e.g. the NCCL gradient exchange: communicators must exist before capture
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(graph):
            # optional host->device uploads captured INTO the graph: the small per-step inputs (camera, joint
            # rotations) first, the large upstream gradient on a forked stream that is joined right before backward,
            # so its PCIe time hides behind the forward kernels
            join = None
            for dst, src in (uploads or []):
                dst.copy_(src, non_blocking=True)
            late = [(d, h) for d, h in ((dL_dimage, dL_host), (target, target_host)) if h is not None]
            if late:
                main = torch.cuda.current_stream(self.device)
                up = torch.cuda.Stream(self.device)
                up.wait_stream(main)
                with torch.cuda.stream(up):
                    for d, h in late:
                        d.copy_(h, non_blocking=True)
                join = lambda: main.wait_stream(up)  # noqa: E731
            out, grads = self.step_grads(view, dL_dimage, compact_sp_W, before_backward=join, **kw)
            if epilogue is not None:
                epilogue(out, grads)
        self.launches_per_step = _lib.launch_count() - n0  # kernels of libskgs_b200.so inside one replay
        torch.cuda.synchronize(self.device)
        out['_header_words'] = words
        out['_capacity'] = kw['fixed_capacity']
        if not hasattr(self, '_graph_words'):
            self._graph_words = []
        self._graph_words.append(words)
        return graph, out, grads

    def overflowed(self) -> bool:
        """True if the most recent (completed) replay of ANY graph captured from this HotPath exceeded its binning
        capacity: that replay's image and gradients are invalid.  Call after a synchronisation point."""
        return any(int(w[3]) != 0 for w in getattr(self, '_graph_words', []))

    def zero_grad(self):
        for t in list(self.params.values()) + [self.sp_radius, self.sp_weight]:
            t.grad = None

    def grads(self) -> Dict[str, Tensor]:
        return {n: t.grad for n, t in self.params.items() if t.grad is not None}
