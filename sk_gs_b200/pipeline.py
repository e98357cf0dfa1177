"""The hot path end to end on one GPU: parameters -> FK + LBS -> assembly -> rasterize (-> backward).

Mirrors what `SkeletonGaussianSplatting.render` does per view in the `sk` stage
(/root/reference/networks/sk_gs.py:1206-1242 -> forward :1160-1204 -> sk_stage :1109-1150 -> render_gs_offical),
with the joint rotations as leaf parameters by default.  The step before the path (joint-rotation network, SURVEY.md 8f-1,
`joint_mlp=True`) and the two steps after it - photometric loss (8f-2, networks/sk_gs.py:1524-1529) and Adam (8f-3,
networks/gaussian_splatting.py:445-453) - are optional stages of `step_grads` / `sk_gs_b200.train.TrainLoop`; the
benchmarked metric (BASELINE.json) excludes them.

A training step of the reference loops over its views and lets autograd sum their gradients (sk_gs.py:1220); here
`step_views` / `capture_step` take a LIST of views: every view runs the whole path, its gradients land in a flat arena
(sk_gs_b200.dist.GradArena) and the arenas of the 2nd, 3rd ... view are folded into the first (skgs_accumulate_f32)."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Union

import torch
from torch import Tensor

from . import _lib
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


def accumulate_(dst: Tensor, src: Tensor):
    """dst += src over flat fp32 buffers (libskgs_b200.so, current stream)."""
    assert dst.is_contiguous() and src.is_contiguous() and dst.numel() == src.numel()
    st = torch.cuda.current_stream(dst.device).cuda_stream
    _lib.check(_lib.lib().skgs_accumulate_f32(dst.data_ptr(), src.data_ptr(), dst.numel(), st), 'skgs_accumulate_f32')
    return dst


def maximum_(dst: Tensor, src: Tensor):
    """dst = max(dst, src) over int32 radii (libskgs_b200.so, current stream)."""
    assert dst.dtype == torch.int32 and src.dtype == torch.int32 and dst.numel() == src.numel()
    st = torch.cuda.current_stream(dst.device).cuda_stream
    _lib.check(_lib.lib().skgs_max_i32(dst.data_ptr(), src.data_ptr(), dst.numel(), st), 'skgs_max_i32')
    return dst


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
        self._graph_words: List[Tensor] = []

    # ---------------------------------------------------------------------------------------------- autograd (drop-in)
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
        """The reference's call sequence through the drop-in operators (fk_lbs -> assemble -> render_gs_offical), with
        autograd."""
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

    # --------------------------------------------------------------------------------------- hand-driven (capturable)
    def forward_raw(self, view: int, fixed_capacity: Optional[int] = None, header_words: Optional[Tensor] = None,
                    fused: Optional[bool] = None):
        """FK + LBS -> assembly -> rasterize of one view with the operators driven by hand (no autograd graph).
        Returns (outputs dict, context for backward_raw).  `fused` (default: whenever possible - LBS mode W and a known
        binning capacity): the per-Gaussian part runs as ONE kernel (skgs_deform_forward_geometry), bit-identical to
        the three operator kernels."""
        from . import diff_gaussian_rasterization as DGR
        from .fk_lbs import assemble_forward_raw, fk_lbs_forward_raw
        rs = self.settings[view]
        P = self.params['xyz'].shape[0]
        can_fuse = self.mode == 'W' and P > 0 and (fixed_capacity is not None or DGR.capacity_known(
            self.device, P, int(rs.image_width), int(rs.image_height)))
        if fused is None:
            fused = can_fuse
        if fused and not can_fuse:
            raise RuntimeError('fused forward needs LBS mode W and a known binning capacity')
        if fused:
            return self._forward_fused(view, fixed_capacity, header_words)
        with torch.no_grad():
            p = self.params
            W = self.mode == 'W'
            cm = None
            if self.mlp is not None:
                from .deform_net import joint_mlp_forward_raw
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
        out = {'images': color, 'depths': depth, 'alpha': alpha, 'radii': radii, 'visibility_filter': None,
               'loss_terms': None, '_raster_state': st, 'viewspace_points': None,
               '_sk': (d_xyz, d_rot, d_scale, sk_T, sk_d_rot, sk_d_scale, p['g_tr'], weights, indices)}
        return out, (cm, c1, c2, st)

    def _forward_fused(self, view: int, fixed_capacity, header_words):
        """fk_table_kernel + deform_preprocess_kernel + sort + compositing; builds the same backward contexts as the
        operator-by-operator path."""
        from . import diff_gaussian_rasterization as DGR
        from .fk_lbs import _Ctx, _skeleton
        with torch.no_grad():
            p, dev = self.params, self.device
            cm = None
            if self.mlp is not None:
                from .deform_net import joint_mlp_forward_raw
                if not hasattr(self, '_mlp_buffers'):
                    self._mlp_buffers = {}
                (sk_r, sk_d_rot, sk_d_scale), cm = joint_mlp_forward_raw(self.mlp.cfg, p['theta'].data, p['joints'],
                                                                         self.t, out=self._mlp_buffers.setdefault(view, {}))
            else:
                sk_r, sk_d_rot, sk_d_scale = p['sk_r'], p['sk_d_rot'], p['sk_d_scale']
            f = DGR._f32c
            xyz, joints = f(p['xyz'].detach()), f(p['joints'].detach())
            sk_r, sk_d_rot, sk_d_scale = f(sk_r.detach()), f(sk_d_rot.detach()), f(sk_d_scale.detach())
            g_tr = f(p['g_tr'].detach().reshape(-1))
            sp_W = f(p['sp_W'].detach())
            scaling, rotation, opacity = f(p['scaling'].detach()), f(p['rotation'].detach()), f(p['opacity'].detach())
            parents = self.parents.to(device=dev, dtype=torch.int32).contiguous()
            P, M, K = xyz.shape[0], joints.shape[0], self.K
            sk = _skeleton(joints, sk_r, sk_d_rot, sk_d_scale, g_tr, parents, self.root, K, 'W', sp_W, None, None, 1.0,
                           None)
            new = lambda *s, **kw: torch.empty(*s, device=dev, **kw)  # noqa: E731
            points, scales, rotations, opac = new(P, 3), new(P, 3), new(P, 4), new(P, 1)
            d_rot, weights, sk_T = new(P, 4), new(P, K), new(M, 7)
            indices = new(P, K, dtype=torch.int64)
            ws = new(_lib.lib().skgs_fk_lbs_workspace_bytes(M), dtype=torch.uint8)
            sh = p['shs'] if 'shs' in p else torch.cat((p['f_dc'], p['f_rest']), dim=1)
            deform = dict(sk=sk, xyz=xyz, scaling=scaling, rotation=rotation, opacity=opacity, d_rot=d_rot,
                          weights=weights, indices=indices, sk_T=sk_T, workspace=ws)
            color, depth, alpha, radii, st = DGR.rasterize_forward(
                self.settings[view], points, opac, shs=sh, scales=scales, rotations=rotations, quat_wxyz=False,
                fixed_capacity=fixed_capacity, header_words=header_words, deform=deform)
            # contexts of the two operator backwards (same fields their forward would have saved)
            c1 = _Ctx([False, True, True, True, True, True, True, False, False] + [False] * 6)
            c1.sk = sk
            c1.keep = (xyz, joints, sk_r, sk_d_rot, sk_d_scale, g_tr, sp_W, None, None, parents, None, sk_T, weights,
                       indices)
            c1.mode = 'W'
            c2 = _Ctx([True] * 7)
            c2.keep = (scaling, rotation, opacity, d_rot)
            c2.has = (True, True, True)
        out = {'images': color, 'depths': depth, 'alpha': alpha, 'radii': radii, 'visibility_filter': None,
               'loss_terms': None, '_raster_state': st, 'viewspace_points': None, '_fused': True,
               '_sk': (None, d_rot, None, sk_T, sk_d_rot, sk_d_scale, p['g_tr'], weights, indices)}
        return out, (cm, c1, c2, st)

    def backward_raw(self, ctx, dL_dimage: Tensor, compact_sp_W: bool = False, arena=None, mid_backward=None,
                     fused: bool = True):
        """The backward calls in reverse.  Returns ({parameter name: gradient}, join for mid_backward).  `fused`
        (default): the assembly backward runs inside the rasterizer's per-Gaussian backward kernel
        (skgs_raster_assemble_backward) instead of as a kernel of its own."""
        from . import diff_gaussian_rasterization as DGR
        from .fk_lbs import assemble_backward_raw, fk_lbs_backward_raw
        cm, c1, c2, st = ctx
        with torch.no_grad():
            # with an arena (sk_gs_b200.dist.GradArena) every final gradient is written straight into its slot of the
            # flat all-reduce buffer: no packing copies before the exchange
            A = (lambda n: arena.view(n) if n in arena.offsets else None) if arena is not None else (lambda n: None)
            if fused:
                scaling, rotation, opacity, d_rot = c2.keep
                # mid_backward hands the rasterizer-side arena blocks ('xyz' and 'rotation' among them) to an all-reduce
                # that runs WHILE the LBS backward reads dL/d_xyz and dL/d_rot: those two then need private copies
                ga = DGR.rasterize_assemble_backward(
                    st, scaling, rotation, opacity, d_rot, dL_dimage,
                    out={'xyz': A('xyz'), 'means2D': A('viewspace_points'), 'shs': A('shs'), 'scaling': A('scaling'),
                         'rotation': A('rotation'), 'opacity': A('opacity')},
                    private_lbs_inputs=mid_backward is not None and arena is not None)
                g = {'means3D': ga['xyz'], 'means2D': ga['means2D'], 'shs': ga['shs']}
                dscaling, drotation, dopacity = ga['scaling'], ga['rotation'], ga['opacity']
                dd_xyz, dd_rot, dd_scale = ga['dd_xyz'], ga['dd_rot'], ga['dd_scale']
            else:
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
                from .deform_net import joint_mlp_backward_raw
                d_theta, d_joints_net = joint_mlp_backward_raw(cm, d_sk_r, d_sk_d_rot, d_sk_d_scale,
                                                               out={'theta': A('theta')})
                d_joints.add_(d_joints_net)
        grads = {'xyz': dxyz, 'scaling': dscaling, 'rotation': drotation, 'opacity': dopacity, 'shs': g['shs'],
                 'f_dc': g['shs'][:, :1], 'f_rest': g['shs'][:, 1:], 'sp_W': d_sp_W, 'joints': d_joints, 'sk_r': d_sk_r,
                 'sk_d_rot': d_sk_d_rot, 'sk_d_scale': d_sk_d_scale, 'g_tr': d_g_tr, 'viewspace_points': g['means2D'],
                 'sp_radius': d_sp_radius, 'sp_weight': d_sp_weight, 'theta': d_theta}
        return grads, join_mid

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
        out, ctx = self.forward_raw(view, fixed_capacity, header_words)
        join_after = after_forward(out['radii']) if after_forward is not None else None  # e.g. radii MAX, side stream
        if before_backward is not None:
            before_backward()  # e.g. join the stream that uploads dL_dimage / the target while the forward runs
        if target is not None:
            from .losses import image_loss_raw
            if not hasattr(self, '_loss_buffers'):
                self._loss_buffers = {}
            with torch.no_grad():
                out['loss_terms'], dL_dimage = image_loss_raw(out['images'], target,
                                                              out=self._loss_buffers.setdefault(view, {}), **(loss or {}))
        grads, join_mid = self.backward_raw(ctx, dL_dimage, compact_sp_W, arena, mid_backward)
        if join_after is not None:
            join_after()
        if join_mid is not None:
            join_mid()
        return out, grads

    def step_views(self, views: Sequence[int], dLs: Sequence[Optional[Tensor]], arena, scratch, compact_sp_W: bool = False,
                   targets: Optional[Sequence[Tensor]] = None, loss: Optional[dict] = None,
                   capacities: Optional[Sequence[int]] = None, words: Optional[Sequence[Tensor]] = None,
                   before_backward=None):
        """One multi-view step on this rank: the whole path for every view in `views`, gradients SUMMED in `arena`
        (view 0 writes its slots directly, view v > 0 writes `scratch` - an arena of the same layout - which is then
        folded in), screen radii MAXed.  Returns (outputs of the last view with 'radii' = max over views, grads = views
        of `arena`)."""
        assert len(views) >= 1 and arena is not None
        radii_max = None
        out = None
        for k, v in enumerate(views):
            a = arena if k == 0 else scratch
            out, _ = self.step_grads(v, dLs[k], compact_sp_W, arena=a, target=None if targets is None else targets[k],
                                     loss=loss, fixed_capacity=None if capacities is None else capacities[k],
                                     header_words=None if words is None else words[k],
                                     before_backward=before_backward if k == 0 else None)
            if k == 0:
                radii_max = out['radii']
            else:
                accumulate_(arena.flat, scratch.flat)
                maximum_(radii_max, out['radii'])
        out = dict(out)
        out['radii'] = radii_max
        return out, arena.unpack()

    # ------------------------------------------------------------------------------------------------ CUDA graphs
    def measure_capacity(self, view: int, headroom: float) -> int:
        """Binning capacity for a captured graph of `view`: headroom x the R of an eager forward at the current
        parameters (+ slack)."""
        from . import diff_gaussian_rasterization as DGR
        self.forward_raw(view, fused=False)
        torch.cuda.synchronize(self.device)
        R = int(DGR.last_header_words(self.device)[0]) & 0xffffffff
        return int(R * headroom) + 4096

    def capture_step(self, view: Union[int, Sequence[int]], dL_dimage, headroom: float = 1.3,
                     compact_sp_W: bool = False, uploads=None, dL_host=None, epilogue=None, arena=None,
                     after_forward=None, mid_backward=None, target=None, loss: Optional[dict] = None,
                     target_host=None, scratch=None):
        """Capture forward + backward of one view - or of a LIST of views (multi-view step, needs `arena` + `scratch`;
        `dL_dimage` / `target` / `dL_host` / `target_host` are then lists) - into a CUDA graph: static shapes, fixed
        binning capacity per view = headroom x the R observed in an eager warm-up.  Returns (graph, outputs, grads);
        replay with graph.replay(), results appear in the returned tensors.  The capacities belong to THIS graph (nothing
        process-wide changes; eager calls keep sizing their arena from R).  Every captured view owns a pinned
        header-word buffer (outputs['_header_words'], a list for several views): after a replay has finished,
        `self.overflowed()` tells whether R exceeded the capacity of any graph captured from this HotPath, and
        outputs['_raster_state'].overflow_ptr is the device-side flag (skgs_adam_step honours it)."""
        from . import diff_gaussian_rasterization as DGR
        multi = not isinstance(view, int)
        views = list(view) if multi else [view]
        as_list = (lambda x: list(x) if multi else [x])
        dLs, targets = as_list(dL_dimage), (None if target is None else as_list(target))
        dL_hosts = [None] * len(views) if dL_host is None else as_list(dL_host)
        target_hosts = [None] * len(views) if target_host is None else as_list(target_host)
        caps = [self.measure_capacity(v, headroom) for v in views]
        words = [DGR.new_header_words() for _ in views]  # pinned memory must be allocated before the capture starts

        def run(before_backward=None):
            if multi:
                return self.step_views(views, dLs, arena, scratch, compact_sp_W, targets, loss, caps, words,
                                       before_backward)
            return self.step_grads(views[0], dLs[0], compact_sp_W, arena=arena, after_forward=after_forward,
                                   mid_backward=mid_backward, target=None if targets is None else targets[0], loss=loss,
                                   fixed_capacity=caps[0], header_words=words[0], before_backward=before_backward)

        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(3):
                o_, g_ = run()
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
            late = [(d, h) for d, h in list(zip(dLs, dL_hosts)) + list(zip(targets or [], target_hosts))
                    if h is not None]
            if late:
                main = torch.cuda.current_stream(self.device)
                up = torch.cuda.Stream(self.device)
                up.wait_stream(main)
                with torch.cuda.stream(up):
                    for d, h in late:
                        d.copy_(h, non_blocking=True)
                join = lambda: main.wait_stream(up)  # noqa: E731
            out, grads = run(before_backward=join)
            if epilogue is not None:
                epilogue(out, grads)
        self.launches_per_step = _lib.launch_count() - n0  # kernels of libskgs_b200.so inside one replay
        torch.cuda.synchronize(self.device)
        out['_header_words'] = words if multi else words[0]
        out['_capacity'] = caps if multi else caps[0]
        self._graph_words.extend(words)
        return graph, out, grads

    def capture_render(self, view: int, headroom: float = 1.3):
        """Forward only (FK + LBS + assembly + rasterize) of one view as a CUDA graph: the reference's FPS protocol
        (test.py:102-123) without the host in the loop.  Returns (graph, outputs)."""
        from . import diff_gaussian_rasterization as DGR
        cap = self.measure_capacity(view, headroom)
        words = DGR.new_header_words()
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(2):
                self.forward_raw(view, cap, words)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(graph):
            out, _ = self.forward_raw(view, cap, words)
        self.launches_per_render = _lib.launch_count() - n0
        torch.cuda.synchronize(self.device)
        out['_header_words'] = words
        out['_capacity'] = cap
        self._graph_words.append(words)
        return graph, out

    def overflowed(self) -> bool:
        """True if the most recent (completed) replay of ANY graph captured from this HotPath exceeded its binning
        capacity: that replay's image and gradients are invalid.  Call after a synchronisation point."""
        return any(int(w[3]) != 0 for w in self._graph_words)

    def zero_grad(self):
        for t in list(self.params.values()) + [self.sp_radius, self.sp_weight]:
            t.grad = None

    def grads(self) -> Dict[str, Tensor]:
        return {n: t.grad for n, t in self.params.items() if t.grad is not None}
