"""Generate golden vectors from the REFERENCE's own pure-torch helpers (run in the authoring container, where
/root/reference exists; the .npz files are committed, the GPU box never needs the reference).

The reference cannot be imported as-is (lietorch / pytorch3d / diff_gaussian_rasterization ... are absent):
missing third-party packages are replaced by inert stub modules so that the modules holding the pure-torch formulas can
be imported UNMODIFIED from /root/reference.  Functions exercised (all torch, CPU):
  networks/encoders/sphere_harmonics.py : eval_sh, RGB2SH                       -> sh.npz
  networks/GS_utils.py                   : build_rotation, compute_cov2D        -> cov.npz  (cov3D via R S S R^T as
                                           build_covariance_from_scaling_rotation :44-82, whose "cuda" literal is bypassed)
  my_ext/ops_3d/rigid.py                 : quaternion_to_Rt                     -> fk.npz
  networks/sk_gs.py                      : find_root, skeleton_warp, skeleton_warp_v0   -> fk.npz
  my_ext/ops_3d/coord_trans_opencv.py    : perspective                          -> cam.npz
  networks/losses/ssim.py, image_loss.py : SSIM_Loss, ImageLoss (+ autograd)    -> loss.npz
  torch.optim.Adam (the reference's optimizer, gaussian_splatting.py:445-453)   -> adam.npz
  networks/sk_gs.py                      : SimpleDeformationNetwork ('freq_torch' encoders, fp64) -> deform_net.npz
  networks/sk_gs.py                      : calc_LBS_weight, warp (sp-stage; lietorch SE3/SO3 and pytorch3d knn_points
                                           replaced by functional stand-ins written from lie.h:59-64,246) -> sp_stage.npz
  networks/gaussian_splatting.py         : add_densification_stats, densify (clone + split), prune, reset_opacity with
                                           change_optimizer on a real torch.optim.Adam                -> densify.npz
"""
import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get('REF', '/root/reference')
OUT = os.path.dirname(os.path.abspath(__file__))
STUBS = ['lietorch', 'pytorch3d', 'pykdtree', 'plyfile', 'torchmetrics', 'dearpygui', 'imageio', 'matplotlib', 'cv2',
         'diff_gaussian_rasterization', 'open3d', 'trimesh', 'lpips', 'skimage', 'PIL', 'tqdm_joblib', 'kornia',
         'tensorboard', 'tensorboardX', 'pymeshlab', 'xatlas', 'nvdiffrast', 'tinycudann', 'seaborn', 'termcolor',
         'prettytable', 'torch_scatter', 'einops_exts', 'pyrender', 'mcubes', 'pysdf', 'OpenGL', 'glfw', 'moderngl']


class _Anything(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__') and name.endswith('__'):
            raise AttributeError(name)
        return _Dummy()


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]  # used as a bare decorator
        return _Dummy()

    def __getattr__(self, name):
        return _Dummy()

    def __mro_entries__(self, bases):
        return (object,)

    def __iter__(self):
        return iter(())


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split('.')[0] in STUBS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Anything(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def main():
    sys.meta_path.insert(0, _StubFinder())
    sys.path.insert(0, REF)
    os.chdir(REF)
    g = torch.Generator().manual_seed(20241017)
    # ---- SH
    sh_mod = importlib.import_module('networks.encoders.sphere_harmonics')
    P = 257
    sh = torch.randn(P, 3, 16, generator=g) * 0.3
    dirs = torch.nn.functional.normalize(torch.randn(P, 3, generator=g), dim=-1)
    sh_out = {f'rgb_deg{d}': sh_mod.eval_sh(d, sh, dirs).numpy() for d in range(4)}
    rgb = torch.rand(P, 3, generator=g)
    np.savez(os.path.join(OUT, 'sh.npz'), sh=sh.numpy(), dirs=dirs.numpy(), rgb=rgb.numpy(),
             rgb2sh=sh_mod.RGB2SH(rgb).numpy(), **sh_out)
    # ---- covariances
    gs = importlib.import_module('networks.GS_utils')
    q = torch.randn(P, 4, generator=g)
    s = torch.exp(torch.randn(P, 3, generator=g) * 0.5 - 3.0)
    R = gs.build_rotation(q)  # normalises q, xyzw
    L = R @ torch.diag_embed(s)
    cov = L @ L.transpose(1, 2)
    cov6 = torch.stack([cov[:, 0, 0], cov[:, 0, 1], cov[:, 0, 2], cov[:, 1, 1], cov[:, 1, 2], cov[:, 2, 2]], -1)
    pts = torch.randn(P, 3, generator=g) * 0.5
    Tw2v = torch.eye(4)
    ang = 0.3
    Tw2v[:3, :3] = torch.tensor([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    Tw2v[:3, 3] = torch.tensor([0.1, -0.2, 4.0])
    fx, fy, tx, ty = 1100.0, 1050.0, 0.36, 0.30
    cov2d = gs.compute_cov2D(pts, cov6, Tw2v, fx, fy, tx, ty)
    np.savez(os.path.join(OUT, 'cov.npz'), q=q.numpy(), s=s.numpy(), R=R.numpy(), cov6=cov6.numpy(), pts=pts.numpy(),
             Tw2v=Tw2v.numpy(), fx=fx, fy=fy, tanfovx=tx, tanfovy=ty, cov2d=cov2d.numpy())
    # ---- FK: matrix formulation of the reference + its tree builder
    rigid = importlib.import_module('my_ext.ops_3d.rigid')
    sk = importlib.import_module('networks.sk_gs')
    fk = {}
    for idx, M in enumerate((3, 8, 24, 33, 64)):
        if idx == 3:
            father = torch.tensor([-1] + list(range(M - 1)))  # chain
        else:
            father = torch.tensor([-1] + [int(torch.randint(0, j, (1,), generator=g)) for j in range(1, M)])
        parents, depth, root = sk.find_root(father)
        joints = torch.randn(M, 3, generator=g).double() * 0.4
        r = torch.nn.functional.normalize(torch.randn(M, 4, generator=g).double(), dim=-1)
        gq = torch.nn.functional.normalize(torch.randn(4, generator=g).double(), dim=-1)
        gt = torch.randn(3, generator=g).double() * 0.1
        Rm = rigid.quaternion_to_Rt(r)  # [M,4,4] rotation only
        local = Rm.clone()
        local[:, :3, 3] = joints - (Rm[:, :3, :3] @ joints[:, :, None])[:, :, 0]  # rotation about the joint position
        G = rigid.quaternion_to_Rt(gq, gt)
        T_jump = sk.skeleton_warp(local, G, parents, root)
        T_v0 = sk.skeleton_warp_v0(local, G, parents, root)
        fk.update({f'father{idx}': father.numpy(), f'parents{idx}': parents.numpy(), f'depth{idx}': depth.numpy(),
                   f'root{idx}': np.int64(root), f'joints{idx}': joints.numpy(), f'r{idx}': r.numpy(),
                   f'g_tr{idx}': torch.cat([gt, gq]).numpy(), f'T_jump{idx}': T_jump.numpy(), f'T_v0{idx}': T_v0.numpy()})
    np.savez(os.path.join(OUT, 'fk.npz'), n=np.int64(5), **fk)
    # ---- camera
    cv = importlib.import_module('my_ext.ops_3d.coord_trans_opencv')
    Tv2c = cv.perspective(fovy=0.6911, n=0.01, f=1000.0, size=(800, 800))
    Tv2c2 = cv.perspective(fovy=0.5, n=0.01, f=1000.0, size=(1920, 1080))
    np.savez(os.path.join(OUT, 'cam.npz'), Tv2c_800=Tv2c.numpy(), Tv2c_1080p=Tv2c2.numpy())
    make_loss()
    make_adam()
    make_deform_net()
    make_sp_stage()
    make_densify()
    print('golden vectors written to', OUT)


def make_loss():
    """The reference's loss modules on random images, fp32 (what the reference runs) and fp64 (tight check of the oracle);
    the image is [1,H,W,3] as at networks/sk_gs.py:1527, weights of exps/default.yaml:83-84."""
    ssim_mod = importlib.import_module('networks.losses.ssim')
    img_mod = importlib.import_module('networks.losses.image_loss')
    g = torch.Generator().manual_seed(20241017 + 100)
    out = {}
    cases = [(37, 45), (64, 80), (11, 9)]
    for i, (H, W) in enumerate(cases):
        gt = torch.rand(1, H, W, 3, generator=g)
        img = (gt + 0.2 * torch.randn(1, H, W, 3, generator=g)).clamp(0, 1) if i != 1 else torch.rand(1, H, W, 3, generator=g)
        if i == 0:
            img[0, :4, :5] = gt[0, :4, :5]  # exact zeros of the L1 term
        for dt, tag in ((torch.float32, 'f32'), (torch.float64, 'f64')):
            x = img.detach().clone().to(dt).requires_grad_(True)
            y = gt.to(dt)
            l1 = img_mod.ImageLoss(method='l1')(x, y)
            mse = img_mod.ImageLoss(method='mse')(x, y)
            ss = ssim_mod.SSIM_Loss()(x, y)
            g_l1, = torch.autograd.grad(l1, x, retain_graph=True)
            g_mse, = torch.autograd.grad(mse, x, retain_graph=True)
            g_ss, = torch.autograd.grad(ss, x, retain_graph=True)
            total = 0.8 * l1 + 0.2 * ss
            g_total, = torch.autograd.grad(total, x)
            out.update({f'l1_{tag}_{i}': l1.detach().numpy(), f'mse_{tag}_{i}': mse.detach().numpy(),
                        f'ssim_{tag}_{i}': ss.detach().numpy(), f'total_{tag}_{i}': total.detach().numpy(),
                        f'g_l1_{tag}_{i}': g_l1.numpy(), f'g_mse_{tag}_{i}': g_mse.numpy(),
                        f'g_ssim_{tag}_{i}': g_ss.numpy(), f'g_total_{tag}_{i}': g_total.numpy()})
        out.update({f'img_{i}': img.numpy(), f'gt_{i}': gt.numpy()})
    np.savez_compressed(os.path.join(OUT, 'loss.npz'), n=np.int64(len(cases)), **out)


def make_deform_net():
    """The reference's joint-rotation network (networks/sk_gs.py:134-164) with the configuration of
    exps/default.yaml:48-55 but the pure-torch encoder variant (the CUDA one cannot run here), in float64, small width so
    the fixture stays small; outputs and gradients w.r.t. joints and every parameter."""
    sk = importlib.import_module('networks.sk_gs')
    torch.manual_seed(20241017 + 300)
    out = {}
    cases = [dict(M=16, width=32, depth=8, skips=(4,)), dict(M=5, width=16, depth=4, skips=(1, 2))]
    for ci, c in enumerate(cases):
        net = sk.SimpleDeformationNetwork(p_in_channels=3, t_in_channels=1, out_channels=[4, 4, 3], width=c['width'],
                                          depth=c['depth'], skips=c['skips'], pos_enc_p='freq_torch',
                                          pos_enc_p_cfg=dict(degree=10), pos_enc_t='freq_torch',
                                          pos_enc_t_cfg=dict(degree=6)).double()
        for m in net.dynamic_net.last:  # the reference's 1e-6 init would put every gradient at noise level
            torch.nn.init.normal_(m.weight, std=0.05)
            torch.nn.init.normal_(m.bias, std=0.05)
        joints = (torch.randn(c['M'], 3, dtype=torch.float64) * 0.4).requires_grad_(True)
        t = torch.tensor([0.37], dtype=torch.float64)
        o_r, o_rot, o_s = net(joints, t)
        g_r, g_rot, g_s = torch.randn_like(o_r), torch.randn_like(o_rot), torch.randn_like(o_s)
        params = list(net.dynamic_net.net.parameters()) + list(net.dynamic_net.last.parameters())
        grads = torch.autograd.grad([o_r, o_rot, o_s], [joints] + params, [g_r, g_rot, g_s])
        out.update({f'joints{ci}': joints.detach().numpy(), f't{ci}': t.numpy(), f'o_r{ci}': o_r.detach().numpy(),
                    f'o_rot{ci}': o_rot.detach().numpy(), f'o_s{ci}': o_s.detach().numpy(), f'g_r{ci}': g_r.numpy(),
                    f'g_rot{ci}': g_rot.numpy(), f'g_s{ci}': g_s.numpy(), f'd_joints{ci}': grads[0].numpy(),
                    f'cfg{ci}': np.array([c['M'], c['width'], c['depth']] + list(c['skips']))})
        layers = list(net.dynamic_net.net) + list(net.dynamic_net.last)
        for li, layer in enumerate(layers):
            out[f'w{ci}_{li}'] = layer.weight.detach().numpy()
            out[f'b{ci}_{li}'] = layer.bias.detach().numpy()
        for li in range(len(layers)):
            out[f'dw{ci}_{li}'] = grads[1 + 2 * li].numpy()
            out[f'db{ci}_{li}'] = grads[2 + 2 * li].numpy()
    np.savez_compressed(os.path.join(OUT, 'deform_net.npz'), n=np.int64(len(cases)), **out)


# ---- functional stand-ins for the two un-vendored packages `warp` / `calc_LBS_weight` call into
def _rot(q, p):
    """my_ext/_C/include/lie.h:59-64 (the rotation lietorch's SO3 / SE3 act applies; q used as stored)."""
    v, w = q[..., :3], q[..., 3:]
    v, p = torch.broadcast_tensors(v, p)
    uv = 2 * torch.linalg.cross(v, p)
    return p + w * uv + torch.linalg.cross(v, uv)


class _SO3Stub:
    def __init__(self, data):
        self.data = data

    @classmethod
    def InitFromVec(cls, v):
        return cls(v)

    def __getitem__(self, i):
        return type(self)(self.data[i])

    def act(self, p):
        return _rot(self.data, p)

    def vec(self):
        return self.data


class _SE3Stub(_SO3Stub):
    def act(self, p):  # lie.h:246
        return _rot(self.data[..., 3:], p) + self.data[..., :3]


def _knn_points_stub(p1, p2, l1=None, l2=None, K=1):
    """pytorch3d.ops.knn_points: K smallest squared distances, ascending (differentiable)."""
    d2 = ((p1[:, :, None, :] - p2[:, None, :, :]) ** 2).sum(-1)
    idx = torch.argsort(d2, dim=-1, stable=True)[..., :K]
    return torch.gather(d2, -1, idx), idx, None


def make_sp_stage():
    """`calc_LBS_weight` (networks/sk_gs.py:751-774) + `warp` (:776-828) of the reference class, called unbound on a
    namespace that carries the attributes they read, the way `sp_stage` (:843-855) calls them; float64.  Gradients are
    taken w.r.t. the PRE-normalisation rotation (sp_stage normalises at :847), where lietorch's tangent-projected
    gradient and plain autograd agree."""
    sk = importlib.import_module('networks.sk_gs')
    sk.SE3, sk.SO3, sk.knn_points = _SE3Stub, _SO3Stub, _knn_points_stub
    cls = sk.SkeletonGaussianSplatting
    g = torch.Generator().manual_seed(20241017 + 400)
    out, n = {}, 0
    for mode in ('W', 'kernel', 'weighted_kernel', 'dist'):
        for method in ('LBS', 'LBS_c', 'largest'):
            for sep_rot in (True, False):
                P, M, K = 61, 13, 4
                dd = dict(dtype=torch.float64)
                points = torch.randn(P, 3, generator=g, **dd) * 0.5
                leaf = lambda *sh, scale=1.0: (torch.randn(*sh, generator=g, **dd) * scale).requires_grad_()  # noqa: E731
                sp_points, sp_t, raw_r = leaf(M, 3, scale=0.5), leaf(M, 3, scale=0.1), leaf(M, 4, scale=0.3)
                raw_g, sp_scale = leaf(M, 4, scale=0.3), leaf(M, 3, scale=0.05)
                sp_W, sp_radius, sp_weight = leaf(P, M), leaf(M, scale=0.3), leaf(M)
                bias = torch.tensor([0, 0, 0, 1.0], **dd)
                sp_r = torch.nn.functional.normalize(raw_r + bias, dim=-1)
                sp_rot = torch.nn.functional.normalize(raw_g + bias, dim=-1) if sep_rot else None
                self = types.SimpleNamespace(
                    num_knn=K, sk_is_init=True, training=True,
                    _sp_radius=sp_radius if 'kernel' in mode else None,
                    _sp_weight=sp_weight if mode == 'weighted_kernel' else None,
                    kernel_radius=torch.exp(sp_radius), kernel_weight=torch.sigmoid(sp_weight),
                    sp_W=sp_W if mode == 'W' else None)
                w, idx = cls.calc_LBS_weight(self, points, sp_points, None, None, temperature=1.0)
                if method == 'largest':  # :850-851
                    self.p2sp = torch.gather(idx, -1, w.argmax(dim=-1, keepdim=True))[:, 0]
                d_points, d_rotation, d_scales, spT = cls.warp(self, points, sp_points, sp_t, sp_r, sp_rot, sp_scale, w,
                                                               idx, method)
                cot = [torch.randn(t.shape, generator=g, **dd) for t in (d_points, d_rotation, d_scales, spT)]
                loss = sum((t * c).sum() for t, c in zip((d_points, d_rotation, d_scales, spT), cot))
                leaves = {'sp_points': sp_points, 'sp_t': sp_t, 'raw_r': raw_r, 'sp_scale': sp_scale}
                if sep_rot:
                    leaves['raw_g'] = raw_g
                if mode == 'W':
                    leaves['sp_W'] = sp_W
                if 'kernel' in mode:
                    leaves['sp_radius'] = sp_radius
                if mode == 'weighted_kernel':
                    leaves['sp_weight'] = sp_weight
                grads = torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)
                grads = [torch.zeros_like(l_) if g_ is None else g_ for g_, l_ in zip(grads, leaves.values())]
                pre = f'c{n}_'
                out[pre + 'cfg'] = np.array([mode, method, str(int(sep_rot))])
                for k_, v_ in dict(points=points, sp_points=sp_points, sp_t=sp_t, raw_r=raw_r, raw_g=raw_g,
                                   sp_scale=sp_scale, sp_W=sp_W, sp_radius=sp_radius, sp_weight=sp_weight, w=w, idx=idx,
                                   d_points=d_points, d_rotation=d_rotation, d_scales=d_scales, spT=spT,
                                   cot0=cot[0], cot1=cot[1], cot2=cot[2], cot3=cot[3]).items():
                    out[pre + k_] = v_.detach().numpy()
                for k_, v_ in zip(leaves, grads):
                    out[pre + 'grad_' + k_] = v_.numpy()
                n += 1
    np.savez_compressed(os.path.join(OUT, 'sp_stage.npz'), n=np.int64(n), K=np.int64(4), **out)


def make_densify():
    """The reference's densification bookkeeping (networks/gaussian_splatting.py:503-665) run UNMODIFIED on a small
    Gaussian set with a real torch.optim.Adam holding non-trivial moments.  The methods are borrowed from the class and
    bound to a light object that carries the attributes they touch (the full constructor needs the whole framework);
    torch.normal is replaced by `noise * std` with a recorded standard-normal table so that the split samples can be
    reproduced (:601-603)."""
    gsm = importlib.import_module('networks.gaussian_splatting')
    GS = gsm.GaussianSplatting

    class Shim:
        param_names_map = dict(GS.param_names_map)
        use_so3 = False
        scaling_activation_inverse = staticmethod(torch.log)
        points = property(lambda self: self._xyz)
        get_scaling = property(lambda self: torch.exp(self._scaling))
        get_opacity = property(lambda self: torch.sigmoid(self._opacity))

    for fn in ('add_densification_stats', 'prune_points', 'densification_postfix', 'densify_and_split',
               'densify_and_clone', 'densify', 'prune', 'reset_opacity'):
        setattr(Shim, fn, GS.__dict__[fn])
    Shim.change_optimizer = staticmethod(GS.__dict__['change_optimizer'].__func__)

    g = torch.Generator().manual_seed(20241017 + 500)
    out, n = {}, 0
    extent = 2.0
    cases = [dict(densify=True, prune=False, screen=None), dict(densify=True, prune=True, screen=20.0),
             dict(densify=False, prune=True, screen=20.0), dict(densify=True, prune=True, screen=None),
             dict(densify=False, prune=True, screen=None)]
    for case in cases:
        P = 400
        m = Shim()
        m._xyz = torch.nn.Parameter(torch.randn(P, 3, generator=g) * 0.5)
        m._features_dc = torch.nn.Parameter(torch.randn(P, 1, 3, generator=g))
        m._features_rest = torch.nn.Parameter(torch.randn(P, 15, 3, generator=g) * 0.1)
        m._scaling = torch.nn.Parameter(torch.randn(P, 3, generator=g) * 1.0 - 3.6)
        m._rotation = torch.nn.Parameter(torch.randn(P, 4, generator=g))
        m._opacity = torch.nn.Parameter(torch.randn(P, 1, generator=g) * 3.0 - 2.0)
        names = Shim.param_names_map
        opt = torch.optim.Adam([{'params': [getattr(m, k)], 'lr': 1e-3, 'name': v} for k, v in names.items()],
                               eps=1e-15)
        for k in names:  # one optimizer step so that the moments are not all zero
            getattr(m, k).grad = torch.randn(getattr(m, k).shape, generator=g)
        opt.step()
        opt.zero_grad(set_to_none=True)
        m.xyz_gradient_accum, m.denom, m.max_radii2D = torch.zeros(P, 1), torch.zeros(P, 1), torch.zeros(P)
        # three steps of statistics (:669-675)
        stat_in = []
        for _ in range(3):
            radii = torch.randint(0, 40, (P,), generator=g, dtype=torch.int32) * (torch.rand(P, generator=g) > 0.3)
            vs = torch.zeros(P, 3, requires_grad=True)
            vs.grad = torch.randn(P, 3, generator=g) * 4e-4
            mask = radii > 0
            m.max_radii2D[mask] = torch.max(m.max_radii2D[mask], radii[mask].float())
            m.add_densification_stats(vs, mask)
            stat_in.append((radii.clone(), vs.grad.clone()))
        pre = f'c{n}_'
        before = {v: getattr(m, k).detach().clone() for k, v in names.items()}
        state0 = {v: opt.state[getattr(m, k)] for k, v in names.items()}
        for v_, t_ in before.items():
            out[pre + 'in_' + v_] = t_.numpy()
            out[pre + 'in_m_' + v_] = state0[v_]['exp_avg'].clone().numpy()
            out[pre + 'in_v_' + v_] = state0[v_]['exp_avg_sq'].clone().numpy()
        for j, (r_, g_) in enumerate(stat_in):
            out[pre + f'radii{j}'], out[pre + f'vsgrad{j}'] = r_.numpy(), g_.numpy()
        out[pre + 'stat_accum'], out[pre + 'stat_denom'] = m.xyz_gradient_accum.clone().numpy(), m.denom.clone().numpy()
        out[pre + 'stat_radii'] = m.max_radii2D.clone().numpy()
        noise = torch.randn(2 * P, 3, generator=g)
        real_normal = torch.normal
        torch.normal = lambda mean=None, std=None, **kw: noise[:std.shape[0]] * std
        try:
            with torch.no_grad():
                if case['densify']:
                    m.densify(opt, max_grad=0.0002, extent=extent, densify_percent_dense=0.01)
                if case['prune']:
                    m.prune(opt, min_opacity=0.005, extent=extent, max_screen_size=case['screen'],
                            prune_percent_dense=0.1)
        finally:
            torch.normal = real_normal
        out[pre + 'noise'] = noise.numpy()
        out[pre + 'cfg'] = np.array([float(case['densify']), float(case['prune']), case['screen'] or 0.0, extent])
        for k, v_ in names.items():
            t_ = getattr(m, k)
            out[pre + 'out_' + v_] = t_.detach().numpy()
            out[pre + 'out_m_' + v_] = opt.state[t_]['exp_avg'].numpy()
            out[pre + 'out_v_' + v_] = opt.state[t_]['exp_avg_sq'].numpy()
        out[pre + 'out_accum'], out[pre + 'out_denom'] = m.xyz_gradient_accum.numpy(), m.denom.numpy()
        out[pre + 'out_radii'] = m.max_radii2D.numpy()
        with torch.no_grad():
            m.reset_opacity(opt)
        out[pre + 'reset_opacity'] = m._opacity.detach().numpy()
        out[pre + 'reset_m'] = opt.state[m._opacity]['exp_avg'].numpy()
        n += 1
    np.savez_compressed(os.path.join(OUT, 'densify.npz'), n=np.int64(n), **out)


def make_adam():
    """torch.optim.Adam trajectories with the reference's hyper-parameters (exps/default.yaml:121-125) and two param
    groups with different lr (networks/gaussian_splatting.py:447-452)."""
    g = torch.Generator().manual_seed(20241017 + 200)
    shapes = [(301, 3), (77, 16), (5000,)]
    lrs = [1.6e-4, 2.5e-3, 5e-2]
    params = [torch.nn.Parameter(torch.randn(*s, generator=g)) for s in shapes]
    opt = torch.optim.Adam([{'params': [p], 'lr': lr} for p, lr in zip(params, lrs)], lr=1e-3, eps=1e-15,
                           betas=(0.9, 0.999))
    out = {f'p0_{i}': p.detach().clone().numpy() for i, p in enumerate(params)}
    steps = 4
    for t in range(steps):
        for i, p in enumerate(params):
            scale = 10.0 ** (-2 * i)
            p.grad = torch.randn(*shapes[i], generator=g) * scale
            if t == 2 and i == 0:
                p.grad[::3] = 0  # rows without gradient still move (momentum)
            out[f'g{t}_{i}'] = p.grad.clone().numpy()
        opt.step()
        for i, p in enumerate(params):
            st = opt.state[p]
            out[f'p{t + 1}_{i}'] = p.detach().clone().numpy()
            out[f'm{t + 1}_{i}'] = st['exp_avg'].clone().numpy()
            out[f'v{t + 1}_{i}'] = st['exp_avg_sq'].clone().numpy()
    np.savez_compressed(os.path.join(OUT, 'adam.npz'), steps=np.int64(steps), n=np.int64(len(shapes)),
                        lrs=np.array(lrs), **out)


if __name__ == '__main__':
    main()
