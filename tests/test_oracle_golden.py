"""The CPU oracle against golden vectors produced by the REFERENCE's own pure-torch helpers
(tests/golden/make_golden.py imports them unmodified from /root/reference).  Runs without a GPU."""
import os

import numpy as np
import torch

from oracle import fk_lbs as OF
from oracle import raster as OR
from sk_gs_b200 import scene as S

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def test_sh_colour_matches_reference_eval_sh():
    """reference networks/encoders/sphere_harmonics.py:131-185 (eval_sh) + 0.5, clamp >= 0."""
    d = np.load(os.path.join(G, 'sh.npz'))
    shs = np.ascontiguousarray(d['sh'].transpose(0, 2, 1))  # [P,3,16] -> rasterizer layout [P,16,3]
    for deg in range(4):
        rgb, cl = OR.sh_to_rgb(deg, shs, d['dirs'] * 3.7)  # the oracle normalises the direction itself
        ref = d[f'rgb_deg{deg}'] + 0.5
        assert np.abs(rgb - np.maximum(ref, 0)).max() <= 2e-6
        assert np.array_equal(cl.astype(bool)[np.abs(ref) > 1e-5], (ref < 0)[np.abs(ref) > 1e-5])
    # RGB2SH convention used by the synthetic scenes (sphere_harmonics.py:188-189)
    assert np.abs(d['rgb2sh'] - (d['rgb'] - 0.5) / 0.28209479177387814).max() <= 1e-6


def test_cov3d_matches_reference_gs_utils():
    """reference networks/GS_utils.py:44-82: Sigma = (R S)(R S)^T with R from the normalised xyzw quaternion."""
    d = np.load(os.path.join(G, 'cov.npz'))
    qn = d['q'] / np.linalg.norm(d['q'], axis=1, keepdims=True)  # the kernels do not normalise (colmap.cu:129)
    cov6 = OR.cov3d(d['s'], qn, quat_wxyz=False)
    assert np.abs(cov6 - d['cov6']).max() <= 1e-6 * max(1.0, np.abs(d['cov6']).max())
    wxyz = np.ascontiguousarray(qn[:, [3, 0, 1, 2]])
    assert np.array_equal(OR.cov3d(d['s'], wxyz, quat_wxyz=True), cov6)
    # NB: the reference's python `compute_cov2D` (GS_utils.py:102-125) evaluates (W J)^T V (W J), the row-major
    # "OpenGL" variant of gaussian_preprocess.cu, not the colmap/upstream EWA form (J W) V (J W)^T that this path
    # implements (gaussian_preprocess_colmap.cu:99-109); it is therefore not used as a pin.  The EWA form is checked
    # against an independent float64 autograd restatement in tests/test_oracle_selfcheck.py.
    assert 'cov2d' in d


def _to_mat(T7):
    q, t = T7[:, 3:], T7[:, :3]
    x, y, z, w = q.unbind(-1)
    R = torch.stack([1 - 2 * y * y - 2 * z * z, 2 * x * y - 2 * w * z, 2 * w * y + 2 * x * z,
                     2 * x * y + 2 * w * z, 1 - 2 * x * x - 2 * z * z, 2 * y * z - 2 * w * x,
                     2 * x * z - 2 * w * y, 2 * w * x + 2 * y * z, 1 - 2 * x * x - 2 * y * y], -1).view(-1, 3, 3)
    M = torch.eye(4, dtype=T7.dtype).repeat(T7.shape[0], 1, 1)
    M[:, :3, :3] = R
    M[:, :3, 3] = t
    return M


def test_fk_matches_reference_skeleton_warp_and_find_root():
    """reference networks/sk_gs.py:50-103 (find_root), :167-190 (skeleton_warp_v0 / skeleton_warp, 4x4 matrices)."""
    d = np.load(os.path.join(G, 'fk.npz'))
    for i in range(int(d['n'])):
        father = torch.from_numpy(d[f'father{i}'])
        parents, depth, root = OF.build_tree(father)
        assert root == int(d[f'root{i}'])
        assert np.array_equal(parents.numpy(), d[f'parents{i}'])
        assert np.array_equal(depth.numpy(), d[f'depth{i}'])
        p2, d2, r2 = S.find_root_table([int(v) for v in father])
        assert r2 == root and np.array_equal(p2.numpy(), parents.numpy()) and np.array_equal(d2.numpy(), depth.numpy())
        joints, r, g_tr = (torch.from_numpy(d[f'{k}{i}']) for k in ('joints', 'r', 'g_tr'))
        local = OF.local_transforms(joints, r)
        for T7 in (OF.skeleton_warp_jump(local, g_tr, parents, root),
                   OF.skeleton_warp_serial(local, g_tr, parents[:, 0], root)):
            M = _to_mat(T7)
            assert np.abs(M.numpy() - d[f'T_jump{i}']).max() <= 1e-12
            assert np.abs(M.numpy() - d[f'T_v0{i}']).max() <= 1e-12


def _sp_case(d, i, dtype=torch.float64, device='cpu'):
    """Inputs of case i of sp_stage.npz as leaves + the keyword arguments of oracle.fk_lbs.sp_stage / sp_lbs.sp_warp."""
    mode, method, sep = (str(v) for v in d[f'c{i}_cfg'])
    t = {k: torch.from_numpy(d[f'c{i}_{k}']).to(device=device, dtype=dtype) for k in
         ('points', 'sp_points', 'sp_t', 'raw_r', 'raw_g', 'sp_scale', 'sp_W', 'sp_radius', 'sp_weight')}
    leaves = {k: t[k].clone().requires_grad_() for k in t if k != 'points'}
    bias = torch.tensor([0, 0, 0, 1.0], dtype=dtype, device=device)
    sp_r = torch.nn.functional.normalize(leaves['raw_r'] + bias, dim=-1)          # networks/sk_gs.py:847
    sp_rot = torch.nn.functional.normalize(leaves['raw_g'] + bias, dim=-1) if sep == '1' else None  # :848
    kw = dict(K=int(d['K']), mode=mode, method=method, sp_W=leaves['sp_W'] if mode == 'W' else None,
              sp_radius=leaves['sp_radius'] if 'kernel' in mode else None,
              sp_weight=leaves['sp_weight'] if mode == 'weighted_kernel' else None, temperature=1.0)
    return t['points'], leaves, sp_r, sp_rot, kw


def test_sp_stage_matches_reference_warp_and_calc_lbs_weight():
    """reference networks/sk_gs.py:751-774 (calc_LBS_weight) + :776-828 (warp), all 4 weight modes x 3 warp methods x
    sep_rot on/off: outputs and the gradients of every input (rotation: w.r.t. the pre-normalisation vector)."""
    d = np.load(os.path.join(G, 'sp_stage.npz'))
    for i in range(int(d['n'])):
        points, leaves, sp_r, sp_rot, kw = _sp_case(d, i)
        outs = OF.sp_stage(points, leaves['sp_points'], leaves['sp_t'], sp_r, sp_rot, leaves['sp_scale'], **kw)
        d_points, d_rotation, d_scales, spT, w, idx = outs
        assert np.array_equal(idx.numpy(), d[f'c{i}_idx'])
        for name, got in (('w', w), ('d_points', d_points), ('d_rotation', d_rotation), ('d_scales', d_scales),
                          ('spT', spT)):
            assert np.abs(got.detach().numpy() - d[f'c{i}_{name}']).max() <= 1e-12, (i, name)
        loss = sum((o * torch.from_numpy(d[f'c{i}_cot{j}'])).sum() for j, o in
                   enumerate((d_points, d_rotation, d_scales, spT)))
        names = [k[len(f'c{i}_grad_'):] for k in d.files if k.startswith(f'c{i}_grad_')]
        grads = torch.autograd.grad(loss, [leaves[n] for n in names], allow_unused=True)
        for n, g in zip(names, grads):
            ref = d[f'c{i}_grad_{n}']
            got = np.zeros_like(ref) if g is None else g.numpy()
            assert np.abs(got - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max()), (i, n)


DENSIFY_NAMES = ('xyz', 'f_dc', 'f_rest', 'scaling', 'rotation', 'opacity')


def _densify_case(d, i):
    """Inputs of case i of densify.npz: ({name: (param, exp_avg, exp_avg_sq)}, stats, config keywords, noise)."""
    pre = f'c{i}_'
    t = {n: tuple(torch.from_numpy(d[pre + k + n]) for k in ('in_', 'in_m_', 'in_v_')) for n in DENSIFY_NAMES}
    dens, prune, screen, extent = (float(v) for v in d[pre + 'cfg'])
    kw = dict(do_densify=bool(dens), do_prune=bool(prune), grad_threshold=0.0002, densify_extent=0.01 * extent,
              min_opacity=0.005, max_screen_size=screen, prune_extent=0.1 * extent)
    return t, kw, torch.from_numpy(d[pre + 'noise'])


def test_densify_matches_reference_adaptive_control():
    """reference networks/gaussian_splatting.py:503-513 (statistics), :640-645 densify = clone :624-638 + split :589-622,
    :653-660 prune, :662-665 reset_opacity, with change_optimizer (:515-563) acting on a real torch.optim.Adam."""
    from oracle import densify as OD
    d = np.load(os.path.join(G, 'densify.npz'))
    for i in range(int(d['n'])):
        pre = f'c{i}_'
        t, kw, noise = _densify_case(d, i)
        P = t['xyz'][0].shape[0]
        accum, denom, radii = torch.zeros(P), torch.zeros(P), torch.zeros(P)
        for j in range(3):
            OD.add_densification_stats(accum, denom, radii, torch.from_numpy(d[pre + f'radii{j}']),
                                       torch.from_numpy(d[pre + f'vsgrad{j}']))
        assert np.array_equal(accum.numpy(), d[pre + 'stat_accum'][:, 0])
        assert np.array_equal(denom.numpy(), d[pre + 'stat_denom'][:, 0])
        assert np.array_equal(radii.numpy(), d[pre + 'stat_radii'])
        out, accum, denom, radii = OD.densify_and_prune(t, accum, denom, radii, noise=noise, **kw)
        for n in DENSIFY_NAMES:
            for k, j in (('out_', 0), ('out_m_', 1), ('out_v_', 2)):
                ref = d[pre + k + n]
                assert out[n][j].shape == ref.shape, (i, n, k)
                assert np.abs(out[n][j].numpy() - ref).max() <= 1e-6, (i, n, k)
        assert np.array_equal(accum.numpy(), d[pre + 'out_accum'].reshape(-1))
        assert np.array_equal(radii.numpy(), d[pre + 'out_radii'])
        assert np.abs(OD.reset_opacity(out['opacity'][0]).numpy() - d[pre + 'reset_opacity']).max() <= 1e-6
        assert np.abs(d[pre + 'reset_m']).max() == 0


def test_camera_matches_reference_perspective():
    """reference my_ext/ops_3d/coord_trans_opencv.py:203-239."""
    d = np.load(os.path.join(G, 'cam.npz'))
    assert np.abs(S.perspective_opencv(0.6911, 800, 800).numpy() - d['Tv2c_800']).max() <= 1e-6
    assert np.abs(S.perspective_opencv(0.5, 1920, 1080).numpy() - d['Tv2c_1080p']).max() <= 1e-6


def test_exp_and_msb_specification():
    """orc_exp is the fully specified exp of the compositing stage: <= 8 ulp(1) from exp on [-6, 0]; getHigherMsb as the
    reference (gaussian_rasterizer_forward.cu:30-42)."""
    rng = np.random.default_rng(0)
    for lo, bound in ((6.0, 8 * 2.0 ** -24), (20.0, 2e-6)):  # alpha >= 1/255 needs power >= -5.54
        x = -rng.random(4000).astype(np.float32) * np.float32(lo)
        e = OR.orc_exp(x)
        ref = np.exp(x.astype(np.float64))
        assert np.max(np.abs(e - ref) / ref) <= bound
    assert float(OR.orc_exp(np.float32(0.0))) == 1.0
    L = OR.lib()
    for n, want in [(1, 1), (2, 2), (625, 10), (1024, 11), (2500, 12), (4096, 13), (8160, 13), (65536, 17)]:
        assert L.orc_higher_msb(n) == want
