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
