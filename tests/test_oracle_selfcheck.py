"""Self-consistency of the oracle (no GPU): the hand-derived backward of oracle/raster_oracle.c (which restates the
reference's backward kernels) against float64 torch autograd of an INDEPENDENT dense restatement of the forward;
FK pointer-jumping vs the serial recurrence; LBS weight modes; stable-sort / range invariants of the binning stage."""
import numpy as np
import pytest
import torch

from oracle import fk_lbs as OF
from oracle import raster as OR
from sk_gs_b200 import scene as S
from skgs_test_util import np32, oracle_deform, oracle_settings, rel_err
from torch_raster_ref import render as torch_render


def _small_scene(P=70, W=64, H=48, seed=3):
    sc = S.make_scene('c1', P=P, seed=seed)
    cam = sc.cameras[0]
    cam.W, cam.H = W, H
    cam.tanfovy = cam.tanfovx * H / W
    sc.scaling = sc.scaling + 1.2  # larger splats so that the tiny image is well covered
    net, _, _ = oracle_deform(sc)
    return sc, cam, {k: v.detach() for k, v in net.items()}


@pytest.mark.parametrize('seed', [3, 4])
def test_backward_matches_float64_autograd(seed):
    sc, cam, net = _small_scene(seed=seed)
    s = oracle_settings(cam)
    img, g, b = OR.render_forward(s, np32(net['points']), np32(net['opacity']), np32(net['scales']),
                                  np32(net['rotations']), np32(net['sh_features']))
    leaves = {k: net[k].double().clone().requires_grad_(True) for k in ('points', 'scales', 'rotations', 'opacity',
                                                                         'sh_features')}
    color, depth, alpha, radius = torch_render(
        leaves['points'], leaves['scales'], leaves['rotations'], leaves['opacity'].reshape(-1), leaves['sh_features'],
        cam.viewmatrix.double(), cam.projmatrix.double(), cam.campos.double(), cam.W, cam.H, cam.tanfovx, cam.tanfovy,
        cam.bg.double())
    # forward agreement (different arithmetic: float64 + libm exp vs the specified fp32 path)
    assert np.array_equal(radius.numpy().astype(np.int32), g.radii)
    assert np.abs(color.detach().numpy() - img.color).max() <= 2e-5
    assert np.abs(alpha.detach().numpy() - img.alpha).max() <= 2e-5
    assert np.abs(depth.detach().numpy() - img.depth).max() <= 2e-4
    rng = np.random.default_rng(seed)
    dC = rng.standard_normal(img.color.shape).astype(np.float32)
    dD = rng.standard_normal(img.depth.shape).astype(np.float32)
    dA = rng.standard_normal(img.alpha.shape).astype(np.float32)
    loss = (color * torch.from_numpy(dC).double()).sum() + (depth * torch.from_numpy(dD).double()).sum() + \
        (alpha * torch.from_numpy(dA).double()).sum()
    loss.backward()
    og = OR.render_backward(s, g, b, img, dC, np32(net['points']), np32(net['scales']), np32(net['rotations']),
                            np32(net['sh_features']), dL_ddepth=dD, dL_dalpha=dA)
    for got, ref, tol in [(og.dL_dmeans3D, leaves['points'].grad, 2e-3), (og.dL_dscales, leaves['scales'].grad, 2e-3),
                          (og.dL_drotations, leaves['rotations'].grad, 2e-3),
                          (og.dL_dopacity, leaves['opacity'].grad.reshape(-1), 2e-3),
                          (og.dL_dsh, leaves['sh_features'].grad, 2e-3)]:
        assert rel_err(got, ref.numpy()) <= tol, rel_err(got, ref.numpy())


def test_binning_invariants():
    sc, cam, net = _small_scene(P=400, W=160, H=96, seed=5)
    s = oracle_settings(cam)
    img, g, b = OR.render_forward(s, np32(net['points']), np32(net['opacity']), np32(net['scales']),
                                  np32(net['rotations']), np32(net['sh_features']))
    assert b.R == int(g.tiles_touched.sum()) == len(b.keys)
    assert np.all(np.diff(b.keys.astype(np.uint64).astype(object)) >= 0)
    # the sort is a stable permutation of the emission order
    order = np.argsort(b.keys_unsorted, kind='stable')
    assert np.array_equal(b.keys_unsorted[order], b.keys) and np.array_equal(b.vals_unsorted[order], b.point_list)
    tiles = (b.keys >> np.uint64(32)).astype(np.int64)
    for t in np.unique(tiles):
        idx = np.nonzero(tiles == t)[0]
        assert b.ranges[t, 0] == idx[0] and b.ranges[t, 1] == idx[-1] + 1
    empty = np.setdiff1d(np.arange(b.ranges.shape[0]), np.unique(tiles))
    assert np.all(b.ranges[empty] == 0)
    # depth bits of positive floats sort like the floats
    d = g.depths[b.point_list]
    for t in np.unique(tiles)[:20]:
        r0, r1 = b.ranges[t]
        assert np.all(np.diff(d[r0:r1]) >= 0)


def test_fk_jump_equals_serial_and_lbs_modes():
    for name, P in (('c1', 500), ('c4', 300)):
        sc = S.make_scene(name, P=P, seed=12)
        local = OF.local_transforms(sc.joints.double(), sc.sk_r.double())
        a = OF.skeleton_warp_jump(local, sc.g_tr.double(), sc.parents.long(), sc.root)
        b = OF.skeleton_warp_serial(local, sc.g_tr.double(), sc.parents[:, 0].long(), sc.root)
        assert (a - b).abs().max().item() <= 1e-12
        for mode in ('W', 'kernel', 'weighted_kernel', 'dist'):
            w, idx = OF.lbs_weights(sc.xyz, sc.joints, sc.K, mode, sp_W=sc.sp_W, sp_radius=sc.sp_radius,
                                    sp_weight=sc.sp_weight)
            assert torch.allclose(w.sum(-1), torch.ones(P), atol=1e-5) and (w >= 0).all()
            d2 = ((sc.xyz[:, None] - sc.joints[None]) ** 2).sum(-1)
            assert torch.equal(idx, d2.topk(sc.K, largest=False, sorted=True).indices)


def test_empty_and_degenerate_inputs():
    sc, cam, net = _small_scene(P=8, seed=6)
    s = oracle_settings(cam)
    z = np.zeros
    img, g, b = OR.render_forward(s, z((0, 3), np.float32), z((0, 1), np.float32), z((0, 3), np.float32),
                                  z((0, 4), np.float32), z((0, 16, 3), np.float32))
    assert b.R == 0 and np.all(img.color == 1.0) and np.all(img.alpha == 0)
    # zero-size Gaussians still get the 0.3 px low-pass and a radius; opacity 0 never contributes
    img, g, b = OR.render_forward(s, np32(net['points']), np32(net['opacity']) * 0, np32(net['scales']) * 0,
                                  np32(net['rotations']), np32(net['sh_features']))
    assert np.all(g.radii[g.radii > 0] == 3) and np.all(img.alpha == 0)  # ceil(3*sqrt(0.3+sqrt(0.1)))
