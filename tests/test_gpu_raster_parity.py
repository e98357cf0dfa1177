"""GPU parity of the rasterizer against the CPU oracle, stage by stage, through the C ABI.

Bars (BASELINE.json north_star): radii, tile counts, keys (emission order and sorted), point_list and tile ranges
BIT-EXACT; image <= 1e-5 abs (here it is bit-exact by construction, see oracle/raster_oracle.c header); gradients
<= 1e-4 relative to the largest magnitude of each tensor (fp32 atomics vs fp64 oracle accumulation).
"""
import numpy as np
import pytest
import torch

from oracle import raster as OR
from sk_gs_b200 import diff_gaussian_rasterization as DGR
from sk_gs_b200 import scene as S
from sk_gs_b200.pipeline import raster_settings_for
from skgs_test_util import arena_view, np32, oracle_deform, oracle_settings, rel_err

pytestmark = pytest.mark.gpu

IMG_ATOL = 1e-5
GRAD_RTOL = 1e-4


def _inputs(name, P=None, views=1, seed=None):
    sc = S.make_scene(name, P=P, views=views, seed=seed)
    net, _, _ = oracle_deform(sc)
    return sc, {k: v.detach() for k, v in net.items()}


def _gpu_forward(sc, net, view=0, debug_flags=0, sh_degree=3):
    dev = torch.device('cuda:0')
    rs = raster_settings_for(sc.cameras[view], dev, sh_degree)
    color, depth, alpha, radii, st = DGR.rasterize_forward(
        rs, net['points'].to(dev), net['opacity'].to(dev), shs=net['sh_features'].to(dev), scales=net['scales'].to(dev),
        rotations=net['rotations'].to(dev), quat_wxyz=False, debug_flags=debug_flags)
    torch.cuda.synchronize()
    return color, depth, alpha, radii, st


def _oracle_forward(sc, net, view=0, sh_degree=3):
    s = oracle_settings(sc.cameras[view], sh_degree)
    img, g, b = OR.render_forward(s, np32(net['points']), np32(net['opacity']), np32(net['scales']),
                                  np32(net['rotations']), np32(net['sh_features']))
    return s, img, g, b


def _geom_arrays(st):
    lay, P = st.layout, st.P
    f32, i32 = torch.float32, torch.int32
    return dict(
        means2D=arena_view(st.geom, lay.means2D, f32, 2 * P).view(P, 2).cpu().numpy(),
        depths=arena_view(st.geom, lay.depths, f32, P).cpu().numpy(),
        cov3D=arena_view(st.geom, lay.cov3D, f32, 6 * P).view(P, 6).cpu().numpy(),
        conic_opacity=arena_view(st.geom, lay.conic_opacity, f32, 4 * P).view(P, 4).cpu().numpy(),
        rgbd=arena_view(st.geom, lay.rgbd, f32, 4 * P).view(P, 4).cpu().numpy(),
        clamped=arena_view(st.geom, lay.clamped, torch.uint8, P).cpu().numpy(),
        tiles_touched=arena_view(st.geom, lay.tiles_touched, i32, P).cpu().numpy().astype(np.uint32),
        point_offsets=arena_view(st.geom, lay.point_offsets, i32, P).cpu().numpy().astype(np.uint32),
        header=arena_view(st.geom, lay.header, i32, 4).cpu().numpy().astype(np.uint32),
    )


CASES = [('c1', 2000, 11), ('c1', None, None), ('c2', 20000, 5), ('c2', None, None)]


@pytest.mark.parametrize('name,P,seed', CASES)
def test_preprocess_bit_exact(name, P, seed):
    sc, net = _inputs(name, P, seed=seed)
    _, _, _, radii, st = _gpu_forward(sc, net)
    _, _, g, b = _oracle_forward(sc, net)
    ga = _geom_arrays(st)
    radii = radii.cpu().numpy()
    assert np.array_equal(radii, g.radii)
    assert np.array_equal(ga['tiles_touched'], g.tiles_touched)
    vis = g.radii > 0
    assert vis.sum() > 0.9 * len(vis)
    # bit-exact comparison of every float the later stages consume
    for a, o in [(ga['means2D'], g.means2D), (ga['depths'], g.depths), (ga['cov3D'], g.cov3D),
                 (ga['conic_opacity'], g.conic_opacity), (ga['rgbd'][:, :3], g.rgb), (ga['rgbd'][:, 3], g.depths)]:
        assert np.array_equal(a[vis].view(np.uint32), o[vis].view(np.uint32))
    cl = np.stack([(ga['clamped'] >> c) & 1 for c in range(3)], 1).astype(np.uint8)
    assert np.array_equal(cl[vis], g.clamped[vis])
    assert np.array_equal(ga['point_offsets'], b.offsets)
    assert int(ga['header'][0]) == b.R and int(ga['header'][1]) == int(vis.sum()) and int(ga['header'][3]) == 0
    assert st.num_rendered == b.R


def _sorted(st):
    keys, plist = st.sorted_lists()
    return keys.cpu().numpy().view(np.uint64), plist.cpu().numpy().view(np.uint32)


@pytest.mark.parametrize('path', ['fused-emission', 'split'])
@pytest.mark.parametrize('name,P,seed', CASES)
def test_binning_bit_exact(name, P, seed, path):
    """Sorted keys, point_list and tile ranges against the oracle's stable sort, through both emission paths: keys
    emitted by the preprocess kernel (capacity known from the previous call of the shape) and by duplicate_keys_kernel
    from stored geometry (first call of a shape: R is read between the stages)."""
    sc, net = _inputs(name, P, seed=seed)
    _, _, g, b = _oracle_forward(sc, net)
    key = (0, net['points'].shape[0], sc.cameras[0].W, sc.cameras[0].H)
    if path == 'split':
        DGR._capacity.last.pop(key, None)
    else:
        DGR._capacity.put(key, b.R)
    color, _, _, _, st = _gpu_forward(sc, net)
    lay, R = st.layout, b.R
    keys, plist = _sorted(st)
    h = st.header()
    assert int(h.num_rendered) == R and int(h.overflow) == 0
    tiles = b.ranges.shape[0]
    ranges = arena_view(st.img, lay.ranges, torch.int32, 2 * tiles).view(tiles, 2).cpu().numpy().view(np.uint32)
    assert np.array_equal(keys, b.keys)
    assert np.array_equal(plist, b.point_list)
    assert np.array_equal(ranges, b.ranges)
    # both binning paths feed the same compositing kernel: identical image bits
    _, img, _, _ = _oracle_forward(sc, net)
    assert np.array_equal(color.cpu().numpy().view(np.uint32), img.color.view(np.uint32))


@pytest.mark.parametrize('P,longest', [(12000, 8192), (60000, 3 * 16384)])
def test_sort_long_tiles_and_depth_ties(P, longest):
    """Very long tile lists and duplicated Gaussians: exactly equal depths must come out in ascending Gaussian id
    (= emission order of the stable reference sort).  > 8192 entries: the 16384-entry shared-memory configuration of the
    per-tile sort; > 3 x 16384: sorted chunks + merge passes."""
    sc = S.make_scene('c1', P=P, seed=17)
    cam = sc.cameras[0]
    cam.W, cam.H = 48, 40
    cam.tanfovy = cam.tanfovx * cam.H / cam.W
    sc.scaling = sc.scaling + 2.0
    sc.xyz[P // 2:] = sc.xyz[:P // 2]  # exact duplicates (same skinning weights too) -> exact depth ties
    sc.sp_W[P // 2:] = sc.sp_W[:P // 2]
    net, _, _ = oracle_deform(sc)
    net = {k: v.detach() for k, v in net.items()}
    _, img, g, b = _oracle_forward(sc, net)
    assert (b.ranges[:, 1] - b.ranges[:, 0]).max() > longest
    d = g.depths[b.point_list]
    assert (np.diff(d) == 0).sum() > 1000
    for _ in range(2):  # first call of the shape: split path; second: keys emitted by the preprocess kernel
        color, _, _, _, st = _gpu_forward(sc, net)
        keys, plist = _sorted(st)
        assert np.array_equal(plist, b.point_list) and np.array_equal(keys, b.keys)
        assert np.array_equal(color.cpu().numpy().view(np.uint32), img.color.view(np.uint32))


def test_large_tile_grid_paths():
    """2560 x 1440 = 160 x 90 = 14 400 tiles: more than the shared-memory capacity of the plan kernel (the difference
    grid is summed in place in global memory) and of the scatter kernel's per-chunk tile counters (one atomic per
    entry): keys, point_list, ranges and the image are still bit-identical to the oracle."""
    sc = S.make_scene('c1', P=20000, seed=23)
    cam = sc.cameras[0]
    cam.W, cam.H = 2560, 1440
    cam.tanfovy = cam.tanfovx * cam.H / cam.W
    net, _, _ = oracle_deform(sc)
    net = {k: v.detach() for k, v in net.items()}
    _, img, g, b = _oracle_forward(sc, net)
    assert b.ranges.shape[0] == 160 * 90
    for _ in range(2):  # split path, then keys emitted by the preprocess kernel
        color, _, _, _, st = _gpu_forward(sc, net)
        keys, plist = _sorted(st)
        lay = st.layout
        tiles = b.ranges.shape[0]
        ranges = arena_view(st.img, lay.ranges, torch.int32, 2 * tiles).view(tiles, 2).cpu().numpy().view(np.uint32)
        assert np.array_equal(keys, b.keys) and np.array_equal(plist, b.point_list)
        assert np.array_equal(ranges, b.ranges)
        assert np.array_equal(color.cpu().numpy().view(np.uint32), img.color.view(np.uint32))


@pytest.mark.parametrize('name,P,seed', CASES)
def test_composite_forward(name, P, seed):
    sc, net = _inputs(name, P, seed=seed)
    color, depth, alpha, radii, st = _gpu_forward(sc, net)
    _, img, g, b = _oracle_forward(sc, net)
    H, W = img.alpha.shape
    lay = st.layout
    n_contrib = arena_view(st.img, lay.n_contrib, torch.int32, H * W).view(H, W).cpu().numpy().view(np.uint32)
    final_T = arena_view(st.img, lay.final_T, torch.float32, H * W).view(H, W).cpu().numpy()
    assert np.array_equal(n_contrib, img.n_contrib)
    assert np.array_equal(final_T.view(np.uint32), img.final_T.view(np.uint32))
    c, d, a = color.cpu().numpy(), depth.cpu().numpy()[0], alpha.cpu().numpy()[0]
    assert np.abs(c - img.color).max() <= IMG_ATOL
    assert np.abs(a - img.alpha).max() <= IMG_ATOL
    assert np.abs(d - img.depth).max() <= IMG_ATOL * max(1.0, float(img.depth.max()))
    # stronger than required: the arithmetic is fully specified, so the image is reproduced bit for bit
    assert np.array_equal(c.view(np.uint32), img.color.view(np.uint32))
    assert np.array_equal(d.view(np.uint32), img.depth.view(np.uint32))
    assert img.alpha.max() > 0.5  # the scene actually covers pixels


@pytest.mark.parametrize('name,P,seed,with_aux', [('c1', 2000, 11, True), ('c1', None, None, False),
                                                  ('c2', 20000, 5, True), ('c2', None, None, False)])
def test_backward(name, P, seed, with_aux):
    sc, net = _inputs(name, P, seed=seed)
    color, depth, alpha, radii, st = _gpu_forward(sc, net)
    s, img, g, b = _oracle_forward(sc, net)
    H, W = img.alpha.shape
    rng = np.random.default_rng(3)
    dC = (rng.standard_normal((3, H, W)) / (3 * H * W)).astype(np.float32)
    dD = (rng.standard_normal((H, W)) / (H * W)).astype(np.float32) if with_aux else None
    dA = (rng.standard_normal((H, W)) / (H * W)).astype(np.float32) if with_aux else None
    og = OR.render_backward(s, g, b, img, dC, np32(net['points']), np32(net['scales']), np32(net['rotations']),
                            np32(net['sh_features']), dL_ddepth=dD, dL_dalpha=dA)
    dev = color.device
    gg = DGR.rasterize_backward(st, torch.from_numpy(dC).to(dev),
                                None if dD is None else torch.from_numpy(dD).to(dev)[None],
                                None if dA is None else torch.from_numpy(dA).to(dev)[None])
    torch.cuda.synchronize()
    pairs = [('means3D', og.dL_dmeans3D), ('means2D', og.dL_dmeans2D), ('shs', og.dL_dsh), ('opacities', og.dL_dopacity),
             ('scales', og.dL_dscales), ('rotations', og.dL_drotations)]
    for k, ref in pairs:
        got = gg[k].cpu().numpy().reshape(ref.shape)
        assert np.isfinite(got).all(), k
        assert np.abs(ref).max() > 0, k
        assert rel_err(got, ref) <= GRAD_RTOL, (k, rel_err(got, ref))
    vis = g.radii > 0
    assert np.all(gg['means3D'].cpu().numpy()[~vis] == 0)


def test_precomputed_colors_and_cov3D():
    """colors_precomp / cov3D_precomp inputs (exactly-one-of rule, networks/renderer/gaussian_render.py:250-255)."""
    sc, net = _inputs('c1', 3000, seed=21)
    s = oracle_settings(sc.cameras[0])
    g0 = OR.preprocess_fwd(s, np32(net['points']), np32(net['opacity']), np32(net['scales']), np32(net['rotations']),
                           np32(net['sh_features']))
    colors, cov = g0.rgb.copy(), g0.cov3D.copy()
    img, g, b = OR.render_forward(s, np32(net['points']), np32(net['opacity']), colors_precomp=colors,
                                  cov3D_precomp=cov)
    dev = torch.device('cuda:0')
    rs = raster_settings_for(sc.cameras[0], dev)
    color, depth, alpha, radii, st = DGR.rasterize_forward(rs, net['points'].to(dev), net['opacity'].to(dev),
                                                           colors_precomp=torch.from_numpy(colors).to(dev),
                                                           cov3D_precomp=torch.from_numpy(cov).to(dev))
    assert np.array_equal(radii.cpu().numpy(), g.radii)
    assert np.array_equal(color.cpu().numpy().view(np.uint32), img.color.view(np.uint32))
    H, W = img.alpha.shape
    dC = (np.random.default_rng(1).standard_normal((3, H, W)) / (3 * H * W)).astype(np.float32)
    og = OR.render_backward(s, g, b, img, dC, np32(net['points']), cov3D_precomp=cov)
    gg = DGR.rasterize_backward(st, torch.from_numpy(dC).to(dev))
    assert rel_err(gg['colors_precomp'].cpu().numpy(), og.dL_dcolors) <= GRAD_RTOL
    assert rel_err(gg['cov3D_precomp'].cpu().numpy(), og.dL_dcov3D) <= GRAD_RTOL
    assert rel_err(gg['means3D'].cpu().numpy(), og.dL_dmeans3D) <= GRAD_RTOL
    with pytest.raises(RuntimeError):
        DGR.rasterize_forward(rs, net['points'].to(dev), net['opacity'].to(dev))  # neither shs nor colors


def test_edge_cases_empty_and_culled():
    dev = torch.device('cuda:0')
    sc = S.make_scene('c1', P=16, seed=2)
    rs = raster_settings_for(sc.cameras[0], dev)
    # P == 0 (gaussian_rasterizer_forward.cu:298-313): background image, empty radii
    z = torch.zeros
    color, depth, alpha, radii, st = DGR.rasterize_forward(rs, z(0, 3, device=dev), z(0, 1, device=dev),
                                                           shs=z(0, 16, 3, device=dev), scales=z(0, 3, device=dev),
                                                           rotations=z(0, 4, device=dev))
    assert radii.numel() == 0 and torch.all(color == 1.0) and torch.all(alpha == 0)
    # everything behind the camera -> R == 0 (:236)
    net, _, _ = oracle_deform(sc)
    pts = net['points'].detach().to(dev) * 0 + sc.cameras[0].campos.to(dev) * 2.0
    color, depth, alpha, radii, st = DGR.rasterize_forward(rs, pts, net['opacity'].detach().to(dev),
                                                           shs=net['sh_features'].detach().to(dev),
                                                           scales=net['scales'].detach().to(dev),
                                                           rotations=net['rotations'].detach().to(dev),
                                                           quat_wxyz=False)
    assert st.num_rendered == 0 and torch.all(radii == 0) and torch.all(color == 1.0)
    g = DGR.rasterize_backward(st, torch.ones_like(color))
    assert all(torch.all(v == 0) for v in g.values() if v is not None)


def test_sh_degrees_and_ragged_image():
    """active SH degree < 3 uses the first (D+1)^2 coefficients; image size not a multiple of 16."""
    sc = S.make_scene('c1', P=1500, seed=4)
    for cam in sc.cameras:
        cam.W, cam.H = 203, 117
        cam.tanfovy = cam.tanfovx * cam.H / cam.W
    net, _, _ = oracle_deform(sc)
    net = {k: v.detach() for k, v in net.items()}
    for D in (0, 1, 2):
        color, depth, alpha, radii, st = _gpu_forward(sc, net, sh_degree=D)
        _, img, g, b = _oracle_forward(sc, net, sh_degree=D)
        assert np.array_equal(radii.cpu().numpy(), g.radii)
        assert np.array_equal(color.cpu().numpy().view(np.uint32), img.color.view(np.uint32))


def test_capacity_overflow_recovers():
    """An under-estimated binning arena is detected (header.overflow) and the render stage is re-run."""
    sc, net = _inputs('c1', 4000, seed=8)
    dev = torch.device('cuda:0')
    key = (0, 4000, sc.cameras[0].W, sc.cameras[0].H)
    DGR._capacity.put(key, 10)  # absurdly small estimate
    color, depth, alpha, radii, st = _gpu_forward(sc, net)
    _, img, g, b = _oracle_forward(sc, net)
    assert st.num_rendered == b.R and st.R_cap >= b.R
    assert np.array_equal(color.cpu().numpy().view(np.uint32), img.color.view(np.uint32))
