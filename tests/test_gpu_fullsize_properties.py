"""Full BASELINE.json sizes on the GPU, checked through size-independent properties (the oracle would take too long):
sortedness and consistency of the binning products, determinism, exact homogeneity of the backward in the upstream
gradient, range checks - plus the B2 (in-tree API) entry points against B1."""
import pytest
import torch

from sk_gs_b200 import diff_gaussian_rasterization as DGR
from sk_gs_b200 import scene as S
from sk_gs_b200.pipeline import HotPath
from sk_gs_b200.renderer import rasterize_gaussians_b2, rasterize_gaussians_backward_b2
from skgs_test_util import arena_view

pytestmark = pytest.mark.gpu


def _forward(hp, view=0, flags=1):
    with torch.no_grad():
        net, _ = hp.deform()
        return net, DGR.rasterize_forward(hp.settings[view], net['points'], net['opacity'], shs=net['sh_features'],
                                          scales=net['scales'], rotations=net['rotations'], quat_wxyz=False,
                                          debug_flags=flags)


@pytest.mark.parametrize('name', ['c2', 'c4'])
def test_binning_products_are_consistent_at_full_size(name):
    cfg = S.CONFIGS[name]
    hp = HotPath(S.make_scene(cfg, views=1), 'cuda:0', requires_grad=False)
    net, (color, depth, alpha, radii, st) = _forward(hp)
    torch.cuda.synchronize()
    lay, R, P = st.layout, st.num_rendered, cfg.P
    tiles = ((cfg.W + 15) // 16) * ((cfg.H + 15) // 16)
    keys, plist = st.sorted_lists()
    assert keys.numel() == R
    plist = plist.long()
    touched = arena_view(st.geom, lay.tiles_touched, torch.int32, P).long()
    ranges = arena_view(st.img, lay.ranges, torch.int32, 2 * tiles).view(tiles, 2).long()
    depths = arena_view(st.geom, lay.depths, torch.float32, P)
    assert R == int(touched.sum()) and R > 4 * P
    assert bool((keys[1:] >= keys[:-1]).all())                           # sorted (positive 44-bit keys)
    assert torch.equal(torch.bincount(plist, minlength=P), touched)       # every Gaussian appears tiles_touched times
    assert bool(((radii > 0) == (touched > 0)).all())
    tile_of = keys >> 32
    assert bool((tile_of < tiles).all())
    cnt = torch.bincount(tile_of, minlength=tiles)
    assert torch.equal(ranges[:, 1] - ranges[:, 0], cnt)                  # ranges partition the list
    nz = cnt > 0
    assert torch.equal(ranges[nz, 0], (torch.cumsum(cnt, 0) - cnt)[nz])
    assert torch.equal((keys & 0xffffffff).int().view(torch.float32), depths[plist])  # key payload = depth bits of its Gaussian
    n_contrib = arena_view(st.img, lay.n_contrib, torch.int32, cfg.H * cfg.W).view(cfg.H, cfg.W).long()
    per_pixel_len = cnt.view((cfg.H + 15) // 16, (cfg.W + 15) // 16).repeat_interleave(16, 0).repeat_interleave(16, 1)
    assert bool((n_contrib <= per_pixel_len[:cfg.H, :cfg.W]).all())
    assert bool(torch.isfinite(color).all()) and float(alpha.min()) >= 0 and float(alpha.max()) < 1
    assert float(alpha.mean()) > 0.1


def test_determinism_and_homogeneity_c2():
    cfg = S.CONFIGS['c2']
    hp = HotPath(S.make_scene(cfg, views=1), 'cuda:0', requires_grad=False)
    net, (c1, d1, a1, r1, st1) = _forward(hp, flags=0)
    net, (c2, d2, a2, r2, st2) = _forward(hp, flags=0)
    assert torch.equal(c1, c2) and torch.equal(d1, d2) and torch.equal(r1, r2)   # forward is bit-deterministic
    dL = torch.randn(3, cfg.H, cfg.W, device='cuda') / (3 * cfg.H * cfg.W)
    g1 = DGR.rasterize_backward(st1, dL)
    g2 = DGR.rasterize_backward(st2, 2.0 * dL)                                 # scaling by 2 is exact in fp32
    for k in ('means3D', 'shs', 'scales', 'rotations', 'opacities', 'means2D'):
        a, b = g1[k], g2[k]
        scale = float(a.abs().max())
        assert scale > 0 and float((2 * a - b).abs().max()) <= 2e-5 * scale, k   # only the atomics' order differs
    # a second backward on the SAME state: the kernels restore their own invariants (accumulators re-zeroed by
    # preprocess_bwd, work ticket reset), no memset in between
    g3 = DGR.rasterize_backward(st1, dL)
    for k in ('means3D', 'shs', 'scales', 'rotations', 'opacities', 'means2D'):
        assert float((g1[k] - g3[k]).abs().max()) <= 2e-5 * float(g1[k].abs().max()), k


def test_forward_only_3m_gaussians_1080p():
    """config 5 shape: 3M Gaussians, 64 joints, 1920x1080, forward only."""
    cfg = S.CONFIGS['c5']
    hp = HotPath(S.make_scene(cfg, views=1), 'cuda:0', requires_grad=False)
    with torch.no_grad():
        out = hp.render(0)
    torch.cuda.synchronize()
    w = DGR.last_header_words('cuda:0')
    assert int(w[3]) == 0 and int(w[0]) > 10_000_000
    img = out['images']
    assert img.shape == (3, 1080, 1920) and bool(torch.isfinite(img).all())
    assert float(img.min()) >= 0 and float(out['alpha'].max()) < 1 and float(out['alpha'].mean()) > 0.3
    assert int((out['radii'] > 0).sum()) > 2_500_000


def test_b2_entry_points_match_b1():
    sc = S.make_scene('c1', P=5000, seed=3)
    hp = HotPath(sc, 'cuda:0', requires_grad=False)
    with torch.no_grad():
        net, _ = hp.deform()
    rs = hp.settings[0]
    n, color, opacity, radii, geom, binning, img, extra = rasterize_gaussians_b2(
        rs.image_height, rs.image_width, rs.tanfovx, rs.tanfovy, rs.sh_degree, 1.0, False, False, True, rs.viewmatrix,
        rs.projmatrix, rs.campos, net['points'], net['opacity'], net['sh_features'], net['scales'], net['rotations'],
        None, torch.Tensor([]), torch.Tensor([]))
    c1, d1, a1, r1, st = DGR.rasterize_forward(rs, net['points'], net['opacity'], shs=net['sh_features'],
                                               scales=net['scales'], rotations=net['rotations'], quat_wxyz=False)
    assert n == st.num_rendered and torch.equal(radii, r1) and torch.equal(opacity, a1[0])
    bg = rs.bg.view(3, 1, 1)
    assert float((color + (1 - opacity[None]) * bg - c1).abs().max()) <= 1e-6   # reference blends bg in Python
    dC = torch.randn_like(color) / color.numel()
    out = rasterize_gaussians_backward_b2(1.0, rs.tanfovx, rs.tanfovy, rs.sh_degree, False, True, rs.viewmatrix,
                                          rs.projmatrix, rs.campos, net['points'], None, None, net['scales'],
                                          net['rotations'], None, net['sh_features'], n, radii, opacity, dC,
                                          torch.zeros_like(opacity), None, None, None, None, geom, binning, img)
    assert out[3].shape == (5000, 3) and bool(torch.isfinite(out[5]).all()) and float(out[5].abs().max()) > 0
    # the state travels in three plain uint8 tensors, like the reference's geomBuffer / binningBuffer / imgBuffer
    assert all(t.dtype == torch.uint8 and t.is_cuda for t in (geom, binning, img))
    _, _, _, _, st0 = DGR.rasterize_forward(rs._replace(bg=None), net['points'], net['opacity'], shs=net['sh_features'],
                                            scales=net['scales'], rotations=net['rotations'], quat_wxyz=False)
    gb1 = DGR.rasterize_backward(st0, dC, None, torch.zeros(1, *opacity.shape, device='cuda'))  # B2 has no background
    for got, want in ((out[0], gb1['means2D']), (out[2], gb1['opacities']), (out[3], gb1['means3D']),
                      (out[5], gb1['shs']), (out[6], gb1['scales']), (out[7], gb1['rotations'])):
        assert float((got - want).abs().max()) <= 2e-5 * float(want.abs().max())
    with pytest.raises(RuntimeError):
        rasterize_gaussians_b2(rs.image_height, rs.image_width, rs.tanfovx, rs.tanfovy, 3, 1.0, False, False, True,
                               rs.viewmatrix, rs.projmatrix, rs.campos, net['points'], net['opacity'],
                               net['sh_features'], net['scales'], net['rotations'], torch.ones(5000, 2, device='cuda'),
                               torch.Tensor([]), torch.Tensor([]))
