"""GPU parity against the REFERENCE's OWN rasterizer at FULL size: my_ext/_C/src/nerf/gaussian_*.cu compiled UNMODIFIED
from /root/reference into oracle/_ref/ (oracle/build_ref.sh), driven like networks/renderer/gaussian_render.py:51-188
with colmap=True.  Two builds of the same sources:

* `_ref_raster_nofma` (-fmad=false): the reference's arithmetic as its source text states it.  Against it the north-star
  bars are asserted as written (BASELINE.json): radii, tiles touched, sorted keys, point_list and tile ranges - read
  back from the reference's geomBuffer / binningBuffer / imgBuffer - BIT-EXACT; image <= 1e-5 abs; gradients <= 1e-4
  relative.  The per-Gaussian floats that feed the later stages (means2D, depths, conic, cov3D, rgb) are compared bit
  for bit too.  What cannot be bit-identical is exp(): the reference calls CUDA's expf (MUFU-based, not reproducible on
  a CPU), this repo a fully specified polynomial (csrc/common.cuh `skgs_exp`, shared with the CPU oracle).  A 1-ulp
  difference can flip one of the three threshold tests of a (pixel, Gaussian) pair (alpha < 1/255, power > 0,
  T' < 1e-4); such pairs are COUNTED (pixels whose n_contrib or colour disagree), bounded, and the 1e-5 bar is asserted
  on all other pixels.
  Gradients: the in-tree extension recovers the transmittance its backward starts from as `T_final = 1 - out_opacity`
  with out_opacity = 1 - T (gaussian_render.cu:215, the upstream module keeps T itself) - a cancellation that costs it
  up to ~1e-4 of relative accuracy in EVERY gradient (measured below against the fp64-accumulating oracle: reference
  1.1e-4, this repo 3e-7 on dL/dsh).  The product keeps the exact T (upstream semantics); with settings.debug bit 2 its
  backward recovers T_final the reference's way, and THAT is what the 1e-4 bar is asserted on - it isolates the one
  line the two implementations differ by.  The default mode is asserted against the oracle (test_gpu_raster_parity.py)
  and, here, to stay within the reference's own distance from the oracle.
* `_ref_raster` (nvcc defaults, FMA contraction): what a user of the reference runs.  nvcc's choice of contractions is
  a compiler artefact, so radii may flip on a handful of Gaussians whose 3*sqrt(lambda) sits within an ulp of an
  integer; tolerance-level agreement is asserted.

Third-party semantics stay "parity unpinned" (SURVEY 8c): the depth / alpha outputs of upstream
diff_gaussian_rasterization have no counterpart in the in-tree extension (alpha is checked against its out_opacity).
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ref_ext
from sk_gs_b200 import diff_gaussian_rasterization as DGR
from sk_gs_b200 import scene as S
from sk_gs_b200.pipeline import raster_settings_for
from sk_gs_b200.renderer import render_gs_offical
from skgs_test_util import arena_view, oracle_deform, rel_err

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_ext.available(), reason='oracle/_ref not built')]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
IMG_ATOL = 1e-5     # north star: images within 1e-5 abs
GRAD_RTOL = 1e-4    # north star: gradients within 1e-4 rel (fp32)

# (scene, P override, seed override, view): c1 and the three full-size shapes the north star names - c2 100K @ 800^2,
# c4 300K @ 1024^2, one of c3's eight views (200K @ 512^2)
FULL = [('c1', None, None, 0), ('c2', None, None, 0), ('c3', None, None, 3), ('c4', None, None, 1)]


def _report(name, data):
    d = os.path.join(ROOT, 'gpurun_out')
    if os.path.isdir(d):
        with open(os.path.join(d, 'parity_vs_reference.jsonl'), 'a') as f:
            f.write(json.dumps(dict(case=name, **data)) + '\n')


def _scene_inputs(name, P, seed, view):
    sc = S.make_scene(name, P=P, seed=seed, views=view + 1)
    net, _, _ = oracle_deform(sc)
    return sc, sc.cameras[view], {k: v.detach() for k, v in net.items()}


def _ours_raw(cam, net, dev):
    """Product through the non-autograd entry (keeps the arenas for inspection), background off so that the colour
    buffer is comparable with the reference's (which blends the background in Python)."""
    rs = raster_settings_for(cam, dev)._replace(bg=None)
    color, depth, alpha, radii, st = DGR.rasterize_forward(
        rs, net['points'].to(dev), net['opacity'].to(dev), shs=net['sh_features'].to(dev),
        scales=net['scales'].to(dev), rotations=net['rotations'].to(dev), quat_wxyz=False)
    torch.cuda.synchronize()
    return color, depth, alpha, radii, st


def _grad_errors(got, ref):
    """(max-norm relative error, fraction of elements outside |a-b| <= rtol*|b| + rtol*rms(b), number of elements whose
    error exceeds rtol * max|ref|)."""
    got, ref = np.asarray(got, np.float64).ravel(), np.asarray(ref, np.float64).ravel()
    scale = max(np.abs(ref).max(), 1e-30)
    rms = np.sqrt((ref ** 2).mean())
    err = np.abs(got - ref)
    bad = err > GRAD_RTOL * np.abs(ref) + GRAD_RTOL * rms
    return float(err.max() / scale), float(bad.mean()), int((err > GRAD_RTOL * scale).sum())


@pytest.mark.skipif(not ref_ext.available(nofma=True), reason='oracle/_ref/_ref_raster_nofma not built')
@pytest.mark.parametrize('name,P,seed,view', FULL)
def test_bit_exact_state_and_north_star_bars_vs_reference_nofma(name, P, seed, view):
    dev = torch.device('cuda:0')
    sc, cam, net = _scene_inputs(name, P, seed, view)
    color, depth, alpha, radii, st = _ours_raw(cam, net, dev)
    ref_in = {k: v.to(dev).requires_grad_(True) for k, v in net.items()}
    r = ref_ext.render(ref_in['points'], ref_in['opacity'], ref_in['scales'], ref_in['rotations'],
                       ref_in['sh_features'], cam, nofma=True, keep_buffers=True)
    torch.cuda.synchronize()
    rb = ref_ext.parse_buffers(r['buffers'])
    Pn, R = st.P, int(r['buffers']['num_rendered'])
    lay = st.layout
    rep = dict(P=Pn, R=R, W=cam.W, H=cam.H)

    # ---- integers: bit-exact
    radii_m, radii_r = radii.cpu().numpy(), r['radii'].cpu().numpy()
    rep['radii_mismatch'] = int((radii_m != radii_r).sum())
    vis = radii_r > 0
    tt = arena_view(st.geom, lay.tiles_touched, torch.int32, Pn).cpu().numpy().view(np.uint32)
    rep['tiles_touched_mismatch'] = int((tt != rb['tiles_touched']).sum())
    rep['num_rendered'] = [int(st.num_rendered), R]
    # ---- the floats later stages consume: bit for bit on visible Gaussians
    f32 = torch.float32
    mine = dict(means2D=arena_view(st.geom, lay.means2D, f32, 2 * Pn).view(Pn, 2).cpu().numpy(),
                depths=arena_view(st.geom, lay.depths, f32, Pn).cpu().numpy(),
                cov3D=arena_view(st.geom, lay.cov3D, f32, 6 * Pn).view(Pn, 6).cpu().numpy(),
                conic_opacity=arena_view(st.geom, lay.conic_opacity, f32, 4 * Pn).view(Pn, 4).cpu().numpy(),
                rgb=arena_view(st.geom, lay.rgbd, f32, 4 * Pn).view(Pn, 4).cpu().numpy()[:, :3])
    for k in ('means2D', 'depths', 'cov3D', 'conic_opacity', 'rgb'):
        a, b = np.ascontiguousarray(mine[k][vis]), np.ascontiguousarray(rb[k][vis])
        rep[k + '_bit_mismatch'] = int((a.view(np.uint32) != b.view(np.uint32)).sum())
        rep[k + '_maxabs'] = float(np.abs(a - b).max()) if a.size else 0.0
    cl = arena_view(st.geom, lay.clamped, torch.uint8, Pn).cpu().numpy()
    cl3 = np.stack([(cl >> c) & 1 for c in range(3)], 1).astype(np.uint8)
    rep['clamped_mismatch'] = int((cl3[vis] != (rb['clamped'][vis] != 0)).sum())
    # ---- binning: sorted keys, point_list, ranges
    if int(st.num_rendered) == R:
        keys_t, plist_t = st.sorted_lists()
        keys, plist = keys_t.cpu().numpy().view(np.uint64), plist_t.cpu().numpy().view(np.uint32)
        rep['keys_mismatch'] = int((keys != rb['keys']).sum())
        rep['point_list_mismatch'] = int((plist != rb['point_list']).sum())
    tiles = rb['ranges'].shape[0]
    ranges = arena_view(st.img, lay.ranges, torch.int32, 2 * tiles).view(tiles, 2).cpu().numpy().view(np.uint32)
    rep['ranges_mismatch'] = int((ranges != rb['ranges']).any(axis=1).sum())
    # ---- image: 1e-5 abs except pixels with a flipped (pixel, Gaussian) threshold test
    H, W = cam.H, cam.W
    ncm = arena_view(st.img, lay.n_contrib, torch.int32, H * W).view(H, W).cpu().numpy().view(np.uint32)
    img_m, img_r = color.cpu().numpy(), r['color_nobg'].detach().cpu().numpy()
    diff = np.abs(img_m - img_r).max(axis=0)
    over = diff > IMG_ATOL
    rep['image_max_diff'] = float(diff.max())
    rep['image_mean_diff'] = float(diff.mean())
    rep['pixels_over_1e-5'] = int(over.sum())
    rep['n_contrib_mismatch_pixels'] = int((ncm != rb['n_contrib']).sum())
    rep['image_max_diff_excluding_flips'] = float(diff[~over].max())
    a_diff = np.abs(alpha.cpu().numpy()[0] - r['opacity'].detach().cpu().numpy())
    rep['alpha_pixels_over_1e-5'] = int((a_diff > IMG_ATOL).sum())
    # ---- gradients
    dC = (torch.randn(3, H, W, generator=torch.Generator().manual_seed(1)) / (3 * H * W)).to(dev)
    gg = DGR.rasterize_backward(st, dC)
    r['color_nobg'].backward(dC)
    torch.cuda.synchronize()
    pairs = [('means3D', 'points'), ('scales', 'scales'), ('rotations', 'rotations'), ('opacities', 'opacity'),
             ('shs', 'sh_features')]
    gerr = {}
    for mk, rk in pairs:
        gerr[mk] = _grad_errors(gg[mk].cpu().numpy(), ref_in[rk].grad.cpu().numpy().reshape(gg[mk].shape))
    gerr['means2D'] = _grad_errors(gg['means2D'].cpu().numpy(), r['viewspace_points'].grad.cpu().numpy())
    rep['grad_rel_maxnorm_default_mode'] = {k: v[0] for k, v in gerr.items()}
    # the same backward with the reference's T_final recovery (settings.debug bit 2)
    ge = DGR.rasterize_backward(st, dC, debug_flags=4)
    gerr = {mk: _grad_errors(ge[mk].cpu().numpy(), ref_in[rk].grad.cpu().numpy().reshape(ge[mk].shape))
            for mk, rk in pairs}
    gerr['means2D'] = _grad_errors(ge['means2D'].cpu().numpy(), r['viewspace_points'].grad.cpu().numpy())
    rep['grad_rel_maxnorm'] = {k: v[0] for k, v in gerr.items()}
    rep['grad_frac_elements_outside'] = {k: v[1] for k, v in gerr.items()}
    rep['grad_elements_over_bar'] = {k: v[2] for k, v in gerr.items()}
    # how reproducible is each side?  fp32 atomics in a nondeterministic order: run both backward passes a second time
    ref_in2 = {k: v.to(dev).requires_grad_(True) for k, v in net.items()}
    r2 = ref_ext.render(ref_in2['points'], ref_in2['opacity'], ref_in2['scales'], ref_in2['rotations'],
                        ref_in2['sh_features'], cam, nofma=True)
    r2['color_nobg'].backward(dC)
    gg2 = DGR.rasterize_backward(st, dC)
    torch.cuda.synchronize()
    rep['grad_selfnoise_reference'] = {mk: _grad_errors(ref_in2[rk].grad.cpu().numpy(), ref_in[rk].grad.cpu().numpy())[0]
                                       for mk, rk in pairs}
    rep['grad_selfnoise_ours'] = {mk: _grad_errors(gg2[mk].cpu().numpy(), gg[mk].cpu().numpy())[0] for mk, _ in pairs}
    if Pn <= 100000:  # both against the CPU oracle, which accumulates every sum in fp64 (the closest thing to exact)
        from oracle import raster as OR
        from skgs_test_util import np32, oracle_settings
        s_o = oracle_settings(cam)
        s_o.bg = np.zeros(3, np.float32)  # both GPU sides rendered without background
        img_o, g_o, b_o = OR.render_forward(s_o, np32(net['points']), np32(net['opacity']), np32(net['scales']),
                                            np32(net['rotations']), np32(net['sh_features']))
        og = OR.render_backward(s_o, g_o, b_o, img_o, dC.cpu().numpy(), np32(net['points']), np32(net['scales']),
                                np32(net['rotations']), np32(net['sh_features']))
        exact = {'means3D': og.dL_dmeans3D, 'scales': og.dL_dscales, 'rotations': og.dL_drotations,
                 'opacities': og.dL_dopacity, 'shs': og.dL_dsh}
        rep['grad_vs_fp64_oracle_ours'] = {mk: _grad_errors(gg[mk].cpu().numpy(), exact[mk])[0] for mk, _ in pairs}
        rep['grad_vs_fp64_oracle_reference'] = {mk: _grad_errors(ref_in[rk].grad.cpu().numpy(), exact[mk])[0]
                                                for mk, rk in pairs}
    _report(f'{name}-view{view}-nofma', rep)

    # ================================================================================================ assertions
    assert rep['radii_mismatch'] == 0, rep
    assert rep['tiles_touched_mismatch'] == 0 and int(st.num_rendered) == R, rep
    for k in ('means2D', 'depths', 'conic_opacity', 'cov3D'):
        assert rep[k + '_bit_mismatch'] == 0, (k, rep)
    assert rep['rgb_maxabs'] <= 1e-6 and rep['clamped_mismatch'] <= 2, rep  # SH polynomial: association may differ
    assert rep['keys_mismatch'] == 0 and rep['point_list_mismatch'] == 0 and rep['ranges_mismatch'] == 0, rep
    # threshold flips (expf vs the specified polynomial): a few pixels per million, each off by at most one
    # alpha = 1/255 contribution (colours can reach ~1.6) or one early-termination (T < 1e-4) step
    flips_allowed = max(8, int(4e-5 * H * W))
    assert rep['pixels_over_1e-5'] <= flips_allowed, rep
    assert rep['image_max_diff'] <= 1.7 / 255.0, rep
    assert rep['image_max_diff_excluding_flips'] <= IMG_ATOL, rep
    assert rep['alpha_pixels_over_1e-5'] <= flips_allowed, rep
    # like for like (T_final recovered the reference's way): every gradient element within 1e-4 of the tensor's scale,
    # except the few entries of the Gaussians that touch a pixel with a flipped threshold test (counted above)
    flipped = rep['pixels_over_1e-5'] + rep['n_contrib_mismatch_pixels']
    for k, e in rep['grad_rel_maxnorm'].items():
        assert e <= GRAD_RTOL or (flipped > 0 and rep['grad_elements_over_bar'][k] <= 16 * flipped), (k, rep)
    for k, f in rep['grad_frac_elements_outside'].items():
        assert f <= 2e-3, (k, rep)
    for k, e in rep['grad_rel_maxnorm_default_mode'].items():  # exact T_final: off by the reference's own cancellation
        assert e <= 4e-4, (k, rep)
    if 'grad_vs_fp64_oracle_ours' in rep:
        for k, e in rep['grad_vs_fp64_oracle_ours'].items():
            assert e <= GRAD_RTOL and e <= rep['grad_vs_fp64_oracle_reference'][k] + 2e-6, (k, rep)


@pytest.mark.parametrize('name,P,seed,view', [('c1', None, None, 0), ('c2', None, None, 0)])
def test_forward_backward_vs_reference_extension_default_build(name, P, seed, view):
    """The FMA-contracted build a user of the reference runs, through the drop-in autograd entry point."""
    dev = torch.device('cuda:0')
    sc, cam, net = _scene_inputs(name, P, seed, view)
    mine = {k: v.to(dev).requires_grad_(True) for k, v in net.items()}
    ref = {k: v.to(dev).requires_grad_(True) for k, v in net.items()}
    out = render_gs_offical(raster_settings=raster_settings_for(cam, dev), **mine)
    r = ref_ext.render(ref['points'], ref['opacity'], ref['scales'], ref['rotations'], ref['sh_features'], cam)
    radii_m, radii_r = out['radii'].cpu().numpy(), r['radii'].cpu().numpy()
    flips = int((radii_m != radii_r).sum())
    img_m, img_r = out['images'].detach().cpu().numpy(), r['images'].detach().cpu().numpy()
    diff = np.abs(img_m - img_r)
    alpha_r = r['opacity'].detach().cpu().numpy()
    H, W = cam.H, cam.W
    dC = torch.randn(3, H, W, generator=torch.Generator().manual_seed(1)).to(dev) / (3 * H * W)
    out['images'].backward(dC)
    r['images'].backward(dC)
    errs = {k: rel_err(mine[k].grad.cpu().numpy(), ref[k].grad.cpu().numpy())
            for k in ('points', 'scales', 'rotations', 'opacity', 'sh_features')}
    errs['viewspace'] = rel_err(out['viewspace_points'].grad.cpu().numpy(), r['viewspace_points'].grad.cpu().numpy())
    rep = dict(radii_flips=flips, image_max=float(diff.max()), image_mean=float(diff.mean()),
               pixels_over_1e_5=int((diff.max(axis=0) > IMG_ATOL).sum()), grad_rel=errs)
    _report(f'{name}-view{view}-fma', rep)
    assert flips <= max(2, len(radii_m) // 5000), rep
    assert diff.mean() <= 2e-6, rep
    assert (diff > 1e-4).mean() <= 2e-4, rep  # pixels under a radius / threshold flip
    assert np.abs(out['alpha'].detach().cpu().numpy()[0] - alpha_r).mean() <= 2e-6, rep
    for k, e in errs.items():
        assert e <= 2e-3, (k, rep)


@pytest.mark.skipif(not ref_ext.available(nofma=True), reason='oracle/_ref/_ref_raster_nofma not built')
def test_b2_entry_points_against_the_reference_functions_they_replace():
    """Boundary B2: `rasterize_gaussians` / `rasterize_gaussians_backward` of the in-tree extension
    (gaussian_rasterizer_forward.cu:260-315, gaussian_rasterizer_backwrad.cu:200-261) and this repo's drop-ins, called
    with the SAME positional arguments; the state crosses forward -> backward in three uint8 tensors on both sides."""
    from sk_gs_b200.renderer import rasterize_gaussians_b2, rasterize_gaussians_backward_b2
    dev = torch.device('cuda:0')
    sc, cam, net = _scene_inputs('c1', None, None, 0)
    n = {k: v.to(dev) for k, v in net.items()}
    empty = torch.Tensor([])
    m = ref_ext.module(nofma=True)
    fargs = (cam.H, cam.W, cam.tanfovx, cam.tanfovy, 3, 1.0, False, False, True, cam.viewmatrix.to(dev),
             cam.projmatrix.to(dev), cam.campos.to(dev), n['points'], n['opacity'], n['sh_features'], n['scales'],
             n['rotations'], None, empty, empty)
    R_r, color_r, op_r, radii_r, gb_r, bb_r, ib_r, _ = m.rasterize_gaussians(*fargs)
    R_m, color_m, op_m, radii_m, gb_m, bb_m, ib_m, _ = rasterize_gaussians_b2(*fargs)
    assert R_m == R_r and torch.equal(radii_m, radii_r)
    assert all(t.dtype == torch.uint8 for t in (gb_m, bb_m, ib_m))
    d = (color_m - color_r).abs().amax(0)
    assert int((d > IMG_ATOL).sum()) <= 8 and float(d[d <= IMG_ATOL].max()) <= IMG_ATOL
    assert int(((op_m - op_r).abs() > IMG_ATOL).sum()) <= 8
    H, W = cam.H, cam.W
    dC = (torch.randn(3, H, W, generator=torch.Generator().manual_seed(5)) / (3 * H * W)).to(dev)
    dO = (torch.randn(H, W, generator=torch.Generator().manual_seed(6)) / (H * W)).to(dev)

    def bargs(R, radii, op, gb, bb, ib):
        return (1.0, cam.tanfovx, cam.tanfovy, 3, False, True, cam.viewmatrix.to(dev), cam.projmatrix.to(dev),
                cam.campos.to(dev), n['points'], empty, None, n['scales'], n['rotations'], empty, n['sh_features'], R,
                radii, op, dC, dO, None, None, None, None, gb, bb, ib)
    out_r = m.rasterize_gaussians_backward(*bargs(R_r, radii_r, op_r, gb_r, bb_r, ib_r))
    out_m = rasterize_gaussians_backward_b2(*bargs(R_m, radii_m, op_m, gb_m, bb_m, ib_m))
    out_e = rasterize_gaussians_backward_b2(*bargs(R_m, radii_m, op_m, gb_m, bb_m, ib_m), _debug_flags=4)
    names = ['dL_dmeans2D', 'dL_dcolors', 'dL_dopacity', 'dL_dmeans3D', 'dL_dcov3D', 'dL_dsh', 'dL_dscales',
             'dL_drotations']
    for k, (a, b) in zip(names, zip(out_m, out_r)):
        if k in ('dL_dcolors', 'dL_dcov3D'):
            continue  # intermediate products of the reference (per-Gaussian colour / covariance), not outputs here
        assert a is not None, k
        assert rel_err(a.cpu().numpy().reshape(-1), b.cpu().numpy().reshape(-1)[:a.numel()]) <= 4e-4, k
    for k, (a, b) in zip(names, zip(out_e, out_r)):  # T_final recovered the reference's way: the north-star bar
        if k in ('dL_dcolors', 'dL_dcov3D'):
            continue
        assert rel_err(a.cpu().numpy().reshape(-1), b.cpu().numpy().reshape(-1)[:a.numel()]) <= GRAD_RTOL, k
