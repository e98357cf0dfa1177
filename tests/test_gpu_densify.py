"""GPU parity of the densification bookkeeping (csrc/densify.cu, sk_gs_b200/densify.py) against
(a) the results of the REFERENCE's own add_densification_stats / densify / prune / reset_opacity + change_optimizer
    (tests/golden/densify.npz, produced by tests/golden/make_golden.py from /root/reference/networks/
    gaussian_splatting.py:503-665 on a real torch.optim.Adam), and
(b) the oracle restatement at the benchmark's size;
and the loop it exists for: a captured TrainLoop that keeps training across a change of the Gaussian count.

Bars: every copied value (parameters, Adam moments, order of the survivors) bit-exact; split samples and re-scaled
log-scales within 1e-6 (fp32 expf / logf / 3x3 product against torch's); statistics within 1 ulp of torch.norm."""
import os

import numpy as np
import pytest
import torch

from oracle import densify as OD
from sk_gs_b200 import scene as S
from sk_gs_b200.densify import (AdaptiveControl, DensifyStats, add_densification_stats, check_interval,
                                densify_and_prune, reset_opacity)
from sk_gs_b200.pipeline import HotPath
from sk_gs_b200.train import TrainLoop
from test_oracle_golden import DENSIFY_NAMES, _densify_case

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
DEV = 'cuda:0'


def _merge_sh(t):
    """The product keeps the SH coefficients as one [P,16,3] tensor (f_dc | f_rest)."""
    out = {n: t[n] for n in ('xyz', 'scaling', 'rotation', 'opacity')}
    out['shs'] = tuple(torch.cat([a, b], dim=1) for a, b in zip(t['f_dc'], t['f_rest']))
    return out


def test_golden_cases_of_the_reference_functions():
    d = np.load(os.path.join(G, 'densify.npz'))
    for i in range(int(d['n'])):
        pre = f'c{i}_'
        t, kw, noise = _densify_case(d, i)
        P = t['xyz'][0].shape[0]
        stats = DensifyStats(P, DEV)
        for j in range(3):
            add_densification_stats(stats, torch.from_numpy(d[pre + f'radii{j}']).to(DEV),
                                    torch.from_numpy(d[pre + f'vsgrad{j}']).to(DEV))
        assert np.abs(stats.grad_accum.cpu().numpy() - d[pre + 'stat_accum'][:, 0]).max() <= 1e-9
        assert np.array_equal(stats.denom.cpu().numpy(), d[pre + 'stat_denom'][:, 0])
        assert np.array_equal(stats.max_radii2D.cpu().numpy(), d[pre + 'stat_radii'])
        # decisions must not depend on the 1-ulp difference of the norm: use the reference's own statistics
        stats.grad_accum.copy_(torch.from_numpy(d[pre + 'stat_accum'][:, 0]))
        gpu = {n: tuple(x.to(DEV).contiguous() for x in tr) for n, tr in _merge_sh(t).items()}
        res = densify_and_prune(gpu, stats, noise=noise.to(DEV), **kw)
        ref = _merge_sh({n: tuple(torch.from_numpy(d[pre + k + n]) for k in ('out_', 'out_m_', 'out_v_'))
                         for n in DENSIFY_NAMES})
        assert res.counts['n_new'] == ref['xyz'][0].shape[0], (i, res.counts)
        kind = res.kind.cpu()
        for n, tr in ref.items():
            for j in range(3):
                got, want = res.tensors[n][j].cpu(), tr[j]
                assert got.shape == want.shape, (i, n, j)
                copied = torch.ones(len(kind), dtype=torch.bool) if n not in ('xyz', 'scaling') or j > 0 else kind < 2
                assert torch.equal(got[copied], want[copied]), (i, n, j)
                assert float((got - want).abs().max()) <= 1e-6, (i, n, j)
        assert np.array_equal(res.stats.grad_accum.cpu().numpy(), d[pre + 'out_accum'].reshape(-1))
        assert np.array_equal(res.stats.denom.cpu().numpy(), d[pre + 'out_denom'].reshape(-1))
        assert np.array_equal(res.stats.max_radii2D.cpu().numpy(), d[pre + 'out_radii'])
        o, m, v = res.tensors['opacity']
        reset_opacity(o, m, v)
        assert float((o.cpu() - torch.from_numpy(d[pre + 'reset_opacity'])).abs().max()) <= 1e-6
        assert float(m.abs().max()) == 0 and float(v.abs().max()) == 0


@pytest.mark.parametrize('do_densify,do_prune,screen', [(True, True, 20.0), (True, False, 0.0), (False, True, 20.0)])
def test_benchmark_size_against_oracle(do_densify, do_prune, screen):
    """P = 100 000 with the skinning table as a seventh per-Gaussian tensor (param_names_map of the skeleton model,
    networks/sk_gs.py:471): same survivors in the same order as the oracle's four rounds of cat / mask indexing."""
    P, M = 100_000, 32
    g = torch.Generator().manual_seed(11)
    r = lambda *s, scale=1.0, shift=0.0: torch.randn(*s, generator=g) * scale + shift  # noqa: E731
    params = dict(xyz=r(P, 3, scale=0.5), shs=r(P, 16, 3), scaling=r(P, 3, shift=-3.6), rotation=r(P, 4),
                  opacity=r(P, 1, scale=3.0, shift=-2.0), sp_W=r(P, M))
    t = {n: (p, r(*p.shape, scale=0.01), r(*p.shape, scale=0.01).abs()) for n, p in params.items()}
    accum = (torch.rand(P, generator=g) * 1.2e-3)
    denom = torch.randint(0, 4, (P,), generator=g).float()
    radii = torch.randint(0, 40, (P,), generator=g).float()
    noise = torch.randn(2 * P, 3, generator=g)
    kw = dict(do_densify=do_densify, do_prune=do_prune, grad_threshold=0.0002, densify_extent=0.02, min_opacity=0.005,
              max_screen_size=screen, prune_extent=0.2)
    want, w_accum, w_denom, w_radii = OD.densify_and_prune(t, accum.clone(), denom.clone(), radii.clone(), noise=noise,
                                                           **kw)
    stats = DensifyStats(P, DEV)
    stats.grad_accum, stats.denom, stats.max_radii2D = accum.to(DEV), denom.to(DEV), radii.to(DEV)
    res = densify_and_prune({n: tuple(x.to(DEV) for x in tr) for n, tr in t.items()}, stats, noise=noise.to(DEV), **kw)
    assert res.counts['n_new'] == want['xyz'][0].shape[0]
    kind = res.kind.cpu()
    for n in t:
        for j in range(3):
            got, ref = res.tensors[n][j].cpu(), want[n][j]
            assert got.shape == ref.shape, (n, j)
            if n in ('xyz', 'scaling') and j == 0:
                assert torch.equal(got[kind < 2], ref[kind < 2])
                assert float((got - ref).abs().max()) <= 2e-6, n
            else:
                assert torch.equal(got, ref), (n, j)
    assert torch.equal(res.stats.grad_accum.cpu(), w_accum) and torch.equal(res.stats.max_radii2D.cpu(), w_radii)
    assert torch.equal(res.stats.denom.cpu(), w_denom)
    # the plan is a permutation-with-repeats of the sources in the documented order
    src = res.src.cpu().long()
    nk, nc = res.counts['n_keep'], res.counts['n_clone']
    assert torch.all(src[:nk][1:] > src[:nk][:-1]) and torch.all(kind[:nk] == 0)
    assert torch.all(kind[nk:nk + nc] == 1)
    ns = res.counts['n_split']
    assert torch.equal(src[nk + nc:nk + nc + ns], src[nk + nc + ns:]) and torch.all(kind[nk + nc + ns:] == 3)


def test_interval_rule_matches_reference_check_interval_v2():
    """my_ext/utils/utils.py:126-146 with close='()' on the default schedule [100, 500, 25000]."""
    hits = [s for s in range(0, 26000) if check_interval(s, 100, 500, 25_000)]
    assert hits[0] == 600 and hits[-1] == 24_900 and len(hits) == 244
    assert not check_interval(500, 100, 500, 25_000) and not check_interval(25_000, 100, 500, 25_000)
    assert check_interval(6000, 3000, 3000, -1) and not check_interval(3000, 3000, 3000, -1)
    assert not check_interval(100, 0, 0, -1)


def test_train_loop_keeps_training_across_densification():
    """A captured TrainLoop with AdaptiveControl: statistics accumulate inside the graph; on the scheduled step the
    Gaussian set is rebuilt (parameters + moments), the graph is re-captured for the new P and replay() continues.  The
    same schedule run with eager steps gives the same Gaussian counts and the same loss trajectory."""
    cfg = dict(densify_interval=(4, 0, 100), prune_interval=(4, 0, 100), opacity_reset_interval=(7, 0, -1),
               densify_grad_threshold=2e-6, prune_max_screen_size=0)
    results = {}
    for mode in ('graph', 'eager'):
        sc = S.make_scene('c1', P=6000, views=1, seed=5)
        hp = HotPath(sc, DEV, requires_grad=False, merged_sh=True)
        target = torch.rand(3, sc.cfg.H, sc.cfg.W, generator=torch.Generator().manual_seed(1)).to(DEV)
        loop = TrainLoop(hp)
        ctl = AdaptiveControl(loop, cameras_extent=4.0, cfg=cfg, seed=3)
        if mode == 'graph':
            loop.capture(0, target, headroom=3.0)
        losses, counts = [], []
        for step in range(10):
            out = loop.replay() if mode == 'graph' else loop.step(0, target)
            torch.cuda.synchronize()
            losses.append(float(out['loss_terms'][2]))
            if step == 2:
                assert float(ctl.stats.denom.max()) == 3.0  # three steps of statistics, none from warm-up / capture
            ctl.after_step(step)
            counts.append(hp.params['xyz'].shape[0])
            assert loop.exp_avg['xyz'].shape == hp.params['xyz'].shape
            assert hp.params['sp_W'].shape[0] == counts[-1] and ctl.stats.P == counts[-1]
        results[mode] = (losses, counts, ctl.history)
        assert counts[3] != 6000 and counts[2] == 6000      # step index 3 -> step 4: densify + prune ran
        assert all(np.isfinite(losses))
        assert not hp.overflowed()
    assert results['graph'][1] == results['eager'][1]
    assert np.abs(np.array(results['graph'][0]) - np.array(results['eager'][0])).max() <= 1e-4


def test_edge_cases_nothing_selected_everything_pruned_single_gaussian():
    """Degenerate plans: no Gaussian qualifies (identity), every Gaussian is pruned (empty set), P = 1."""
    g = torch.Generator().manual_seed(2)
    for P in (1, 777):
        params = dict(xyz=torch.randn(P, 3, generator=g), shs=torch.randn(P, 16, 3, generator=g),
                      scaling=torch.full((P, 3), -4.0), rotation=torch.randn(P, 4, generator=g),
                      opacity=torch.full((P, 1), 2.0))
        t = {n: (p.to(DEV), torch.ones_like(p).to(DEV), torch.ones_like(p).to(DEV)) for n, p in params.items()}
        stats = DensifyStats(P, DEV)  # denom = 0 -> NaN gradient -> 0: nothing is hot
        res = densify_and_prune(t, stats, True, True, grad_threshold=0.0002, densify_extent=0.02, min_opacity=0.005,
                                max_screen_size=20.0, prune_extent=0.2)
        assert res.counts == dict(n_keep=P, n_clone=0, n_split=0, n_selected=0, n_new=P)
        for n in t:
            assert all(torch.equal(a, b) for a, b in zip(res.tensors[n], t[n])), n
        assert torch.equal(res.src.cpu(), torch.arange(P, dtype=torch.int32))
        # opacity below the threshold everywhere: the set becomes empty
        t['opacity'] = (torch.full((P, 1), -20.0, device=DEV), t['opacity'][1], t['opacity'][2])
        res = densify_and_prune(t, stats, False, True, min_opacity=0.005)
        assert res.counts['n_new'] == 0 and res.tensors['xyz'][0].shape == (0, 3) and res.stats.P == 0
    # statistics: invisible Gaussians (radii == 0) are left alone
    stats = DensifyStats(5, DEV)
    add_densification_stats(stats, torch.tensor([0, 3, 0, 7, 0], dtype=torch.int32, device=DEV),
                            torch.tensor([[3.0, 4.0, 9.0]] * 5, device=DEV))
    assert stats.denom.tolist() == [0, 1, 0, 1, 0] and stats.max_radii2D.tolist() == [0, 3, 0, 7, 0]
    assert stats.grad_accum.tolist() == [0, 5, 0, 5, 0]
    with pytest.raises(RuntimeError):
        add_densification_stats(stats, torch.zeros(5, dtype=torch.int64, device=DEV), torch.zeros(5, 3, device=DEV))
    with pytest.raises(RuntimeError):
        densify_and_prune({'xyz': (torch.zeros(4, 3, device=DEV), None, None)}, DensifyStats(4, DEV), True, True)
