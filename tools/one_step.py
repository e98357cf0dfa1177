"""Run a few hot-path steps of one workload (used under ncu / for CPU-side profiling)."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sk_gs_b200 import scene as S
from sk_gs_b200.pipeline import HotPath

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='c2')
ap.add_argument('--steps', type=int, default=3)
ap.add_argument('--cprofile', action='store_true')
ap.add_argument('--fwd-only', action='store_true')
a = ap.parse_args()
cfg = S.CONFIGS[a.workload]
sc = S.make_scene(cfg, views=1)
hp = HotPath(sc, 'cuda:0')
dL = (torch.randn(3, cfg.H, cfg.W) / (3 * cfg.H * cfg.W)).cuda()


def run(n):
    for _ in range(n):
        hp.zero_grad()
        if a.fwd_only:
            with torch.no_grad():
                hp.render(0)
        else:
            hp.step(0, dL)
    torch.cuda.synchronize()


run(2)
if a.cprofile:
    import cProfile, pstats
    pr = cProfile.Profile()
    pr.enable()
    t0 = time.perf_counter()
    run(a.steps)
    dt = time.perf_counter() - t0
    pr.disable()
    print(f'{dt / a.steps * 1e3:.3f} ms per step (wall)')
    pstats.Stats(pr).sort_stats('cumulative').print_stats(35)
else:
    run(a.steps)
