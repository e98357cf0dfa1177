"""A few hand-driven hot-path steps of one workload (for ncu): step_grads, i.e. the kernels of the benchmarked graph."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sk_gs_b200 import scene as S
from sk_gs_b200.pipeline import HotPath

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='c2')
ap.add_argument('--steps', type=int, default=4)
a = ap.parse_args()
cfg = S.CONFIGS[a.workload]
hp = HotPath(S.make_scene(cfg, views=1), 'cuda:0', merged_sh=True, requires_grad=False)
dL = (torch.randn(3, cfg.H, cfg.W) / (3 * cfg.H * cfg.W)).cuda()
for _ in range(a.steps):
    hp.step_grads(0, dL, compact_sp_W=True)
torch.cuda.synchronize()
