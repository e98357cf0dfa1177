import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sk_gs_b200 import scene as S, _lib
from sk_gs_b200.pipeline import HotPath
cfg = S.CONFIGS['c2']
hp = HotPath(S.make_scene(cfg, views=1), 'cuda:0')
dL = (torch.randn(3, cfg.H, cfg.W) / (3 * cfg.H * cfg.W)).cuda()
def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n, (time.perf_counter() - t0) / n * 1e3
def eager():
    hp.zero_grad(); hp.step(0, dL)
print('eager ms (gpu, wall):', timeit(eager))
g, out, grads = hp.capture_step(0, dL)
print('launches per step', hp.launches_per_step)
print('graph ms (gpu, wall):', timeit(g.replay))
hp.zero_grad(); hp.step(0, dL); torch.cuda.synchronize()
print('grad check', {n: float((grads[n] - hp.params[n].grad).abs().max() / (hp.params[n].grad.abs().max() + 1e-30)) for n in hp.params}, 'overflow', hp.overflowed())
def manual():
    hp.step_grads(0, dL)
print('manual eager ms (gpu, wall):', timeit(manual))
_lib.profile_enable(True)
eager(); torch.cuda.synchronize()
print({k: round(v[1], 1) for k, v in _lib.profile_collect().items()})
_lib.profile_enable(False)
