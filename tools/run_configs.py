"""Run every BASELINE.json config shape once on one GPU (1 view of each): sanity + timing table."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sk_gs_b200 import scene as S
from sk_gs_b200 import diff_gaussian_rasterization as DGR
from sk_gs_b200.pipeline import HotPath
names = sys.argv[1:] or ['c1', 'c2', 'c3', 'c4', 'ns', 'c5']
for name in names:
    cfg = S.CONFIGS[name]
    t0 = time.time()
    hp = HotPath(S.make_scene(cfg, views=1), 'cuda:0', requires_grad=cfg.backward)
    dL = (torch.randn(3, cfg.H, cfg.W) / (3 * cfg.H * cfg.W)).cuda()
    def step():
        if cfg.backward:
            return hp.step_grads(0, dL)
        with torch.no_grad():
            return hp.render(0), None
    for _ in range(3):
        out, g = step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    a.record()
    for _ in range(n):
        out, g = step()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    w = DGR.last_header_words('cuda:0')
    img = out['images']
    ok = bool(torch.isfinite(img).all()) and (g is None or all(torch.isfinite(v).all() for v in g.values() if v is not None))
    print(f'{cfg.name:28s} P={cfg.P:8d} {cfg.W}x{cfg.H} R={int(w[0]):9d} overflow={int(w[3])} '
          f'{"fwd+bwd" if cfg.backward else "fwd    "} {ms:8.3f} ms/view  finite={ok} alpha_mean={float(out["alpha"].mean()):.3f} '
          f'mem={torch.cuda.max_memory_allocated() / 2**30:.2f} GiB (setup {time.time() - t0:.0f}s)', flush=True)
    del hp, out, g
    torch.cuda.empty_cache()
