"""Debug: per-work-item statistics of composite_fwd_kernel (cycles, batches, survivors, hits)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sk_gs_b200 import scene as S, _lib
from sk_gs_b200.pipeline import HotPath
wl = sys.argv[1] if len(sys.argv) > 1 else 'c2'
cfg = S.CONFIGS[wl]
hp = HotPath(S.make_scene(cfg, views=1), 'cuda:0')
tiles = ((cfg.W + 15) // 16) * ((cfg.H + 15) // 16)
with torch.no_grad():
    hp.render(0); hp.render(0)
    stats = torch.zeros(tiles * 8, 6, dtype=torch.int64, device='cuda')
    L = _lib.lib(); L.skgs_debug_set_item_stats.argtypes = [ctypes.c_void_p]
    L.skgs_debug_set_item_stats(stats.data_ptr())
    hp.render(0); torch.cuda.synchronize()
    L.skgs_debug_set_item_stats(None)
s = stats.cpu().numpy()
cyc = s[:, 5]
print('items', len(s), 'sum cycles', cyc.sum(), 'max', cyc.max(), 'mean', cyc.mean())
print('entries', s[:, 1].sum(), 'batches', s[:, 2].sum(), 'survivors', s[:, 3].sum(), 'hit-lanes', s[:, 4].sum())
o = np.argsort(-cyc)[:12]
for i in o:
    t, tot, nb, ns, nh, c = s[i]
    print(f'item {i:5d} tile {t:5d} total {tot:6d} batches {nb:4d} surv {ns:6d} hitlanes {nh:7d} cycles {c:8d}  cyc/entry {c/max(tot,1):6.1f} cyc/surv {c/max(ns,1):6.1f}')
print('first items in order:'); 
for i in range(6):
    t, tot, nb, ns, nh, c = s[i]; print(f'  item {i} tile {t} total {tot} cycles {c}')
