"""Per-tile phase timeline of the LAST radix pass executed (debug hook skgs_debug_set_sort_trace)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sk_gs_b200 import _lib, scene as S
from sk_gs_b200.pipeline import HotPath
L = _lib.lib()
L.skgs_debug_set_sort_trace.argtypes = [ctypes.c_void_p]
cfg = S.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else 'c2']
hp = HotPath(S.make_scene(cfg, views=1), 'cuda:0', merged_sh=True, requires_grad=False)
for _ in range(3):
    hp.forward_raw(0)
torch.cuda.synchronize()
tr = torch.zeros(4096 * 8, dtype=torch.int64, device='cuda')
L.skgs_debug_set_sort_trace(tr.data_ptr())
hp.forward_raw(0)
torch.cuda.synchronize()
L.skgs_debug_set_sort_trace(None)
t = tr.cpu().numpy().reshape(-1, 8)[:, :6]
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
rel = (t - t0) / 1000.0
names = ['start', 'keys+zero', 'ranked', 'published', 'lookback', 'scattered']
print('tiles', len(t), '(last executed pass)')
print('phase end times relative to the first CTA start [us]: mean / max')
for k, n in enumerate(names):
    print(f'  {n:10s} {rel[:, k].mean():7.2f} {rel[:, k].max():7.2f}')
d = np.diff(rel, axis=1)
print('phase durations [us] mean / p90 / max:')
for k, n in enumerate(names[1:]):
    print(f'  {n:10s} {d[:, k].mean():7.2f} {np.percentile(d[:, k], 90):7.2f} {d[:, k].max():7.2f}')
order = np.argsort(t[:, 0])
print('look-back duration by tile index (every 16th):', [round(float(d[i, 3]), 2) for i in range(0, len(t), 16)])
