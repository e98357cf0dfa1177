"""A/B the variants in sk_gs_b200/variants/: per-kernel times (profile events) + graph step time for one workload."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
wl = sys.argv[1] if len(sys.argv) > 1 else 'c2'
code = r'''
import os, sys
sys.path.insert(0, %r)
import torch
from sk_gs_b200 import scene as S, _lib
from sk_gs_b200.pipeline import HotPath
cfg = S.CONFIGS[%r]
hp = HotPath(S.make_scene(cfg, views=1), 'cuda:0', merged_sh=True, requires_grad=False)
dL = (torch.randn(3, cfg.H, cfg.W) / (3 * cfg.H * cfg.W)).cuda()
for _ in range(3): hp.step_grads(0, dL)
torch.cuda.synchronize()
_lib.profile_enable(True)
for _ in range(5): hp.step_grads(0, dL)
torch.cuda.synchronize()
pr = _lib.profile_collect(); _lib.profile_enable(False)
g, out, grads = hp.capture_step(0, dL)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(10): g.replay()
a.record()
for _ in range(100): g.replay()
b.record(); torch.cuda.synchronize()
sort = sum(v[1] for k, v in pr.items() if k.startswith('tile_')) / 5
print('%%-22s pdl=%%s graph %%.1f us | pre %%.1f binning %%.1f plan %%.1f comp_fwd %%.1f comp_bwd %%.1f pre_bwd %%.1f' %% (os.environ.get('VNAME'), os.environ.get('SKGS_PDL', '1'), a.elapsed_time(b) * 10, (pr.get('preprocess_scan_kernel') or pr.get('deform_preprocess_kernel'))[1] / 5, sort, pr['tile_plan_kernel'][1] / 5, pr['composite_fwd_kernel'][1] / 5, pr['composite_bwd_kernel'][1] / 5, pr['preprocess_bwd_kernel'][1] / 5))
''' % (ROOT, wl)
libs = {os.path.basename(l)[8:-3]: l for l in sorted(glob.glob(os.path.join(ROOT, 'sk_gs_b200', 'variants', 'libskgs_*.so')))}
runs = [(n, {}) for n in libs]
for name, extra in runs:
    env = dict(os.environ, SKGS_LIB=libs[name], VNAME=name + ''.join(f' {k[5:]}={v}' for k, v in extra.items()), **extra)
    r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True)
    print(r.stdout.strip() or r.stderr.strip().splitlines()[-1], flush=True)
