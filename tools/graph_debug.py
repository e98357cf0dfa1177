import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sk_gs_b200 import scene as S, _lib
from sk_gs_b200 import diff_gaussian_rasterization as DGR
from sk_gs_b200.pipeline import HotPath
from sk_gs_b200.fk_lbs import fk_lbs, assemble
cfg = S.CONFIGS['c1']
hp = HotPath(S.make_scene(cfg, views=1), 'cuda:0')
dL = (torch.randn(3, cfg.H, cfg.W) / (3 * cfg.H * cfg.W)).cuda()
hp.step(0, dL); torch.cuda.synchronize()
R = int(DGR.last_header_words(hp.device)[0]); DGR.set_fixed_capacity(R * 2)
for _ in range(3):
    hp.zero_grad(); hp.step(0, dL)
torch.cuda.synchronize()
def trial(name, fn, mode='global'):
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode=mode):
            fn()
        g.replay(); torch.cuda.synchronize()
        print(name, mode, 'OK')
    except Exception as e:
        import traceback; traceback.print_exc()
        print(name, mode, 'FAILED', str(e).splitlines()[0])
        torch.cuda.synchronize()
with torch.no_grad():
    net, sk = hp.deform()
net = {k: v.detach() for k, v in net.items()}
def raster_fwd_bwd():
    c, d, a, r, st = DGR.rasterize_forward(hp.settings[0], net['points'], net['opacity'], shs=net['sh_features'], scales=net['scales'], rotations=net['rotations'], quat_wxyz=False)
    DGR.rasterize_backward(st, dL)
p = hp.params
def lbs_fwd_bwd():
    out = fk_lbs(p['xyz'], p['joints'], p['sk_r'], p['sk_d_rot'], p['sk_d_scale'], p['g_tr'], hp.parents, hp.root, K=5, mode='W', sp_W=p['sp_W'])
    torch.autograd.grad(out[0].sum() + out[1].sum() + out[2].sum(), [p['joints'], p['sk_r'], p['sp_W']])
def asm_fwd_bwd():
    o = assemble(p['xyz'], p['scaling'], p['rotation'], p['opacity'])
    torch.autograd.grad(o[0].sum() + o[1].sum() + o[2].sum() + o[3].sum(), [p['xyz'], p['scaling'], p['rotation'], p['opacity']])
def full():
    hp.step_grads(0, dL)
def control():
    y = (p['xyz'] * 2).sum()
    torch.autograd.grad(y, [p['xyz']])
import sys as _s
which = _s.argv[1]
for name, fn in [(n, f) for n, f in [('control', control), ('raster', raster_fwd_bwd), ('lbs', lbs_fwd_bwd), ('assemble', asm_fwd_bwd), ('full', full)] if n == which]:
    fn(); torch.cuda.synchronize()
    trial(name, fn)
