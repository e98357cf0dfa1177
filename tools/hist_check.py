import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from sk_gs_b200 import scene as S, _lib
from sk_gs_b200.pipeline import HotPath
cfg = S.CONFIGS['c2']
hp = HotPath(S.make_scene(cfg, views=1), 'cuda:0')
from sk_gs_b200 import diff_gaussian_rasterization as DGR
with torch.no_grad():
    net, _ = hp.deform()
    c, d, a, r, st = DGR.rasterize_forward(hp.settings[0], net['points'], net['opacity'], shs=net['sh_features'], scales=net['scales'], rotations=net['rotations'], quat_wxyz=False)
torch.cuda.synchronize()
lay = st.layout
h = st.binning[lay.sort_hist:lay.sort_hist + 8 * 256 * 4].view(torch.int32).view(8, 256).cpu().numpy()
print('R', st.num_rendered)
for p in range(6):
    nz = np.nonzero(h[p])[0]
    print('pass', p, 'nonzero bins', len(nz), 'max', h[p].max(), 'sum', h[p].sum(), nz[:6])
_lib.profile_enable(True)
with torch.no_grad():
    DGR.rasterize_forward(hp.settings[0], net['points'], net['opacity'], shs=net['sh_features'], scales=net['scales'], rotations=net['rotations'], quat_wxyz=False)
torch.cuda.synchronize()
print(_lib.profile_collect())
