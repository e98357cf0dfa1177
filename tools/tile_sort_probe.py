"""Per-kernel times of one forward (tile-segmented binning) + the debug counters of the per-tile sort."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sk_gs_b200 import _lib, scene as S
from sk_gs_b200.pipeline import HotPath
name = sys.argv[1] if len(sys.argv) > 1 else 'c2'
sc = S.make_scene(name, views=1)
hp = HotPath(sc, 'cuda:0', merged_sh=True, requires_grad=False)
for _ in range(3):
    out, ctx = hp.forward_raw(0)
torch.cuda.synchronize()
st = out['_raster_state']
lay = st.layout
cnt = st.img[lay.work_counters:lay.work_counters + 32].view(torch.int32).cpu().tolist()
print(name, 'counters', cnt, 'R', int(st.header().num_rendered))
_lib.profile_enable(True)
for _ in range(5):
    out, ctx = hp.forward_raw(0)
torch.cuda.synchronize()
prof = _lib.profile_collect()
_lib.profile_enable(False)
print({k: round(v[1] / v[0], 1) for k, v in prof.items()})
