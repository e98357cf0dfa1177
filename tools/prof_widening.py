"""One pass over the widening-row kernels (sp-stage LBS, densification) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sk_gs_b200.densify import DensifyStats, add_densification_stats, densify_and_prune, reset_opacity
from sk_gs_b200.sp_lbs import sp_warp_backward_raw, sp_warp_forward_raw

dev = torch.device('cuda:0')
P, M, K = 100_000, 512, 5
g = torch.Generator().manual_seed(7)
r = lambda *s, scale=1.0, shift=0.0: (torch.randn(*s, generator=g) * scale + shift).to(dev)  # noqa: E731
bias = torch.tensor([0, 0, 0, 1.0], device=dev)
points, sp_points, sp_t = r(P, 3, scale=0.5), r(M, 3, scale=0.5), r(M, 3, scale=0.05)
sp_r = torch.nn.functional.normalize(r(M, 4, scale=0.2) + bias, dim=-1)
sp_rot = torch.nn.functional.normalize(r(M, 4, scale=0.2) + bias, dim=-1)
sp_scale, sp_W = r(M, 3, scale=0.01), r(P, M)
cots = [r(P, 3), r(P, 4), r(P, 3)]
for _ in range(2):
    out, ctx = sp_warp_forward_raw(points, sp_points, sp_t, sp_r, sp_rot, sp_scale, K=K, mode='W', sp_W=sp_W)
    sp_warp_backward_raw(ctx, *cots, compact_sp_W=True)
params = dict(xyz=r(P, 3, scale=0.5), shs=r(P, 16, 3), scaling=r(P, 3, shift=-3.6), rotation=r(P, 4),
              opacity=r(P, 1, scale=3.0, shift=-2.0), sp_W=r(P, 32))
trip = {n: (p, torch.zeros_like(p), torch.zeros_like(p)) for n, p in params.items()}
st = DensifyStats(P, dev)
for _ in range(2):
    add_densification_stats(st, torch.randint(0, 40, (P,), generator=g).int().to(dev), r(P, 3, scale=4e-4))
    res = densify_and_prune(trip, st, True, True, grad_threshold=0.0002, densify_extent=0.02, min_opacity=0.005,
                            max_screen_size=20.0, prune_extent=0.2)
o, m, v = res.tensors['opacity']
reset_opacity(o, m, v)
torch.cuda.synchronize()
print('P_new', res.counts)
