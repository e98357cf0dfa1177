"""Time skgs_adam_step alone on the parameter shapes of workload c2 (which tensor classes cost what)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sk_gs_b200.optim import adam_step_raw

dev = 'cuda:0'
P, M, K = 100_000, 32, 5
g = torch.Generator().manual_seed(0)
mk = lambda *s: torch.randn(*s, generator=g).to(dev)  # noqa
dense = {'xyz': (P, 3), 'shs': (P, 16, 3), 'opacity': (P, 1), 'scaling': (P, 3), 'rotation': (P, 4)}
idx = torch.stack([torch.randperm(M, generator=g)[:K] for _ in range(2000)]).repeat(P // 2000, 1).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def bench(name, params, grads, lrs, knn=None):
    m = [torch.zeros_like(p) for p in params]
    v = [torch.zeros_like(p) for p in params]
    for _ in range(3):
        adam_step_raw(params, grads, m, v, lrs, 1, knn_indices=knn)
    torch.cuda.synchronize()
    ts = []
    for it in range(20):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        adam_step_raw(params, grads, m, v, lrs, it + 2, knn_indices=knn)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    n = sum(p.numel() for p in params)
    us = ts[len(ts) // 2]
    print(f'{os.environ.get("SKGS_LIB", "default")[-24:]:24s} {name:28s} {n:9d} params  {us:8.1f} us  {28 * n / us / 1e3:8.1f} GB/s')


ps = [mk(*s) for s in dense.values()]
bench('dense 5 tensors', ps, [mk(*s) for s in dense.values()], [1e-3, (1e-3, 1e-4, 48, 3), 1e-3, 1e-3, 1e-3])
bench('dense shs only (period)', [ps[1]], [mk(P, 16, 3)], [(1e-3, 1e-4, 48, 3)])
bench('dense shs only (plain lr)', [ps[1]], [mk(P, 16, 3)], [1e-3])
w = mk(P, M)
bench('sp_W compact [P,K] grad', [w], [mk(P, K)], [1e-3], knn=[idx])
bench('sp_W dense grad', [w], [mk(P, M)], [1e-3])
one = mk(64 << 20 >> 2)
bench('one 64 MiB tensor', [one], [torch.randn_like(one)], [1e-3])
