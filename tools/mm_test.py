"""Correctness + timing of SymmGradArena.allreduce (NVLS multimem kernel) against NCCL; run under torchrun."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from sk_gs_b200.dist import GradArena, SymmGradArena
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); local = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
dev = torch.device('cuda', local)
P, K, M = 100000, 5, 32
shapes = {'shs': (P, 16, 3), 'xyz': (P, 3), 'viewspace_points': (P, 3), 'scaling': (P, 3), 'rotation': (P, 4),
          'opacity': (P, 1), 'sp_W': (P, K), 'joints': (M, 3), 'sk_r': (M, 4), 'sk_d_rot': (M, 4), 'sk_d_scale': (M, 3), 'g_tr': (7,)}
a = SymmGradArena(shapes, dev, order=list(shapes))
b = GradArena(shapes, dev, order=list(shapes))
g = torch.Generator(device='cuda').manual_seed(100 + rank)
x = torch.randn(a.flat.numel(), device=dev, generator=g)
a.flat.copy_(x); b.flat.copy_(x)
a.allreduce(); b.allreduce()
torch.cuda.synchronize()
err = float((a.flat - b.flat).abs().max()); ref = float(b.flat.abs().max())
if rank == 0: print('multimem' if a.multimem else 'NCCL fallback', 'max |multimem - nccl| =', err, 'scale', ref, 'numel', a.flat.numel(), flush=True)
assert err <= 1e-5 * ref
# split exchange on two streams, as bench.py does it
a.flat.copy_(x); b.flat.copy_(x)
split = a.block_start('sp_W')
side = torch.cuda.Stream(dev)
main = torch.cuda.current_stream(dev)
side.wait_stream(main)
with torch.cuda.stream(side):
    a.allreduce_range(0, split, channel=0, exit_barrier=False)  # covered by the second range's barriers (after the join)
a.flat[split:].mul_(1.0)  # something on the main stream meanwhile
main.wait_stream(side)
a.allreduce_range(split, a.flat_padded.numel(), channel=1)
b.allreduce()
torch.cuda.synchronize()
err2 = float((a.flat - b.flat).abs().max())
if rank == 0: print('split exchange max err', err2, flush=True)
assert err2 <= 1e-5 * ref


def bench(name, fn, iters=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    if rank == 0: print(f'{name:28s} {s.elapsed_time(e) / iters * 1000:8.1f} us', flush=True)
bench('multimem arena allreduce', a.allreduce)
bench('nccl arena allreduce', b.allreduce)
from sk_gs_b200 import _lib
st = torch.cuda.current_stream().cuda_stream
bench('2 barriers only', lambda: (a.handle.barrier(channel=0), a.handle.barrier(channel=1)))
bench('multimem kernel only', lambda: _lib.lib().skgs_multimem_allreduce(a.handle.multicast_ptr, a.flat_padded.numel(), rank, world, st))
dist.destroy_process_group()
