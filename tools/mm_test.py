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
    a.allreduce_range(0, split, channel=0)
a.flat[split:].mul_(1.0)  # something on the main stream meanwhile
main.wait_stream(side)
a.allreduce_range(split, a.flat_padded.numel(), channel=1)
b.allreduce()
torch.cuda.synchronize()
err2 = float((a.flat - b.flat).abs().max())
if rank == 0: print('split exchange max err', err2, flush=True)
assert err2 <= 1e-5 * ref


# the one-launch exchange: both cross-GPU barriers inside the reduce kernel; split over two streams like bench.py,
# the first range without exit barrier (covered by the second, issued after the join)
print_sync = rank == 0
if a.synced:
    for rep in range(3):  # several epochs: the control words advance on the device
        xr = torch.randn(a.flat.numel(), device=dev, generator=g)
        a.flat.copy_(xr); b.flat.copy_(xr)
        torch.cuda.synchronize(); dist.barrier()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            a.allreduce_range_synced(0, split, slot=0, exit_barrier=False, max_blocks=148)
        a.flat[split:].mul_(1.0)
        main.wait_stream(side)
        a.allreduce_range_synced(split, a.flat_padded.numel(), slot=1, exit_barrier=True)
        b.allreduce()
        torch.cuda.synchronize()
        err3 = float((a.flat - b.flat).abs().max()); ref3 = float(b.flat.abs().max())
        assert not a.sync_error(), 'a peer never arrived'
        assert err3 <= 1e-5 * ref3, (rep, err3, ref3)
    if print_sync: print('synced split exchange max err', err3, flush=True)
else:
    if print_sync: print('synced split exchange max err n/a (no signal pads)', flush=True)


# the same split exchange inside a captured CUDA graph, replayed with new inputs (what bench.py does)
if a.synced:
    for exit_a in (False, True):
        xin = torch.zeros(a.flat.numel(), device=dev)
        s2 = torch.cuda.Stream(dev)
        def body():
            a.flat.copy_(xin)
            m_ = torch.cuda.current_stream(dev)
            s2.wait_stream(m_)
            with torch.cuda.stream(s2):
                a.allreduce_range_synced(0, split, slot=3, exit_barrier=exit_a, max_blocks=0)
            a.flat[split:].mul_(1.0)
            m_.wait_stream(s2)
            a.allreduce_range_synced(split, a.flat_padded.numel(), slot=4, exit_barrier=True)
        sw = torch.cuda.Stream(dev)
        sw.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(sw):
            for _ in range(3):
                body()
        torch.cuda.current_stream(dev).wait_stream(sw)
        torch.cuda.synchronize(); dist.barrier()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            body()
        worst = [0.0, 0.0]
        for rep in range(6):
            xr = torch.randn(a.flat.numel(), device=dev, generator=g)
            xin.copy_(xr); b.flat.copy_(xr)
            gr.replay()
            b.allreduce()
            torch.cuda.synchronize()
            d = (a.flat - b.flat).abs()
            worst[0] = max(worst[0], float(d[:split].max())); worst[1] = max(worst[1], float(d[split:].max()))
        if rank == 0: print(f'graph synced exchange (exit barrier on first range: {exit_a}) max err range A {worst[0]:.3g} range B {worst[1]:.3g} ctrl {a._ctrl[3:5].tolist()}', flush=True)


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
if a.synced:
    bench('synced (1 launch, 2 barriers)', lambda: a.allreduce_range_synced(0, a.flat_padded.numel(), slot=2))
bench('multimem kernel only', lambda: _lib.lib().skgs_multimem_allreduce(a.handle.multicast_ptr, a.flat_padded.numel(), rank, world, st))
dist.destroy_process_group()
