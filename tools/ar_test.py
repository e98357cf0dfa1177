"""Compare all-reduce implementations for the gradient arena size (26.8 MB fp32)."""
import os, sys, time
import torch, torch.distributed as dist
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); local = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
n = 6_700_000
def bench(name, fn, iters=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    t = a.elapsed_time(b) / iters
    if rank == 0: print(f'{name:32s} {t*1000:8.1f} us  algbw {n*4/t/1e6:7.1f} GB/s', flush=True)
x = torch.randn(n, device='cuda')
bench('nccl all_reduce', lambda: dist.all_reduce(x))
bench('nccl all_reduce 2 chunks', lambda: (dist.all_reduce(x[:n//2]), dist.all_reduce(x[n//2:])))
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(n, dtype=torch.float32, device='cuda')
    hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    t.normal_()
    for op in ('two_shot_all_reduce_', 'multimem_all_reduce_', 'one_shot_all_reduce'):
        try:
            f = getattr(torch.ops.symm_mem, op)
            bench('symm_mem ' + op, lambda: f(t, 'sum', dist.group.WORLD.group_name))
        except Exception as e:
            if rank == 0: print(op, 'failed:', str(e).splitlines()[0][:150])
    if rank == 0: print('has multicast', hdl.multicast_ptr != 0 if hasattr(hdl, 'multicast_ptr') else None)
except Exception as e:
    if rank == 0: print('symm_mem unavailable:', str(e).splitlines()[0][:200])
dist.destroy_process_group()
