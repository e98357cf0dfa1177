"""Host-side logic of the view-sharded data-parallel path with world_size 2 on CPU (gloo)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sk_gs_b200.dist import GradArena, allreduce_max_, shard_views


def test_shard_views_partitions_exactly():
    for V in (0, 1, 4, 7, 8, 64):
        for world in (1, 2, 3, 4, 8):
            parts = [shard_views(V, world, r) for r in range(world)]
            flat = [v for p in parts for v in p]
            assert flat == list(range(V))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        shard_views(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, V, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        P, M = 50, 6
        shapes = {'xyz': (P, 3), 'f_rest': (P, 15, 3), 'sp_W': (P, M), 'joints': (M, 3), 'g_tr': (7,),
                  'viewspace_points': (P, 3)}
        arena = GradArena(shapes, 'cpu')
        views = shard_views(V, world, rank)
        first = True
        for v in views:  # per-view gradients are a deterministic function of the view index
            g = torch.Generator().manual_seed(100 + v)
            grads = {n: torch.randn(*s, generator=g) for n, s in shapes.items()}
            grads['g_tr'] = None if v % 2 else grads['g_tr']  # absent gradients count as zero
            arena.pack(grads, accumulate=not first)
            first = False
        if not views:
            arena.pack({}, accumulate=False)
        works = arena.allreduce(scale=1.0 / V, chunks=3, async_op=True)
        for w in works:
            w.wait()
        radii = torch.tensor([rank + 1, 5 - rank, 0], dtype=torch.int32)
        allreduce_max_(radii)
        res = {n: t.clone() for n, t in arena.unpack().items()}
        res['radii'] = radii
        out[rank] = res
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('V', [2, 5])
def test_grad_arena_allreduce_world2(V):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, V, out), nprocs=world, join=True)
    P, M = 50, 6
    shapes = {'xyz': (P, 3), 'f_rest': (P, 15, 3), 'sp_W': (P, M), 'joints': (M, 3), 'g_tr': (7,),
              'viewspace_points': (P, 3)}
    want = {n: torch.zeros(*s) for n, s in shapes.items()}
    for v in range(V):
        g = torch.Generator().manual_seed(100 + v)
        grads = {n: torch.randn(*s, generator=g) for n, s in shapes.items()}
        if v % 2:
            grads['g_tr'] = torch.zeros(7)
        for n in want:
            want[n] += grads[n] / V
    for rank in range(world):
        for n in want:
            assert torch.allclose(out[rank][n], want[n], atol=1e-6), (rank, n)
        assert out[rank]['radii'].tolist() == [2, 5, 0]
    # the big per-Gaussian blocks come first in the arena (they are reduced first)
    arena = GradArena(shapes, 'cpu')
    assert arena.names[0] == 'f_rest' and arena.offsets['f_rest'][0] == 0
