"""View-sharded data parallelism for the hot path.

Views (or reposing poses) of a step are independent given replicated parameters - the reference loops over them
serially (/root/reference/networks/sk_gs.py:1220) and sums their densification statistics
(networks/gaussian_splatting.py:509-512).  Here V views are split over the ranks of one NVLink/NVSwitch box, every rank
runs the whole hot path for its views, and there is exactly ONE exchange step per training iteration: an all-reduce
(SUM) of the Gaussian + skeleton gradients laid out in one flat fp32 arena, plus an all-reduce (MAX) of the screen radii
(max-radius tracking, networks/sk_gs.py:1992-1996).  Forward-only rendering needs no collective at all.

Plumbing is torch.distributed (NCCL on GPUs, gloo in the CPU tests); nothing here launches a kernel of its own.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


def shard_views(num_views: int, world: int, rank: int) -> List[int]:
    """Contiguous block partition of range(num_views); the first (num_views % world) ranks get one extra view."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f'bad rank/world {rank}/{world}')
    base, extra = divmod(num_views, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


class GradArena:
    """One flat fp32 buffer holding every gradient that has to be summed over ranks.

    Layout: the (large) per-Gaussian blocks first - SH gradients are 81 % of the bytes and are final right after
    preprocess-backward, so `allreduce(chunks=...)` can reduce them while the LBS / FK backward of the same step is
    still running on the compute stream - then the per-joint blocks, then the screen-space statistic."""

    def __init__(self, shapes: Dict[str, Sequence[int]], device, order: Optional[Sequence[str]] = None,
                 allocate: bool = True):
        self.names = list(order) if order is not None else sorted(shapes, key=lambda n: -int(torch.Size(shapes[n]).numel()))
        self.shapes = {n: torch.Size(shapes[n]) for n in self.names}
        self.offsets: Dict[str, Tuple[int, int]] = {}
        o = 0
        for n in self.names:
            k = self.shapes[n].numel()
            self.offsets[n] = (o, o + k)
            o = (o + k + 3) // 4 * 4  # every block starts on a 16-byte boundary (vector all-reduce of sub-ranges)
        self.numel = o
        self.flat = torch.zeros(o, dtype=torch.float32, device=device) if allocate else None

    def view(self, name: str) -> Tensor:
        a, b = self.offsets[name]
        return self.flat[a:b].view(self.shapes[name])

    def pack(self, grads: Dict[str, Optional[Tensor]], accumulate: bool = False):
        """Copy (or add, for several views per rank) the gradients into the arena; missing / None entries count as 0."""
        for n in self.names:
            g = grads.get(n)
            v = self.view(n)
            if g is None:
                if not accumulate:
                    v.zero_()
            elif accumulate:
                v.add_(g.reshape(v.shape))
            else:
                v.copy_(g.reshape(v.shape))

    def allreduce(self, scale: float = 1.0, group=None, chunks: int = 1, async_op: bool = False):
        """SUM over ranks (after multiplying by `scale`, e.g. 1/V for a mean over views).  With chunks > 1 the arena is
        reduced in that many contiguous pieces so that NCCL can start on the first piece early."""
        if scale != 1.0:
            self.flat.mul_(scale)
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return []
        n = self.flat.numel()
        chunks = max(int(chunks), 1)
        step = max((n + chunks - 1) // chunks, 1)
        works = []
        for a in range(0, n, step):
            w = dist.all_reduce(self.flat[a:a + step], op=dist.ReduceOp.SUM, group=group, async_op=async_op)
            if async_op:
                works.append(w)
        return works

    def unpack(self) -> Dict[str, Tensor]:
        return {n: self.view(n) for n in self.names}


class SymmGradArena(GradArena):
    """GradArena whose flat buffer is symmetric memory (same allocation on every GPU of the box, one NVLS multicast
    address).  `allreduce()` is then ONE kernel of libskgs_b200.so (multimem.ld_reduce + multimem.st: the NVSwitch sums
    and broadcasts, 1/N of the arena crosses each GPU's links) bracketed by two signal-pad barriers - instead of NCCL's
    ring (2(N-1)/N of the arena per GPU plus protocol latency).  torch's symmetric-memory module is used only for the
    plumbing: allocation, rendezvous, barriers.  Falls back to the NCCL path of the base class when multicast is not
    available (`self.multimem` tells which one is active)."""

    def __init__(self, shapes, device, order=None, group=None):
        super().__init__(shapes, device, order, allocate=False)
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        n = (self.numel + 3) // 4 * 4
        buf = symm_mem.empty(n, dtype=torch.float32, device=device)
        self.handle = symm_mem.rendezvous(buf, self.group.group_name)
        buf.zero_()
        self.flat_padded = buf
        self.flat = buf[:self.numel]
        self.multimem = bool(self.handle.multicast_ptr)
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)

    def allreduce(self, scale: float = 1.0, group=None, chunks: int = 1, async_op: bool = False):
        if not self.multimem:  # NCCL path of the base class, on the group this arena was built for
            return super().allreduce(scale, group if group is not None else self.group, chunks, async_op)
        if scale != 1.0:
            self.flat.mul_(scale)
        self.allreduce_range(0, self.flat_padded.numel())
        return []

    def allreduce_range(self, start: int, stop: int, channel: int = 0, exit_barrier: bool = True):
        """SUM over ranks of flat[start:stop] (both multiples of 4) on the CURRENT stream; `channel` selects the pair of
        signal-pad barriers so that two ranges can be in flight on different streams.  `exit_barrier=False`: the caller
        issues another allreduce_range (with both barriers) on every rank AFTER joining this stream - its entry barrier
        is passed only when every rank's kernel of this call has completed, which is all the exit barrier guarantees."""
        from . import _lib
        if not self.multimem:
            dist.all_reduce(self.flat_padded[start:stop], group=self.group)
            return
        assert start % 4 == 0 and stop % 4 == 0 and 0 <= start <= stop <= self.flat_padded.numel()
        st = torch.cuda.current_stream(self.flat.device).cuda_stream
        self.handle.barrier(channel=2 * channel)      # every rank's producers of this range have finished
        _lib.check(_lib.lib().skgs_multimem_allreduce(self.handle.multicast_ptr + 4 * start, stop - start, self.rank,
                                                      self.world, st), 'skgs_multimem_allreduce')
        if exit_barrier:
            self.handle.barrier(channel=2 * channel + 1)  # every slice has been reduced and broadcast

    def block_start(self, name: str) -> int:
        return self.offsets[name][0]


def allreduce_max_(t: Tensor, group=None) -> Tensor:
    """In-place MAX over ranks (screen radii)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t
