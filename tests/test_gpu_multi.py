"""Needs >= 2 GPUs on the box (skipped otherwise): the in-switch multimem all-reduce of SymmGradArena against NCCL,
including the split exchange on two streams, under torchrun with 2 ranks."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_multimem_allreduce_matches_nccl_world2():
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', '29561', os.path.join(ROOT, 'tools', 'mm_test.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert 'split exchange max err' in r.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_benchmark_step_exchange_matches_nccl_world2():
    """The captured data-parallel step of bench.py (gradient exchange overlapped with the LBS backward) leaves the same
    arena as a plain NCCL all-reduce of the per-rank gradients, over several replays - the check that caught the
    aliasing race of round 2 (the LBS backward reading arena blocks the overlapped all-reduce was already summing)."""
    import json
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', '29563', os.path.join(ROOT, 'bench.py'), '--gpus', '2', '--steps', '10',
           '--warmup', '3', '--headline-only']
    env = dict(os.environ, SKGS_CHECK_REPS='8')
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith('{')][-1]
    chk = json.loads(line)['exchange_check']
    assert chk['ok'] and chk['replays_checked'] == 8 and chk['radii_max_equal'], chk
