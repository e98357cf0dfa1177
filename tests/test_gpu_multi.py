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
    assert 'split exchange max err' in r.stdout and 'synced split exchange max err' in r.stdout
