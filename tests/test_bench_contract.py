"""bench.py's JSON contract, exercised without a GPU through the reference arm on the host (the oracle port, a bounded
sample): every key the driver reads must be present and well-typed.  The GPU arm prints the same keys plus `roofline`,
`clocks`, `gpu_launches` (checked on the box by the driver)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--ref-device', 'cpu',
                          '--workload', 'c1', '--steps', '1', '--warmup', '0', '--cpu-steps', '1'],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
              'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'impl', 'cpu_baseline'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['higher_is_better'] is True and d['vs_baseline'] is None
    assert d['unit'] == 'steps/s' and d['value'] > 0 and d['n_gpus'] == 1 and d['scaling'] == 'weak'
    assert 'workload' in d['config'] and 'model' not in d['config']
    assert set(d['e2e']) >= {'value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step'}
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    cb = d['cpu_baseline']
    assert cb['kind'] in ('reference', 'port') and cb['cores'] >= 1 and cb['sample'] and cb['value'] == d['value']


def test_product_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '1', '--warmup', '0', '--no-cpu'],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0  # no CPU fallback of the product path
    assert not [l for l in out.stdout.splitlines() if l.startswith('{') and '"value"' in l]


def test_fixed_view_workloads_run_only_where_they_shard_evenly():
    """c3 (8 views per step) runs on 1, 2, 4 and 8 ranks; c4 (4 views) on 1, 2 and 4 - on 8 ranks it is skipped instead
    of leaving four ranks without work inside a collective (which dead-locked a round-2 run)."""
    sys.path.insert(0, ROOT)
    import bench
    assert [n for n in (1, 2, 4, 8) if bench.shards_evenly(8, n)] == [1, 2, 4, 8]
    assert [n for n in (1, 2, 3, 4, 8) if bench.shards_evenly(4, n)] == [1, 2, 4]
    assert all(bench.shards_evenly(None, n) for n in (1, 2, 3, 8))


def test_watchdog_prints_what_it_has_and_exits():
    code = ("import sys, time; sys.path.insert(0, %r); import bench; "
            "bench._STATE['line'] = {'metric': 'm', 'value': 1.0}; bench._STATE['section'] = 'workload c4'; "
            "bench._start_watchdog(0.3, 0); time.sleep(30)") % ROOT
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith('{')][-1])
    assert d['value'] == 1.0 and d['aborted']['section'] == 'workload c4'
    code = ("import sys, time; sys.path.insert(0, %r); import bench; bench._start_watchdog(0.3, 0); time.sleep(30)") % ROOT
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 124 and 'watchdog' in out.stdout
