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
