"""Pretty-print the JSON line(s) bench.py wrote to stdin."""
import json, sys
for line in sys.stdin:
    line = line.strip()
    if not line.startswith('{'):
        continue
    d = json.loads(line)
    print(f"impl={d.get('impl','ours')} value={d['value']} {d['unit']} ms/step={d['ms_per_step']} n_gpus={d['n_gpus']} "
          f"e2e={d['e2e']['value']} launches={d.get('gpu_launches')} clocks={d.get('clocks')}")
    if d.get('roofline'):
        print('roofline:', {k: d['roofline'][k] for k in ('kernel', 'achieved', 'frac')})
    tot = 0.0
    for k, v in (d.get('kernels') or {}).items():
        tot += v['us_per_step']
        print(f"  {k:26s} {v['us_per_step']:9.1f} us/step  x{v['launches_per_step']:<4} {v['algorithmic_GBps']:8.1f} GB/s  {v['frac_of_peak']:.3f}")
    if tot:
        print(f"  {'sum of kernels':26s} {tot:9.1f} us/step")
    if d.get('full_iteration'):
        print('full_iteration:', d['full_iteration'])
    if d.get('render_fps'):
        print('render_fps:', d['render_fps']['value'])
    if d.get('cpu_baseline'):
        print('cpu_baseline:', d['cpu_baseline']['value'], d['cpu_baseline']['unit'], 'cores', d['cpu_baseline']['cores'], d['cpu_baseline']['kind'])
