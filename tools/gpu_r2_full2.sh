cd $GRAFT_REPO_ROOT
start=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29580 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r2_full2.out 2> gpurun_out/r2_full2.err
echo rc=$? wall=$(( $(date +%s) - start ))s
grep '^{' gpurun_out/r2_full2.out > gpurun_out/r2_full2.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_full2.json').read())
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['exchange_check']['ok'])
for k,v in (d.get('workloads') or {}).items(): print(k, {a:b for a,b in v.items() if a in ('value','unit','ms_per_step','scaling','views_per_rank','error')}, (v.get('exchange_check') or {}).get('ok'))
PY
tail -3 gpurun_out/r2_full2.err
