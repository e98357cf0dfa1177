cd $GRAFT_REPO_ROOT
for n in 8 4; do
start=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2958$n bench.py --gpus $n --steps 100 --warmup 10 > gpurun_out/r2_full$n.out 2> gpurun_out/r2_full$n.err
echo n=$n rc=$? wall=$(( $(date +%s) - start ))s
grep '^{' gpurun_out/r2_full$n.out > gpurun_out/r2_full$n.json
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_full$n.json').read())
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['exchange_check']['ok'])
for k,v in (d.get('workloads') or {}).items(): print(k, {a:b for a,b in v.items() if a in ('value','unit','ms_per_step','scaling','views_per_rank','error')}, (v.get('exchange_check') or {}).get('ok'))
PY
done
