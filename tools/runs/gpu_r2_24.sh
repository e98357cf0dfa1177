cd $GRAFT_REPO_ROOT
for n in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29570 bench.py --gpus $n --steps 100 --warmup 10 --headline-only 2> gpurun_out/r2_scale_$n.err | grep '^{' > gpurun_out/r2_scale_$n.json
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_scale_$n.json').read())
c=d.get('exchange_check') or {}
print($n, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'check', c.get('ok'), c.get('max_rel_err_vs_nccl_allreduce'), d['config'].get('parallelism'))
PY
done
