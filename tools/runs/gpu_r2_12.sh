cd $GRAFT_REPO_ROOT
for ex in synced barriers; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --steps 100 --warmup 10 --headline-only --exchange $ex 2>&1 | grep '^{' > gpurun_out/r2_bench2_$ex.json
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench2_$ex.json').read())
print('$ex', d['value'], d['ms_per_step'], d.get('exchange_check'))
PY
done
