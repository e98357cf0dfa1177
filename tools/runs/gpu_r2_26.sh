cd $GRAFT_REPO_ROOT
for noex in 1 0; do
SKGS_BENCH_NO_EXCHANGE=$noex timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29570 bench.py --gpus 8 --steps 100 --warmup 10 --headline-only 2> gpurun_out/r2_s8_noex$noex.err | grep '^{' > gpurun_out/r2_s8_noex$noex.json
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_s8_noex$noex.json').read())
c=d.get('exchange_check') or {}
print('noex=$noex', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'check', c.get('ok'))
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 tools/mm_test.py 2>&1 | grep -v "^W\|Warning\|warn\|\*\*\*\|OMP_NUM" | tail -8
