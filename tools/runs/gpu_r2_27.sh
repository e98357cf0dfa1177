cd $GRAFT_REPO_ROOT
for b in 148 296 444 592; do
echo "== SKGS_MM_BLOCKS=$b"
SKGS_MM_BLOCKS=$b timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 tools/mm_test.py 2>&1 | grep "multimem kernel only\|multimem arena"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29570 bench.py --gpus 8 --steps 100 --warmup 10 --headline-only 2> gpurun_out/r2_s8b.err | grep '^{' > gpurun_out/r2_s8b.json
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_s8b.json').read())
c=d.get('exchange_check') or {}
print('N=8', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'check', c.get('ok'), c.get('max_rel_err_vs_nccl_allreduce'))
PY
