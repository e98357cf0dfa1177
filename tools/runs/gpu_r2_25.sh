cd $GRAFT_REPO_ROOT
for pdl in 1 0; do
SKGS_CHECK_REPS=10 SKGS_PDL=$pdl timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 4 --steps 50 --warmup 5 --headline-only 2>&1 | grep '^{' > gpurun_out/r2_b4_pdl$pdl.json
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_b4_pdl$pdl.json').read())
c=d.get('exchange_check')
print('pdl=$pdl', d['value'], d['ms_per_step'], c['ok'], c['max_rel_err_vs_nccl_allreduce'], c['bad_blocks_rank0'])
PY
done
