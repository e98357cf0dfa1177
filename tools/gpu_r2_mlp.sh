cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_joint_mlp.py tests/test_gpu_loss_adam.py -x -q --timeout=300 2>&1 | tail -4
echo "== mega"; timeout 120 python tools/mlp_bench.py 2>&1 | tail -3
echo "== launches"; SKGS_MLP_MEGA=0 timeout 120 python tools/mlp_bench.py 2>&1 | tail -3
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu --no-workloads 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['full_iteration']['value'], d['full_iteration']['ms_per_iteration'])
print({k: round(v['us_per_step'],1) for k,v in d['kernels'].items() if 'joint' in k})"
