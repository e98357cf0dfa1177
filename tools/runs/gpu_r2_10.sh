cd $GRAFT_REPO_ROOT
nvidia-smi -L | head -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tools/mm_test.py 2>&1 | grep -v "^W\|Warning\|warn" | tail -15 | tee gpurun_out/r2_mm_test_2gpu.txt
for ex in synced barriers; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --steps 50 --warmup 5 --headline-only --exchange $ex > gpurun_out/r2_bench_2gpu_$ex.json 2> gpurun_out/r2_bench_2gpu_$ex.err; tail -2 gpurun_out/r2_bench_2gpu_$ex.err
python tools/show_bench.py < gpurun_out/r2_bench_2gpu_$ex.json 2>/dev/null | head -3
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_2gpu_$ex.json') if l.startswith('{')][0]); print('$ex', d['value'], d['ms_per_step'], d.get('exchange_check'))"
done
timeout 300 python bench.py --steps 50 --warmup 5 --headline-only > gpurun_out/r2_bench_1gpu_10.json 2>/dev/null; python tools/show_bench.py < gpurun_out/r2_bench_1gpu_10.json | head -3
