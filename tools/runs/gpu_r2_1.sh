cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -q --timeout=240 -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/r2_gputest_1.log
tail -40 gpurun_out/r2_gputest_1.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/r2_bench_1.json 2> gpurun_out/r2_bench_1.err; tail -5 gpurun_out/r2_bench_1.err
python tools/show_bench.py < gpurun_out/r2_bench_1.json 2>/dev/null | head -60
