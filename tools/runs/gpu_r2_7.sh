cd $GRAFT_REPO_ROOT
rm -f gpurun_out/parity_vs_reference.jsonl
timeout 1500 python -m pytest tests -m gpu -q --timeout=240 -p no:cacheprovider > gpurun_out/r2_gputest_7.log 2>&1
grep -E "passed|failed|FAILED|Error" gpurun_out/r2_gputest_7.log | tail -15
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --no-workloads > gpurun_out/r2_bench_7.json 2> gpurun_out/r2_bench_7.err; tail -3 gpurun_out/r2_bench_7.err
python tools/show_bench.py < gpurun_out/r2_bench_7.json 2>/dev/null | head -40
