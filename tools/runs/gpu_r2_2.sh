cd $GRAFT_REPO_ROOT
rm -f gpurun_out/parity_vs_reference.jsonl
timeout 900 python -m pytest tests/test_gpu_reference_ext.py tests/test_gpu_fullsize_properties.py -m gpu -q --timeout=240 -p no:cacheprovider > gpurun_out/r2_gputest_2.log 2>&1
grep -E "passed|failed" gpurun_out/r2_gputest_2.log | tail -3
timeout 600 python bench.py --steps 50 --warmup 5 --cpu-steps 5 > gpurun_out/r2_bench_2.json 2> gpurun_out/r2_bench_2.err; tail -5 gpurun_out/r2_bench_2.err
python tools/show_bench.py < gpurun_out/r2_bench_2.json 2>/dev/null | head -70
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_ref_2.json 2> gpurun_out/r2_bench_ref_2.err; tail -5 gpurun_out/r2_bench_ref_2.err
python tools/show_bench.py < gpurun_out/r2_bench_ref_2.json 2>/dev/null | head
