cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_raster_parity.py tests/test_gpu_fused_path.py tests/test_gpu_fullsize_properties.py -m gpu -q --timeout=240 -p no:cacheprovider > gpurun_out/r2_gputest_9.log 2>&1
grep -E "passed|failed|FAILED|Error" gpurun_out/r2_gputest_9.log | tail -8
python tools/sort_trace.py c2 2>&1 | tee gpurun_out/r2_sort_trace2.txt
timeout 600 python tools/ab_variants.py c2 2>&1 | tee gpurun_out/r2_ab_9.txt
