cd $GRAFT_REPO_ROOT
python tools/pdl_probe.py 2>&1 | tee gpurun_out/r2_pdl_probe.txt
timeout 900 python tools/ab_variants.py c2 2>&1 | tee gpurun_out/r2_ab_5.txt
timeout 300 python -m pytest tests/test_gpu_raster_parity.py tests/test_gpu_reference_ext.py -m gpu -q --timeout=240 -p no:cacheprovider > gpurun_out/r2_gputest_5.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/r2_gputest_5.log | tail -12
