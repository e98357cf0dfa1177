cd $GRAFT_REPO_ROOT
python tools/sort_trace.py c2 2>&1 | tee gpurun_out/r2_sort_trace.txt
timeout 600 python tools/ab_variants.py c2 2>&1 | tee gpurun_out/r2_ab_8.txt
