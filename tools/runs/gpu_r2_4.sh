cd $GRAFT_REPO_ROOT
rm -f gpurun_out/parity_vs_reference.jsonl
timeout 1200 python -m pytest tests -m gpu -q --timeout=240 -p no:cacheprovider > gpurun_out/r2_gputest_4.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/r2_gputest_4.log | tail -12
timeout 600 python tools/ab_variants.py c2 2>&1 | tee gpurun_out/r2_ab_4.txt
