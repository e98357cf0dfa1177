cd $GRAFT_REPO_ROOT
timeout 1700 python -m pytest tests -m gpu -q --timeout=300 -x 2>&1 | tail -4
timeout 900 python bench.py 2> gpurun_out/r2_23_bench.err | grep '^{' > gpurun_out/r2_23_bench.json
python tools/show_bench.py < gpurun_out/r2_23_bench.json 2>&1 | tail -60
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 2> gpurun_out/r2_23_ref.err | grep '^{' > gpurun_out/r2_23_ref.json
python tools/show_bench.py < gpurun_out/r2_23_ref.json 2>&1 | tail -40
