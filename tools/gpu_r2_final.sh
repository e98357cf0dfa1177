cd $GRAFT_REPO_ROOT
timeout 1700 python -m pytest tests -m gpu -q --timeout=300 2>&1 | tail -3
timeout 900 python bench.py 2> gpurun_out/r2_final_bench.err | grep '^{' > gpurun_out/r2_final_bench.json
python tools/show_bench.py < gpurun_out/r2_final_bench.json 2>&1 | head -16
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 2> gpurun_out/r2_final_ref.err | grep '^{' > gpurun_out/r2_final_ref.json
python tools/show_bench.py < gpurun_out/r2_final_ref.json 2>&1 | head -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
