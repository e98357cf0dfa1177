cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_sp_stage.py tests/test_gpu_densify.py tests/test_gpu_fused_path.py -q --timeout=300 2>&1 | tail -30
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu --no-iteration 2> gpurun_out/r2_14_bench.err | grep '^{' > gpurun_out/r2_14_bench.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_14_bench.json').read())
print(d['value'], d['ms_per_step'], d['e2e'])
print(json.dumps(d.get('widening'), indent=1))
PY
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 2> gpurun_out/r2_14_ref.err | grep '^{' > gpurun_out/r2_14_ref.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_14_ref.json').read())
print(d['value'], d['ms_per_step'])
print(json.dumps(d.get('widening'), indent=1))
PY
tail -5 gpurun_out/r2_14_bench.err gpurun_out/r2_14_ref.err
