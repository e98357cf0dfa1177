cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_raster_parity.py tests/test_gpu_reference_ext.py tests/test_gpu_fused_path.py tests/test_gpu_fullsize_properties.py -x -q --timeout=300 2>&1 | tail -15
for mode in tile radix; do
SKGS_SORT=$mode timeout 600 python bench.py --steps 100 --warmup 10 --headline-only 2> gpurun_out/r2_15_$mode.err | grep '^{' > gpurun_out/r2_15_$mode.json
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_15_$mode.json').read())
print('$mode', d['value'], d['ms_per_step'], d['e2e']['value'])
print({k: round(v['us'],1) if isinstance(v, dict) and 'us' in v else v for k, v in d.get('kernels', {}).items()})
PY
done
tail -3 gpurun_out/r2_15_tile.err
