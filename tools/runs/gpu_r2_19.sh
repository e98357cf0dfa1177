cd $GRAFT_REPO_ROOT
for mode in tile; do
SKGS_SORT=$mode timeout 600 python bench.py --steps 100 --warmup 10 --headline-only 2> gpurun_out/r2_15_$mode.err | grep '^{' > gpurun_out/r2_15_$mode.json
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_15_$mode.json').read())
print('$mode', d['value'], d['ms_per_step'], d['e2e']['value'])
print({k: round(v['us_per_step'],1) for k, v in d.get('kernels', {}).items()})
PY
done
