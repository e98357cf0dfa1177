cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_raster_parity.py -x -q --timeout=300 2>&1 | tail -2
python tools/tile_sort_probe.py c2 2>&1 | tail -1
python tools/tile_sort_probe.py ns 2>&1 | tail -1
timeout 300 python bench.py --steps 100 --warmup 10 --headline-only 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
