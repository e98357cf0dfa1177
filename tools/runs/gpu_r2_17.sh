cd $GRAFT_REPO_ROOT
python tools/tile_sort_probe.py c2 2>&1 | tail -3
SKGS_CELL_STRIDE=8 python tools/tile_sort_probe.py c2 2>&1 | tail -2
SKGS_CELL_STRIDE=32 python tools/tile_sort_probe.py c2 2>&1 | tail -2
python tools/tile_sort_probe.py ns 2>&1 | tail -2
SKGS_SORT=radix python tools/tile_sort_probe.py ns 2>&1 | tail -2
