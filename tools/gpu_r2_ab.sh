cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_raster_parity.py -x -q --timeout=300 -k "large_tile_grid or binning" 2>&1 | tail -2
timeout 900 python tools/ab_variants.py 2>&1 | tail -12
