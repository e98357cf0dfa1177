cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_raster_parity.py -x -q --timeout=300 2>&1 | tail -3
python tools/tile_sort_probe.py c2 2>&1 | tail -2
python tools/tile_sort_probe.py ns 2>&1 | tail -2
timeout 600 ncu --set full --import-source on --clock-control none -k regex:tile_sort_kernel -c 2 -o gpurun_out/r2_tile_sort python tools/tile_sort_probe.py c2 > gpurun_out/ncu_ts.log 2>&1
tail -3 gpurun_out/ncu_ts.log
