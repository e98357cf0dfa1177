cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_sp_stage.py -x -q --timeout=300 2>&1 | tail -15
timeout 1500 python -m pytest tests -m gpu -q --timeout=300 --deselect tests/test_gpu_sp_stage.py 2>&1 | tail -15
