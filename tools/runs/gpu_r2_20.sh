cd $GRAFT_REPO_ROOT
timeout 1700 python -m pytest tests -m gpu -q --timeout=300 -x 2>&1 | tail -8
