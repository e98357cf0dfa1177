cd $GRAFT_REPO_ROOT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tools/mm_test.py 2>&1 | grep -v "^W\|Warning\|warn\|\*\*\*\|OMP_NUM" | tail -15 | tee gpurun_out/r2_mm_test_2gpu_b.txt
