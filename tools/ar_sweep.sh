#!/bin/bash
# NCCL all-reduce of the 26.8 MB gradient arena under different algorithm / protocol choices
N=${1:-8}
for cfg in "" "NCCL_ALGO=NVLS" "NCCL_ALGO=Ring" "NCCL_ALGO=Tree" "NCCL_ALGO=NVLS NCCL_NVLS_CHUNKSIZE=524288" "NCCL_PROTO=LL128" "NCCL_MIN_NCHANNELS=32" "NCCL_ALGO=Ring NCCL_MIN_NCHANNELS=32"; do
  echo "== $cfg"
  env $cfg python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 tools/ar_test.py 2>&1 | grep -E "nccl all_reduce  |two_shot"
done
