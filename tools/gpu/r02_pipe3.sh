#!/bin/bash
N=${1:-8}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 "$@" 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$\|NCCL version"; }
run tools/profile_pipeline.py 1:3 1:23 1:11 1:7 1:15:1 | grep "layout\|per-rank"
