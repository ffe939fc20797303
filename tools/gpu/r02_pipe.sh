#!/bin/bash
N=${1:-2}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 "$@" 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$\|NCCL version"; }
echo "== sliced uploads, own results"; CWQ_SLICE=1 CWQ_RESULTS=own run tools/profile_pipeline.py 4 1:6:1 1 | grep -A12 layout
echo "== full uploads, own results"; CWQ_SLICE=0 CWQ_RESULTS=own run tools/profile_pipeline.py 4 | head -1
echo "== full uploads, all results"; CWQ_SLICE=0 CWQ_RESULTS=all run tools/profile_pipeline.py 4 | head -1
echo "== sliced uploads, all results"; CWQ_SLICE=1 CWQ_RESULTS=all run tools/profile_pipeline.py 4 | head -1
