#!/bin/bash
mkdir -p gpurun_out
N=${1:-1}
IMAGES=${2:-64}
if [ "$N" = "1" ]; then
  timeout 900 python tools/bench_greedy_4k.py --images $IMAGES > gpurun_out/r02_greedy4k_n$N.json 2> gpurun_out/r02_greedy4k_n$N.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 tools/bench_greedy_4k.py --images $IMAGES > gpurun_out/r02_greedy4k_n$N.json 2> gpurun_out/r02_greedy4k_n$N.err
fi
grep "^{" gpurun_out/r02_greedy4k_n$N.json | cut -c1-700
tail -3 gpurun_out/r02_greedy4k_n$N.err | grep -v "OMP_NUM\|^\*\*\*"
