#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/verify_multi_gpu.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
tail -c 400 gpurun_out/r02_bench_n$N.err | grep -v "^\*\*\*\|OMP_NUM"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_n$N.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('n_gpus','ms_per_step','value','result_digest','digest_matches_single','gpu_launches')})
print(d['e2e'])
print(d['roofline']['kernel_ms'], d['roofline']['kernel_share_of_step'])
PY
