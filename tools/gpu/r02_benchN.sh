#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_n$N.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('n_gpus','ms_per_step','value','result_digest','digest_matches_single','gpu_launches')})
print('e2e ms', d['e2e']['ms_per_step'], 'kernel ms', d['roofline']['kernel_ms'], 'share', d['roofline']['kernel_share_of_step'])
PY
