#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/one.py <<'PY'
import os, sys
import numpy as np, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from compression_without_quantization_b200 import engine as E
from compression_without_quantization_b200.synthetic import synthetic_posterior
dev = E.cuda_device()
a = synthetic_posterior(1 << 20, 77)
t = [torch.from_numpy(x).to(dev) for x in a]
specs = [E.SegmentSpec(1 << 20, 42, 12, 4, 10, True)]
for _ in range(3):
    job = E.ISJob(*t, specs).launch(7)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:"k_front_tiles|k_front_groups|k_front_chain|k_finish_dims|k_pack_elias" -s 10 -c 5 -o gpurun_out/r02_front_full -f python /tmp/one.py > gpurun_out/ncu_front.log 2>&1
tail -3 gpurun_out/ncu_front.log
ls -la gpurun_out/
