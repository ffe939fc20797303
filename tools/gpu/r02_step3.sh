#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_c2_front2.csv python tools/bench_configs.py --only c2 --reps 1 > gpurun_out/ncu_c2.log 2>&1
tail -2 gpurun_out/ncu_c2.log
timeout 200 python tools/gpu/dbg_b2b.py
