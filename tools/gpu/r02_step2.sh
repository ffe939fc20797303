#!/bin/bash
mkdir -p gpurun_out
python tools/bench_configs.py --only c1,c2 --reps 20 > gpurun_out/r02_configs_front.log 2>&1
cat gpurun_out/r02_configs_front.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_c2_front.csv python tools/bench_configs.py --only c2 --reps 1 > gpurun_out/ncu_c2.log 2>&1
tail -3 gpurun_out/ncu_c2.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -2
