#!/bin/bash
# round-2 baseline: GPU tests of the tree, per-config timings and a launch list of the small-K configs
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/bench_configs.py --only c1,c2 --reps 5 > gpurun_out/r02_configs_base.log 2>&1
cat gpurun_out/r02_configs_base.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_c2_base.csv python tools/bench_configs.py --only c2 --reps 1 > gpurun_out/ncu_c2.log 2>&1
tail -3 gpurun_out/ncu_c2.log
