#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
tail -c 300 gpurun_out/r02_bench_n1.err
python tools/ncu_is_score.py > gpurun_out/ncu_is_score.log 2>&1
tail -c 600 gpurun_out/ncu_is_score.log
