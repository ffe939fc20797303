#!/bin/bash
# final round-2 measurements on N GPUs of one box: bench.py (BASELINE configs[3]) and configs[4] (64 x 3840x2160, greedy)
mkdir -p gpurun_out
N=${1:-1}
WHAT=${2:-both}
if [ "$N" = "1" ]; then
  if [ "$WHAT" != "greedy" ]; then
    python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_final_reference_arm.json 2> gpurun_out/r02_final_reference_arm.err
    python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_final_bench_n1.json 2> gpurun_out/r02_final_bench_n1.err
    python tools/ncu_is_score.py > gpurun_out/ncu_is_score.log 2>&1
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_bench.log 2>&1
    python tools/profile_small.py > gpurun_out/r02_final_profile_small.log 2>&1
    python tools/profile_small.py --batch >> gpurun_out/r02_final_profile_small.log 2>&1
  fi
  if [ "$WHAT" != "bench" ]; then
    timeout 900 python tools/bench_greedy_4k.py --images 64 > gpurun_out/r02_final_greedy4k_n1.json 2> gpurun_out/r02_final_greedy4k_n1.err
  fi
else
  run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 "$@"; }
  if [ "$WHAT" != "greedy" ]; then
    run bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_final_bench_n$N.json 2> gpurun_out/r02_final_bench_n$N.err
    run tools/verify_multi_gpu.py 2>&1 | grep "^world" > gpurun_out/r02_final_verify_n$N.log
  fi
  if [ "$WHAT" != "bench" ]; then
    run tools/bench_greedy_4k.py --images 64 > gpurun_out/r02_final_greedy4k_n$N.json 2> gpurun_out/r02_final_greedy4k_n$N.err
  fi
fi
for f in gpurun_out/r02_final_*_n$N.json; do grep "^{" $f | cut -c1-300; done
cat gpurun_out/r02_final_verify_n$N.log 2>/dev/null
exit 0
