#!/bin/bash
# Second GPU check of the round: collator kernels (second pass) parity + micro-bench + ncu capture, the bench in its final
# default configuration, the step's launch list, and the 13B shape on one GPU.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a gpurun_out/timeline2.log; }
stamp "start"
timeout 420 python -m pytest tests/test_collate_gpu.py -q -m gpu --durations=5 > gpurun_out/t_collate2.log 2>&1; stamp "collate tests rc=$?"
timeout 200 python tools/bench_collate.py > gpurun_out/bench_collate2.log 2> gpurun_out/bench_collate2.err; stamp "bench_collate rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"resample|collate_responses|mask_fill|image_sum|image_mean" -c 30 \
  -o gpurun_out/r01_ncu_collate_v2 python tools/bench_collate.py --iters 1 > gpurun_out/ncu_collate2.log 2>&1; stamp "ncu collate rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; stamp "bench final rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 16000 --csv --log-file gpurun_out/launches_step_final.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/launches_step_final.log 2>&1; stamp "launch list rc=$?"
timeout 300 python bench.py --model 13b --pairs 4 --resp-len 1024 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_13b.log 2> gpurun_out/bench_13b.err; stamp "bench 13b rc=$?"
stamp "done"
