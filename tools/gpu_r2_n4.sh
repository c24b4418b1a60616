#!/bin/bash
# Round 2 (gpurun --gpus 4): the N=4 bench line of the final code (N=2 and N=8 were run in calls W and R).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29761 bench.py --gpus 4 --steps 4 --warmup 3 --no-cpu-baseline --no-rollout > gpurun_out/bench_r2_4gpu.log 2> gpurun_out/bench_r2_4gpu.err; echo "rc=$?"
tail -n 1 gpurun_out/bench_r2_4gpu.log | cut -c1-260; grep -o '"collective": "[^"]*"' gpurun_out/bench_r2_4gpu.log | cut -c1-200; tail -n 2 gpurun_out/bench_r2_4gpu.err | cut -c1-200
