#!/bin/bash
# Round 2, call W (gpurun --gpus 2): final code, data-parallel sanity: hardware DP check (pytest launcher) and the N=2 bench line.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
OUT=gpurun_out
timeout 500 python -m pytest tests/test_dist_gpu.py -m gpu -q -x > $OUT/t_r2w_dist.log 2>&1; echo "dist rc=$?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29751 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-rollout > $OUT/bench_r2w_2gpu.log 2> $OUT/bench_r2w_2gpu.err; echo "bench rc=$?"
tail -n 2 $OUT/t_r2w_dist.log; tail -n 1 $OUT/bench_r2w_2gpu.log | cut -c1-300; grep -o '"collective": "[^"]*"' $OUT/bench_r2w_2gpu.log | cut -c1-250
