#!/bin/bash
# Round 2, call E (gpurun --gpus 2): the NVLink all-reduce kernel (csrc/collective.cu) on hardware: kernel-level exactness on both
# transports + the 2-GPU DP step against the 1-GPU step (tests/dist_gpu_check.py), then the 2-GPU bench with it and with NCCL.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
OUT=gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a $OUT/r2e_timeline.log; }
nvidia-smi topo -m > $OUT/r2e_topo.txt 2>&1
timeout 500 python -m pytest tests/test_dist_gpu.py -m gpu -q -x > $OUT/t_r2e_dist.log 2>&1; stamp "dist test rc=$?"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711"
timeout 400 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-rollout > $OUT/bench_r2e_nvl.log 2> $OUT/bench_r2e_nvl.err; stamp "bench 2gpu nvl rc=$?"
OPADPO_DP_WIRE=bf16 timeout 400 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-rollout > $OUT/bench_r2e_bf16wire.log 2> $OUT/bench_r2e_bf16wire.err; stamp "bench 2gpu bf16 wire rc=$?"
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-rollout > $OUT/bench_r2e_1gpu.log 2> $OUT/bench_r2e_1gpu.err; stamp "bench 1gpu same box rc=$?"
tail -5 $OUT/t_r2e_dist.log; cat $OUT/dist_gpu_check.json | head -60; tail -2 $OUT/bench_r2e_nvl.log | cut -c1-1800; tail -3 $OUT/bench_r2e_nvl.err; tail -2 $OUT/bench_r2e_bf16wire.log | cut -c1-300; tail -1 $OUT/bench_r2e_1gpu.log | cut -c1-300
stamp done
