#!/bin/bash
# Round 2, call F (gpurun --gpus 2): sensitivity of the 2-GPU step to how the NVLink all-reduce kernel is scheduled
# (overlapped under the backward with few CTAs vs one wide launch afterwards), NCCL on the same box for reference.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
OUT=gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a $OUT/r2f_timeline.log; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711"
B="bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-rollout"
run() { name=$1; shift; env "$@" timeout 300 $TR $B > $OUT/bench_r2f_$name.log 2> $OUT/bench_r2f_$name.err; stamp "$name rc=$?"; }
run nooverlap256 OPADPO_DP_OVERLAP=0 OPADPO_DP_TAIL_BLOCKS=256
run overlap_b4 OPADPO_DP_BLOCKS=4
run overlap_b16_g8 OPADPO_DP_BLOCKS=16 OPADPO_DP_GROUP_LAYERS=8
run overlap_b48 OPADPO_DP_BLOCKS=48
run nccl OPADPO_DP_COLLECTIVE=nccl
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-rollout > $OUT/bench_r2f_1gpu.log 2> $OUT/bench_r2f_1gpu.err; stamp "1gpu rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r2f_*.log')):
    for l in open(f):
        if l.startswith('{"metric"'):
            d=json.loads(l); print(f.split('bench_r2f_')[-1], d['n_gpus'], round(d['ms_per_step'],1), round(d['value'],2), round(d['roofline']['gemm_ms_per_step'],1), d['clocks']['sm_mhz'], d['config']['collective'][:100])
PY
stamp done
