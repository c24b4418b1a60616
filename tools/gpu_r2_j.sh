#!/bin/bash
# Round 2, call J (gpurun --gpus 8): the data-parallel path at the full NVLink domain. Hardware DP check at world 8, the headline
# bench at N=8 with the own NVLink kernel and with NCCL, BASELINE configs[4] (13B, R=1024) at N=8, and the same box's 1-GPU step.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
OUT=gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a $OUT/r2j_timeline.log; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29731"
B="bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --no-rollout"
timeout 300 $TR tests/dist_gpu_check.py > $OUT/dist8_r2j.log 2>&1; stamp "dist check world 8 rc=$?"
cp $OUT/dist_gpu_check.json $OUT/dist_gpu_check_world8.json 2>/dev/null
timeout 300 $TR $B > $OUT/bench_r2j_nvl8.log 2> $OUT/bench_r2j_nvl8.err; stamp "bench 8 gpu nvl rc=$?"
OPADPO_DP_COLLECTIVE=nccl timeout 300 $TR $B > $OUT/bench_r2j_nccl8.log 2> $OUT/bench_r2j_nccl8.err; stamp "bench 8 gpu nccl rc=$?"
timeout 400 $TR bench.py --gpus 8 --model 13b --resp-len 1024 --pairs 4 --steps 3 --warmup 3 --no-cpu-baseline --no-rollout > $OUT/bench_r2j_13b8.log 2> $OUT/bench_r2j_13b8.err; stamp "bench 13B 8 gpu rc=$?"
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-rollout > $OUT/bench_r2j_1gpu.log 2> $OUT/bench_r2j_1gpu.err; stamp "bench 1 gpu same box rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r2j_*.log')):
    for l in open(f):
        if l.startswith('{"metric"'):
            d=json.loads(l); print(f.split('bench_r2j_')[-1], d['n_gpus'], round(d['ms_per_step'],1), round(d['value'],2), round(d['roofline']['gemm_ms_per_step'],1), d['clocks']['sm_mhz'], d['config']['collective'][:110])
PY
grep -o '"ok": [a-z]*' $OUT/dist_gpu_check_world8.json; tail -n 3 $OUT/dist8_r2j.log | cut -c1-300
stamp done
