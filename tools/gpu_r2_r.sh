#!/bin/bash
# Round 2, call R (gpurun --gpus 8): final data-parallel numbers of the round's code (side-stream weight gradients, concurrent
# reference pass, own NVLink all-reduce): world-8 DP check, N=8 bench, the same with the all-reduce skipped (what max-over-ranks
# timing alone costs), and one GPU of the same box.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
OUT=gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a $OUT/r2r_timeline.log; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29741"
B="bench.py --gpus 8 --steps 6 --warmup 3 --no-cpu-baseline --no-rollout"
timeout 300 $TR tests/dist_gpu_check.py > $OUT/dist8_r2r.log 2>&1; stamp "dist check world 8 rc=$?"
cp $OUT/dist_gpu_check.json $OUT/dist_gpu_check_world8.json 2>/dev/null
timeout 300 $TR $B > $OUT/bench_r2r_nvl8.log 2> $OUT/bench_r2r_nvl8.err; stamp "bench 8 gpu nvl rc=$?"
timeout 300 $TR $B --diag-no-collective > $OUT/bench_r2r_nocoll8.log 2> $OUT/bench_r2r_nocoll8.err; stamp "bench 8 gpu no collective rc=$?"
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-rollout > $OUT/bench_r2r_1gpu.log 2> $OUT/bench_r2r_1gpu.err; stamp "bench 1 gpu same box rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r2r_*.log')):
    for l in open(f):
        if l.startswith('{"metric"'):
            d=json.loads(l); print(f.split('bench_r2r_')[-1], d['n_gpus'], round(d['ms_per_step'],1), round(d['value'],2), round(d['roofline']['gemm_ms_per_step'],1), d['clocks']['sm_mhz'], d['config']['collective'][:110])
PY
grep -o '"ok": [a-z]*' $OUT/dist_gpu_check_world8.json; tail -n 2 $OUT/bench_r2r_nvl8.err | cut -c1-300
stamp done
