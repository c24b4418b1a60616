#!/bin/bash
# Round 2, call N (1 GPU): LoRA weight gradients on a side stream: parity suites (incl. the A/B test) and a same-box bench A/B.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
OUT=gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a $OUT/r2n_timeline.log; }
timeout 400 python -m pytest tests/test_engine_gpu.py tests/test_pipeline_gpu.py tests/test_gemm_gpu.py -q -m gpu > $OUT/t_r2n_fast.log 2>&1; stamp "engine suites rc=$?"
B="bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-rollout"
timeout 300 python $B --wgrad-stream 1 > $OUT/bench_r2n_side.log 2> $OUT/bench_r2n_side.err; stamp "bench side rc=$?"
timeout 300 python $B --wgrad-stream 0 > $OUT/bench_r2n_main.log 2> $OUT/bench_r2n_main.err; stamp "bench main rc=$?"
timeout 300 python $B --wgrad-stream 1 > $OUT/bench_r2n_side2.log 2> $OUT/bench_r2n_side2.err; stamp "bench side again rc=$?"
timeout 300 python $B --wgrad-stream 0 > $OUT/bench_r2n_main2.log 2> $OUT/bench_r2n_main2.err; stamp "bench main again rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r2n_*.log')):
    for l in open(f):
        if l.startswith('{"metric"'):
            d=json.loads(l); print(f.split('bench_r2n_')[-1], round(d['ms_per_step'],1), round(d['e2e']['ms_per_step'],1), round(d['roofline']['gemm_ms_per_step'],1), d['clocks']['sm_mhz'], d['config']['loss'], d['config']['grad_norm'])
PY
tail -n 5 $OUT/t_r2n_fast.log
stamp done
