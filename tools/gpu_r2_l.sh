#!/bin/bash
# Round 2, call L (1 GPU): reference pass on a side stream next to the policy forward (OPADPO_CONCURRENT_REF=1) against the
# sequential step on the same box; RMSNorm backward with the residual gradient prefetched; engine suites.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
OUT=gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a $OUT/r2l_timeline.log; }
timeout 400 python -m pytest tests/test_gemm_gpu.py tests/test_engine_gpu.py tests/test_pipeline_gpu.py tests/test_attention_gpu.py -q -m gpu > $OUT/t_r2l_fast.log 2>&1; stamp "engine suites rc=$?"
B="bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-rollout"
timeout 300 python $B > $OUT/bench_r2l_seq.log 2> $OUT/bench_r2l_seq.err; stamp "bench sequential rc=$?"
OPADPO_CONCURRENT_REF=1 timeout 300 python $B > $OUT/bench_r2l_conc.log 2> $OUT/bench_r2l_conc.err; stamp "bench concurrent ref rc=$?"
timeout 300 python $B > $OUT/bench_r2l_seq2.log 2> $OUT/bench_r2l_seq2.err; stamp "bench sequential again rc=$?"
OPADPO_CONCURRENT_REF=1 timeout 300 python $B > $OUT/bench_r2l_conc2.log 2> $OUT/bench_r2l_conc2.err; stamp "bench concurrent ref again rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r2l_*.log')):
    for l in open(f):
        if l.startswith('{"metric"'):
            d=json.loads(l); print(f.split('bench_r2l_')[-1], round(d['ms_per_step'],1), round(d['e2e']['ms_per_step'],1), round(d['roofline']['gemm_ms_per_step'],1), d['clocks']['sm_mhz'], d['config']['loss'], d['config']['grad_norm'])
PY
tail -n 3 $OUT/t_r2l_fast.log
stamp done
