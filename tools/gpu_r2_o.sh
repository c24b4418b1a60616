#!/bin/bash
# Round 2, call O (1 GPU): delta = rowsum(dO*O) computed inside the dQ kernel (attn_delta_kernel removed).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
OUT=gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a $OUT/r2o_timeline.log; }
timeout 300 python -m pytest tests/test_attention_gpu.py -q -m gpu > $OUT/t_r2o_attn.log 2>&1; stamp "attention suite rc=$?"
timeout 120 python tools/bench_attn.py > $OUT/bench_attn_r2o.log 2>&1; stamp "bench_attn rc=$?"
timeout 400 python -m pytest tests/test_engine_gpu.py tests/test_pipeline_gpu.py -q -m gpu > $OUT/t_r2o_fast.log 2>&1; stamp "engine suites rc=$?"
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-rollout > $OUT/bench_r2o.log 2> $OUT/bench_r2o.err; stamp "bench rc=$?"
tail -n 3 $OUT/t_r2o_attn.log $OUT/t_r2o_fast.log; cat $OUT/bench_attn_r2o.log; tail -n 1 $OUT/bench_r2o.log | cut -c1-300
stamp done
