#!/bin/bash
# Round 2, call P (1 GPU): inverse RoPE of dq / dk fused into the attention backward epilogues.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
OUT=gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a $OUT/r2p_timeline.log; }
timeout 300 python -m pytest tests/test_attention_gpu.py -q -m gpu > $OUT/t_r2p_attn.log 2>&1; stamp "attention suite rc=$?"
timeout 120 python tools/bench_attn.py > $OUT/bench_attn_r2p.log 2>&1; stamp "bench_attn rc=$?"
timeout 400 python -m pytest tests/test_engine_gpu.py tests/test_pipeline_gpu.py -q -m gpu > $OUT/t_r2p_fast.log 2>&1; stamp "engine suites rc=$?"
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-rollout > $OUT/bench_r2p.log 2> $OUT/bench_r2p.err; stamp "bench rc=$?"
tail -n 3 $OUT/t_r2p_attn.log $OUT/t_r2p_fast.log; cat $OUT/bench_attn_r2p.log; tail -n 1 $OUT/bench_r2p.log | cut -c1-300
stamp done
