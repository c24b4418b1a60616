#!/bin/bash
# Round 2, call Y (1 GPU): dQ kernel loads its O tile on its own barrier behind the first K/V stage.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
OUT=gpurun_out
timeout 300 python -m pytest tests/test_attention_gpu.py tests/test_engine_gpu.py -q -m gpu > $OUT/t_r2y.log 2>&1; echo "suites rc=$?"
timeout 100 python tools/bench_attn.py > $OUT/bench_attn_r2y.log 2>&1; cat $OUT/bench_attn_r2y.log; tail -n 2 $OUT/t_r2y.log
