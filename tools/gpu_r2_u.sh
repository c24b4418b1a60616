#!/bin/bash
# Round 2, call U (1 GPU): last code changes (32-bit index arithmetic in the SwiGLU kernels, checked side-stream plumbing):
# the suites that cover them, smoke, and the default bench line of the round's final code.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
OUT=gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a $OUT/r2u_timeline.log; }
timeout 500 python -m pytest tests/test_gemm_gpu.py tests/test_engine_gpu.py tests/test_pipeline_gpu.py tests/test_attention_gpu.py tests/test_collate_gpu.py -q -m gpu > $OUT/t_r2u_fast.log 2>&1; stamp "suites rc=$?"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_r2u.log 2>&1; stamp "smoke rc=$?"
timeout 500 python bench.py > $OUT/bench_r2u.log 2> $OUT/bench_r2u.err; stamp "default bench rc=$?"
tail -n 3 $OUT/t_r2u_fast.log; tail -n 1 $OUT/smoke_r2u.log; tail -n 1 $OUT/bench_r2u.log | cut -c1-300
stamp done
