#!/bin/bash
# Round 2, call B: the optimised attention kernels (parity + micro-bench + ncu), the new sampler / suppress / slot tests,
# the headline-path full-depth parity with delta diagnostics, and the bench with the rollout object.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a gpurun_out/r2b_timeline.log; }
timeout 300 python -m pytest tests/test_attention_gpu.py tests/test_gemm_gpu.py tests/test_engine_gpu.py tests/test_pipeline_gpu.py -q -m gpu > gpurun_out/t_r2b_fast.log 2>&1; stamp "fast suites rc=$?"
timeout 120 python tools/bench_attn.py > gpurun_out/bench_attn_r2b.log 2>&1; stamp "bench_attn rc=$?"
timeout 120 python tools/bench_attn.py --stacked >> gpurun_out/bench_attn_r2b.log 2>&1; stamp "bench_attn stacked rc=$?"
timeout 400 python -m pytest tests/test_fullsize_gpu.py -q -m gpu -k "headline or 7b_" > gpurun_out/t_r2b_full.log 2>&1; stamp "fullsize rc=$?"
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2b.log 2> gpurun_out/bench_r2b.err; stamp "bench rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 4 -c 4 -o gpurun_out/r02_ncu_attn python tools/bench_attn.py --reps 1 > gpurun_out/ncu_attn.log 2>&1; stamp "ncu attn rc=$?"
tail -5 gpurun_out/t_r2b_fast.log gpurun_out/t_r2b_full.log; cat gpurun_out/bench_attn_r2b.log
stamp done
