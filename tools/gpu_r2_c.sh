#!/bin/bash
# Round 2, call C: pipelined attention kernels (fwd S double buffer, dq deep pipeline), weight-streaming decode GEMM +
# fused q|k|v reducer: parity, micro-benches, the bench with the rollout object, launch list of one decode step, ncu of attention.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a gpurun_out/r2c_timeline.log; }
timeout 400 python -m pytest tests/test_attention_gpu.py tests/test_gemm_gpu.py tests/test_engine_gpu.py tests/test_pipeline_gpu.py -q -m gpu > gpurun_out/t_r2c_fast.log 2>&1; stamp "fast suites rc=$?"
timeout 120 python tools/bench_attn.py > gpurun_out/bench_attn_r2c.log 2>&1; stamp "bench_attn rc=$?"
timeout 300 python -m pytest tests/test_fullsize_gpu.py -q -m gpu -k "7b_ or truncated" > gpurun_out/t_r2c_full.log 2>&1; stamp "fullsize (no oracle depth) rc=$?"
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2c.log 2> gpurun_out/bench_r2c.err; stamp "bench rc=$?"
timeout 200 python bench.py --mode rollout --no-cpu-baseline > gpurun_out/bench_rollout_r2c.log 2> gpurun_out/bench_rollout_r2c.err; stamp "bench rollout rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_decode_r2c.csv python tools/bench_rollout.py --new 3 --runs 1 > gpurun_out/ncu_decode_r2c.log 2>&1; stamp "launch list decode rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 4 -c 4 -o gpurun_out/r02_ncu_attn_v2 python tools/bench_attn.py --reps 1 > gpurun_out/ncu_attn_v2.log 2>&1; stamp "ncu attn rc=$?"
tail -4 gpurun_out/t_r2c_fast.log gpurun_out/t_r2c_full.log; cat gpurun_out/bench_attn_r2c.log; cat gpurun_out/bench_rollout_r2c.log | cut -c1-1500
stamp done
