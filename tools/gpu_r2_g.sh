#!/bin/bash
# Round 2, call G (1 GPU): transposed dK/dV attention-backward kernel and the 4-key-unrolled decode attention: parity suites,
# attention micro-bench, step and rollout bench, ncu --set full of the three attention kernels (with source), launch list.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
OUT=gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a $OUT/r2g_timeline.log; }
timeout 300 python -m pytest tests/test_attention_gpu.py -q -m gpu -x > $OUT/t_r2g_attn.log 2>&1; stamp "attention suite rc=$?"
timeout 120 python tools/bench_attn.py > $OUT/bench_attn_r2g.log 2>&1; stamp "bench_attn rc=$?"
timeout 400 python -m pytest tests/test_gemm_gpu.py tests/test_engine_gpu.py tests/test_pipeline_gpu.py -q -m gpu > $OUT/t_r2g_fast.log 2>&1; stamp "engine suites rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-rollout > $OUT/bench_r2g.log 2> $OUT/bench_r2g.err; stamp "bench rc=$?"
timeout 200 python bench.py --mode rollout --no-cpu-baseline > $OUT/bench_rollout_r2g.log 2> $OUT/bench_rollout_r2g.err; stamp "bench rollout rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_ -c 5 -o $OUT/r02_ncu_attn_v3 python tools/bench_attn.py --reps 1 > $OUT/ncu_attn_v3.log 2>&1; stamp "ncu attn rc=$?"
tail -4 $OUT/t_r2g_attn.log $OUT/t_r2g_fast.log; cat $OUT/bench_attn_r2g.log; tail -1 $OUT/bench_r2g.log | cut -c1-400; tail -1 $OUT/bench_rollout_r2g.log | cut -c1-300
stamp done
