#!/bin/bash
# Round 2, call H (1 GPU): forward attention with P in tensor memory (tcgen05.mma A operand from TMEM) + split K/V rings.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
OUT=gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a $OUT/r2i_timeline.log; }
timeout 300 python -m pytest tests/test_attention_gpu.py -q -m gpu > $OUT/t_r2i_attn.log 2>&1; stamp "attention suite rc=$?"
timeout 120 python tools/bench_attn.py > $OUT/bench_attn_r2i.log 2>&1; stamp "bench_attn rc=$?"
OPADPO_ATTN_FWD=ss timeout 120 python tools/bench_attn.py --only fwd > $OUT/bench_attn_r2i_ss.log 2>&1; stamp "bench_attn legacy fwd rc=$?"
timeout 400 python -m pytest tests/test_gemm_gpu.py tests/test_engine_gpu.py tests/test_pipeline_gpu.py -q -m gpu > $OUT/t_r2i_fast.log 2>&1; stamp "engine suites rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-rollout > $OUT/bench_r2i.log 2> $OUT/bench_r2i.err; stamp "bench rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_ -c 9 -o $OUT/r02_ncu_attn_v5 python tools/bench_attn.py --reps 1 > $OUT/ncu_attn_v5.log 2>&1; stamp "ncu attn rc=$?"
tail -n 4 $OUT/t_r2i_attn.log $OUT/t_r2i_fast.log; cat $OUT/bench_attn_r2i.log $OUT/bench_attn_r2i_ss.log; tail -n 1 $OUT/bench_r2i.log | cut -c1-400
stamp done
