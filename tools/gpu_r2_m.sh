#!/bin/bash
# Round 2, call M (1 GPU): concurrent-reference default + its test, fresh --set full captures of the dominant GEMM launches
# (merged gate|up forward with the SwiGLU epilogue, policy gate|up forward, one layer's dgrad launches), default bench line.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
OUT=gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a $OUT/r2m_timeline.log; }
timeout 400 python -m pytest tests/test_engine_gpu.py tests/test_pipeline_gpu.py -q -m gpu > $OUT/t_r2m_fast.log 2>&1; stamp "engine suites rc=$?"
B="python bench.py --profile --steps 1 --warmup 1 --no-cpu-baseline --no-rollout"
FULL="--set full --clock-control none --kernel-name-base demangled"
OPADPO_CONCURRENT_REF=0 timeout 300 ncu $FULL -k "regex:gemm_kernel<\(bool\)0, \(bool\)0, \(int\)6," -s 40 -c 2 -o /tmp/p_gu_swiglu $B > $OUT/ncu_gu_swiglu.log 2>&1; stamp "ncu merged gate|up rc=$?"
OPADPO_CONCURRENT_REF=0 timeout 300 ncu $FULL -k "regex:gemm_kernel<\(bool\)0, \(bool\)1, \(int\)0," -s 363 -c 11 -o /tmp/p_dgrad $B > $OUT/ncu_dgrad.log 2>&1; stamp "ncu dgrad layer rc=$?"
OPADPO_CONCURRENT_REF=0 timeout 300 ncu $FULL -k "regex:gemm_kernel<\(bool\)1, \(bool\)1, \(int\)1," -s 360 -c 11 -o /tmp/p_wgrad $B > $OUT/ncu_wgrad.log 2>&1; stamp "ncu wgrad layer rc=$?"
for n in gu_swiglu dgrad wgrad; do ncu -i /tmp/p_$n.ncu-rep --page raw --csv > $OUT/r02_ncu_full_$n.csv 2>/dev/null; done
timeout 500 python bench.py > $OUT/bench_r2m.log 2> $OUT/bench_r2m.err; stamp "default bench rc=$?"
tail -n 3 $OUT/t_r2m_fast.log; tail -n 1 $OUT/bench_r2m.log | cut -c1-500; ls -la $OUT/r02_ncu_full_gu_swiglu.csv $OUT/r02_ncu_full_dgrad.csv $OUT/r02_ncu_full_wgrad.csv
stamp done
