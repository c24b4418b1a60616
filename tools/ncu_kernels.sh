#!/bin/bash
# ncu --set full captures of the non-GEMM kernels of the step and of the rollout decode kernels (run under gpurun).
# Windows are chosen by launch index among the matching kernels: forward of a decoder layer, backward of a decoder
# layer, and one decode step. Reports are reduced to raw CSV on the box (gpurun_out/ has a 64 MiB return limit).
set -u
OUT=gpurun_out
RE='regex:attn_fwd|attn_bwd|attn_delta|swiglu|rmsnorm|rope_kernel|layernorm'
COMMON="--set full --clock-control none"
timeout 400 ncu $COMMON -k "$RE" -s 1560 -c 12 -o /tmp/prof_fwd python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_k1.log 2>&1
timeout 400 ncu $COMMON -k "$RE" -s 2060 -c 16 -o /tmp/prof_bwd python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_k2.log 2>&1
timeout 400 ncu $COMMON --kernel-name-base demangled -k "regex:decode_attention|partial_|greedy_pick|gemm_kernel<\(bool\)0, \(bool\)0, \(int\)4" -s 400 -c 14 -o /tmp/prof_decode python tools/bench_rollout.py --new 3 --runs 1 > $OUT/ncu_k4.log 2>&1
for n in fwd bwd decode; do
  ncu -i /tmp/prof_$n.ncu-rep --page raw --csv > $OUT/ncu_full_${n}_kernels_r01.csv 2>/dev/null
done
ls -la $OUT | grep ncu_full
