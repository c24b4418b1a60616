#!/bin/bash
# ncu --set full captures of the non-GEMM kernels of the step and of the rollout decode kernels (run under gpurun).
# Windows are chosen by launch index among the matching kernels: forward of a decoder layer, backward of a decoder
# layer, the optimiser/loss kernels, and one decode step.
set -u
OUT=gpurun_out
RE='regex:attn_fwd|attn_bwd|attn_delta|swiglu|rmsnorm|rope_kernel|layernorm'
COMMON="--set full --clock-control none --import-source on"
timeout 400 ncu $COMMON -k "$RE" -s 1560 -c 12 -o $OUT/prof_fwd_kernels_r01 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_k1.log 2>&1
timeout 400 ncu $COMMON -k "$RE" -s 2060 -c 16 -o $OUT/prof_bwd_kernels_r01 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_k2.log 2>&1
timeout 400 ncu $COMMON -k "regex:adamw|sumsq|dpo_pair|logprob_finalize|embed_splice" -s 8 -c 7 -o $OUT/prof_opt_kernels_r01 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_k3.log 2>&1
timeout 400 ncu $COMMON --kernel-name-base demangled -k "regex:decode_attention|partial_|greedy_pick|gemm_kernel<\(bool\)0, \(bool\)0, \(int\)4" -s 400 -c 14 -o $OUT/prof_decode_kernels_r01 python tools/bench_rollout.py --new 3 --runs 1 > $OUT/ncu_k4.log 2>&1
ls -la $OUT | grep prof_
