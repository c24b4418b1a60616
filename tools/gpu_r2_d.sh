#!/bin/bash
# Round 2, call D: evidence. Launch list of one training step and of the rollout, ncu --set full rows of the kernels the
# round-1 verdict listed as unprofiled (dgrad / wgrad / lm_head GEMMs, fused-epilogue GEMMs, AdamW, sumsq, dpo_pair,
# embed_splice, sampler, decode kernels). Reports are reduced to raw CSV on the box (64 MiB return limit).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
OUT=gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a $OUT/r2d_timeline.log; }
B="python bench.py --profile --steps 1 --warmup 1 --no-cpu-baseline --no-rollout"
FULL="--set full --clock-control none --kernel-name-base demangled"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:opadpo --csv --log-file $OUT/launches_step_r2d.csv $B > $OUT/ncu_step_list.log 2>&1; stamp "launch list step rc=$?"
timeout 300 ncu $FULL -k "regex:adamw_kernel|sumsq_kernel|dpo_pair_kernel|embed_splice|logprob_finalize|norm_from_sumsq|fold_rows" -c 14 -o /tmp/p_small $B > $OUT/ncu_small.log 2>&1; stamp "ncu small kernels rc=$?"
timeout 300 ncu $FULL -k "regex:gemm_kernel<\(bool\)0, \(bool\)0, \(int\)[23]," -c 3 -o /tmp/p_lmhead $B > $OUT/ncu_lmhead.log 2>&1; stamp "ncu lm_head rc=$?"
timeout 300 ncu $FULL -k "regex:gemm_kernel<\(bool\)0, \(bool\)1, \(int\)0,|gemm_kernel<\(bool\)1, \(bool\)1, \(int\)1," -s 48 -c 16 -o /tmp/p_bwdgemm $B > $OUT/ncu_bwdgemm.log 2>&1; stamp "ncu dgrad/wgrad rc=$?"
timeout 300 ncu $FULL -k "regex:gemm_kernel<\(bool\)0, \(bool\)0, \(int\)[0156]," -s 40 -c 12 -o /tmp/p_fwdgemm $B > $OUT/ncu_fwdgemm.log 2>&1; stamp "ncu fwd gemms rc=$?"
timeout 300 ncu $FULL -k "regex:swiglu|rmsnorm|rope_kernel|attn_delta" -s 200 -c 10 -o /tmp/p_elem $B > $OUT/ncu_elem.log 2>&1; stamp "ncu elementwise rc=$?"
timeout 300 ncu $FULL -k "regex:decode_attention|decode_gemm|partial_|sample_pick|greedy_pick" -s 460 -c 14 -o /tmp/p_decode python tools/bench_rollout.py --new 3 --runs 1 --sample > $OUT/ncu_decode.log 2>&1; stamp "ncu decode rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:opadpo --csv --log-file $OUT/launches_rollout_r2d.csv python tools/bench_rollout.py --new 3 --runs 1 > $OUT/ncu_rollout_list.log 2>&1; stamp "launch list rollout rc=$?"
for n in small lmhead bwdgemm fwdgemm elem decode; do
  ncu -i /tmp/p_$n.ncu-rep --page raw --csv > $OUT/r02_ncu_full_$n.csv 2>/dev/null
done
ls -la $OUT | grep -E "r02_ncu_full|launches_.*r2d"
stamp done
