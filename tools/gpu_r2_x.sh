#!/bin/bash
# Round 2, call X (1 GPU): ncu --set full of the final attention kernels at the step's shape.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
OUT=gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_ -c 7 -o $OUT/r02_ncu_attn_v6 python tools/bench_attn.py --reps 1 > $OUT/ncu_attn_v6.log 2>&1; echo "ncu rc=$?"
ncu -i $OUT/r02_ncu_attn_v6.ncu-rep --page raw --csv > $OUT/r02_ncu_attn_v6.csv 2>/dev/null
timeout 100 python tools/bench_attn.py > $OUT/bench_attn_r2x.log 2>&1; cat $OUT/bench_attn_r2x.log
