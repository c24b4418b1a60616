#!/bin/bash
# Round 2, call K (1 GPU): the whole GPU suite as the driver runs it (incl. full-depth parity), smoke, default bench line.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
OUT=gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a $OUT/r2k_timeline.log; }
timeout 900 python -m pytest tests/ -x -q -m gpu --durations=8 > $OUT/t_r2k_all.log 2>&1; stamp "pytest -m gpu rc=$?"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_r2k.log 2>&1; stamp "smoke rc=$?"
timeout 500 python bench.py > $OUT/bench_r2k.log 2> $OUT/bench_r2k.err; stamp "default bench rc=$?"
tail -n 15 $OUT/t_r2k_all.log; tail -n 2 $OUT/smoke_r2k.log; tail -n 1 $OUT/bench_r2k.log | cut -c1-600
stamp done
