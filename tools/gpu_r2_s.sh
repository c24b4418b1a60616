#!/bin/bash
# Round 2, call S (1 GPU): final state of the round: whole GPU suite as the driver runs it, smoke, the default bench line
# (train + rollout object + cpu_baseline), the reference arm, and the launch list of the last step.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
OUT=gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a $OUT/r2s_timeline.log; }
timeout 900 python -m pytest tests/ -x -q -m gpu --durations=5 > $OUT/t_r2s_all.log 2>&1; stamp "pytest -m gpu rc=$?"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_r2s.log 2>&1; stamp "smoke rc=$?"
timeout 500 python bench.py > $OUT/bench_r2s.log 2> $OUT/bench_r2s.err; stamp "default bench rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:opadpo --launch-skip 3600 --csv --log-file $OUT/launches_step_r2s.csv python bench.py --profile --steps 1 --warmup 1 --no-cpu-baseline --no-rollout > $OUT/ncu_step_list_r2s.log 2>&1; stamp "launch list rc=$?"
tail -n 12 $OUT/t_r2s_all.log; tail -n 1 $OUT/smoke_r2s.log; tail -n 1 $OUT/bench_r2s.log | cut -c1-400; wc -l $OUT/launches_step_r2s.csv
stamp done
