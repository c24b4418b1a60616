#!/bin/bash
# Third GPU check: collator kernels after the third pass (thread-per-column horizontal kernel, pooled scratch).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a gpurun_out/timeline3.log; }
stamp "start"
timeout 300 python -m pytest tests/test_collate_gpu.py -q -m gpu > gpurun_out/t_collate3.log 2>&1; stamp "collate tests rc=$?"
timeout 120 python tools/bench_collate.py > gpurun_out/bench_collate3.log 2> gpurun_out/bench_collate3.err; stamp "bench_collate rc=$?"
timeout 200 ncu --set full --clock-control none -k regex:"resample|collate_responses|mask_fill|image_sum" --launch-skip 12 -c 6 \
  -o gpurun_out/r01_ncu_collate_v3 python tools/bench_collate.py --iters 1 > gpurun_out/ncu_collate3.log 2>&1; stamp "ncu collate rc=$?"
stamp "done"
