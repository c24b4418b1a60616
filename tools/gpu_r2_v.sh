#!/bin/bash
# Round 2, call V (1 GPU): 13B-width truncated-depth parity case.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
OUT=gpurun_out
timeout 500 python -m pytest tests/test_fullsize_gpu.py -q -m gpu -k "13b_width or true_width" --durations=3 > $OUT/t_r2v.log 2>&1; echo "rc=$?"
tail -n 25 $OUT/t_r2v.log
