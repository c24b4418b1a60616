#!/bin/bash
# Round 2, call Z (1 GPU): single-GPU test of opadpo_engine_adopt_grad_buffer + the engine suite.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_engine_gpu.py -q -m gpu > gpurun_out/t_r2z.log 2>&1; echo "rc=$?"; tail -n 12 gpurun_out/t_r2z.log
