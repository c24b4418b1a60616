#!/bin/bash
# First GPU call of a round: the tests that were written after the previous round's GPU budget ran out (opt-in guard), the
# whole GPU suite in one process as the driver runs it, smoke, the bench in its default configuration, the OPA and collator
# micro-benches. Logs under gpurun_out/ (copy what is to be judged into profiles/).
#   gpurun --timeout 900 -- 'bash tools/gpu_round_start.sh'
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a gpurun_out/round_start_timeline.log; }
stamp "start"; nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv | tee -a gpurun_out/round_start_timeline.log
OPADPO_RUN_UNVERIFIED=1 timeout 300 python -m pytest tests/test_engine_gpu.py tests/test_gemm_gpu.py -q -m gpu \
  -k "reference_alignment or early_eos or rope_epilogue or fused_swiglu or swiglu_epilogue" \
  > gpurun_out/t_unverified.log 2>&1; stamp "opt-in tests rc=$?"
timeout 600 python -m pytest tests/ -x -q -m gpu --durations=8 > gpurun_out/t_all_gpu.log 2>&1; stamp "full gpu suite rc=$?"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; stamp "smoke rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; stamp "bench rc=$?"
timeout 200 python bench.py --steps 5 --warmup 3 --fuse-rope 1 --fuse-swiglu 1 --no-cpu-baseline > gpurun_out/bench_fused_epilogues.log 2> gpurun_out/bench_fused_epilogues.err; stamp "bench fused epilogues rc=$?"
timeout 200 python tools/bench_opa.py --entropy 1 --no-cpu-baseline > gpurun_out/bench_opa_entropy.log 2> gpurun_out/bench_opa_entropy.err; stamp "bench_opa rc=$?"
timeout 120 python tools/bench_collate.py > gpurun_out/bench_collate.log 2> gpurun_out/bench_collate.err; stamp "bench_collate rc=$?"
tail -3 gpurun_out/t_unverified.log gpurun_out/t_all_gpu.log gpurun_out/smoke.log
stamp "done"
