#!/bin/bash
# One gpurun call: new-row parity tests first, then the suites the edits touch, the bench (reference pass unmerged and
# merged), the collator micro-bench, launch lists + one ncu capture of the collator kernels, the remaining suites last.
# Everything is logged under gpurun_out/ so a cut-off call still leaves what finished.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
export PYTHONUNBUFFERED=1
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a gpurun_out/timeline.log; }
stamp "start"; nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv | tee -a gpurun_out/timeline.log
timeout 420 python -m pytest tests/test_collate_gpu.py -q -m gpu --durations=5 > gpurun_out/t_collate.log 2>&1; stamp "collate tests rc=$?"
timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_gemm_gpu.py -q -m gpu --durations=8 > gpurun_out/t_engine_gemm.log 2>&1; stamp "engine+gemm tests rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; stamp "bench default rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --merge-frozen 1 --no-cpu-baseline > gpurun_out/bench_merge_frozen.log 2> gpurun_out/bench_merge_frozen.err; stamp "bench merge_frozen rc=$?"
timeout 200 python tools/bench_collate.py > gpurun_out/bench_collate.log 2> gpurun_out/bench_collate.err; stamp "bench_collate rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"resample|collate_responses|mask_fill|image_mean" -c 8 \
  -o gpurun_out/r01_ncu_collate python tools/bench_collate.py --iters 1 > gpurun_out/ncu_collate.log 2>&1; stamp "ncu collate rc=$?"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; stamp "smoke rc=$?"
timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_fullsize_gpu.py -q -m gpu --durations=8 > gpurun_out/t_attn_fullsize.log 2>&1; stamp "attention+fullsize tests rc=$?"
stamp "done"
