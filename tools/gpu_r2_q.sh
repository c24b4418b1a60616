#!/bin/bash
# Round 2, call Q (1 GPU): SwiGLU backward in the down-projection dgrad epilogue + inverse RoPE in the attention backward epilogues:
# operator and engine tests, same-box bench A/B of each.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
OUT=gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a $OUT/r2q_timeline.log; }
timeout 400 python -m pytest tests/test_gemm_gpu.py tests/test_engine_gpu.py tests/test_pipeline_gpu.py tests/test_attention_gpu.py -q -m gpu > $OUT/t_r2q_fast.log 2>&1; stamp "suites rc=$?"
B="bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-rollout"
timeout 300 python $B > $OUT/bench_r2q_fused.log 2> $OUT/bench_r2q_fused.err; stamp "bench fused rc=$?"
timeout 300 python $B --engine-opt fuse_swiglu_bwd=0 > $OUT/bench_r2q_noswiglu.log 2> $OUT/bench_r2q_noswiglu.err; stamp "bench swiglu_bwd standalone rc=$?"
timeout 300 python $B > $OUT/bench_r2q_fused2.log 2> $OUT/bench_r2q_fused2.err; stamp "bench fused again rc=$?"
timeout 300 python $B --engine-opt fuse_swiglu_bwd=0 > $OUT/bench_r2q_noswiglu2.log 2> $OUT/bench_r2q_noswiglu2.err; stamp "bench swiglu_bwd standalone again rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r2q_*.log')):
    for l in open(f):
        if l.startswith('{"metric"'):
            d=json.loads(l); print(f.split('bench_r2q_')[-1], round(d['ms_per_step'],1), round(d['e2e']['ms_per_step'],1), round(d['roofline']['gemm_ms_per_step'],1), d['gpu_launches'], d['clocks']['sm_mhz'], d['config']['loss'])
PY
tail -n 6 $OUT/t_r2q_fast.log
stamp done
