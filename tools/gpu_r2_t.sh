#!/bin/bash
# Round 2, call T (1 GPU): inverse RoPE in the attention backward epilogues with the table rows staged through shared memory
# (the first version read them per lane from global memory: +330 us per layer on the kernels' exposed tails).
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
OUT=gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a $OUT/r2t_timeline.log; }
timeout 400 python -m pytest tests/test_engine_gpu.py tests/test_pipeline_gpu.py tests/test_attention_gpu.py -q -m gpu > $OUT/t_r2t_fast.log 2>&1; stamp "suites rc=$?"
B="bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-rollout"
timeout 300 python $B > $OUT/bench_r2t_fused.log 2> $OUT/bench_r2t_fused.err; stamp "bench fused rc=$?"
timeout 300 python $B --engine-opt fuse_rope_bwd=0 > $OUT/bench_r2t_standalone.log 2> $OUT/bench_r2t_standalone.err; stamp "bench standalone rope bwd rc=$?"
timeout 300 python $B > $OUT/bench_r2t_fused2.log 2> $OUT/bench_r2t_fused2.err; stamp "bench fused again rc=$?"
timeout 300 python $B --engine-opt fuse_rope_bwd=0 > $OUT/bench_r2t_standalone2.log 2> $OUT/bench_r2t_standalone2.err; stamp "bench standalone again rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k "regex:attn_bwd|rope_kernel" --launch-skip 40 -c 6 --csv --log-file $OUT/launches_attn_r2t.csv python bench.py --profile --steps 1 --warmup 1 --no-cpu-baseline --no-rollout > $OUT/ncu_attn_list_r2t.log 2>&1; stamp "attn launch list rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r2t_*.log')):
    for l in open(f):
        if l.startswith('{"metric"'):
            d=json.loads(l); print(f.split('bench_r2t_')[-1], round(d['ms_per_step'],1), round(d['e2e']['ms_per_step'],1), round(d['roofline']['gemm_ms_per_step'],1), d['gpu_launches'], d['clocks']['sm_mhz'])
PY
grep -E "attn_bwd|rope" $OUT/launches_attn_r2t.csv | awk -F'","' '{print $5, $(NF)}' | cut -c1-120 | tail -6
tail -n 3 $OUT/t_r2t_fast.log
stamp done
