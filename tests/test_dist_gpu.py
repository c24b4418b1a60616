"""Data-parallel step on real GPUs: launches tests/dist_gpu_check.py under torchrun with one process per
GPU (2 ranks) when at least two GPUs are visible; skipped on a single-GPU box. `gpurun --gpus 2 -- 'python -m pytest
tests/test_dist_gpu.py -m gpu -q'` runs it; the report lands in gpurun_out/dist_gpu_check.json."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_gpu_dp_step_equals_single_gpu_step_on_concatenated_batch():
    port = 29600 + os.getpid() % 1000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_gpu_check.py")]
    res = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    tail = res.stdout[-4000:]
    assert res.returncode == 0, tail
    line = [l for l in res.stdout.splitlines() if l.startswith("DIST_GPU_CHECK ")]
    assert line, tail
    rep = json.loads(line[-1][len("DIST_GPU_CHECK "):])
    assert rep["ok"] and all(rep["checks"].values()), rep
