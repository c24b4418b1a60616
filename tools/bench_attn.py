"""Attention kernels alone at the step's shape (7B, 8 packed samples: S = 703 + 2*512 = 1727, 32 heads x 128, causal +
segment mask + one left-padded row), through the C-ABI operators. CUDA events, L2 flushed between timed launches.
Prints one JSON line per kernel family: ms, executed TFLOP/s (element-level visible pairs: forward 2 GEMM units,
backward 7 in the two-kernel design) and the algorithmic figure (backward 5 units) against the measured bf16 peak.
  python tools/bench_attn.py [--reps 5] [--n 8] [--stacked]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opa_dpo_b200 import ops


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--n", type=int, default=8)
    ap.add_argument("--stacked", action="store_true", help="the reference's K stacked rows (N = 16, S = 1215, no segment mask)")
    ap.add_argument("--only", default="", help="fwd | bwd")
    a = ap.parse_args()
    dev = torch.device("cuda")
    H, HD, S0, R, K = 32, 128, 703, 512, 2
    if a.stacked:
        N, S, pre, seg = a.n * K, S0 + R, 0, 0
        pairs = N * S * (S + 1) / 2
    else:
        N, S, pre, seg = a.n, S0 + K * R, S0, R
        pairs = N * (S0 * (S0 + 1) / 2 + K * (R * S0 + R * (R + 1) / 2))
    d = H * HD
    g = torch.Generator(device=dev).manual_seed(1)
    qkv = (torch.randn(N * S, 3 * d, generator=g, device=dev) * 0.5).to(torch.bfloat16)
    q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
    o = torch.empty(N * S, d, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(N * S, H, device=dev, dtype=torch.float32)
    dO = (torch.randn(N * S, d, generator=g, device=dev) * 0.1).to(torch.bfloat16)
    dqkv = torch.empty(N * S, 3 * d, device=dev, dtype=torch.bfloat16)
    delta = torch.empty(N * S, H, device=dev, dtype=torch.float32)
    mask = torch.ones(N, S, dtype=torch.uint8, device=dev)
    mask[0, :17] = 0                                     # one left-padded query, as in a real batch
    kbits = ops.pack_mask(mask)
    scale = HD ** -0.5
    kw = dict(N=N, S=S, H=H, HD=HD, causal=True, scale=scale, kbits=kbits, prefix_len=pre, segment_len=seg)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def timed(fn):
        ts = []
        for _ in range(a.reps + 2):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts = sorted(ts[2:])
        return ts[len(ts) // 2]

    unit = 2.0 * pairs * HD * H
    peak = 1653.1
    pk = os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["bf16_tflops"])
    shape = dict(N=N, S=S, H=H, HD=HD, prefix=pre, segment=seg)
    if a.only in ("", "fwd"):
        ms = timed(lambda: ops.attn_fwd(q, k, v, o, lse, **kw))
        print(json.dumps({"kernel": "attn_fwd", **shape, "ms": ms, "tflops": 2 * unit / ms / 1e9,
                          "frac_of_measured_burst_peak": 2 * unit / ms / 1e9 / peak}))
    else:
        ops.attn_fwd(q, k, v, o, lse, **kw)
    if a.only in ("", "bwd"):
        ms = timed(lambda: ops.attn_bwd(q, k, v, o, lse, dO, dqkv[:, :d], dqkv[:, d:2 * d], dqkv[:, 2 * d:], delta, **kw))
        print(json.dumps({"kernel": "attn_bwd (dq incl. delta + dkdv)", **shape, "ms": ms, "executed_tflops": 7 * unit / ms / 1e9,
                          "algorithmic_tflops": 5 * unit / ms / 1e9,
                          "algorithmic_frac_of_measured_burst_peak": 5 * unit / ms / 1e9 / peak}))


if __name__ == "__main__":
    main()
