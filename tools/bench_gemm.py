"""Times the tcgen05 GEMM on the hot-path shapes (CUDA events, L2 flushed between iterations)."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opa_dpo_b200 import ops


def bench(fn, iters=10, warmup=3):
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(warmup):
        fn()
    times = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        times.append(s.elapsed_time(e))
    times.sort()
    return times[len(times) // 2]


def main():
    shapes = [(8192, 8192, 8192), (13816, 4096, 4096), (13816, 12288, 4096), (13816, 22016, 4096), (13816, 4096, 11008)]
    res = []
    for M, N, K in shapes:
        a = torch.randn((M, K), device="cuda", dtype=torch.bfloat16)
        b = torch.randn((N, K), device="cuda", dtype=torch.bfloat16)
        out = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
        ms = bench(lambda: ops.gemm(a, b, out, cluster=0))
        ms_cl = bench(lambda: ops.gemm(a, b, out, cluster=1))
        ms_pair = bench(lambda: ops.gemm(a, b, out, cluster=2))
        ms_ref = bench(lambda: torch.matmul(a, b.t(), out=out))
        tf = 2.0 * M * N * K / ms / 1e9
        res.append(dict(M=M, N=N, K=K, ms=ms, tflops=tf, cluster_ms=ms_cl, cluster_tflops=2.0 * M * N * K / ms_cl / 1e9,
                        pair_tflops=2.0 * M * N * K / ms_pair / 1e9, cublas_ms=ms_ref, cublas_tflops=2.0 * M * N * K / ms_ref / 1e9))
        print(json.dumps(res[-1]), flush=True)
        # dgrad form (B MN-major)
        bt = torch.randn((K, N), device="cuda", dtype=torch.bfloat16)
        ms = bench(lambda: ops.gemm(a, bt, out, b_mn=True, cluster=0))
        ms_cl = bench(lambda: ops.gemm(a, bt, out, b_mn=True, cluster=1))
        ms_pair = bench(lambda: ops.gemm(a, bt, out, b_mn=True, cluster=2))
        print(json.dumps(dict(kind="b_mn", M=M, N=N, K=K, ms=ms, tflops=2.0 * M * N * K / ms / 1e9, cluster_ms=ms_cl,
                              cluster_tflops=2.0 * M * N * K / ms_cl / 1e9, pair_tflops=2.0 * M * N * K / ms_pair / 1e9)),
              flush=True)
    # wgrad form
    for T, Mo, No, splits in [(19440, 256, 4096, 5), (19440, 4096, 256, 5), (19440, 768, 4096, 2)]:
        a = torch.randn((T, Mo), device="cuda", dtype=torch.bfloat16)
        b = torch.randn((T, No), device="cuda", dtype=torch.bfloat16)
        out = torch.zeros((Mo, No), device="cuda", dtype=torch.float32)
        ms = bench(lambda: ops.gemm(a, b, out, a_mn=True, b_mn=True, epilogue=ops.EPI_F32_ATOMIC, splits=splits))
        print(json.dumps(dict(kind="wgrad", T=T, M=Mo, N=No, splits=splits, ms=ms, tflops=2.0 * T * Mo * No / ms / 1e9)), flush=True)


if __name__ == "__main__":
    main()
