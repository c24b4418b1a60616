"""Collator-stage kernels timed on the GPU at the repo's batch shapes, with the reference's own CPU
path (Pillow + numpy, torch) timed beside them on the box's host cores. One JSON line per kernel:
  algorithmic bytes = bytes read once + bytes written once (stated per line), GB/s = bytes / CUDA-event time,
  frac = GB/s / measured HBM peak (MEASURED_PEAKS.json, else the profiling guide's fallback).
Usage: python tools/bench_collate.py [--images 64] [--rows 64] [--iters 20]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6533.0, "fallback"


def timed(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        flush.fill_(1.0)                      # > L2: every timed launch starts with a cold L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=64)
    ap.add_argument("--rows", type=int, default=64)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    from opa_dpo_b200 import collate as K
    from opa_dpo_b200 import _lib
    import ctypes
    peak, src = hbm_peak()
    dev = torch.device("cuda", 0)
    flush = torch.empty(64 * 1024 * 1024, device=dev)     # 256 MB > 126 MB L2
    rs = np.random.RandomState(0)
    lines = []

    # ---- images: 480x640 uint8 -> expand2square -> 336x336 f32 (data_utils_dpo.py:311-335)
    n, S = args.images, 336
    imgs = [rs.randint(0, 256, (480, 640, 3)).astype(np.uint8) for _ in range(n)]
    lib = _lib.load()
    flat = torch.from_numpy(np.concatenate([i.reshape(-1) for i in imgs])).to(dev)
    offsets = np.cumsum([0] + [i.size for i in imgs]).astype(np.int64)
    hs, ws = np.full(n, 480, np.int32), np.full(n, 640, np.int32)
    mean, std = np.asarray(K.CLIP_MEAN, np.float32), np.asarray(K.CLIP_STD, np.float32)
    bg = np.asarray([int(x * 255) for x in K.CLIP_MEAN], np.uint8)
    out = torch.empty((n, 3, S, S), device=dev)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def run_pre():
        _lib.check(lib.opadpo_op_preprocess_images(flat.data_ptr(), p(offsets), p(hs), p(ws), n, S, p(mean), p(std), p(bg), 1,
                                                   out.data_ptr(), st()), "preprocess")
    ms = timed(run_pre, args.iters, flush)
    bytes_alg = n * (480 * 640 * 3 + 3 * S * S * 4)
    t0 = time.perf_counter()
    cpu_n = min(n, 16)
    try:
        from PIL import Image
        for im in imgs[:cpu_n]:                       # the reference's own path: Pillow resize + numpy rescale / normalise
            L = max(im.shape[:2])
            canvas = Image.new("RGB", (L, L), tuple(int(b) for b in bg))
            canvas.paste(Image.fromarray(im), ((L - im.shape[1]) // 2, (L - im.shape[0]) // 2))
            r = np.asarray(canvas.resize((S, S), resample=Image.BICUBIC))
            _ = (((r.astype(np.float64) * (1 / 255)).astype(np.float32) - mean) / std).transpose(2, 0, 1)
        cpu_kind = "reference"
    except ImportError:
        from oracle import collate_oracle as C
        cpu_n = 1
        C.clip_preprocess(imgs[0], S, True)
        cpu_kind = "port"
    cpu_s = (time.perf_counter() - t0) / cpu_n
    lines.append(dict(kernel="resample_h + resample_v (preprocess_images)", unit="images/s", value=n / (ms * 1e-3), ms=ms,
                      launches=2, images=n, shape="480x640x3 u8 -> 3x336x336 f32, pad to square",
                      roofline=dict(bound="hbm", achieved=bytes_alg / (ms * 1e-3) / 1e9, peak=peak, unit="GB/s",
                                    frac=bytes_alg / (ms * 1e-3) / 1e9 / peak, peak_source=src,
                                    algorithmic_bytes="per image: H*W*3 read + 3*336*336*4 written = 2.28 MB"),
                      cpu_baseline=dict(value=1.0 / cpu_s, unit="images/s", cores=1, kind=cpu_kind,
                                        sample=f"{cpu_n} images through Pillow resize + numpy normalise, one thread")))

    # ---- CoPO masks on 336x336 images (dpo_trainer.py:83-109)
    x = torch.randn(n, 3, S, S, device=dev)
    idx = torch.stack([torch.randperm(S * S)[: int(S * S * 0.3)] for _ in range(n)]).to(dev)
    mo = torch.empty_like(x)
    means = torch.empty(n, device=dev)

    def run_mask():
        _lib.check(lib.opadpo_op_mask_images(x.data_ptr(), mo.data_ptr(), n, 3, S, S, idx.data_ptr(), idx.shape[1], 1, 0,
                                             means.data_ptr(), st()), "mask")
    ms = timed(run_mask, args.iters, flush)
    bytes_alg = n * 3 * S * S * 4 * 2 + idx.numel() * 8
    xc = x[: min(n, 16)].cpu()
    t0 = time.perf_counter()
    for i in range(xc.shape[0]):                          # mask_single_image as written in the reference
        im = xc[i:i + 1].clone()
        mv = im.mean()
        im.view(3, -1)[:, torch.randperm(S * S)[: int(S * S * 0.3)]] = mv
    cpu_s = (time.perf_counter() - t0) / xc.shape[0]
    lines.append(dict(kernel="image_mean + copy_pixels + mask_fill (mask_images, random 0.3)", unit="images/s",
                      value=n / (ms * 1e-3), ms=ms, launches=3, images=n,
                      roofline=dict(bound="hbm", achieved=bytes_alg / (ms * 1e-3) / 1e9, peak=peak, unit="GB/s",
                                    frac=bytes_alg / (ms * 1e-3) / 1e9 / peak, peak_source=src,
                                    algorithmic_bytes="per image: 3*336*336*4 read + the same written + 8 B per masked pixel index"),
                      cpu_baseline=dict(value=1.0 / cpu_s, unit="images/s", cores=torch.get_num_threads(), kind="reference",
                                        sample=f"{xc.shape[0]} images, mask_single_image's torch ops on the host")))

    # ---- response assembly: rows x R=896 from ~30 report entries each (data_utils_dpo.py:139-235)
    b, R = args.rows, 896
    pack = K.SegmentPack()
    for i in range(b):
        pack.begin_row()
        for j in range(30):
            pack.add([1] + rs.randint(3, 32000, rs.randint(5, 40)).tolist(), 2.0, 3.0, K.SEG_KEEP_BOS if j == 0 else K.SEG_STRIP_SPACE)
        pack.end_row()
    arrs = {k: torch.from_numpy(v).to(dev) for k, v in pack.arrays().items()}
    ids = torch.empty((b, R), device=dev, dtype=torch.int64)
    mask = torch.empty((b, R), device=dev, dtype=torch.uint8)
    sc, rel = torch.empty((b, R), device=dev), torch.empty((b, R), device=dev)
    ov = torch.empty((b,), device=dev, dtype=torch.int32)
    seg = _lib.Segments(arrs["tokens"].data_ptr(), arrs["seg_offsets"].data_ptr(), arrs["seg_score"].data_ptr(),
                        arrs["seg_relation"].data_ptr(), arrs["seg_flags"].data_ptr(), arrs["row_offsets"].data_ptr(), b)

    def run_resp():
        _lib.check(lib.opadpo_op_collate_responses(ctypes.byref(seg), R, 0, 2, 3, ids.data_ptr(), mask.data_ptr(), sc.data_ptr(),
                                                   rel.data_ptr(), ov.data_ptr(), st()), "collate")
    ms = timed(run_resp, args.iters, flush)
    bytes_alg = arrs["tokens"].numel() * 4 + b * R * (8 + 1 + 4 + 4)
    from oracle import collate_oracle as C
    rows = []
    a = pack.arrays()
    for i in range(b):
        rows.append([(a["tokens"][a["seg_offsets"][j]:a["seg_offsets"][j + 1]], float(a["seg_score"][j]), float(a["seg_relation"][j]),
                      int(a["seg_flags"][j])) for j in range(a["row_offsets"][i], a["row_offsets"][i + 1])])
    t0 = time.perf_counter()
    C.collate_responses(rows, R, 0, 2, 3)
    cpu_s = time.perf_counter() - t0
    lines.append(dict(kernel="collate_responses", unit="rows/s", value=b / (ms * 1e-3), ms=ms, launches=1, rows=b, R=R,
                      roofline=dict(bound="hbm (launch-latency bound at this size)", achieved=bytes_alg / (ms * 1e-3) / 1e9,
                                    peak=peak, unit="GB/s", frac=bytes_alg / (ms * 1e-3) / 1e9 / peak, peak_source=src,
                                    algorithmic_bytes="per row: 4 B per token read + R*(8+1+4+4) B written = 15.2 KB"),
                      cpu_baseline=dict(value=b / cpu_s, unit="rows/s", cores=1, kind="port",
                                        sample=f"{b} rows through the numpy oracle (torch.cat chain restated)")))
    for ln in lines:
        print(json.dumps(ln), flush=True)


if __name__ == "__main__":
    main()
