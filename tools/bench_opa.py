"""OPA stage (LoRA-SFT) on one B200: samples/s of one optimisation step — policy forward over one
response per sample, token-mean NLL [+ the entropy regulariser's second, image-masked forward], LoRA backward,
clip + AdamW — with the same timing rules as bench.py (>= 3 warm-up steps, CUDA events, inputs larger than L2, clocks
sampled under load) and the reference's CPU path timed beside it on a bounded sample.
  python tools/bench_opa.py [--samples 8] [--resp-len 512] [--entropy 1] [--steps 5]
Algorithmic FLOPs per sample (flops.py conventions): forward + backward of ONE sequence (x2 with the regulariser)."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402  (peaks, clock sampler, Q_LEN / IMG_POS)


def cpu_sample(threads: int, R: int, entropy: bool):
    """The reference path on host cores (oracle port): one sample at true 7B widths, depth truncated to 2 + 2 layers."""
    import torch
    import torch.nn.functional as F
    from oracle import llava_oracle as O
    from opa_dpo_b200.engine import EngineConfig
    from opa_dpo_b200.flops import flops_per_pair
    torch.set_num_threads(threads)
    oc = O.OracleConfig(**{**O.LLAVA_7B, "n_layers": 2, "v_layers": 2})
    w = O.init_weights(oc, seed=1)
    lp = O.init_lora(oc, 2)
    images, q, qm, resp = O.synthetic_batch(oc, 1, B.Q_LEN, R, ["standard_response"], ragged=False, image_pos=B.IMG_POS)
    lpg = {m: (A.clone().requires_grad_(not m.startswith(O.VT)), Bm.clone().requires_grad_(not m.startswith(O.VT)))
           for m, (A, Bm) in lp.items()}
    ids = torch.cat([q, resp["standard_response"]], 1)
    mask = torch.cat([qm.bool(), resp["standard_response"] != 0], 1)
    labels = resp["standard_response"].clone()
    labels[labels == 0] = -100
    t0 = time.perf_counter()

    def one(img):
        logits = O.llava_logits(oc, w, lpg, ids, mask, img)[:, -R - 1:-1]
        ls = torch.log_softmax(logits, -1)
        return F.cross_entropy(logits.reshape(-1, oc.vocab), labels.reshape(-1), ignore_index=-100), -(ls.exp() * ls).sum(-1)
    loss, h = one(images)
    if entropy:
        _, hm = one(images * 0.5)
        loss = loss + 0.1 * (-(hm - h)).mean()
    loss.backward()
    dt = time.perf_counter() - t0
    f = flops_per_pair(EngineConfig(n_layers=2, v_layers=2), B.Q_LEN, R)
    return dt, (f["fwd_per_seq"] + f["bwd_per_seq"]) * (2 if entropy else 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=8)
    ap.add_argument("--resp-len", type=int, default=512)
    ap.add_argument("--entropy", type=int, default=1, help="1: entropy regulariser on (second forward on masked images)")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    import torch
    from types import SimpleNamespace
    from opa_dpo_b200 import collate
    from opa_dpo_b200.dpo import DPOArgs
    from opa_dpo_b200.engine import EngineConfig
    from opa_dpo_b200.flops import flops_per_pair
    from opa_dpo_b200.opa import OPAArgs, OPAStepper
    from opa_dpo_b200.optim import B200AdamW
    from opa_dpo_b200.policy import AutoregressivePolicy, B200LlavaModel
    if not torch.cuda.is_available():
        raise SystemExit("bench_opa.py needs a CUDA device")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    cfg, R, b = EngineConfig(), args.resp_len, args.samples
    model = B200LlavaModel(cfg, dev)
    model.engine.random_init(seed=1234)
    model.make_trainable("lora_policy")
    tok = SimpleNamespace(pad_token_id=cfg.pad_id, eos_token_id=cfg.eos_id)
    policy = AutoregressivePolicy(DPOArgs(response_len=R), model, tok, adapter_name="lora_policy")
    opt = B200AdamW(model, lr=2e-5)          # run/train_opa.sh learning rate
    stepper = OPAStepper(policy, opt, args=OPAArgs(entropy_loss=bool(args.entropy), entropy_loss_coef=0.1))
    g = torch.Generator().manual_seed(1234)
    h_images = torch.randn(b, 3, cfg.v_image, cfg.v_image, generator=g).pin_memory()
    h_q = torch.randint(3, cfg.vocab, (b, B.Q_LEN), generator=g)
    h_q[:, B.IMG_POS] = cfg.image_token
    h_q = h_q.pin_memory()
    h_qm = torch.ones(b, B.Q_LEN, dtype=torch.uint8).pin_memory()
    h_r = torch.randint(3, cfg.vocab, (b, R), generator=g)
    h_r[:, -1] = cfg.eos_id
    h_r = h_r.pin_memory()
    h2d = h_images.numel() * 4 + h_q.numel() * 8 + h_qm.numel() + h_r.numel() * 8

    def step(e2e: bool):
        images, q, qm, r = (t.to(dev, non_blocking=True) for t in (h_images, h_q, h_qm, h_r)) if e2e else d_in
        masked = collate.mask_images(images, 0.3, "random") if args.entropy else None     # opa_trainer.py:66-72
        loss, norm = stepper.step(images, q, qm, {"standard_response": r}, masked_images=masked)
        return float(loss.item()) if e2e else loss

    d_in = tuple(t.to(dev) for t in (h_images, h_q, h_qm, h_r))
    for _ in range(max(3, args.warmup)):
        step(False)
    torch.cuda.synchronize()
    clocks = B.ClockSampler(0)
    clocks.start()
    l0 = model.engine.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step(False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    launches = model.engine.launch_count() - l0
    clk = clocks.stop()
    step(True)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        lh = step(True)
    t1.record()
    torch.cuda.synchronize()
    ms_e2e = t0.elapsed_time(t1) / args.steps
    f = flops_per_pair(cfg, B.Q_LEN, R)
    per_sample = (f["fwd_per_seq"] + f["bwd_per_seq"]) * (2 if args.entropy else 1)
    peaks = B._peaks()
    tflops = per_sample * b / (ms * 1e-3) / 1e12
    line = {"metric": "samples/sec LLaVA-1.5-7B OPA LoRA-SFT step", "value": b / (ms * 1e-3), "unit": "samples/s", "n_gpus": 1,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"LLaVA-1.5-7B LoRA(r=256) OPA-SFT step, 336^2 img, Q=128, R={R}, token-mean NLL"
                                   + (" + entropy regulariser (second forward on CoPO-masked images, both differentiated)" if args.entropy else "")
                                   + " + LoRA backward + clip + AdamW",
                       "samples_per_gpu": b, "l2": "inputs larger than L2", "loss": float(loss.item())},
            "e2e": {"value": b / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e, "loss": lh},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": tflops, "peak": peaks["burst"], "unit": "TFLOP/s", "frac": tflops / peaks["burst"],
                         "peak_source": peaks["source"] + " burst", "frac_of_sustained": tflops / peaks["sustained"],
                         "tflop_per_sample": per_sample / 1e12,
                         "note": "step-level: algorithmic FLOPs of the whole step / step time (the GEMM family's own roofline is bench.py's)"},
            "clocks": clk}
    if not args.no_cpu_baseline:
        try:
            threads = os.cpu_count() or 1
            dt, fl = cpu_sample(threads, R, bool(args.entropy))
            line["cpu_baseline"] = {"value": (fl / dt) / per_sample, "unit": "samples/s", "cores": threads, "kind": "port",
                                    "sample": f"1 sample, fp32 oracle at true 7B widths, depth truncated to 2 LLaMA + 2 CLIP layers "
                                              f"({dt:.1f} s); extrapolated to full depth by algorithmic FLOPs"}
        except Exception as ex:
            line["cpu_baseline"] = {"value": None, "unit": "samples/s", "kind": "port", "sample": f"failed: {type(ex).__name__}: {ex}"}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
