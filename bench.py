#!/usr/bin/env python
"""bench.py — preference-pairs/s of the OPA-DPO step (BASELINE.json metric) on N B200s of one node.

A "step" = one pass of the hot path over one batch of synthetic preference pairs per GPU:
  reference-adapter forward (chosen+rejected) -> policy forward (chosen+rejected) -> fused DPO loss -> LoRA backward
  -> [one NCCL all-reduce of the LoRA-grad bucket, N>1] -> grad-norm clip + AdamW.
Workload = BASELINE.json configs[1]: LLaVA-1.5-7B, LoRA r=256 on every Linear but lm_head, 336^2 images, query 128,
512-token chosen/rejected responses, bf16, random-init weights of that architecture (no checkpoints offline).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs P_per_gpu] [--impl reference]

`value`  : device-resident inputs, CUDA-event timed, max over ranks.
`e2e`    : same step through the public nn.Module API with HOST (pinned) inputs copied every step + loss read back.
`roofline`: tensor-pipe bound; achieved = executed GEMM FLOPs / summed CUDA-event kernel time of every tcgen05 GEMM
           launch of one instrumented step (opadpo_prof_begin/end), peak from MEASURED_PEAKS.json.
`cpu_baseline` / `--impl reference`: the oracle (CPU fp32 port of the reference path) on the host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

Q_LEN, R_LEN, IMG_POS = 128, 512, 35
MERGE_FROZEN_DEFAULT = 1   # measured: 621 vs 635 ms/step, parity test green (profiles/r01_bench_s3_*.log); --merge-frozen 0 runs it unmerged


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(burst=float(d["bf16_tflops"]), sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    hbm=float(d["hbm_gbs"]), source="measured")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback")


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_sample(threads: int, n_layers: int = 2, v_layers: int = 2, R: int = R_LEN, repeats: int = 1):
    """The reference path on host cores: oracle (verbatim restatement of rl_models/dpo_trainer over a restated
    LLaVA) at TRUE 7B widths but truncated depth (n_layers LLaMA / v_layers CLIP layers), one preference pair,
    fp32: reference fwd x2 + policy fwd x2 + DPO loss + autograd backward. Returns (seconds per step, algorithmic
    FLOPs of the truncated step)."""
    import torch
    from oracle import llava_oracle as O
    from opa_dpo_b200.engine import EngineConfig
    from opa_dpo_b200.flops import flops_per_pair
    torch.set_num_threads(threads)
    oc = O.OracleConfig(**{**O.LLAVA_7B, "n_layers": n_layers, "v_layers": v_layers})
    w = O.init_weights(oc, seed=1)
    lp, lr = O.init_lora(oc, 2), O.init_lora(oc, 3)
    images, q, qm, resp = O.synthetic_batch(oc, 1, Q_LEN, R, ["chosen_response", "rejected_response"], ragged=False,
                                            image_pos=IMG_POS)
    lc = O.LossConfig(beta=0.1)
    times = []
    for _ in range(repeats):
        lpg = {m: (A.clone().requires_grad_(not m.startswith(O.VT)), B.clone().requires_grad_(not m.startswith(O.VT)))
               for m, (A, B) in lp.items()}
        t0 = time.perf_counter()
        with torch.no_grad():
            ref = O.policy_forward(oc, w, lr, images, q, qm, resp)
        pol = O.policy_forward(oc, w, lpg, images, q, qm, resp)
        loss, _, _ = O.pair_loss(lc, pol["chosen_response_logprobs"], pol["rejected_response_logprobs"],
                                 ref["chosen_response_logprobs"], ref["rejected_response_logprobs"])
        loss.backward()
        times.append(time.perf_counter() - t0)
    ec = EngineConfig(n_layers=n_layers, v_layers=v_layers)
    return min(times), flops_per_pair(ec, Q_LEN, R)["per_pair"]


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from opa_dpo_b200.engine import EngineConfig
    from opa_dpo_b200.flops import flops_per_pair
    threads = os.cpu_count() or 1
    full = flops_per_pair(EngineConfig(), Q_LEN, R_LEN)["per_pair"]
    steps, warm = max(1, min(args.steps, 5)), max(0, min(args.warmup, 1))
    ts = []
    fl = None
    for i in range(warm + steps):
        t, fl = cpu_reference_sample(threads)
        if i >= warm:
            ts.append(t)
    t = statistics.median(ts)
    cpu_flops = fl / t
    value = cpu_flops / full
    sample = ("1 pair, R=512, fp32, true 7B widths, depth truncated to 2 LLaMA + 2 CLIP layers; pairs/s extrapolated to "
              "full depth by algorithmic FLOPs (%.1f TFLOP sample vs %.1f TFLOP full pair)" % (fl / 1e12, full / 1e12))
    line = {
        "impl": "reference", "metric": "preference-pairs/sec LLaVA-1.5-7B OPA-DPO", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "LLaVA-1.5-7B LoRA(r=256) OPA-DPO step, 336^2 img, Q=128, R=512 (CPU oracle port of the reference path)"},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": sample,
                         "cpu_tflops": cpu_flops / 1e12},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.path = gpu_index, None, f"/tmp/opadpo_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for ln in f:
                p = [x.strip() for x in ln.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    mx.append(float(p[2]))
                except ValueError:
                    continue
                for nm, v in zip(names, p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        try:
            os.remove(self.path)
        except OSError:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=8, help="preference pairs per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="7b", choices=["7b", "13b"], help="7b = BASELINE configs[1] (headline); 13b = configs[4] shape")
    ap.add_argument("--resp-len", type=int, default=R_LEN)
    ap.add_argument("--no-share-prefix", action="store_true", help="run the reference's K stacked rows (no prefix sharing)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--merge-frozen", type=int, default=int(os.environ.get("OPADPO_MERGE_FROZEN", str(MERGE_FROZEN_DEFAULT))),
                    help="1: the frozen reference adapter's forwards run on decoder weights merged once (no LoRA GEMMs)")
    ap.add_argument("--fuse-swiglu", type=int, default=int(os.environ.get("OPADPO_FUSE_SWIGLU", "1")),
                    help="1: SwiGLU in the gate|up GEMM epilogue of the merged reference pass")
    ap.add_argument("--fuse-rope", type=int, default=int(os.environ.get("OPADPO_FUSE_ROPE", "1")),
                    help="1: RoPE in the q|k|v GEMM epilogue")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from opa_dpo_b200.engine import EngineConfig
    from opa_dpo_b200.policy import AutoregressivePolicy, B200LlavaModel
    from opa_dpo_b200.dpo import DPOArgs, DPOStepper
    from opa_dpo_b200.flops import flops_per_pair
    from opa_dpo_b200 import _lib
    import ctypes
    from types import SimpleNamespace

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the hot path has no CPU implementation); "
                         "use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    cfg = EngineConfig() if args.model == "7b" else EngineConfig.llava_13b()
    R = args.resp_len
    model = B200LlavaModel(cfg, dev)
    if args.no_share_prefix:
        model.engine.set_option("share_prefix", 0)
    model.engine.set_option("merge_frozen", args.merge_frozen)
    model.engine.set_option("fuse_rope", int(bool(args.fuse_rope)))
    model.engine.set_option("fuse_swiglu", int(bool(args.fuse_swiglu)))
    model.engine.random_init(seed=1234)
    model.make_trainable("lora_policy")
    dargs = DPOArgs(beta=0.1, response_len=R, learning_rate=1e-6, max_grad_norm=1.0)
    tok = SimpleNamespace(pad_token_id=cfg.pad_id, eos_token_id=cfg.eos_id)
    policy = AutoregressivePolicy(dargs, model, tok, adapter_name="lora_policy")
    ref_policy = AutoregressivePolicy(dargs, model, tok, adapter_name="lora_ref_policy")
    stepper = DPOStepper(policy, ref_policy, dargs, world_size=world)

    b = args.pairs
    g = torch.Generator().manual_seed(1234 + rank)
    h_images = torch.randn(b, 3, cfg.v_image, cfg.v_image, generator=g).pin_memory()
    h_queries = torch.randint(3, cfg.vocab, (b, Q_LEN), generator=g)
    h_queries[:, IMG_POS] = cfg.image_token
    h_queries = h_queries.pin_memory()
    h_qmask = torch.ones(b, Q_LEN, dtype=torch.uint8).pin_memory()
    h_resp = {}
    for n in ("chosen_response", "rejected_response"):
        t = torch.randint(3, cfg.vocab, (b, R), generator=g)
        t[:, -1] = cfg.eos_id
        h_resp[n] = t.pin_memory()
    h2d_bytes = h_images.numel() * 4 + h_queries.numel() * 8 + h_qmask.numel() + sum(t.numel() * 8 for t in h_resp.values())

    def to_dev():
        return dict(images=h_images.to(dev, non_blocking=True), queries=h_queries.to(dev, non_blocking=True),
                    queries_attn_masks=h_qmask.to(dev, non_blocking=True),
                    **{k: v.to(dev, non_blocking=True) for k, v in h_resp.items()})

    d_batch = to_dev()
    torch.cuda.synchronize()

    def step_device():
        return stepper.step(d_batch, "chosen_response", "rejected_response")

    def step_e2e():
        loss, norm = stepper.step(to_dev(), "chosen_response", "rejected_response")
        return float(loss.item())   # device -> host read of the step's result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out

    for _ in range(args.warmup):
        step_device()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = model.engine.launch_count()
    ms_total, (loss, norm) = timed(step_device, args.steps)
    launches = model.engine.launch_count() - l0
    clk = clocks.stop() if rank == 0 else None
    step_e2e()
    ms_e2e, loss_host = timed(step_e2e, args.steps)

    # instrumented step: CUDA events around every tcgen05 GEMM launch (the dominant kernel family)
    lib = _lib.load()
    barrier()
    lib.opadpo_prof_begin()
    step_device()
    gm, gf, gl = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
    lib.opadpo_prof_end(ctypes.byref(gm), ctypes.byref(gf), ctypes.byref(gl))
    barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    ms_step = ms_total / args.steps
    pairs_per_step = b * world
    value = pairs_per_step / (ms_step / 1e3)
    e2e_value = pairs_per_step / (ms_e2e / args.steps / 1e3)
    fl = flops_per_pair(cfg, Q_LEN, R)
    peaks = _peaks()
    traffic = None
    prof = os.path.join(ROOT, "profiles", "r01_ncu_full_gemm_fwd_layer.json")
    if args.model == "7b" and os.path.exists(prof):
        try:  # dram bytes of the dominant launch (fused gate|up forward GEMM) from the committed ncu --set full capture
            with open(prof) as f:
                traffic = json.load(f).get("dominant_launch_dram_bytes")
        except Exception:
            traffic = None
    gemm_tflops = gf.value / (gm.value * 1e-3) / 1e12 if gm.value > 0 else None
    step_tflops = fl["per_pair"] * b / (ms_step * 1e-3) / 1e12   # per GPU, algorithmic
    line = {
        "metric": "preference-pairs/sec LLaVA-1.5-7B OPA-DPO", "value": value, "unit": "pairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"LLaVA-1.5-{args.model.upper()} LoRA(r=256, all Linear but lm_head) OPA-DPO step, 336^2 img, Q=128, "
                               f"R={R}, 4 forwards (policy+frozen ref x chosen+rejected) + LoRA backward + clip + AdamW",
                   "pairs_per_gpu": b, "global_pairs": pairs_per_step, "seq_len": Q_LEN - 1 + cfg.n_patches + R,
                   "prefix_sharing": (not args.no_share_prefix), "merged_reference_adapter": bool(args.merge_frozen),
                   "fused_rope_epilogue": bool(args.fuse_rope), "fused_swiglu_epilogue": bool(args.fuse_swiglu),
                   "parallelism": f"dp{world}", "l2": "inputs larger than L2 (weights 13.5 GB + activations >> 126 MB)",
                   "collective": "NCCL all-reduce(sum) of the fp32 LoRA-grad bucket once per step, issued as 33 per-layer slices "
                                 "overlapped with the backward" if world > 1 else "none",
                   "loss": float(loss.item()), "grad_norm": float(norm.item())},
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps, "loss": loss_host},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "achieved": gemm_tflops, "peak": peaks["sustained"], "unit": "TFLOP/s",
                     "frac": (gemm_tflops / peaks["sustained"]) if gemm_tflops else None, "traffic": traffic,
                     "traffic_note": "dram read+write bytes of ONE dominant launch (gate|up forward GEMM, M=13816 N=22016 K=4096+256; "
                                     "algorithmic 0.92 GB) from profiles/r01_ncu_full_gemm_fwd_layer.json",
                     "kernel": "opadpo::gemm_kernel (tcgen05; CTA-pair cta_group::2 256x256x64 tiles for the large GEMMs, 128x256x64 otherwise; all epilogues)", "peak_source": peaks["source"] + " sustained",
                     "gemm_launches_per_step": int(gl.value), "gemm_ms_per_step": gm.value,
                     "gemm_share_of_step": gm.value / ms_step,
                     "step_algorithmic_tflops_per_gpu": step_tflops,
                     "step_executed_gemm_tflop": gf.value / 1e12,
                     "step_frac_of_burst_peak": step_tflops / peaks["burst"],
                     "step_frac_of_sustained_peak": step_tflops / peaks["sustained"],
                     "tflop_per_pair": fl["per_pair"] / 1e12},
        "clocks": clk,
    }
    if not args.no_cpu_baseline and world == 1:
        try:
            threads = os.cpu_count() or 1
            t, sfl = cpu_reference_sample(threads)
            cpu_flops = sfl / t
            line["cpu_baseline"] = {
                "value": cpu_flops / fl["per_pair"], "unit": "pairs/s", "cores": threads, "kind": "port",
                "sample": "1 pair, R=512, fp32 oracle at true 7B widths, depth truncated to 2 LLaMA + 2 CLIP layers "
                          f"({t:.1f} s, {sfl / 1e12:.1f} TFLOP); extrapolated to full depth by algorithmic FLOPs",
                "cpu_tflops": cpu_flops / 1e12}
        except Exception as ex:  # the baseline is informative; never lose the GPU line over it
            line["cpu_baseline"] = {"value": None, "unit": "pairs/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"failed: {type(ex).__name__}: {ex}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
