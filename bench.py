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
`rollout` : (N=1) BASELINE configs[3] next to the training line — greedy on-policy rollout, batch 32, 336^2 image, query 128
           (prefill 703 positions), 256 new tokens: tokens/s end to end (host ids/images in, host ids out), decode
           tokens/s, the decode step's HBM roofline (weights once + KV cache per step vs the measured copy peak), kernel
           launches per decode step, and a bounded CPU-port sample. `--mode rollout` prints it as the line itself.
`cpu_baseline` / `--impl reference`: the oracle (CPU fp32 port of the reference path) on the host cores, bounded sample;
           the reference arm also times ONE full-depth configs[0] step (2 pairs, R=128, all 32+23 layers, fwd+bwd).
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


def cpu_full_depth_config0_step(threads: int, pairs: int = 2, R: int = 128):
    """BASELINE configs[0] as the reference runs it on a CPU: the FULL-depth LLaVA-1.5-7B (32 LLaMA + 23 executed CLIP
    layers, LoRA r=256), `pairs` preference pairs with R-token responses, fp32: reference forward x2, policy forward x2,
    DPO loss, autograd backward through the policy adapter (SURVEY.md §8d asked for exactly this measurement). Returns
    (seconds for the step, algorithmic FLOPs of it, seconds spent building the weights)."""
    import torch
    from oracle import llava_oracle as O
    from opa_dpo_b200.engine import EngineConfig
    from opa_dpo_b200.flops import flops_per_pair
    torch.set_num_threads(threads)
    t_init = time.perf_counter()
    oc = O.OracleConfig(**O.LLAVA_7B)
    w = O.init_weights(oc, seed=1)
    lp, lr = O.init_lora(oc, 2), O.init_lora(oc, 3)
    t_init = time.perf_counter() - t_init
    images, q, qm, resp = O.synthetic_batch(oc, pairs, Q_LEN, R, ["chosen_response", "rejected_response"], ragged=False,
                                            image_pos=IMG_POS)
    lpg = {m: (A.clone().requires_grad_(not m.startswith(O.VT)), B.clone().requires_grad_(not m.startswith(O.VT)))
           for m, (A, B) in lp.items()}
    t0 = time.perf_counter()
    with torch.no_grad():
        ref = O.policy_forward(oc, w, lr, images, q, qm, resp)
    # gradient checkpointing per decoder layer, as the reference trains (opadpo_train.py:324-327): same numbers, and the
    # step's host memory stays near the 28 GB of fp32 weights instead of ~100 GB of kept activations
    pol = O.policy_forward(oc, w, lpg, images, q, qm, resp, checkpoint_layers=True)
    loss, _, _ = O.pair_loss(O.LossConfig(beta=0.1), pol["chosen_response_logprobs"], pol["rejected_response_logprobs"],
                             ref["chosen_response_logprobs"], ref["rejected_response_logprobs"])
    loss.backward()
    t = time.perf_counter() - t0
    return t, flops_per_pair(EngineConfig(), Q_LEN, R)["per_pair"] * pairs, t_init


FULL_DEPTH_NEED_GB = 50.0     # measured peak RSS of the checkpointed configs[0] step (43.5 GB) + margin


def _available_host_gb():
    """Host memory this process may still take: MemAvailable, capped by the cgroup limit when there is one."""
    avail = None
    try:
        with open("/proc/meminfo") as f:
            for ln in f:
                if ln.startswith("MemAvailable:"):
                    avail = int(ln.split()[1]) / 1e6
    except OSError:
        pass
    for lim_p, cur_p in (("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory.current"),
                         ("/sys/fs/cgroup/memory/memory.limit_in_bytes", "/sys/fs/cgroup/memory/memory.usage_in_bytes")):
        try:
            lim = open(lim_p).read().strip()
            if lim != "max" and int(lim) < (1 << 60):
                room = (int(lim) - int(open(cur_p).read().strip())) / 1e9
                avail = room if avail is None else min(avail, room)
        except (OSError, ValueError):
            pass
    return avail


def full_depth_config0_isolated(threads: int, timeout_s: float):
    """cpu_full_depth_config0_step in a child process: a host that cannot hold the 7B fp32 model (the kernel's OOM killer ends
    the child, not this process) or is too slow for the time box costs the reference arm its full-depth figure, never its
    line. Returns the `full_depth_config0` object of the JSON line."""
    avail = _available_host_gb()
    if avail is not None and avail < FULL_DEPTH_NEED_GB:
        return {"skipped": f"needs ~{FULL_DEPTH_NEED_GB:.0f} GB of host memory, {avail:.0f} GB available"}
    cmd = [sys.executable, os.path.abspath(__file__), "--full-depth-child", str(threads)]
    try:
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s)
    except subprocess.TimeoutExpired:
        return {"failed": f"no result within the {timeout_s:.0f} s time box"}
    for ln in reversed(p.stdout.strip().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)
    return {"failed": f"child exited {p.returncode}: {(p.stderr or '').strip()[-200:]}"}


def _full_depth_child(threads: int) -> int:
    import resource
    from opa_dpo_b200.flops import flops_per_pair
    from opa_dpo_b200.engine import EngineConfig
    full = flops_per_pair(EngineConfig(), Q_LEN, R_LEN)["per_pair"]
    t_fd, fl_fd, t_init = cpu_full_depth_config0_step(threads)
    print(json.dumps({"config": "BASELINE configs[0]: full-depth LLaVA-1.5-7B (32 LLaMA + 23 CLIP layers), 2 pairs, R=128, fp32, "
                                "reference fwd x2 + policy fwd x2 + DPO loss + LoRA backward, gradient checkpointing per decoder "
                                "layer as in the reference's recipe",
                      "seconds_per_step": t_fd, "pairs_per_s_at_R128": 2.0 / t_fd, "cpu_tflops": fl_fd / t_fd / 1e12,
                      "pairs_per_s_at_R512_by_flops": (fl_fd / t_fd) / full, "weight_init_seconds": t_init,
                      "peak_rss_gb": resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6}), flush=True)
    return 0


def cpu_rollout_sample(threads: int, batch: int = 4, new: int = 8, n_layers: int = 2, v_layers: int = 2):
    """Bounded CPU sample of the rollout (configs[3]): the oracle's greedy loop (HF generate semantics, KV cache) at true 7B
    widths, depth truncated to n_layers LLaMA / v_layers CLIP layers, `batch` sequences, `new` tokens. Returns seconds
    for prefill + new tokens and seconds for prefill + 1 token (their difference / (new-1) = one decode step)."""
    import torch
    from oracle import llava_oracle as O
    torch.set_num_threads(threads)
    oc = O.OracleConfig(**{**O.LLAVA_7B, "n_layers": n_layers, "v_layers": v_layers})
    w = O.init_weights(oc, seed=1)
    lp = O.init_lora(oc, 2)
    images, q, qm, _ = O.synthetic_batch(oc, batch, Q_LEN, 8, ["chosen_response"], ragged=False, image_pos=IMG_POS)
    t0 = time.perf_counter()
    O.greedy_generate(oc, w, lp, images, q, qm, 1)
    t1 = time.perf_counter()
    O.greedy_generate(oc, w, lp, images, q, qm, new)
    t2 = time.perf_counter()
    return t2 - t1, t1 - t0


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
    full_depth = None
    if not args.no_full_depth and args.gpus == 1:
        # one un-truncated step: configs[0] (2 pairs x 128 tokens), every layer, forward and backward — in a child process
        # behind a memory check and a time box, so that it can only ever add to the line
        full_depth = full_depth_config0_isolated(threads, args.full_depth_timeout)
    elif not args.no_full_depth:
        # the CPU arm does not depend on the GPU count: the N > 1 launches of a scaling run repeat the bounded sample only
        full_depth = {"skipped": "measured by this arm's N=1 launch (bench.py --impl reference --gpus 1); the CPU path is the same at every N"}
    if full_depth and "pairs_per_s_at_R512_by_flops" in full_depth:
        # the line's value is the HIGHER of the two CPU estimates (the one less flattering to the GPU arm)
        value = max(value, full_depth["pairs_per_s_at_R512_by_flops"])
        sample += ("; beside it ONE full-depth configs[0] step measured un-truncated (full_depth_config0); value = the higher of "
                   "the two estimates")
    line = {
        "impl": "reference", "metric": "preference-pairs/sec LLaVA-1.5-7B OPA-DPO", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "LLaVA-1.5-7B LoRA(r=256) OPA-DPO step, 336^2 img, Q=128, R=512 (CPU oracle port of the reference path)"},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": sample,
                         "cpu_tflops": cpu_flops / 1e12, "truncated_depth_estimate": cpu_flops / full,
                         "full_depth_config0": full_depth},
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


# ------------------------------------------------------------------------------------------------ rollout (configs[3])
def measure_rollout(model, cfg, dev, batch: int, new: int, peaks: dict, cpu_sample: bool):
    """BASELINE configs[3]: greedy on-policy rollout (online_generator.py:292-309 with do_sample=False), `batch` sequences,
    336^2 image + 128-token query (prefill 703 positions), `new` tokens, through the public `model.generate` call of the
    trainable policy adapter (merged into the decoder weights once per call, as after every optimizer step).
    e2e: pinned host images / ids copied in, generated ids read back to the host, all inside the timed region."""
    import torch
    eng = model.engine
    g = torch.Generator().manual_seed(3)
    h_images = torch.randn(batch, 3, cfg.v_image, cfg.v_image, generator=g).pin_memory()
    h_q = torch.randint(3, cfg.vocab, (batch, Q_LEN), generator=g)
    h_q[:, IMG_POS] = cfg.image_token
    h_q = h_q.pin_memory()
    h_m = torch.ones(batch, Q_LEN, dtype=torch.uint8).pin_memory()
    model.set_adapter("lora_policy")

    def run(n_new, host=True):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        out = model.generate(inputs=h_q.to(dev, non_blocking=True), images=h_images.to(dev, non_blocking=True),
                             attention_mask=h_m.to(dev, non_blocking=True), do_sample=False, max_new_tokens=n_new,
                             pad_token_id=cfg.pad_id)
        ids = out.cpu() if host else out
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), ids

    run(8)                                                   # warm-up (kernel attributes, graph capture path, pool growth)
    run(8)
    run(new)
    l0 = eng.launch_count()
    run(1)
    l1 = eng.launch_count()
    run(2)
    l2 = eng.launch_count()
    launches_per_decode_step = (l2 - l1) - (l1 - l0)         # one eager decode step (the rest replays it as ONE CUDA graph)
    t_prefill = min(run(1)[0] for _ in range(2))             # multimodal prefill + adapter merge + first token
    t_full, ids = min((run(new) for _ in range(3)), key=lambda x: x[0])
    decode_ms = (t_full - t_prefill) / (new - 1)
    n_lin = cfg.n_layers * (4 * cfg.d_model ** 2 + 3 * cfg.d_model * cfg.d_ffn)
    w_bytes = 2 * (n_lin + cfg.vocab * cfg.d_model)          # merged bf16 decoder weights + lm_head, read once per step
    ctx = Q_LEN - 1 + cfg.n_patches + new / 2                # mean context length over the decode steps
    kv_bytes = 2 * 2 * cfg.n_layers * cfg.d_model * ctx * batch
    gbs = (w_bytes + kv_bytes) / (decode_ms * 1e-3) / 1e9
    out = {
        "workload": f"BASELINE configs[3]: LLaVA-1.5-7B greedy rollout, batch {batch}, 336^2 img, Q=128 (prefill 703), {new} new tokens, 1 GPU",
        "metric": "rollout tokens/s", "value": batch * new / (t_full * 1e-3), "unit": "tokens/s",
        "e2e": {"value": batch * new / (t_full * 1e-3), "unit": "tokens/s",
                "h2d_bytes_per_step": h_images.numel() * 4 + h_q.numel() * 8 + h_m.numel(),
                "d2h_bytes_per_step": batch * (Q_LEN + new) * 8},
        "total_ms": t_full, "prefill_ms": t_prefill, "decode_ms_per_step": decode_ms,
        "decode_tokens_per_s": batch / (decode_ms * 1e-3),
        "kernel_launches_per_decode_step": int(launches_per_decode_step),
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                     "peak_source": peaks["source"] + " copy bandwidth", "weight_bytes_per_step": w_bytes,
                     "kv_bytes_per_step_avg": kv_bytes,
                     "kernel": "decode step = split-K weight-streaming opadpo::gemm_kernel<EPI_F32_PARTIAL> + fused partial reducers + decode_attention"},
        "generated_nonpad": int((ids[:, Q_LEN:] != cfg.pad_id).sum().item()),
    }
    if cpu_sample:
        try:
            threads = os.cpu_count() or 1
            cb, cn, cl = 4, 8, 2
            t_new, t_one = cpu_rollout_sample(threads, cb, cn, cl, cl)
            step_s = max(t_new - t_one, 1e-9) / (cn - 1)
            # linear in depth and batch (the CPU decode step streams the fp32 weights: batch-insensitive in truth, so this
            # flatters the CPU), prefill linear in depth and batch
            est = (t_one * (cfg.n_layers / cl) * (batch / cb)) + (new - 1) * step_s * (cfg.n_layers / cl)
            out["cpu_baseline"] = {"value": batch * new / est, "unit": "tokens/s", "cores": threads, "kind": "port",
                                   "sample": f"oracle greedy loop at true 7B widths, {cl} LLaMA + {cl} CLIP layers, batch {cb}, {cn} new "
                                             f"tokens ({t_new:.1f} s; prefill+1 token {t_one:.1f} s); extrapolated linearly in depth "
                                             "(both phases) and in batch (prefill only)"}
        except Exception as ex:
            out["cpu_baseline"] = {"value": None, "unit": "tokens/s", "cores": os.cpu_count(), "kind": "port",
                                   "sample": f"failed: {type(ex).__name__}: {ex}"}
    return out


# ------------------------------------------------------------------------------------------------ GPU arm
def collective_note(stepper, world):
    """Which all-reduce of the LoRA-grad bucket actually ran (opa_dpo_b200/dist.py decides it collectively at start-up)."""
    if world == 1 or stepper.overlap is None:
        return "none"
    ov = stepper.overlap
    note = f"all-reduce(sum) of the LoRA-grad bucket once per step, overlapped with the backward: {ov.name}; transport {ov.transport}"
    if getattr(ov, "fallback_reason", ""):
        note += f" (NVLink kernel unavailable: {ov.fallback_reason[:200]})"
    if hasattr(ov, "check"):
        ov.check()
    return note


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
    ap.add_argument("--no-full-depth", action="store_true", help="reference arm: skip the full-depth configs[0] CPU step (~4 min)")
    ap.add_argument("--full-depth-timeout", type=float, default=900.0, help="reference arm: time box of the full-depth step (s)")
    ap.add_argument("--full-depth-child", type=int, default=0, metavar="THREADS", help=argparse.SUPPRESS)
    ap.add_argument("--mode", default="train", choices=["train", "rollout"],
                    help="rollout: print BASELINE configs[3] (greedy rollout, batch 32, 256 new tokens) as the line itself")
    ap.add_argument("--no-rollout", action="store_true", help="train mode: skip the rollout object")
    ap.add_argument("--diag-no-collective", action="store_true",
                    help="diagnostic (N > 1): skip the gradient all-reduce, every rank steps on its own shard; the line then shows "
                         "what max-over-ranks timing alone costs against one GPU (not a valid bench value: marked in config)")
    ap.add_argument("--profile", action="store_true",
                    help="for runs UNDER ncu only: warm-up as given (may be < 3), --steps device steps, then exit; prints no bench line")
    ap.add_argument("--rollout-batch", type=int, default=32)
    ap.add_argument("--rollout-new", type=int, default=256)
    ap.add_argument("--merge-frozen", type=int, default=int(os.environ.get("OPADPO_MERGE_FROZEN", str(MERGE_FROZEN_DEFAULT))),
                    help="1: the frozen reference adapter's forwards run on decoder weights merged once (no LoRA GEMMs)")
    ap.add_argument("--fuse-swiglu", type=int, default=int(os.environ.get("OPADPO_FUSE_SWIGLU", "1")),
                    help="1: SwiGLU in the gate|up GEMM epilogue of the merged reference pass")
    ap.add_argument("--engine-opt", action="append", default=[], metavar="NAME=VALUE",
                    help="extra opadpo_engine_set_option calls (A/B runs), e.g. --engine-opt fuse_swiglu_bwd=0")
    ap.add_argument("--wgrad-stream", type=int, default=int(os.environ.get("OPADPO_WGRAD_STREAM", "1")),
                    help="LoRA weight-gradient GEMMs on a side stream, off the backward's critical path (engine option wgrad_stream)")
    ap.add_argument("--fuse-rope", type=int, default=int(os.environ.get("OPADPO_FUSE_ROPE", "1")),
                    help="1: RoPE in the q|k|v GEMM epilogue")
    args = ap.parse_args()
    if args.full_depth_child:
        return _full_depth_child(args.full_depth_child)
    if args.impl == "reference":
        return run_reference_arm(args)
    if not args.profile:
        args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from opa_dpo_b200.engine import EngineConfig
    from opa_dpo_b200.policy import AutoregressivePolicy, B200LlavaModel
    from opa_dpo_b200.dpo import DPOArgs, DPOStepper
    from opa_dpo_b200.flops import executed_attention_flops_per_sample, flops_per_pair
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
    model.engine.set_option("wgrad_stream", int(bool(args.wgrad_stream)))
    for kv in args.engine_opt:
        name, _, val = kv.partition("=")
        model.engine.set_option(name, int(val))
    model.engine.set_option("fuse_swiglu", int(bool(args.fuse_swiglu)))
    model.engine.random_init(seed=1234)
    model.make_trainable("lora_policy")
    dargs = DPOArgs(beta=0.1, response_len=R, learning_rate=1e-6, max_grad_norm=1.0)
    tok = SimpleNamespace(pad_token_id=cfg.pad_id, eos_token_id=cfg.eos_id)
    policy = AutoregressivePolicy(dargs, model, tok, adapter_name="lora_policy")
    ref_policy = AutoregressivePolicy(dargs, model, tok, adapter_name="lora_ref_policy")
    stepper = DPOStepper(policy, ref_policy, dargs, world_size=1 if args.diag_no_collective else world)

    if args.mode == "rollout":
        if rank == 0:
            peaks = _peaks()
            ro = measure_rollout(model, cfg, dev, args.rollout_batch, args.rollout_new, peaks, not args.no_cpu_baseline)
            line = {"metric": "rollout tokens/s LLaVA-1.5-7B greedy (BASELINE configs[3])", "value": ro["value"], "unit": "tokens/s",
                    "n_gpus": 1, "steps": 3, "warmup": 3, "ms_per_step": ro["total_ms"], "higher_is_better": True,
                    "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                    "config": {"workload": ro["workload"]}, "e2e": ro["e2e"], "roofline": ro["roofline"],
                    "gpu_launches": ro["kernel_launches_per_decode_step"], "rollout": ro}
            if "cpu_baseline" in ro:
                line["cpu_baseline"] = ro["cpu_baseline"]
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

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
    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_run": True, "note": "not a bench value", "ms_per_step_under_profiler": ms_total / args.steps,
                              "launches_per_step": int(launches) // max(args.steps, 1)}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return 0
    step_e2e()
    ms_e2e, loss_host = timed(step_e2e, args.steps)

    # instrumented step: CUDA events around every tcgen05 GEMM launch (the dominant kernel family)
    # (run on ONE stream: with the reference pass on its side stream two GEMM launches share the SMs and each launch's
    # events would bracket the other's tiles too)
    lib = _lib.load()
    barrier()
    concurrent = stepper.concurrent_reference
    stepper.concurrent_reference = False
    model.engine.set_option("wgrad_stream", 0)
    lib.opadpo_prof_begin()
    step_device()
    gm, gf, gl = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
    lib.opadpo_prof_end(ctypes.byref(gm), ctypes.byref(gf), ctypes.byref(gl))
    stepper.concurrent_reference = concurrent
    model.engine.set_option("wgrad_stream", args.wgrad_stream)
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
    traffic_alg = None
    prof = os.path.join(ROOT, "profiles", "r02_ncu_full_gemm_dominant.json")
    if args.model == "7b" and os.path.exists(prof):
        try:  # dram bytes of the dominant launch (merged gate|up forward GEMM, SwiGLU epilogue) from this round's ncu --set full
              # capture of the bench's own step (tools/gpu_r2_m.sh); ncu cannot run inside a timed bench, so it is a committed figure
            with open(prof) as f:
                pj = json.load(f)
            traffic, traffic_alg = pj.get("dram_bytes_read_plus_write"), pj.get("algorithmic_bytes")
        except Exception:
            traffic = None
    gemm_tflops = gf.value / (gm.value * 1e-3) / 1e12 if gm.value > 0 else None
    step_tflops = fl["per_pair"] * b / (ms_step * 1e-3) / 1e12   # per GPU, algorithmic
    # what the engine really executes per step and GPU: every tcgen05 GEMM launch (event-instrumented count, K padded to the
    # 64-wide k-block) + the attention kernels (element-level count under the causal / segment mask; one packed forward
    # for the reference adapter, one for the policy, one backward) + CLIP attention of both forwards
    att = executed_attention_flops_per_sample(cfg, Q_LEN, R, 2, packed=not args.no_share_prefix)
    clip_att = cfg.v_layers * 4.0 * (cfg.n_patches + 1) ** 2 * cfg.v_d
    exec_flops = gf.value + b * (2 * att["fwd"] + att["bwd"] + 2 * clip_att)
    exec_tflops = exec_flops / (ms_step * 1e-3) / 1e12
    line = {
        "metric": "preference-pairs/sec LLaVA-1.5-7B OPA-DPO", "value": value, "unit": "pairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"LLaVA-1.5-{args.model.upper()} LoRA(r=256, all Linear but lm_head) OPA-DPO step, 336^2 img, Q=128, "
                               f"R={R}, 4 forwards (policy+frozen ref x chosen+rejected) + LoRA backward + clip + AdamW",
                   "pairs_per_gpu": b, "global_pairs": pairs_per_step, "seq_len": Q_LEN - 1 + cfg.n_patches + R,
                   "prefix_sharing": (not args.no_share_prefix), "merged_reference_adapter": bool(args.merge_frozen),
                   "fused_rope_epilogue": bool(args.fuse_rope), "fused_swiglu_epilogue": bool(args.fuse_swiglu),
                   "lora_weight_gradients": ("side stream, under the backward's HBM-bound kernels" if args.wgrad_stream
                                             else "main stream"),
                   "engine_options": args.engine_opt,
                   "reference_pass": ("side stream, concurrent with the policy forward" if stepper.concurrent_reference
                                      else "same stream, before the policy forward"),
                   "parallelism": f"dp{world}", "l2": "inputs larger than L2 (weights 13.5 GB + activations >> 126 MB)",
                   "collective": ("DIAGNOSTIC RUN: gradient all-reduce skipped, ranks independent (not a bench value)"
                                  if args.diag_no_collective and world > 1 else collective_note(stepper, world)),
                   "loss": float(loss.item()), "grad_norm": float(norm.item())},
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps, "loss": loss_host},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "achieved": gemm_tflops, "peak": peaks["sustained"], "unit": "TFLOP/s",
                     "frac": (gemm_tflops / peaks["sustained"]) if gemm_tflops else None, "traffic": traffic,
                     "traffic_algorithmic": traffic_alg,
                     "traffic_note": "dram read+write bytes of ONE dominant launch (merged gate|up forward GEMM with the SwiGLU epilogue, "
                                     "M=13816 N=22016 K=4096, 1.53 ms, tensor pipe 98.7 % active) against its algorithmic bytes, from "
                                     "profiles/r02_ncu_full_gemm_dominant.json (ncu --set full of this bench's step, round 2)",
                     "kernel": "opadpo::gemm_kernel (tcgen05; CTA-pair cta_group::2 256x256x64 tiles for the large GEMMs, 128x256x64 otherwise; all epilogues)", "peak_source": peaks["source"] + " sustained",
                     "gemm_launches_per_step": int(gl.value), "gemm_ms_per_step": gm.value,
                     "gemm_share_of_step": gm.value / ms_step,
                     "step_algorithmic_tflops_per_gpu": step_tflops,
                     "step_executed_gemm_tflop": gf.value / 1e12,
                     "step_frac_of_burst_peak": step_tflops / peaks["burst"],
                     "step_frac_of_sustained_peak": step_tflops / peaks["sustained"],
                     "frac_note": "frac / achieved = the GEMM family alone (its executed FLOPs / its summed kernel time); "
                                  "step_executed_* = ALL executed tensor FLOPs (GEMMs + attention) / the whole step's time; "
                                  "step_frac_of_* = the reference's ALGORITHMIC FLOPs (no prefix sharing, unmerged LoRA) / step time",
                     "step_executed_tflop": exec_flops / 1e12,
                     "step_executed_attention_tflop": (exec_flops - gf.value) / 1e12,
                     "step_executed_tflops": exec_tflops,
                     "step_executed_frac_burst": exec_tflops / peaks["burst"],
                     "step_executed_frac_sustained": exec_tflops / peaks["sustained"],
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
    if world == 1 and not args.no_rollout and args.model == "7b":
        try:
            line["rollout"] = measure_rollout(model, cfg, dev, args.rollout_batch, args.rollout_new, peaks, not args.no_cpu_baseline)
        except Exception as ex:   # never lose the training line over the second workload
            line["rollout"] = {"failed": f"{type(ex).__name__}: {ex}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
