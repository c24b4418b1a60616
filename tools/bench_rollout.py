"""Rollout throughput (BASELINE.json configs[3]): LLaVA-1.5-7B greedy on-policy rollout, 336^2 image, 128-token query
(prefill 703 positions), B sequences, `new` tokens. Reports tokens/s and the HBM roofline of the decode phase:
algorithmic bytes per decode step = bf16 weights read once (LLM linears + lm_head + LoRA) + KV cache read for the
current context, peak = MEASURED_PEAKS.json hbm_gbs. CUDA events, 1 warm-up run."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opa_dpo_b200 import Engine, EngineConfig, POLICY


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--new", type=int, default=256)
    ap.add_argument("--runs", type=int, default=2)
    ap.add_argument("--sample", action="store_true", help="temperature 1.0 / top_k 30 / top_p 0.95 (online_generate.sh:75-77) instead of greedy")
    args = ap.parse_args()
    cfg = EngineConfig()
    eng = Engine(cfg)
    eng.random_init(seed=1234)
    dev = eng.device
    g = torch.Generator().manual_seed(3)
    images = torch.randn(args.batch, 3, 336, 336, generator=g).to(dev)
    q = torch.randint(3, cfg.vocab, (args.batch, 128), generator=g)
    q[:, 35] = cfg.image_token
    q = q.to(dev)
    qm = torch.ones(args.batch, 128, dtype=torch.uint8, device=dev)

    def timed(new):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        kw = dict(do_sample=True, top_k=30, top_p=0.95, temperature=1.0, seed=7) if args.sample else {}
        out = eng.generate(POLICY, images, q, qm, new, **kw)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), out

    timed(8)
    t_prefill = min(timed(1)[0] for _ in range(args.runs))          # prefill + first token
    t_full, out = min((timed(args.new) for _ in range(args.runs)), key=lambda x: x[0])
    decode_ms = (t_full - t_prefill) / (args.new - 1)
    n_lin = cfg.n_layers * (4 * cfg.d_model ** 2 + 3 * cfg.d_model * cfg.d_ffn)
    n_lora = cfg.n_layers * cfg.lora_r * (8 * cfg.d_model + 3 * (cfg.d_model + cfg.d_ffn))
    w_bytes = 2 * (n_lin + n_lora + cfg.vocab * cfg.d_model)
    ctx = 127 + 576 + args.new / 2
    kv_bytes = 2 * 2 * cfg.n_layers * cfg.d_model * ctx * args.batch
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    gbs = (w_bytes + kv_bytes) / (decode_ms * 1e-3) / 1e9
    print(json.dumps({
        "metric": "rollout tokens/s LLaVA-1.5-7B greedy", "batch": args.batch, "new_tokens": args.new,
        "total_ms": t_full, "prefill_ms": t_prefill, "decode_ms_per_step": decode_ms,
        "tokens_per_s_e2e": args.batch * args.new / (t_full * 1e-3),
        "tokens_per_s_decode": args.batch / (decode_ms * 1e-3),
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                     "weight_bytes_per_step": w_bytes, "kv_bytes_per_step_avg": kv_bytes},
        "generated_nonpad": int((out[:, 128:] != 0).sum().item())}))


if __name__ == "__main__":
    main()
