"""Hardware check of the data-parallel step (reference loop rl_trainer.py:155-175). Launched by
tests/test_dist_gpu.py through torchrun, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P tests/dist_gpu_check.py

Every rank builds the same engine (same seeds), takes its shard of ONE global batch of preference pairs and runs the real
DPOStepper step with the overlapped all-reduce armed (split-K `red.global.add` wgrad epilogues racing the per-layer NCCL
slices started from the backward's host callback). Checked:
  * the all-reduced gradient bucket and the post-AdamW masters are BIT-IDENTICAL on all ranks;
  * against the same step on ONE GPU over the concatenated batch (rank 0 runs it afterwards on a second engine): gradient
    bucket (scaled by 1/world) and master UPDATE cosine >= 0.999 and relative L2 <= 2e-2 (bf16 forward/backward + fp32
    atomics noise; the two computations tile their batches differently), loss equal to the mean of the ranks' losses;
  * the same without the overlap (one bucket-wide all-reduce) gives the same reduced bucket to fp32-atomics noise.
Writes gpurun_out/dist_gpu_check.json on rank 0; exit code 0 only if everything held on every rank."""
import json
import os
import sys
from types import SimpleNamespace

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

NAMES = ("chosen_response", "rejected_response")


def build(cfg, dev, world, lr):
    from opa_dpo_b200.dpo import DPOArgs, DPOStepper
    from opa_dpo_b200.policy import AutoregressivePolicy, B200LlavaModel
    model = B200LlavaModel(cfg, dev)
    model.engine.set_option("merge_frozen", 1)
    model.engine.random_init(seed=4321)
    model.make_trainable("lora_policy")
    args = DPOArgs(beta=0.1, response_len=R, learning_rate=lr, max_grad_norm=1.0)
    tok = SimpleNamespace(pad_token_id=cfg.pad_id, eos_token_id=cfg.eos_id)
    pol = AutoregressivePolicy(args, model, tok, adapter_name="lora_policy")
    ref = AutoregressivePolicy(args, model, tok, adapter_name="lora_ref_policy")
    return model, DPOStepper(pol, ref, args, world_size=world)


def global_batch(cfg, pairs):
    g = torch.Generator().manual_seed(99)
    images = torch.randn(pairs, 3, cfg.v_image, cfg.v_image, generator=g)
    q = torch.randint(3, cfg.vocab, (pairs, Q), generator=g)
    qm = torch.ones(pairs, Q, dtype=torch.uint8)
    for i in range(pairs):
        q[i, :i] = cfg.pad_id
        qm[i, :i] = 0
        q[i, i + 3] = cfg.image_token
    resp = {}
    for n in NAMES:
        t = torch.randint(3, cfg.vocab, (pairs, R), generator=g)
        for i in range(pairs):
            k = R - 5 * i
            t[i, k - 1] = cfg.eos_id
            t[i, k:] = cfg.pad_id
        resp[n] = t
    return dict(images=images, queries=q, queries_attn_masks=qm, **resp)


def shard(batch, lo, hi, dev):
    return {k: v[lo:hi].to(dev) for k, v in batch.items()}


def cos_rel(a, b):
    a, b = a.double(), b.double()
    cos = float((a @ b) / (a.norm() * b.norm()))
    return cos, float((a - b).norm() / b.norm())


Q, R = 24, 96


def main():
    from opa_dpo_b200.engine import EngineConfig
    from opa_dpo_b200.dist import shard_range
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    # small widths, real structure: 6 decoder layers so several all-reduce slices are in flight under the backward
    cfg = EngineConfig(d_model=1024, n_layers=6, n_heads=8, d_ffn=2816, vocab=4096, v_d=256, v_layers=2, v_heads=4, v_ffn=1024,
                       v_image=112, v_patch=14, lora_r=64, lora_alpha=128.0, max_positions=512)
    pairs = 2 * world
    batch = global_batch(cfg, pairs)
    lo, hi = shard_range(pairs, rank, world)
    lr = 1e-3
    ok, report = True, {"world": world, "pairs": pairs}

    def same_on_all_ranks(t):
        hi_, lo_ = t.clone(), t.clone()
        dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
        return bool(torch.equal(hi_, lo_))

    # ---- the collective kernel on its own (csrc/collective.cu): integers are exact in bf16 and in the fp32 sum, so the result
    # must equal the exact sum bit for bit; both transports (in-switch multimem reduction, peer loads), ragged region sizes
    from opa_dpo_b200.dist import NvlGradAllReduce
    report["kernel"] = {}
    for wire in ("fp32", "bf16"):
        for mc in ("1", "0"):
            os.environ["OPADPO_DP_MULTICAST"] = mc
            nel = 8 * 100003
            g = torch.Generator(device=dev).manual_seed(17 + rank)
            x = torch.randint(-16, 16, (nel,), device=dev, generator=g).float()
            want = x.clone()
            dist.all_reduce(want, op=dist.ReduceOp.SUM)
            ar = NvlGradAllReduce(None, x, None, wire=wire)
            x = ar.grad                              # fp32 in place: the symmetric copy of x
            ar._launch(0, 8 * 4096, 3)              # three launches over disjoint ranges, different grids
            ar._launch(8 * 4096, 8 * 50000, 16)
            ar._launch(8 * 54096, nel - 8 * 54096, 128)
            sc = ar.finish()
            torch.cuda.synchronize()
            ar.check()
            exact = bool(torch.equal(x, want))
            # generic values. fp32 in place: the fp32 sum (two ranks: exact; more: summation order). bf16 wire: bf16-rounded
            # inputs, fp32 accumulation, one rounding of the sum (the switch's result is within one bf16 ulp)
            y = torch.randn(nel, device=dev, generator=g)
            wanty = y.clone() if wire == "fp32" else y.to(torch.bfloat16).float()
            dist.all_reduce(wanty, op=dist.ReduceOp.SUM)
            ar.grad.copy_(y)
            y = ar.grad
            ar.armed = False
            ar.finish()                              # nothing launched: one bucket-wide call
            torch.cuda.synchronize()
            ar.check()
            err = float((y - wanty).abs().max() / wanty.abs().max())
            tol = 1e-6 if wire == "fp32" else 2.0 ** -7
            report["kernel"][f"{wire}/{'multimem' if ar.mc_ptr else 'peer'}"] = {
                "requested_multicast": mc, "exact_on_integers": exact, "randn_max_rel_err": err, "tol": tol, "scale": sc,
                "identical_across_ranks": same_on_all_ranks(y)}
            ok = ok and exact and err <= tol and sc == 1.0 / world
            del ar
    os.environ.pop("OPADPO_DP_MULTICAST", None)

    model, stepper = build(cfg, dev, world, lr)
    report["collective"] = {"name": stepper.overlap.name, "transport": stepper.overlap.transport,
                            "fallback_reason": getattr(stepper.overlap, "fallback_reason", "")}
    bf16_wire = getattr(stepper.overlap, "wire", "") == "bf16"
    v = model.engine.buffers()
    n = v["n_train"]
    master0 = v["master"][:n].clone()
    local_batch = shard(batch, lo, hi, dev)

    # ---- the DP step as the product runs it: overlapped slices
    refs = stepper.reference_logprobs(local_batch)
    loss, _ = stepper.pair_loss(local_batch, refs, *NAMES)
    scale = stepper.overlap.finish()
    torch.cuda.synchronize()
    g_dp = v["grad"].clone()
    report["grad_identical_across_ranks"] = same_on_all_ranks(g_dp)
    # ---- the same gradient without the overlap: local bucket, then ONE all-reduce
    model.engine.zero_grad()
    stepper.overlap.armed = False
    out = stepper.policy(**local_batch, temperature=1.0)
    total, _ = stepper.loss.pair(out[NAMES[0] + "_logprobs"], out[NAMES[1] + "_logprobs"],
                                 refs["ref_base_" + NAMES[0] + "_logprobs"], refs["ref_base_" + NAMES[1] + "_logprobs"])
    total.backward()
    dist.all_reduce(v["grad"], op=dist.ReduceOp.SUM)
    torch.cuda.synchronize()
    report["overlap_vs_bucket_cos_rel"] = cos_rel(g_dp, v["grad"])
    # ---- AdamW on the overlapped result
    v["grad"].copy_(g_dp)
    norm = model.engine.optimizer_step(lr, (0.9, 0.999), 1e-8, 0.0, 1.0, scale)
    torch.cuda.synchronize()
    master_dp = v["master"][:n].clone()
    report["master_identical_across_ranks"] = same_on_all_ranks(master_dp)
    losses = [torch.zeros((), device=dev) for _ in range(world)]
    dist.all_gather(losses, loss.float())
    report["scale"] = scale
    report["dp_grad_norm"] = float(norm)
    report["rank_losses"] = [float(x) for x in losses]
    model.engine.close()
    del model, stepper, v
    torch.cuda.empty_cache()

    if rank == 0:
        # ---- the same step on ONE GPU over the concatenated batch
        model1, st1 = build(cfg, dev, 1, lr)
        v1 = model1.engine.buffers()
        assert torch.equal(v1["master"][:n], master0), "same seeds must give the same initial adapter"
        full = shard(batch, 0, pairs, dev)
        refs1 = st1.reference_logprobs(full)
        loss1, _ = st1.pair_loss(full, refs1, *NAMES)
        torch.cuda.synchronize()
        g1 = v1["grad"].clone()
        norm1 = model1.engine.optimizer_step(lr, (0.9, 0.999), 1e-8, 0.0, 1.0, 1.0)
        torch.cuda.synchronize()
        report["grad_cos_rel_vs_single_gpu"] = cos_rel(g_dp * scale, g1)
        report["update_cos_rel_vs_single_gpu"] = cos_rel(master_dp - master0, v1["master"][:n] - master0)
        report["single_gpu_loss"] = float(loss1)
        report["single_gpu_grad_norm"] = float(norm1)
        model1.engine.close()
        gc, gr = report["grad_cos_rel_vs_single_gpu"]
        uc, ur = report["update_cos_rel_vs_single_gpu"]
        oc, orel = report["overlap_vs_bucket_cos_rel"]
        mean_loss = sum(report["rank_losses"]) / world
        checks = {
            "grad bucket bit-identical on all ranks": report["grad_identical_across_ranks"],
            "masters bit-identical on all ranks": report["master_identical_across_ranks"],
            "scale == 1/world": scale == 1.0 / world,
            "DP gradient == single-GPU gradient (cos >= 0.999, rel <= 2e-2)": gc >= 0.999 and gr <= 2e-2,
            "DP AdamW update == single-GPU update (cos >= 0.995)": uc >= 0.995,
            # the NVLink kernel carries bf16 (each rank's slice and the sum are rounded once: <= 2^-8 per element)
            "overlapped slices == one fp32 bucket all-reduce (rel <= 1e-3 fp32 / 5e-3 bf16 wire)":
                oc >= 0.9999 and orel <= (5e-3 if bf16_wire else 1e-3),
            "collective kernel exact on integers and within tolerance on randn, both wires x both transports":
                all(k["exact_on_integers"] and k["randn_max_rel_err"] <= k["tol"] and k["identical_across_ranks"]
                    for k in report["kernel"].values()),
            "mean of rank losses == single-GPU loss (rel <= 1e-3)": abs(mean_loss - report["single_gpu_loss"]) <= 1e-3 * abs(report["single_gpu_loss"]),
            "grad norm agrees (rel <= 2e-2)": abs(report["dp_grad_norm"] - report["single_gpu_grad_norm"]) <= 2e-2 * report["single_gpu_grad_norm"],
        }
        report["checks"] = checks
        ok = all(checks.values())
        report["ok"] = ok
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "dist_gpu_check.json"), "w") as f:
            json.dump(report, f, indent=1)
        print("DIST_GPU_CHECK " + json.dumps(report))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if int(flag.item()) == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
