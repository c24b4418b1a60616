"""Parity at the true LLaVA-1.5-7B dimensions (BASELINE.json configs[0]/[1] shapes).

* true widths, truncated depth (2 LLaMA + 2 CLIP layers), r=256: direct comparison with the CPU fp32 oracle (seconds);
* full 7B depth, device-initialised weights: size-independent properties — bit-exact batch invariance and run-to-run
  determinism of the forward, exact zeros at pads, agreement of the shared-prefix packing with the reference's stacked
  rows, finite LoRA gradients whose norm agrees between the two layouts;
* full 7B depth vs the fp32 CPU oracle (a few minutes of host time, part of the default GPU suite): configs[0] (2 pairs x
  128 tokens, stacked and packed rows) and the headline path itself (R = 512, packed rows, merged reference adapter,
  fused epilogues; three data seeds), DPO loss within 1e-3 relative, per-token statistics written to gpurun_out/.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

NAMES2 = ["chosen_response", "rejected_response"]


def _load(eng, O, oc, w, lp, lr):
    from opa_dpo_b200 import POLICY, REF
    sd = dict(w)
    sd[O.VT + "embeddings.patch_embedding.weight"] = w[O.VT + "embeddings.patch_embedding.weight"].reshape(oc.v_d, -1)
    eng.load_state_dict(sd)
    eng.load_adapter(POLICY, lp)
    eng.load_adapter(REF, lr)


def _oracle_vs_engine(n_layers, v_layers, b, R, grads=True, widths="7b"):
    from oracle import llava_oracle as O
    from opa_dpo_b200 import Engine, EngineConfig, POLICY, REF
    from opa_dpo_b200.dpo import DPOArgs, FusedDPO
    torch.set_num_threads(min(32, os.cpu_count() or 1))
    oc = O.OracleConfig(**{**(O.LLAVA_13B if widths == "13b" else O.LLAVA_7B), "n_layers": n_layers, "v_layers": v_layers})
    ec = EngineConfig(**oc.to_dict())
    w = O.init_weights(oc, seed=21)
    lp, lr = O.init_lora(oc, 22), O.init_lora(oc, 23)
    eng = Engine(ec)
    _load(eng, O, oc, w, lp, lr)
    eng.set_trainable(POLICY)
    images, q, qm, resp = O.synthetic_batch(oc, b, 128, R, NAMES2, left_pad=17, image_pos=35)
    lc = O.LossConfig(beta=0.1)
    lpg = {m: (A.clone().requires_grad_(grads and "vision_model" not in m), B.clone().requires_grad_(grads and "vision_model" not in m))
           for m, (A, B) in lp.items()}
    with torch.no_grad():
        o_ref = O.policy_forward(oc, w, lr, images, q, qm, resp)
    with torch.set_grad_enabled(grads):
        o_pol = O.policy_forward(oc, w, lpg, images, q, qm, resp)
        loss_ref, _, _ = O.pair_loss(lc, o_pol[NAMES2[0] + "_logprobs"], o_pol[NAMES2[1] + "_logprobs"],
                                     o_ref[NAMES2[0] + "_logprobs"], o_ref[NAMES2[1] + "_logprobs"])
    if grads:
        loss_ref.backward()
    dev = eng.device
    args = [images.to(dev), q.to(dev), qm.to(dev), [resp[n].to(dev) for n in NAMES2]]
    ref_lp, _ = eng.forward_logprobs(REF, *args)
    pol_lp, _ = eng.forward_logprobs(POLICY, *args, save_slot=0 if grads else -1)
    for i, n in enumerate(NAMES2):
        for got, ref, tag in ((pol_lp[i], o_pol[n + "_logprobs"].detach(), "policy"), (ref_lp[i], o_ref[n + "_logprobs"], "reference")):
            err = (got.cpu() - ref).abs()
            # bf16 noise of random-init weights grows with the width: mean 0.011-0.013 / max 0.05-0.07 at d = 4096,
            # mean 0.016-0.018 / max 0.07-0.09 at d = 5120
            tol_max, tol_mean = (0.12, 0.025) if widths == "13b" else (0.08, 0.015)
            assert err.max().item() <= tol_max and err.mean().item() <= tol_mean, \
                f"{tag} {n}: max |err| {err.max().item():.4g} mean {err.mean().item():.4g}"
    pl = pol_lp.clone().requires_grad_(True)
    total, _ = FusedDPO(DPOArgs(beta=0.1)).pair(pl[0], pl[1], ref_lp[0], ref_lp[1])
    rel = abs(total.item() - loss_ref.item()) / abs(loss_ref.item())
    assert rel <= 1e-3, f"DPO loss {total.item():.6f} vs oracle {loss_ref.item():.6f} (rel {rel:.3g})"
    if grads:
        total.backward()
        eng.zero_grad()
        eng.backward(0, pl.grad)
        torch.cuda.synchronize()
        g = eng.buffers()["grad"]
        for m in ("model.layers.0.self_attn.q_proj", "model.layers.0.self_attn.v_proj", "model.layers.1.self_attn.o_proj",
                  "model.layers.0.mlp.gate_proj", "model.layers.1.mlp.up_proj", "model.layers.1.mlp.down_proj",
                  "model.mm_projector.0", "model.mm_projector.2"):
            oa, ob = eng.lora_offsets(m)
            for t, off, nm in ((lpg[m][0], oa, "A"), (lpg[m][1], ob, "B")):
                gg = g[off:off + t.numel()].view_as(t).cpu()
                cos = torch.nn.functional.cosine_similarity(gg.flatten(), t.grad.flatten(), dim=0).item()
                assert cos >= 0.99, f"{m}.{nm}: cosine {cos:.4f}"
    eng.close()
    return total.item(), loss_ref.item(), rel


def test_true_width_truncated_depth_vs_oracle():
    _oracle_vs_engine(2, 2, b=1, R=128)


def test_13b_width_truncated_depth_vs_oracle():
    """BASELINE configs[4] dims (LLaVA-1.5-13B: d 5120, 40 heads, FFN 13824: N tiles that do not divide by the CTA-pair
    width, 3 norm vectors per thread, K = 13824 dgrads) at depth 2: per-token log-probs, DPO loss and LoRA gradients vs the
    fp32 oracle, same tolerances as the 7B widths."""
    _oracle_vs_engine(2, 2, b=1, R=128, widths="13b")


@pytest.fixture(scope="module")
def full7b():
    """Full-depth LLaVA-1.5-7B twice: the fp32 CPU oracle's weights (seeded, bf16-representable) and the engine holding the
    same numbers. Built once for the full-depth parity cases below (about a minute of host time)."""
    from oracle import llava_oracle as O
    from opa_dpo_b200 import Engine, EngineConfig, POLICY
    torch.set_num_threads(os.cpu_count() or 1)
    oc = O.OracleConfig(**O.LLAVA_7B)
    ec = EngineConfig(**oc.to_dict())
    w = O.init_weights(oc, seed=21)
    lp, lr = O.init_lora(oc, 22), O.init_lora(oc, 23)
    eng = Engine(ec)
    _load(eng, O, oc, w, lp, lr)
    eng.set_trainable(POLICY)          # as in training: the reference adapter is the frozen one (eligible for merge_frozen)
    yield O, oc, w, lp, lr, eng
    eng.close()


def _full_depth_case(full7b, tag, b, R, data_seed, modes):
    """One batch through the oracle (fp32, the reference's stacked rows) and through the engine in each of `modes`
    = (share_prefix, merge_frozen); returns the relative DPO-loss errors. Per-token statistics are printed and appended
    to gpurun_out/full_parity_stats.txt."""
    import time
    from opa_dpo_b200 import POLICY, REF
    from opa_dpo_b200.dpo import DPOArgs, FusedDPO
    O, oc, w, lp, lr, eng = full7b
    images, q, qm, resp = O.synthetic_batch(oc, b, 128, R, NAMES2, seed=data_seed, left_pad=17, image_pos=35, ragged=False)
    t0 = time.time()
    with torch.no_grad():
        o_ref = O.policy_forward(oc, w, lr, images, q, qm, resp)
        o_pol = O.policy_forward(oc, w, lp, images, q, qm, resp)
        loss_ref, _, _ = O.pair_loss(O.LossConfig(beta=0.1), o_pol[NAMES2[0] + "_logprobs"], o_pol[NAMES2[1] + "_logprobs"],
                                     o_ref[NAMES2[0] + "_logprobs"], o_ref[NAMES2[1] + "_logprobs"])
    t_cpu = time.time() - t0
    dev = eng.device
    args = [images.to(dev), q.to(dev), qm.to(dev), [resp[n].to(dev) for n in NAMES2]]
    out_dir = os.path.join(os.path.dirname(__file__), "..", "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    rels = []
    for share, merge in modes:
        eng.set_option("share_prefix", share)
        eng.set_option("merge_frozen", merge)
        ref_lp, _ = eng.forward_logprobs(REF, *args)
        pol_lp, _ = eng.forward_logprobs(POLICY, *args)
        errs = torch.cat([(pol_lp[i].cpu() - o_pol[n + "_logprobs"]).abs().flatten() for i, n in enumerate(NAMES2)] +
                         [(ref_lp[i].cpu() - o_ref[n + "_logprobs"]).abs().flatten() for i, n in enumerate(NAMES2)])
        total, _ = FusedDPO(DPOArgs(beta=0.1)).pair(pol_lp[0], pol_lp[1], ref_lp[0], ref_lp[1])
        rel = abs(total.item() - loss_ref.item()) / abs(loss_ref.item())
        signed = torch.cat([(pol_lp[i].cpu() - o_pol[n + "_logprobs"]).flatten() for i, n in enumerate(NAMES2)])
        # what the loss sees is the per-token error of delta = (pol_c - ref_c) - (pol_r - ref_r): its mean over the tokens
        # (common-mode errors of a whole sequence do not average out over its tokens) and its spread
        d_eng = (pol_lp[0] - ref_lp[0]) - (pol_lp[1] - ref_lp[1])
        d_orc = (o_pol[NAMES2[0] + "_logprobs"] - o_ref[NAMES2[0] + "_logprobs"]) - (o_pol[NAMES2[1] + "_logprobs"] - o_ref[NAMES2[1] + "_logprobs"])
        e_d = (d_eng.cpu() - d_orc).flatten()
        slope = float((d_eng.cpu().flatten() @ d_orc.flatten()) / (d_orc.flatten() @ d_orc.flatten()))
        seq_means = [float((pol_lp[i].cpu() - o_pol[n + "_logprobs"]).mean()) for i, n in enumerate(NAMES2)] + \
                    [float((ref_lp[i].cpu() - o_ref[n + "_logprobs"]).mean()) for i, n in enumerate(NAMES2)]
        diag = (f"; delta err mean {e_d.mean().item():+.5f} std {e_d.std().item():.4f} (oracle delta std {d_orc.std().item():.3f}, "
                f"slope {slope:.4f}); per-set mean err pol_c/pol_r/ref_c/ref_r " + "/".join(f"{x:+.4f}" for x in seq_means))
        msg = (f"FULL-7B parity [{tag}] b={b} R={R} data_seed={data_seed} share_prefix={share} merge_frozen={merge}: loss "
               f"{total.item():.6f} oracle {loss_ref.item():.6f} rel {rel:.3e}; logprob |err| max {errs.max().item():.4f} mean "
               f"{errs.mean().item():.4f} p99 {errs.quantile(0.99).item():.4f} signed-mean {signed.mean().item():+.5f}; "
               f"oracle {t_cpu:.0f} s on {torch.get_num_threads()} threads" + diag)
        print(msg)
        with open(os.path.join(out_dir, "full_parity_stats.txt"), "a") as f:
            f.write(msg + "\n")
        rels.append(rel)
    eng.set_option("share_prefix", 1)
    eng.set_option("merge_frozen", 0)
    return rels


def test_full_depth_7b_config0_loss_vs_oracle(full7b):
    """BASELINE configs[0]: full-depth LLaVA-1.5-7B, 2 pairs x 128-token responses, fp32 CPU reference (oracle) vs the bf16
    engine in the reference's stacked layout and in the packed one. Bar: DPO loss within 1e-3 relative (north_star)."""
    rels = _full_depth_case(full7b, "configs[0]", 2, 128, 7, [(0, 0), (1, 0)])
    assert max(rels) <= 1e-3, rels


@pytest.mark.parametrize("data_seed", [7, 8, 9])
def test_full_depth_7b_headline_path_loss_vs_oracle(full7b, data_seed):
    """The path bench.py times (configs[1]): depth 32, R = 512, chosen + rejected packed behind the shared prefix
    (share_prefix=1), the frozen reference adapter on merged weights (merge_frozen=1), RoPE / SwiGLU in the GEMM epilogues —
    against the fp32 oracle, which computes the reference's stacked rows with unmerged adapters (rl_models.py:120-132,
    dpo_trainer.py:631). One pair per seed (1024 response tokens); three draws so the margin is not a single sample."""
    (rel,) = _full_depth_case(full7b, "headline", 1, 512, data_seed, [(1, 1)])
    assert rel <= 1e-3, rel


@pytest.fixture(scope="module")
def eng7b():
    from opa_dpo_b200 import Engine, EngineConfig, POLICY
    eng = Engine(EngineConfig())
    eng.random_init(seed=77)
    eng.set_trainable(POLICY)
    yield eng
    eng.close()


def _batch7b(b, R, seed=5, ragged=True):
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(b, 3, 336, 336, generator=g)
    q = torch.randint(3, 32000, (b, 128), generator=g)
    qm = torch.ones(b, 128, dtype=torch.uint8)
    for i in range(b):
        lp = 5 * i
        q[i, :lp] = 0
        qm[i, :lp] = 0
        q[i, lp + 35] = -200
    resp = []
    for k in range(3):
        t = torch.randint(3, 32000, (b, R), generator=g)
        for i in range(b):
            n = int(torch.randint(R // 4, R + 1, (1,), generator=g)) if ragged else R
            t[i, n - 1] = 2
            t[i, n:] = 0
        resp.append(t)
    return images, q, qm, resp


def test_7b_forward_properties(eng7b):
    from opa_dpo_b200 import POLICY, REF
    eng = eng7b
    dev = eng.device
    images, q, qm, resp = _batch7b(3, 512)
    I, Q, M = images.to(dev), q.to(dev), qm.to(dev)
    Rr = [r.to(dev) for r in resp]
    eng.set_option("share_prefix", 0)
    lp_a, ent_a = eng.forward_logprobs(POLICY, I, Q, M, Rr[:2])
    lp_b, _ = eng.forward_logprobs(POLICY, I, Q, M, Rr[:2])
    assert torch.equal(lp_a, lp_b), "forward must be run-to-run deterministic (no atomics on the forward path)"
    assert torch.isfinite(lp_a).all() and torch.isfinite(ent_a).all()
    for k in range(2):
        pad = Rr[k] == 0
        assert (lp_a[k][pad] == 0).all() and (ent_a[k][pad] == 0).all()
        assert (lp_a[k][~pad] < 0).all() and (ent_a[k][~pad] > 0).all()
    # batch invariance: a sample alone, or in another batch order / with another partner response, gives the SAME bits
    lp_1, _ = eng.forward_logprobs(POLICY, I[1:2], Q[1:2], M[1:2], [Rr[0][1:2]])
    assert torch.equal(lp_1[0, 0], lp_a[0, 1]), "row results must not depend on the batch they are computed in"
    perm = torch.tensor([2, 0, 1], device=dev)
    lp_p, _ = eng.forward_logprobs(POLICY, I[perm], Q[perm], M[perm], [Rr[0][perm], Rr[1][perm]])
    assert torch.equal(lp_p[:, 0], lp_a[:, 2])
    # K = 3 responses (repo default) and the frozen reference adapter
    lp_3, _ = eng.forward_logprobs(REF, I, Q, M, Rr)
    assert lp_3.shape == (3, 3, 512) and torch.isfinite(lp_3).all()
    # shared-prefix packing == stacked rows (same math, different tiling): bf16-level agreement, identical pad zeros
    eng.set_option("share_prefix", 1)
    lp_s, ent_s = eng.forward_logprobs(POLICY, I, Q, M, Rr[:2])
    d = (lp_s - lp_a).abs()
    # two bf16 tilings of the same math at depth 32: vs the fp32 oracle each is off by mean 0.034 / max 0.29
    # (profiles/r01_full_parity_stats.txt); between themselves they agree better than that
    assert d.max().item() <= 0.3 and d.mean().item() <= 0.02, f"shared-prefix vs stacked: max {d.max().item():.4g} mean {d.mean().item():.4g}"
    assert ((lp_s == 0) == (lp_a == 0)).all()
    lp_s2, _ = eng.forward_logprobs(POLICY, I, Q, M, Rr[:2])
    assert torch.equal(lp_s, lp_s2)


def test_7b_backward_layout_agreement(eng7b):
    """LoRA gradients of the same loss from the packed (shared prefix) and stacked layouts agree (cosine >= 0.995)."""
    from opa_dpo_b200 import POLICY
    eng = eng7b
    dev = eng.device
    images, q, qm, resp = _batch7b(2, 256, seed=9)
    I, Q, M = images.to(dev), q.to(dev), qm.to(dev)
    Rr = [r.to(dev) for r in resp[:2]]
    g = torch.Generator(device=dev).manual_seed(1)
    dlp = torch.randn((2, 2, 256), generator=g, device=dev) * 1e-3
    grads = []
    for share in (0, 1):
        eng.set_option("share_prefix", share)
        eng.zero_grad()
        eng.forward_logprobs(POLICY, I, Q, M, Rr, save_slot=0)
        eng.backward(0, dlp)
        torch.cuda.synchronize()
        grads.append(eng.buffers()["grad"].clone())
    eng.set_option("share_prefix", 1)
    eng.zero_grad()
    assert torch.isfinite(grads[0]).all() and torch.isfinite(grads[1]).all()
    cos = torch.nn.functional.cosine_similarity(grads[0], grads[1], dim=0).item()
    rel = ((grads[0] - grads[1]).norm() / grads[0].norm()).item()
    assert cos >= 0.995 and rel <= 0.1, f"cos {cos:.5f} rel {rel:.4f}"


def test_7b_rollout_properties(eng7b):
    """Greedy rollout at 7B: deterministic, batch-order invariant, pads only after EOS."""
    from opa_dpo_b200 import POLICY
    eng = eng7b
    images, q, qm, _ = _batch7b(4, 8, seed=3)
    out1 = eng.generate(POLICY, images, q, qm, 24)
    out2 = eng.generate(POLICY, images, q, qm, 24)
    assert torch.equal(out1, out2)
    assert torch.equal(out1[:, :128].cpu(), q)
    perm = torch.tensor([3, 1, 0, 2])
    out_p = eng.generate(POLICY, images[perm], q[perm], qm[perm], 24)
    gen, gen_p = out1[:, 128:].cpu(), out_p[:, 128:].cpu()
    agree = (gen[perm] == gen_p).float().mean().item()
    assert agree >= 0.9, f"rollout should not depend on batch order (agreement {agree:.2f})"
    for row in gen:
        eos = (row == 2).nonzero()
        if eos.numel():
            assert (row[int(eos[0]) + 1:] == 0).all()


def test_7b_merged_reference_adapter_and_entropy_head(eng7b):
    """Full-size properties of the two later additions: (1) the frozen reference adapter on merged weights
    (merge_frozen) agrees with its unmerged pass at bf16 level — two bf16 roundings of the same function at depth 32 —
    is deterministic, and the merged copy is reused across calls; (2) the entropy head's backward (opadpo_backward_ent):
    the gradient is linear in d(entropy) (2x the incoming gradient -> 2x the LoRA gradient, to fp32-atomics noise) and a
    zero d(entropy) next to d(logprob) reproduces the log-prob-only backward."""
    from opa_dpo_b200 import POLICY, REF
    eng = eng7b
    dev = eng.device
    images, q, qm, resp = _batch7b(2, 256, seed=11)
    I, Q, M = images.to(dev), q.to(dev), qm.to(dev)
    Rr = [r.to(dev) for r in resp[:2]]
    eng.set_option("share_prefix", 1)
    eng.set_option("merge_frozen", 0)
    lp_u, ent_u = eng.forward_logprobs(REF, I, Q, M, Rr)
    eng.set_option("merge_frozen", 1)
    lp_m, ent_m = eng.forward_logprobs(REF, I, Q, M, Rr)
    lp_m2, _ = eng.forward_logprobs(REF, I, Q, M, Rr)
    assert torch.equal(lp_m, lp_m2)
    pol_a, _ = eng.forward_logprobs(POLICY, I, Q, M, Rr)       # the trainable adapter never takes the merged path
    eng.set_option("merge_frozen", 0)
    pol_b, _ = eng.forward_logprobs(POLICY, I, Q, M, Rr)
    assert torch.equal(pol_a, pol_b)
    # Merged vs unmerged are two bf16 realisations of the same function: W' = bf16(W + s B A) against fp32 accumulation
    # of x W^T + bf16(s x A^T) B^T. With the synthetic init the LoRA delta is as large as W itself (B ~ N(0, 0.02^2)), so
    # the per-token difference sits at the level of the bf16-vs-fp32 error of either (mean 0.058, max 0.29 at depth 32,
    # profiles/r01_full_parity_stats.txt): bound it by that level. The bar that matters is the loss built on it: the DPO
    # loss with the merged reference pass within 1e-3 relative of the one with the unmerged pass (the north-star bar).
    from opa_dpo_b200.dpo import DPOArgs, FusedDPO
    d = (lp_m - lp_u).abs()
    fused = FusedDPO(DPOArgs(beta=0.1))
    loss_m, _ = fused.pair(pol_a[0], pol_a[1], lp_m[0], lp_m[1])
    loss_u, _ = fused.pair(pol_a[0], pol_a[1], lp_u[0], lp_u[1])
    rel = abs(loss_m.item() - loss_u.item()) / abs(loss_u.item())
    msg = (f"merged vs unmerged reference pass (7B, 2 x 2 x 256 tokens): per-token |dlogprob| max {d.max().item():.4g} "
           f"mean {d.mean().item():.4g}; entropies mean |d| {(ent_m - ent_u).abs().mean().item():.4g}; DPO loss "
           f"{loss_m.item():.6f} vs {loss_u.item():.6f} (rel {rel:.3g})")
    print(msg)
    import os
    if os.path.isdir("gpurun_out"):
        with open("gpurun_out/merged_ref_stats.txt", "w") as f:
            f.write(msg + "\n")
    assert rel <= 1e-3, msg
    assert d.max().item() <= 0.3 and d.mean().item() <= 0.06, msg   # measured: max 0.214, mean 0.0275
    assert ((lp_m == 0) == (lp_u == 0)).all()

    g = torch.Generator(device=dev).manual_seed(3)
    valid = torch.stack(Rr).ne(0).float()
    dent = torch.randn((2, 2, 256), generator=g, device=dev) * 1e-3 * valid
    dlp = torch.randn((2, 2, 256), generator=g, device=dev) * 1e-3 * valid
    grads = {}
    for name, (a, b) in {"ent": (None, dent), "ent2": (None, 2 * dent), "lp": (dlp, None), "lp+0": (dlp, torch.zeros_like(dent))}.items():
        eng.zero_grad()
        eng.forward_logprobs(POLICY, I, Q, M, Rr, save_slot=0)
        eng.backward(0, a, b)
        torch.cuda.synchronize()
        grads[name] = eng.buffers()["grad"].clone()
    eng.zero_grad()
    for v in grads.values():
        assert torch.isfinite(v).all()
    assert grads["ent"].norm().item() > 0
    rel = ((grads["ent2"] - 2 * grads["ent"]).norm() / grads["ent2"].norm()).item()
    assert rel <= 1e-3, f"entropy backward must be linear in d(entropy): rel {rel:.3g}"
    rel0 = ((grads["lp+0"] - grads["lp"]).norm() / grads["lp"].norm()).item()
    assert rel0 <= 1e-3, f"zero d(entropy) must not change the log-prob backward: rel {rel0:.3g}"
