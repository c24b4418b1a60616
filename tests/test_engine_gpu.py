"""Engine parity on the GPU: the CUDA hot path (through the C ABI) vs the CPU fp32 oracle on identical seeded
bf16-representable weights and synthetic image/text batches, at dimensions the oracle finishes in seconds.

Tolerances (bf16 storage, fp32 accumulate): per-token log-prob |err| <= 0.05 absolute with mean |err| <= 0.01; DPO loss
within 1e-3 relative (the north-star bar); LoRA gradients cosine >= 0.99 and relative L2 <= 0.08 per module group;
greedy rollout ids exact (on a peaked lm_head so the top-1 margin exceeds bf16 noise, margin asserted)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

NAMES2 = ["chosen_response", "rejected_response"]
NAMES3 = ["standard_response", "original_generate_response", "AI_pseudo_response"]


def _cfgs(n_layers=2, v_layers=2, **kw):
    from oracle import llava_oracle as O
    from opa_dpo_b200 import EngineConfig
    oc = O.OracleConfig(n_layers=n_layers, v_layers=v_layers, **kw)
    ec = EngineConfig(**{k: v for k, v in oc.to_dict().items()})
    return oc, ec


def _build(oc, ec, peaked=1.0, trainable=True, share_prefix=1):
    from oracle import llava_oracle as O
    from opa_dpo_b200 import Engine, POLICY, REF
    w = O.init_weights(oc, peaked=peaked)
    lp, lr = O.init_lora(oc, 11), O.init_lora(oc, 12)
    eng = Engine(ec)
    eng.set_option("share_prefix", share_prefix)
    sd = dict(w)
    sd[O.VT + "embeddings.patch_embedding.weight"] = w[O.VT + "embeddings.patch_embedding.weight"].reshape(oc.v_d, -1)
    eng.load_state_dict(sd)
    eng.load_adapter(POLICY, lp)
    eng.load_adapter(REF, lr)
    if trainable:
        eng.set_trainable(POLICY)
    return eng, w, lp, lr


def _cmp_lp(got, ref, what, atol=0.05, mean_tol=0.01):
    err = (got.cpu() - ref).abs()
    assert torch.isfinite(got).all(), f"{what}: non-finite"
    assert err.max().item() <= atol and err.mean().item() <= mean_tol, \
        f"{what}: max |err| {err.max().item():.4g} mean {err.mean().item():.4g} (ref mean |x| {ref.abs().mean().item():.3g})"


@pytest.mark.parametrize("n_layers,v_layers,share", [(0, 0, 0), (0, 2, 0), (1, 0, 0), (2, 2, 0), (0, 0, 1), (1, 0, 1), (2, 2, 1)])
def test_forward_logprobs_vs_oracle(n_layers, v_layers, share):
    from oracle import llava_oracle as O
    from opa_dpo_b200 import POLICY, REF, BASE
    oc, ec = _cfgs(n_layers, v_layers)
    eng, w, lp, lr = _build(oc, ec, trainable=False, share_prefix=share)
    images, q, qm, resp = O.synthetic_batch(oc, 3, 16, 24, NAMES2, left_pad=3)
    dev = eng.device
    for adapter, lora in ((POLICY, lp), (REF, lr), (BASE, None)):
        ref = O.policy_forward(oc, w, lora, images, q, qm, resp)
        got_lp, got_ent = eng.forward_logprobs(adapter, images.to(dev), q.to(dev), qm.to(dev), [resp[n].to(dev) for n in NAMES2])
        torch.cuda.synchronize()
        for i, n in enumerate(NAMES2):
            _cmp_lp(got_lp[i], ref[n + "_logprobs"], f"L{n_layers}/V{v_layers} adapter {adapter} {n} logprobs")
            _cmp_lp(got_ent[i], ref[n + "_entropies"], f"L{n_layers}/V{v_layers} adapter {adapter} {n} entropies")
            pad = resp[n] == oc.pad_id
            assert (got_lp[i].cpu()[pad] == 0).all() and (got_ent[i].cpu()[pad] == 0).all(), "pads must be exactly 0"
    eng.close()


def test_forward_temperature_and_copo_attention_mask():
    from oracle import llava_oracle as O
    from opa_dpo_b200 import POLICY
    oc, ec = _cfgs()
    eng, w, lp, lr = _build(oc, ec, trainable=False)
    images, q, qm, resp = O.synthetic_batch(oc, 2, 16, 24, NAMES2, left_pad=2)
    dev = eng.device
    g = torch.Generator().manual_seed(5)
    img_mask = torch.rand(2, oc.n_patches, generator=g) > 0.3
    ref = O.policy_forward(oc, w, lp, images, q, torch.cat([img_mask, qm], 1), resp, temperature=0.7)
    got, _ = eng.forward_logprobs(POLICY, images.to(dev), q.to(dev), qm.to(dev), [resp[n].to(dev) for n in NAMES2],
                                  temperature=0.7, img_mask=img_mask.to(dev))
    for i, n in enumerate(NAMES2):
        _cmp_lp(got[i], ref[n + "_logprobs"], f"copo-attention {n}")
    eng.close()


def _oracle_grads(oc, w, lp, lr, images, q, qm, resp, lc, names):
    from oracle import llava_oracle as O
    lpg = {m: (A.clone().requires_grad_(True), B.clone().requires_grad_(True)) for m, (A, B) in lp.items()}
    with torch.no_grad():
        ref = O.policy_forward(oc, w, lr, images, q, qm, resp)
    pol = O.policy_forward(oc, w, lpg, images, q, qm, resp)
    loss, _, _ = O.pair_loss(lc, pol[names[0] + "_logprobs"], pol[names[1] + "_logprobs"], ref[names[0] + "_logprobs"],
                             ref[names[1] + "_logprobs"])
    loss.backward()
    return loss.detach(), lpg, ref, pol


def test_fused_swiglu_backward_and_rope_backward_equal_the_standalone_kernels():
    """opadpo_backward with the SwiGLU backward inside the down projection's dgrad epilogue and the inverse rotary embedding
    inside the attention backward's epilogues (the second is the default, the first opt-in) against the same backward with the standalone kernels: the LoRA
    gradient bucket agrees to bf16-rounding noise of the intermediate dq / dk (the fused rotation works on the fp32
    accumulator: one rounding instead of two; the option also moves the forward's rotation into its GEMM epilogue) — relative
    L2 <= 1e-2, cosine >= 0.9999."""
    from oracle import llava_oracle as O
    from opa_dpo_b200 import POLICY
    oc, ec = _cfgs(n_layers=3)
    eng, w, lp, lr = _build(oc, ec, share_prefix=1)
    images, q, qm, resp = O.synthetic_batch(oc, 3, 16, 24, NAMES2, left_pad=3)
    dev = eng.device
    args = [images.to(dev), q.to(dev), qm.to(dev), [resp[n].to(dev) for n in NAMES2]]
    eng.set_option("wgrad_stream", 0)
    grads = {}
    for fused in (1, 0):
        eng.set_option("fuse_swiglu_bwd", fused)
        eng.set_option("fuse_rope", fused)
        pol_lp, _ = eng.forward_logprobs(POLICY, *args, save_slot=0)
        d = torch.randn(pol_lp.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(77)) * 0.1
        eng.zero_grad()
        eng.backward(0, d)
        torch.cuda.synchronize()
        grads[fused] = eng.buffers()["grad"].clone()
    a, b = grads[1].double(), grads[0].double()
    assert torch.isfinite(grads[1]).all()
    cos = float((a @ b) / (a.norm() * b.norm()))
    rel = float((a - b).norm() / b.norm())
    assert cos >= 0.9999 and rel <= 1e-2, (cos, rel)
    # the SwiGLU fusion alone is bit-compatible with the standalone kernel (same roundings): fp32-atomics noise only
    eng.set_option("fuse_rope", 1)
    for fused in (1, 0):
        eng.set_option("fuse_swiglu_bwd", fused)
        pol_lp, _ = eng.forward_logprobs(POLICY, *args, save_slot=0)
        eng.zero_grad()
        eng.backward(0, d)
        torch.cuda.synchronize()
        grads[fused] = eng.buffers()["grad"].clone()
    rel = float((grads[1] - grads[0]).norm() / grads[0].norm())
    assert rel <= 2e-5, rel
    eng.close()


def test_adopted_gradient_buffer_receives_the_gradients():
    """opadpo_engine_adopt_grad_buffer (what the data-parallel step uses to accumulate straight into a symmetric-memory
    bucket): the current gradients move into the caller's buffer, the engine's views point at it, the next backward and
    zero_grad act on it, the optimizer reads it; too small / unaligned / non-fp32 buffers are rejected."""
    from oracle import llava_oracle as O
    from opa_dpo_b200 import POLICY
    from opa_dpo_b200._lib import OpadpoError
    oc, ec = _cfgs()
    eng, w, lp, lr = _build(oc, ec, share_prefix=1)
    images, q, qm, resp = O.synthetic_batch(oc, 2, 16, 24, NAMES2, left_pad=2)
    dev = eng.device
    args = [images.to(dev), q.to(dev), qm.to(dev), [resp[n].to(dev) for n in NAMES2]]

    def backward():
        pol_lp, _ = eng.forward_logprobs(POLICY, *args, save_slot=0)
        d = torch.randn(pol_lp.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(3)) * 0.1
        eng.backward(0, d)
        torch.cuda.synchronize()

    eng.zero_grad()
    backward()
    v = eng.buffers()
    n = v["n_train"]
    g0 = v["grad"].clone()
    assert g0.abs().sum().item() > 0
    with pytest.raises(OpadpoError):
        eng.adopt_grad_buffer(torch.zeros(n - 8, device=dev))                       # too small
    with pytest.raises(OpadpoError):
        eng.adopt_grad_buffer(torch.zeros(n, device=dev, dtype=torch.bfloat16))     # not fp32
    buf = torch.full((n + 64,), 7.0, device=dev)
    eng.adopt_grad_buffer(buf)
    v2 = eng.buffers()
    assert v2["grad"].data_ptr() == buf.data_ptr() and v2["grad"].numel() == n
    assert torch.equal(buf[:n], g0) and bool((buf[n:] == 7.0).all())                # gradients carried over, the tail untouched
    eng.zero_grad()
    torch.cuda.synchronize()
    assert buf[:n].abs().sum().item() == 0
    backward()
    rel = ((buf[:n] - g0).norm() / g0.norm()).item()
    assert rel <= 1e-5, rel                                                         # the same gradient, now in the adopted buffer
    master0 = v2["master"][:n].clone()
    norm = eng.optimizer_step(lr=1e-3, max_grad_norm=1.0)
    torch.cuda.synchronize()
    assert abs(norm.item() - g0.norm().item()) <= 1e-4 * g0.norm().item()
    assert not torch.equal(v2["master"][:n], master0)
    eng.close()


def test_side_stream_weight_gradients_equal_main_stream_weight_gradients():
    """opadpo_backward with the LoRA weight-gradient GEMMs on the engine's side stream (default) against the same backward
    with everything on one stream: the same gradient bucket up to the order of the fp32 split-K atomics, three backward
    calls in a row (the side stream's u / dy buffers are recycled layer by layer and call by call), and the layer callback
    still sees every slice exactly once, in top-down order, projector last."""
    from oracle import llava_oracle as O
    from opa_dpo_b200 import POLICY
    oc, ec = _cfgs(n_layers=5)      # deep enough for every buffer to be recycled several times
    eng, w, lp, lr = _build(oc, ec, share_prefix=1)
    images, q, qm, resp = O.synthetic_batch(oc, 3, 16, 24, NAMES2, left_pad=3)
    dev = eng.device
    args = [images.to(dev), q.to(dev), qm.to(dev), [resp[n].to(dev) for n in NAMES2]]
    grads = {}
    seen = []
    eng.set_backward_callback(lambda layer: seen.append(layer))
    for side in (1, 0):
        eng.set_option("wgrad_stream", side)
        acc = []
        for rep in range(3):
            pol_lp, _ = eng.forward_logprobs(POLICY, *args, save_slot=0)
            d = torch.randn(pol_lp.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(40 + rep)) * 0.1
            eng.zero_grad()
            seen.clear()
            eng.backward(0, d)
            torch.cuda.synchronize()
            assert seen == list(range(ec.n_layers - 1, -1, -1)) + [ec.n_layers], seen
            acc.append(eng.buffers()["grad"].clone())
        grads[side] = acc
    eng.set_backward_callback(None)
    for a, b in zip(grads[1], grads[0]):
        assert torch.isfinite(a).all()
        rel = ((a - b).norm() / b.norm()).item()
        assert rel <= 1e-5, rel
    eng.close()


@pytest.mark.parametrize("share", [0, 1])
def test_dpo_step_loss_and_lora_grads_vs_oracle(share):
    """Four-forward DPO step (policy & frozen reference x chosen & rejected): loss within 1e-3 relative, LoRA grads.
    share=1: image+query prefix computed once per sample (packed rows); share=0: the reference's stacked rows."""
    from oracle import llava_oracle as O
    from opa_dpo_b200 import POLICY, REF
    from opa_dpo_b200.dpo import DPOArgs, FusedDPO
    oc, ec = _cfgs()
    eng, w, lp, lr = _build(oc, ec, share_prefix=share)
    images, q, qm, resp = O.synthetic_batch(oc, 3, 16, 24, NAMES2, left_pad=3)
    lc = O.LossConfig(beta=0.1)
    loss_ref, lpg, ref_o, pol_o = _oracle_grads(oc, w, lp, lr, images, q, qm, resp, lc, NAMES2)
    dev = eng.device
    args = [images.to(dev), q.to(dev), qm.to(dev), [resp[n].to(dev) for n in NAMES2]]
    ref_lp, _ = eng.forward_logprobs(REF, *args)
    pol_lp, _ = eng.forward_logprobs(POLICY, *args, save_slot=0)
    fused = FusedDPO(DPOArgs(beta=0.1))
    pl = pol_lp.clone().requires_grad_(True)
    total, out = fused.pair(pl[0], pl[1], ref_lp[0], ref_lp[1])
    total.backward()
    rel = abs(total.item() - loss_ref.item()) / abs(loss_ref.item())
    assert rel <= 1e-3, f"DPO loss {total.item():.6f} vs oracle {loss_ref.item():.6f} (rel {rel:.3g})"
    eng.zero_grad()
    eng.backward(0, pl.grad)
    torch.cuda.synchronize()
    views = eng.buffers()
    grad = views["grad"]
    assert torch.isfinite(grad).all()
    bad = []
    for m, (A, B) in lpg.items():
        if "vision_model" in m:
            assert A.grad is None or A.grad.abs().max() == 0  # tower runs under no_grad
            continue
        oa, ob = eng.lora_offsets(m)
        for nm, t, off in (("A", A, oa), ("B", B, ob)):
            gg = grad[off:off + t.numel()].view_as(t).cpu()
            gr = t.grad
            cos = torch.nn.functional.cosine_similarity(gg.flatten(), gr.flatten(), dim=0).item()
            rl2 = ((gg - gr).norm() / (gr.norm() + 1e-30)).item()
            if not (cos >= 0.99 and rl2 <= 0.08):
                bad.append((m, nm, round(cos, 4), round(rl2, 4), float(gr.norm())))
    assert not bad, f"LoRA grad mismatches (module, which, cos, rel_l2, ref_norm): {bad[:12]} ... {len(bad)} total"

    # optimiser: clip + AdamW first step on the same gradients (fp32 masters)
    n_train = views["n_train"]
    master0 = views["master"][:n_train].clone()
    g0 = grad.clone()
    norm = eng.optimizer_step(lr=1e-3, max_grad_norm=1.0)
    torch.cuda.synchronize()
    (p1,), nref = O.adamw_reference([master0.cpu()], [g0.cpu()], lr=1e-3, max_grad_norm=1.0)
    assert abs(norm.item() - nref.item()) / nref.item() < 1e-4
    assert (views["master"][:n_train].cpu() - p1).abs().max().item() < 1e-6
    eng.close()


def test_policy_module_compute_policy_loss_repo_defaults():
    """The reference's own call pattern: AutoregressivePolicy.forward over 3 responses + CoPO (attention) second
    forward, compute_policy_loss with AncPO (mDPO anchor) and score weights; loss and stats vs the oracle."""
    from types import SimpleNamespace
    from oracle import llava_oracle as O
    from opa_dpo_b200.policy import AutoregressivePolicy, B200LlavaModel
    from opa_dpo_b200.dpo import DPOArgs, FusedDPO
    oc, ec = _cfgs()
    model = B200LlavaModel(ec)
    w = O.init_weights(oc)
    lp, lr = O.init_lora(oc, 11), O.init_lora(oc, 12)
    sd = dict(w)
    sd[O.VT + "embeddings.patch_embedding.weight"] = w[O.VT + "embeddings.patch_embedding.weight"].reshape(oc.v_d, -1)
    model.engine.load_state_dict(sd)
    model.load_adapter_state("lora_policy", {f"base_model.model.{m}.lora_{x}.weight": t for m, ab in lp.items() for x, t in zip("AB", ab)})
    model.load_adapter_state("lora_ref_policy", {f"base_model.model.{m}.lora_{x}.weight": t for m, ab in lr.items() for x, t in zip("AB", ab)})
    model.make_trainable("lora_policy")
    R = 24
    args = DPOArgs(beta=0.1, CoPO=True, CoPO_coef=0.2, AncPO=True, mDPO_anchor=True, detailed_report=True, response_score=True,
                   response_image_relation=True, response_len=R)
    tok = SimpleNamespace(pad_token_id=0, eos_token_id=2)
    policy = AutoregressivePolicy(args, model, tok, adapter_name="lora_policy")
    ref_policy = AutoregressivePolicy(args, model, tok, adapter_name="lora_ref_policy")
    images, q, qm, resp = O.synthetic_batch(oc, 2, 16, R, NAMES3, left_pad=2)
    g = torch.Generator().manual_seed(9)
    img_mask = torch.rand(2, oc.n_patches, generator=g) > 0.3
    masked_qm = torch.cat([img_mask, qm], 1)
    scores = {}
    for n in ("original_generate_response", "AI_pseudo_response"):
        s = torch.tensor([1.0, 1.5, 2.0, 2.5])[torch.randint(0, 4, (2, R), generator=g)] * (resp[n] != 0)
        scores[n + "_scores"] = s
        scores[n + "_image_relations"] = torch.tensor([1.0, 3.0])[torch.randint(0, 2, (2, R), generator=g)] * (resp[n] != 0)
    # oracle
    lc = O.LossConfig(beta=0.1, CoPO=True, CoPO_coef=0.2, AncPO=True, mDPO_anchor=True, detailed_report=True,
                      response_score=True, response_image_relation=True)
    mresp = {"mask_standard_response": resp["standard_response"], "mask_AI_pseudo_response": resp["AI_pseudo_response"]}
    with torch.no_grad():
        o_ref = {"ref_base_" + k: v for k, v in O.policy_forward(oc, w, lr, images, q, qm, resp).items()}
        o_refm = O.policy_forward(oc, w, lr, images, q, masked_qm, mresp)
        o_ref.update({"ref_" + k: v for k, v in o_refm.items()})
        o_pol = O.policy_forward(oc, w, lp, images, q, qm, resp)
        o_polm = O.policy_forward(oc, w, lp, images, q, masked_qm, mresp)
        loss_ref, stats_ref = O.compute_policy_loss(lc, o_pol, o_ref, o_polm, scores)
    # engine, through the reference-shaped modules
    dev = model.engine.device
    with torch.no_grad():
        rollouts = {"ref_base_" + k: v for k, v in ref_policy(images, q, qm, **resp).items()}
        rollouts.update({"ref_" + k: v for k, v in ref_policy(images, q, masked_qm, **mresp).items()})
    rollouts.update(scores)
    policy.train()
    out = policy(images, q, qm, temperature=1.0, mode="policy", **resp)
    out_m = policy(images, q, masked_qm, temperature=1.0, mode="policy", **mresp)
    loss, stats = FusedDPO(args).compute_policy_loss(out, rollouts, out_m)
    rel = abs(loss.item() - loss_ref.item()) / abs(loss_ref.item())
    assert rel <= 1e-3, f"compute_policy_loss {loss.item():.6f} vs oracle {loss_ref.item():.6f} (rel {rel:.3g})"
    for k, v in stats_ref.items():
        assert abs(stats[k].item() - v.item()) <= 2e-3 + 2e-2 * abs(v.item()), (k, stats[k].item(), v.item())
    model.engine.zero_grad()
    loss.backward()
    torch.cuda.synchronize()
    gnorm = model.flat_param.grad.norm().item()
    assert math.isfinite(gnorm) and gnorm > 0
    model.engine.close()


@pytest.mark.parametrize("merge", [1, 0])
def test_greedy_rollout_ids_exact(merge):
    """merge=1: rollout on W + (alpha/r) B A merged once per call (default); merge=0: adapters applied unmerged."""
    from oracle import llava_oracle as O
    from opa_dpo_b200 import POLICY, BASE
    oc, ec = _cfgs()
    eng, w, lp, lr = _build(oc, ec, peaked=40.0, trainable=False)
    eng.set_option("merge_rollout", merge)
    images, q, qm, _ = O.synthetic_batch(oc, 3, 16, 8, NAMES2, left_pad=3)
    ids_ref, margin = O.greedy_generate(oc, w, lp, images, q, qm, 12)
    got = eng.generate(POLICY, images, q, qm, 12).cpu()
    assert margin > 0.05, f"oracle top-1 margin {margin} too small for a bf16 comparison"
    assert torch.equal(got, ids_ref), f"rollout ids differ (margin {margin:.3g}):\n{got[:, 16:]}\n{ids_ref[:, 16:]}"
    trunc = O.truncate_after_eos_with_padding(got[:, 16:], oc.eos_id, oc.pad_id, [1577, 29973])
    assert trunc.shape == (3, 12)
    # the bare base model (adapter -1) is a different policy: its ids follow the oracle without LoRA
    ids_base, margin_b = O.greedy_generate(oc, w, None, images, q, qm, 6)
    got_b = eng.generate(BASE, images, q, qm, 6).cpu()
    if margin_b > 0.05:
        assert torch.equal(got_b, ids_base)
    eng.close()


@pytest.mark.parametrize("div,ls,anchor", [("reverse_kl", 0.0, 0), ("reverse_kl", 0.1, 1), ("js_divergence", 0.0, 2),
                                           ("alpha_divergence", 0.0, 0)])
def test_dpo_pair_kernel_vs_oracle(div, ls, anchor):
    from oracle import llava_oracle as O
    from opa_dpo_b200.dpo import DPOArgs, FusedDPO
    g = torch.Generator().manual_seed(3)
    b, R = 4, 37
    mk = lambda: -torch.rand(b, R, generator=g) * 6
    mask_c = torch.rand(b, R, generator=g) > 0.3
    mask_r = torch.rand(b, R, generator=g) > 0.2
    pc, rc = mk() * mask_c, mk() * mask_c
    pr, rr = mk() * mask_r, mk() * mask_r
    sc = (torch.randint(2, 6, (b, R), generator=g).float() / 2) * mask_c
    sr = (torch.randint(2, 6, (b, R), generator=g).float() / 2) * mask_r
    lc = O.LossConfig(beta=0.1, label_smoothing=ls, f_divergence_type=div)
    pcg, prg = pc.clone().requires_grad_(True), pr.clone().requires_grad_(True)
    l, cr, rj = O.dpo_loss(lc, pcg, prg, rc, rr, sc, sr)
    tot = l.mean() * 0.7
    if anchor == 1:
        tot = tot + (-torch.nn.functional.logsigmoid(-(cr - 0.05))).mean() * 0.9
    elif anchor == 2:
        tot = tot + ((cr - 0.05) ** 2).mean() * 0.9
    tot.backward()
    args = DPOArgs(beta=0.1, label_smoothing=ls, f_divergence_type=div, AncPO=anchor > 0, mDPO_anchor=anchor == 1,
                   Anchor_value=0.05, Anchor_coef=0.9)
    dev = "cuda"
    a, bb = pc.to(dev).requires_grad_(True), pr.to(dev).requires_grad_(True)
    total, out = FusedDPO(args).pair(a, bb, rc.to(dev), rr.to(dev), sc.to(dev), sr.to(dev), loss_coef=0.7,
                                     anchor_sign=-1.0 if anchor == 1 else (1.0 if anchor == 2 else 0.0))
    total.backward()
    assert abs(total.item() - tot.item()) <= 1e-5 * max(1, abs(tot.item())), (total.item(), tot.item())
    assert (a.grad.cpu() - pcg.grad).abs().max().item() <= 1e-6
    assert (bb.grad.cpu() - prg.grad).abs().max().item() <= 1e-6
    assert abs(out[2].item() - O.masked_mean(cr, rc != 0).item()) <= 1e-5
    assert abs(out[3].item() - O.masked_mean(rj, rr != 0).item()) <= 1e-5


def _oracle_step_logits(oc, w, lora, images, q, qm, gen):
    """Oracle logits that predicted every generated token (teacher forcing on the engine's own sample)."""
    from oracle import llava_oracle as O
    ids = torch.cat([q, gen], 1)
    mask = torch.cat([qm.bool(), torch.ones_like(gen, dtype=torch.bool)], 1)
    with torch.no_grad():
        logits = O.llava_logits(oc, w, lora, ids, mask, images)
    T = gen.shape[1]
    return logits[:, -T - 1:-1]


def test_sampling_rollout_topk_topp():
    """Sampling rollout (online_generate.sh:75-77: top_k/top_p/temperature): top_k=1 degenerates to greedy exactly;
    every sampled token lies in the oracle's top-k / nucleus support given the same prefix; runs are reproducible per
    seed and differ across seeds."""
    from oracle import llava_oracle as O
    from opa_dpo_b200 import POLICY
    oc, ec = _cfgs()
    eng, w, lp, lr = _build(oc, ec, peaked=12.0, trainable=False)
    images, q, qm, _ = O.synthetic_batch(oc, 4, 16, 8, NAMES2, left_pad=3)
    greedy = eng.generate(POLICY, images, q, qm, 10).cpu()
    k1 = eng.generate(POLICY, images, q, qm, 10, do_sample=True, top_k=1, top_p=1.0, temperature=0.7, seed=5).cpu()
    assert torch.equal(greedy, k1)
    K, P, T = 5, 0.9, 0.8
    a = eng.generate(POLICY, images, q, qm, 10, do_sample=True, top_k=K, top_p=P, temperature=T, seed=123).cpu()
    a2 = eng.generate(POLICY, images, q, qm, 10, do_sample=True, top_k=K, top_p=P, temperature=T, seed=123).cpu()
    b = eng.generate(POLICY, images, q, qm, 10, do_sample=True, top_k=K, top_p=P, temperature=T, seed=124).cpu()
    assert torch.equal(a, a2) and not torch.equal(a, b)
    gen = a[:, 16:]
    logits = _oracle_step_logits(oc, w, lp, images, q, qm, gen) / T
    probs = torch.softmax(logits, -1)
    finished = torch.zeros(4, dtype=torch.bool)
    for t in range(gen.shape[1]):
        for n in range(4):
            if finished[n]:
                assert gen[n, t] == oc.pad_id
                continue
            order = probs[n, t].argsort(descending=True)
            rank = int((order == gen[n, t]).nonzero()[0])
            assert rank < K + 1, f"row {n} step {t}: sampled token has oracle rank {rank} (top_k {K})"
            topk = probs[n, t][order[:K]]
            before = (topk / topk.sum())[:rank].sum().item()
            assert before < P + 0.05, f"row {n} step {t}: nucleus mass before the token {before:.3f} >= top_p"
            if gen[n, t] == oc.eos_id:
                finished[n] = True
    eng.close()


@pytest.mark.parametrize("top_k,top_p,temperature", [(0, 0.8, 0.9), (100, 1.0, 1.0), (100, 0.6, 1.3), (400, 0.9, 1.0), (600, 0.95, 1.0)])
def test_sampling_general_topk_topp_support(top_k, top_p, temperature):
    """Any top_k (radix select, no 64-entry cap) and top_p without top_k: every sampled token lies in the support HF's
    warpers leave (oracle.hf_warp, pinned to transformers' classes on the CPU) given the same prefix. bf16 logits move
    the boundary slightly, so the oracle support is taken with a little slack (top_k + 2, top_p + 0.03); a large sample
    of first tokens also has to COVER most of the support's mass (a sampler that kept too little would not)."""
    from oracle import llava_oracle as O
    from opa_dpo_b200 import POLICY
    oc, ec = _cfgs()
    eng, w, lp, lr = _build(oc, ec, peaked=12.0, trainable=False)
    images, q, qm, _ = O.synthetic_batch(oc, 4, 16, 8, NAMES2, left_pad=3)
    a = eng.generate(POLICY, images, q, qm, 8, do_sample=True, top_k=top_k, top_p=top_p, temperature=temperature, seed=77).cpu()
    a2 = eng.generate(POLICY, images, q, qm, 8, do_sample=True, top_k=top_k, top_p=top_p, temperature=temperature, seed=77).cpu()
    assert torch.equal(a, a2)
    gen = a[:, 16:]
    logits = _oracle_step_logits(oc, w, lp, images, q, qm, gen)
    slack = O.hf_warp(logits, temperature, min(top_k + 2, oc.vocab) if top_k else 0, min(top_p + 0.03, 1.0))
    finished = torch.zeros(4, dtype=torch.bool)
    for t in range(gen.shape[1]):
        for n in range(4):
            if finished[n]:
                assert gen[n, t] == oc.pad_id
                continue
            assert torch.isfinite(slack[n, t, gen[n, t]]), f"row {n} step {t}: token {int(gen[n, t])} outside the warped support"
            finished[n] |= bool(gen[n, t] == oc.eos_id)
    # coverage of the first token's support over 64 x 16 draws
    B = 64
    imgs, qq, qqm = images[:1].repeat(B, 1, 1, 1), q[:1].repeat(B, 1), qm[:1].repeat(B, 1)
    counts = torch.zeros(oc.vocab)
    for seed in range(16):
        out = eng.generate(POLICY, imgs, qq, qqm, 1, do_sample=True, top_k=top_k, top_p=top_p, temperature=temperature, seed=seed).cpu()
        counts += torch.bincount(out[:, 16], minlength=oc.vocab).float()
    with torch.no_grad():
        ids = torch.cat([q[:1], torch.full((1, 1), 5)], 1)
        l0 = O.llava_logits(oc, w, lp, ids, torch.cat([qm[:1].bool(), torch.ones(1, 1, dtype=torch.bool)], 1), images[:1])[0, -2]
    want = torch.softmax(O.hf_warp(l0[None], temperature, top_k, top_p)[0], -1)
    loose = O.hf_warp(l0[None], temperature, min(top_k + 2, oc.vocab) if top_k else 0, min(top_p + 0.03, 1.0))[0]
    assert counts[~torch.isfinite(loose)].sum() == 0
    emp = counts / counts.sum()
    tv = 0.5 * (emp - want).abs().sum().item()
    assert tv < 0.10 + 0.5 * (want * (1 - want) / counts.sum()).sqrt().sum().item(), f"total variation {tv:.3f}"
    eng.close()


@pytest.mark.parametrize("do_sample", [False, True])
def test_rollout_suppress_eos_at_generation(do_sample):
    """suppress_tokens=[eos] (online_generator.py:298-302, rl_models.py:173-177 under suppress_eos_at_generation): the EOS
    never appears, no row is padded, and the greedy ids equal the oracle's greedy loop with the same suppression — on a
    configuration whose unsuppressed greedy path does emit EOS early."""
    import dataclasses
    from types import SimpleNamespace
    from oracle import llava_oracle as O
    from opa_dpo_b200 import EngineConfig
    from opa_dpo_b200.policy import B200LlavaModel
    from opa_dpo_b200.generation import B200OnlineGenerator
    oc = dataclasses.replace(O.OracleConfig(), eos_id=298)
    ec = EngineConfig(**oc.to_dict())
    model = B200LlavaModel(ec)
    w = O.init_weights(oc, peaked=40.0)
    lp = O.init_lora(oc, 11)
    sd = dict(w)
    sd[O.VT + "embeddings.patch_embedding.weight"] = w[O.VT + "embeddings.patch_embedding.weight"].reshape(oc.v_d, -1)
    model.engine.load_state_dict(sd)
    model.load_adapter_state("lora_policy", {f"base_model.model.{m}.lora_{x}.weight": t for m, ab in lp.items() for x, t in zip("AB", ab)})
    model.set_adapter("lora_policy")
    images, q, qm, _ = O.synthetic_batch(oc, 2, 16, 24, NAMES3, left_pad=2)
    plain, _ = O.greedy_generate(oc, w, lp, images, q, qm, 12)
    assert (plain[:, 16:] == oc.eos_id).any()                      # without suppression the EOS is emitted
    dev = model.engine.device
    out = model.generate(inputs=q.to(dev), images=images.to(dev), attention_mask=qm.to(dev), do_sample=do_sample, max_new_tokens=12,
                         pad_token_id=oc.pad_id, suppress_tokens=[oc.eos_id], top_k=0, top_p=1.0, temperature=1.0).cpu()
    assert not (out[:, 16:] == oc.eos_id).any() and not (out[:, 16:] == oc.pad_id).any()
    if not do_sample:
        want, margin = O.greedy_generate(oc, w, lp, images, q, qm, 12, suppress_tokens=[oc.eos_id])
        assert margin > 0.05 and torch.equal(out, want)
    gen = B200OnlineGenerator(SimpleNamespace(response_len=12, top_p=1.0, top_k=0, temperature=1.0, suppress_eos_at_generation=True),
                              model, SimpleNamespace(pad_token_id=oc.pad_id, eos_token_id=oc.eos_id))
    r = gen.generate_responses(images.to(dev), q.to(dev), qm.to(dev), do_sample=do_sample).cpu()
    assert r.shape == (2, 12) and not (r == oc.eos_id).any()
    model.engine.close()


def test_rollout_decode_gemm_variants_agree():
    """Option decode_swap_ab: the weight-streaming decode GEMM + fused q|k|v reducer against the general split-K kernel
    (option off): the same greedy ids on a peaked head — both equal to the oracle's — with and without a merged adapter."""
    from oracle import llava_oracle as O
    from opa_dpo_b200 import POLICY, BASE
    oc, ec = _cfgs()
    eng, w, lp, lr = _build(oc, ec, peaked=40.0, trainable=False)
    images, q, qm, _ = O.synthetic_batch(oc, 3, 16, 8, NAMES2, left_pad=3)
    ids_ref, margin = O.greedy_generate(oc, w, lp, images, q, qm, 12)
    assert margin > 0.05
    for adapter, want in ((POLICY, ids_ref), (BASE, None)):
        outs = []
        for opt in (1, 0):
            eng.set_option("decode_swap_ab", opt)
            outs.append(eng.generate(adapter, images, q, qm, 12).cpu())
        assert torch.equal(outs[0], outs[1])
        if want is not None:
            assert torch.equal(outs[0], want)
    eng.set_option("decode_swap_ab", 1)
    eng.set_option("merge_rollout", 0)          # unmerged adapter: the LoRA linears keep the general kernel, the head streams
    assert torch.equal(eng.generate(POLICY, images, q, qm, 12).cpu(), ids_ref)
    eng.close()


def test_eval_generation_single_row_greedy():
    """The eval scripts' call shape (eval_llava_rlhf_coco/model_vqa.py:71-290: one image, one prompt, greedy,
    `model.generate(input_ids, images=..., do_sample=False, max_new_tokens=...)`): b = 1, no left padding, ids exact."""
    from oracle import llava_oracle as O
    from opa_dpo_b200 import POLICY
    oc, ec = _cfgs()
    eng, w, lp, lr = _build(oc, ec, peaked=60.0, trainable=False)
    images, q, qm, _ = O.synthetic_batch(oc, 1, 24, 8, NAMES2, seed=8)
    ids_ref, margin = O.greedy_generate(oc, w, lp, images, q, qm, 16)
    got = eng.generate(POLICY, images, q, qm, 16).cpu()
    assert margin > 0.05 and got.shape == (1, 40) and torch.equal(got, ids_ref)
    eng.close()


def test_sampling_full_multinomial_distribution():
    """top_k=0, top_p=1 (AutoregressivePolicy._respond, rl_models.py:166-188): empirical first-token distribution over
    2048 draws vs the oracle's softmax (total variation < 0.08)."""
    from oracle import llava_oracle as O
    from opa_dpo_b200 import POLICY
    oc, ec = _cfgs()
    eng, w, lp, lr = _build(oc, ec, peaked=60.0, trainable=False)
    images, q, qm, _ = O.synthetic_batch(oc, 1, 16, 8, NAMES2)
    B = 64
    imgs, qq, qqm = images.repeat(B, 1, 1, 1), q.repeat(B, 1), qm.repeat(B, 1)
    counts = torch.zeros(oc.vocab)
    for seed in range(32):
        out = eng.generate(POLICY, imgs, qq, qqm, 1, do_sample=True, top_k=0, top_p=1.0, temperature=1.0, seed=seed).cpu()
        counts += torch.bincount(out[:, 16], minlength=oc.vocab).float()
    with torch.no_grad():
        ids = torch.cat([q, torch.full((1, 1), 5)], 1)
        logits = O.llava_logits(oc, w, lp, ids, torch.cat([qm.bool(), torch.ones(1, 1, dtype=torch.bool)], 1), images)[0, -2]
    p = torch.softmax(logits, -1)
    tv = 0.5 * (counts / counts.sum() - p).abs().sum().item()
    assert tv < 0.08 + 0.5 * (p * (1 - p) / counts.sum()).sqrt().sum().item(), f"total variation {tv:.3f}; max p {p.max().item():.3f}"
    eng.close()


def test_checkpoint_roundtrip_peft_layout(tmp_path):
    """save_model / resume_training (dpo_trainer.py:837-931): PEFT-layout adapter files, optimizer.pt, scheduler.pt;
    reload restores LoRA tensors bit-identically (bf16 as the reference stores them), Adam moments, step and LR."""
    import json, os
    from types import SimpleNamespace
    from oracle import llava_oracle as O
    from opa_dpo_b200.policy import AutoregressivePolicy, B200LlavaModel
    from opa_dpo_b200.dpo import DPOArgs, DPOStepper
    from opa_dpo_b200.optim import B200AdamW, cosine_schedule, get_last_checkpoint, load_checkpoint, save_checkpoint
    oc, ec = _cfgs()

    def make():
        model = B200LlavaModel(ec)
        w = O.init_weights(oc)
        sd = dict(w)
        sd[O.VT + "embeddings.patch_embedding.weight"] = w[O.VT + "embeddings.patch_embedding.weight"].reshape(oc.v_d, -1)
        model.engine.load_state_dict(sd)
        for name, seed in (("lora_policy", 11), ("lora_ref_policy", 12)):
            lo = O.init_lora(oc, seed)
            model.load_adapter_state(name, {f"base_model.model.{m}.lora_{x}.weight": t for m, ab in lo.items() for x, t in zip("AB", ab)})
        model.make_trainable("lora_policy")
        return model

    model = make()
    args = DPOArgs(beta=0.1, response_len=24, learning_rate=1e-3)
    tok = SimpleNamespace(pad_token_id=0, eos_token_id=2)
    pol = AutoregressivePolicy(args, model, tok, "lora_policy")
    ref = AutoregressivePolicy(args, model, tok, "lora_ref_policy")
    stepper = DPOStepper(pol, ref, args)
    opt = B200AdamW(model, lr=1e-3)
    sch = cosine_schedule(opt, 2, 10)
    images, q, qm, resp = O.synthetic_batch(oc, 2, 16, 24, NAMES2, left_pad=2)
    batch = dict(images=images, queries=q, queries_attn_masks=qm, **resp)
    for _ in range(2):
        refs = stepper.reference_logprobs(batch)
        stepper.pair_loss(batch, refs, NAMES2[0], NAMES2[1])
        opt.step(); opt.zero_grad(); sch.step()
    ckpt = save_checkpoint(model, opt, sch, str(tmp_path), 2)
    d = os.path.join(ckpt, "adapter_model", "lora_policy")
    cfgj = json.load(open(os.path.join(d, "adapter_config.json")))
    assert cfgj["r"] == oc.lora_r and cfgj["lora_alpha"] == oc.lora_alpha and cfgj["bias"] == "none" and \
        cfgj["task_type"] == "CAUSAL_LM" and "q_proj" in cfgj["target_modules"]
    saved = torch.load(os.path.join(d, "adapter_model.bin"))
    assert "base_model.model.model.layers.0.self_attn.q_proj.lora_A.weight" in saved
    assert saved["base_model.model.model.layers.1.mlp.down_proj.lora_B.weight"].shape == (oc.d_model, oc.lora_r)
    assert os.path.exists(os.path.join(ckpt, "optimizer.pt")) and os.path.exists(os.path.join(ckpt, "scheduler.pt"))
    assert get_last_checkpoint(str(tmp_path)) == ckpt

    model2 = make()
    opt2 = B200AdamW(model2, lr=1e-3)
    sch2 = cosine_schedule(opt2, 2, 10)
    step = load_checkpoint(model2, opt2, sch2, ckpt)
    assert step == 2 and opt2.step_count == 2 and sch2.get_last_lr() == sch.get_last_lr()
    again = model2.adapter_state_dict("lora_policy")
    for k, v in saved.items():
        assert torch.equal(again[k], v), k
    v1, v2 = model.engine.buffers(), model2.engine.buffers()
    assert torch.equal(v1["exp_avg"], v2["exp_avg"]) and torch.equal(v1["exp_avg_sq"], v2["exp_avg_sq"])
    # the fp32 masters come back exactly (optimizer.pt carries them; adapter_model.bin alone would give their bf16 rounding),
    # so a resumed run takes the same next step as the uninterrupted one
    n = v1["n_train"]
    assert torch.equal(v1["master"][:n], v2["master"][:n])
    assert not torch.equal(v1["master"][:n], v1["master"][:n].to(torch.bfloat16).float())     # the rounding would have lost something
    for m, o, sc in ((model, opt, sch), (model2, opt2, sch2)):
        p_, r_ = AutoregressivePolicy(args, m, tok, "lora_policy"), AutoregressivePolicy(args, m, tok, "lora_ref_policy")
        st = DPOStepper(p_, r_, args)
        refs = st.reference_logprobs(batch)
        st.pair_loss(batch, refs, NAMES2[0], NAMES2[1])
        o.step(); o.zero_grad(); sc.step()
    torch.cuda.synchronize()
    # (split-K wgrad atomics land in any order, so two runs agree to fp32 noise, not bit for bit; a resume from the bf16
    # rounding alone would sit ~3e-5 away)
    assert (v1["master"][:n] - v2["master"][:n]).abs().mean().item() < 1e-6
    # scheduler.pt written by the reference is a torch LambdaLR state dict: last_epoch is what is restored
    sch3 = cosine_schedule(opt2, 2, 10)
    sch3.load_state_dict({"last_epoch": 2, "base_lrs": [1e-3], "_step_count": 3, "lr_lambdas": [None], "_last_lr": [0.0]})
    assert sch3.last_epoch == 2 and sch3.warmup == 2 and sch3.total == 10
    # optimizer.pt written by the reference (one state entry per LoRA tensor) is rejected with a clear message
    with pytest.raises(ValueError, match="one state entry per LoRA tensor"):
        opt2.load_state_dict({"state": {0: {"exp_avg": torch.zeros(4)}, 1: {"exp_avg": torch.zeros(4)}}, "param_groups": [{}]})
    model.engine.close(); model2.engine.close()


def test_stepper_concurrent_reference_pass_equals_sequential_step():
    """DPOStepper.step with the frozen-adapter pass enqueued on a side stream next to the policy forward (the default) against
    the same step run sequentially on one stream: both passes are deterministic, so the loss is identical bit for bit; the
    gradient norm and the updated masters agree to fp32-atomics noise (split-K wgrad partials land in any order)."""
    from types import SimpleNamespace
    from oracle import llava_oracle as O
    from opa_dpo_b200.policy import AutoregressivePolicy, B200LlavaModel
    from opa_dpo_b200.dpo import DPOArgs, DPOStepper
    oc, ec = _cfgs()
    args = DPOArgs(beta=0.1, response_len=24, learning_rate=1e-3)
    tok = SimpleNamespace(pad_token_id=0, eos_token_id=2)
    images, q, qm, resp = O.synthetic_batch(oc, 2, 16, 24, NAMES2, left_pad=2)
    out = {}
    for concurrent in (True, False):
        model = B200LlavaModel(ec)
        w = O.init_weights(oc)
        sd = dict(w)
        sd[O.VT + "embeddings.patch_embedding.weight"] = w[O.VT + "embeddings.patch_embedding.weight"].reshape(oc.v_d, -1)
        model.engine.load_state_dict(sd)
        for name, seed in (("lora_policy", 11), ("lora_ref_policy", 12)):
            lo = O.init_lora(oc, seed)
            model.load_adapter_state(name, {f"base_model.model.{m}.lora_{x}.weight": t for m, ab in lo.items() for x, t in zip("AB", ab)})
        model.make_trainable("lora_policy")
        dev = model.engine.device
        batch = dict(images=images.to(dev), queries=q.to(dev), queries_attn_masks=qm.to(dev), **{k: v.to(dev) for k, v in resp.items()})
        st = DPOStepper(AutoregressivePolicy(args, model, tok, "lora_policy"), AutoregressivePolicy(args, model, tok, "lora_ref_policy"), args)
        st.concurrent_reference = concurrent
        losses, norms = [], []
        for _ in range(3):      # three steps: the side stream's buffers are recycled across steps
            loss, norm = st.step(batch, NAMES2[0], NAMES2[1])
            losses.append(float(loss)); norms.append(float(norm))
        torch.cuda.synchronize()
        v = model.engine.buffers()
        out[concurrent] = (losses, norms, v["master"][: v["n_train"]].clone())
        assert (st._side is not None) == concurrent
        model.engine.close()
    (l1, n1, m1), (l0, n0, m0) = out[True], out[False]
    assert l1[0] == l0[0], (l1, l0)                      # first step: identical weights, deterministic forwards
    for a, b in zip(l1, l0):
        assert abs(a - b) <= 1e-5 * abs(b), (l1, l0)
    for a, b in zip(n1, n0):
        assert abs(a - b) <= 1e-3 * abs(b), (n1, n0)
    assert (m1 - m0).abs().mean().item() < 1e-6


def test_opa_sft_loss_and_grads_vs_oracle():
    """OPA stage (LoRA-SFT): token-mean NLL over the response tokens == HF causal-LM loss with prompt labels ignored
    (opa_trainer.py:58-125, entropy_loss off); loss and LoRA gradients vs the oracle's logits + F.cross_entropy."""
    from types import SimpleNamespace
    import torch.nn.functional as F
    from oracle import llava_oracle as O
    from opa_dpo_b200.policy import AutoregressivePolicy, B200LlavaModel
    from opa_dpo_b200.opa import sft_loss
    from opa_dpo_b200.dpo import DPOArgs
    oc, ec = _cfgs()
    model = B200LlavaModel(ec)
    w = O.init_weights(oc)
    lp = O.init_lora(oc, 11)
    sd = dict(w)
    sd[O.VT + "embeddings.patch_embedding.weight"] = w[O.VT + "embeddings.patch_embedding.weight"].reshape(oc.v_d, -1)
    model.engine.load_state_dict(sd)
    model.load_adapter_state("lora_policy", {f"base_model.model.{m}.lora_{x}.weight": t for m, ab in lp.items() for x, t in zip("AB", ab)})
    model.make_trainable("lora_policy")
    R = 24
    policy = AutoregressivePolicy(DPOArgs(response_len=R), model, SimpleNamespace(pad_token_id=0, eos_token_id=2), "lora_policy")
    images, q, qm, resp = O.synthetic_batch(oc, 3, 16, R, ["standard_response"], left_pad=2)
    # oracle: logits of the full row, labels = response tokens, everything else IGNORE_INDEX (HF shift-by-one loss)
    lpg = {m: (A.clone().requires_grad_("vision_model" not in m), B.clone().requires_grad_("vision_model" not in m)) for m, (A, B) in lp.items()}
    ids = torch.cat([q, resp["standard_response"]], 1)
    mask = torch.cat([qm.bool(), resp["standard_response"] != 0], 1)
    logits = O.llava_logits(oc, w, lpg, ids, mask, images)[:, -R - 1:-1]
    labels = resp["standard_response"].clone()
    labels[labels == 0] = -100
    loss_ref = F.cross_entropy(logits.reshape(-1, oc.vocab), labels.reshape(-1), ignore_index=-100)
    loss_ref.backward()
    policy.train()
    out = policy(images, q, qm, **resp)
    loss = sft_loss(out, resp, 0)
    assert abs(loss.item() - loss_ref.item()) <= 1e-3 * abs(loss_ref.item()), (loss.item(), loss_ref.item())
    model.engine.zero_grad()
    loss.backward()
    torch.cuda.synchronize()
    g = model.engine.buffers()["grad"]
    for m in ("model.layers.0.self_attn.q_proj", "model.layers.1.mlp.down_proj", "model.mm_projector.2"):
        oa, ob = model.engine.lora_offsets(m)
        for t, off in ((lpg[m][0], oa), (lpg[m][1], ob)):
            gg = g[off:off + t.numel()].view_as(t).cpu()
            cos = torch.nn.functional.cosine_similarity(gg.flatten(), t.grad.flatten(), dim=0).item()
            assert cos >= 0.99, (m, cos)
    model.engine.close()


def test_activation_slots_are_never_silently_reused():
    """The engine keeps the activations of two undifferentiated forwards. A third grad-enabled forward raises instead of
    overwriting slot 1; a graph dropped without backward hands its slot back; after release_slots() a late backward raises
    instead of differentiating another forward's activations; torch.no_grad() forwards take no slot."""
    from types import SimpleNamespace
    from oracle import llava_oracle as O
    from opa_dpo_b200 import OpadpoError
    from opa_dpo_b200.policy import AutoregressivePolicy, B200LlavaModel
    from opa_dpo_b200.dpo import DPOArgs
    oc, ec = _cfgs()
    model = B200LlavaModel(ec)
    w = O.init_weights(oc)
    lp = O.init_lora(oc, 11)
    sd = dict(w)
    sd[O.VT + "embeddings.patch_embedding.weight"] = w[O.VT + "embeddings.patch_embedding.weight"].reshape(oc.v_d, -1)
    model.engine.load_state_dict(sd)
    model.load_adapter_state("lora_policy", {f"base_model.model.{m}.lora_{x}.weight": t for m, ab in lp.items() for x, t in zip("AB", ab)})
    model.make_trainable("lora_policy")
    policy = AutoregressivePolicy(DPOArgs(response_len=24), model, SimpleNamespace(pad_token_id=0, eos_token_id=2), "lora_policy")
    images, q, qm, resp = O.synthetic_batch(oc, 2, 16, 24, ["standard_response"], left_pad=2)
    key = "standard_response_logprobs"
    a = policy(images, q, qm, **resp)
    b = policy(images, q, qm, **resp)
    assert model._slot_busy == [True, True]
    with torch.no_grad():
        policy(images, q, qm, **resp)                              # inference: no slot needed
    with pytest.raises(OpadpoError, match="both activation slots"):
        policy(images, q, qm, **resp)
    del b                                                          # graph dropped without backward -> slot 1 is free again
    import gc; gc.collect()
    assert model._slot_busy == [True, False]
    c = policy(images, q, qm, **resp)
    assert model._slot_busy == [True, True]
    model.engine.zero_grad()
    a[key].sum().backward()
    c[key].sum().backward()
    torch.cuda.synchronize()
    g1 = model.engine.buffers()["grad"].clone()
    assert model._slot_busy == [False, False] and g1.abs().sum().item() > 0
    d = policy(images, q, qm, **resp)
    model.release_slots()                                          # e.g. zero_grad() after a failed step
    e = policy(images, q, qm, **resp)                              # takes slot 0 again
    with pytest.raises(RuntimeError, match="released or reused"):
        d[key].sum().backward()
    model.engine.zero_grad()
    e[key].sum().backward()
    torch.cuda.synchronize()
    g2 = model.engine.buffers()["grad"]
    assert torch.allclose(g2 * 2, g1, rtol=1e-3, atol=1e-6)        # a and c were the same forward twice
    model.engine.close()


def test_opa_entropy_regulariser_loss_and_grads_vs_oracle():
    """OPA stage with entropy_loss on (opa_trainer.py:58-125): base forward + a second forward on masked images, loss =
    sft + coef * mean_rows(-sum((H_masked - H) * mask) / sum(mask)); the gradient flows through the entropies of BOTH
    forwards (two engine save slots, opadpo_backward_ent). Loss and LoRA gradients vs the oracle's logits + autograd.
    lm_head is scaled x8 so the entropies sit well below ln(vocab) and differ between the two forwards (with N(0, 0.02^2)
    logits the regulariser is ~1e-4 and would test nothing); tolerances: entropy term 0.02 absolute, loss 1e-2 relative
    (bf16 logits of magnitude ~5), gradients cosine >= 0.99 and relative L2 <= 0.1."""
    from types import SimpleNamespace
    import torch.nn.functional as F
    from oracle import llava_oracle as O
    from opa_dpo_b200.policy import AutoregressivePolicy, B200LlavaModel
    from opa_dpo_b200.opa import OPAArgs, OPAStepper
    from opa_dpo_b200.dpo import DPOArgs
    oc, ec = _cfgs()
    model = B200LlavaModel(ec)
    w = O.init_weights(oc, peaked=8.0)
    lp = O.init_lora(oc, 11)
    sd = dict(w)
    sd[O.VT + "embeddings.patch_embedding.weight"] = w[O.VT + "embeddings.patch_embedding.weight"].reshape(oc.v_d, -1)
    model.engine.load_state_dict(sd)
    model.load_adapter_state("lora_policy", {f"base_model.model.{m}.lora_{x}.weight": t for m, ab in lp.items() for x, t in zip("AB", ab)})
    model.make_trainable("lora_policy")
    R, coef = 24, 0.5
    policy = AutoregressivePolicy(DPOArgs(response_len=R), model, SimpleNamespace(pad_token_id=0, eos_token_id=2), "lora_policy")
    images, q, qm, resp = O.synthetic_batch(oc, 3, 16, R, ["standard_response"], left_pad=2)
    masked = images.clone()
    masked[:, :, ::3, ::2] = images.mean()          # any masked copy of the images serves (the masking itself is f-3)
    lpg = {m: (A.clone().requires_grad_("vision_model" not in m), B.clone().requires_grad_("vision_model" not in m)) for m, (A, B) in lp.items()}
    ids = torch.cat([q, resp["standard_response"]], 1)
    mask = torch.cat([qm.bool(), resp["standard_response"] != 0], 1)
    labels = resp["standard_response"].clone()
    labels[labels == 0] = -100
    valid = (labels != -100).float()
    mods = ("model.layers.0.self_attn.q_proj", "model.layers.1.mlp.down_proj", "model.mm_projector.2")

    def ref_pass(img):
        logits = O.llava_logits(oc, w, lpg, ids, mask, img)[:, -R - 1:-1]
        ls = torch.log_softmax(logits, -1)
        return F.cross_entropy(logits.reshape(-1, oc.vocab), labels.reshape(-1), ignore_index=-100), -(ls.exp() * ls).sum(-1)

    def check_grads(cos_min, rel_max, what):
        torch.cuda.synchronize()
        g = model.engine.buffers()["grad"]
        for m in mods:
            oa, ob = model.engine.lora_offsets(m)
            for t, off in ((lpg[m][0], oa), (lpg[m][1], ob)):
                gg = g[off:off + t.numel()].view_as(t).cpu()
                cos = torch.nn.functional.cosine_similarity(gg.flatten(), t.grad.flatten(), dim=0).item()
                rel = ((gg - t.grad).norm() / t.grad.norm()).item()
                assert cos >= cos_min and rel <= rel_max, (what, m, cos, rel)

    # (1) the entropy head alone, one forward: loss = sum(wts * H) with random per-token weights
    gen = torch.Generator().manual_seed(5)
    wts = torch.randn(3, R, generator=gen) * valid
    _, h_ref = ref_pass(images)
    (wts * h_ref).sum().backward()
    policy.train()
    model.engine.zero_grad()
    out = policy(images, q, qm, **resp)
    h = out["standard_response_entropies"]
    assert ((h.detach().cpu() - h_ref.detach() * valid).abs().max().item()) <= 0.08
    (wts.to(h.device) * h).sum().backward()
    check_grads(0.99, 0.1, "entropy head")
    assert model._slot_busy == [False, False]

    # (2) the full regularised OPA loss: two forwards, both differentiated
    for a, b in lpg.values():
        a.grad = None
        b.grad = None
    sft_ref, h_ref = ref_pass(images)
    _, hm_ref = ref_pass(masked)
    ent_ref = (-((hm_ref - h_ref) * valid).sum(-1) / valid.sum(-1)).mean()
    loss_ref = sft_ref + coef * ent_ref
    loss_ref.backward()
    stepper = OPAStepper(policy, None, args=OPAArgs(entropy_loss=True, entropy_loss_coef=coef, entropy_decay_coef=1.0,
                                                    entropy_alignment="label"))
    model.engine.zero_grad()
    loss, logs = stepper.compute_loss(images, q, qm, resp, masked_images=masked)
    assert abs(logs["entropy_loss"].item() - ent_ref.item()) <= 0.02, (logs["entropy_loss"].item(), ent_ref.item())
    assert abs(loss.item() - loss_ref.item()) <= 1e-2 * abs(loss_ref.item()), (loss.item(), loss_ref.item())
    loss.backward()
    check_grads(0.99, 0.1, "sft + entropy regulariser")
    assert model._slot_busy == [False, False]
    model.engine.close()


@pytest.mark.parametrize("share", [0, 1])
def test_reference_pass_on_merged_frozen_adapter(share):
    """Option merge_frozen: the frozen reference adapter's inference forward runs on W' = W + (alpha/r) B A merged once
    (no LoRA GEMMs). Same tolerances vs the oracle as the unmerged pass; the DPO loss built on it stays within 1e-3; the
    merged copy is rebuilt when the adapter is reloaded; the trainable adapter is never merged."""
    from oracle import llava_oracle as O
    from opa_dpo_b200 import POLICY, REF
    from opa_dpo_b200.dpo import DPOArgs, FusedDPO
    oc, ec = _cfgs()
    eng, w, lp, lr = _build(oc, ec, share_prefix=share)
    eng.set_option("merge_frozen", 1)
    images, q, qm, resp = O.synthetic_batch(oc, 3, 16, 24, NAMES2, left_pad=3)
    dev = eng.device
    args = [images.to(dev), q.to(dev), qm.to(dev), [resp[n].to(dev) for n in NAMES2]]
    ref_o = O.policy_forward(oc, w, lr, images, q, qm, resp)
    pol_o = O.policy_forward(oc, w, lp, images, q, qm, resp)
    ref_lp, ref_ent = eng.forward_logprobs(REF, *args)
    ref_lp2, _ = eng.forward_logprobs(REF, *args)               # second call reuses the cached merged weights
    pol_lp, _ = eng.forward_logprobs(POLICY, *args)             # trainable adapter: unmerged path
    torch.cuda.synchronize()
    assert torch.equal(ref_lp, ref_lp2)
    for i, n in enumerate(NAMES2):
        _cmp_lp(ref_lp[i], ref_o[n + "_logprobs"], f"merged ref {n} logprobs")
        _cmp_lp(ref_ent[i], ref_o[n + "_entropies"], f"merged ref {n} entropies")
        _cmp_lp(pol_lp[i], pol_o[n + "_logprobs"], f"policy {n} logprobs")
    loss_o, _, _ = O.pair_loss(O.LossConfig(beta=0.1), pol_o[NAMES2[0] + "_logprobs"], pol_o[NAMES2[1] + "_logprobs"],
                               ref_o[NAMES2[0] + "_logprobs"], ref_o[NAMES2[1] + "_logprobs"])
    total, _ = FusedDPO(DPOArgs(beta=0.1)).pair(pol_lp[0], pol_lp[1], ref_lp[0], ref_lp[1])
    rel = abs(total.item() - loss_o.item()) / abs(loss_o.item())
    assert rel <= 1e-3, f"DPO loss on the merged reference pass: {total.item():.6f} vs {loss_o.item():.6f} (rel {rel:.3g})"
    # the unmerged pass of the same adapter agrees to bf16 noise
    eng.set_option("merge_frozen", 0)
    ref_lp_unmerged, _ = eng.forward_logprobs(REF, *args)
    assert (ref_lp_unmerged - ref_lp).abs().max().item() <= 0.05
    # reloading the adapter invalidates the merged copy: load the POLICY tensors into the REF slot
    eng.set_option("merge_frozen", 1)
    eng.forward_logprobs(REF, *args)
    eng.load_adapter(REF, lp)
    swapped, _ = eng.forward_logprobs(REF, *args)
    torch.cuda.synchronize()
    for i, n in enumerate(NAMES2):
        _cmp_lp(swapped[i], pol_o[n + "_logprobs"], f"reloaded+remerged {n} logprobs")
    eng.close()


def test_trainer_mirror_outer_step_matches_manual_composition():
    """B200DPOTrainer (rollout -> step_with_rollouts with CoPO, grad accumulation, noptepochs=2, shuffled mini-batches)
    against the same step composed by hand from the parity-tested pieces (reference-adapter pass, policy pass,
    FusedDPO.compute_policy_loss, LoRA backward, fused clip+AdamW): the same LoRA masters afterwards. Not bit for bit:
    the LoRA wgrad GEMMs accumulate split-K partials with fp32 atomics, so two runs differ in the last bits of the
    gradient and Adam's first steps (update ~ lr * sign(g)) amplify that at elements whose gradient is ~0; tolerance:
    the two parameter updates have cosine >= 0.999 and 99.5 % of the elements agree to 1e-6."""
    from types import SimpleNamespace
    from torch.utils.data import DataLoader
    from oracle import llava_oracle as O
    from opa_dpo_b200 import collate
    from opa_dpo_b200.policy import AutoregressivePolicy, B200LlavaModel
    from opa_dpo_b200.dpo import DPOArgs, FusedDPO
    from opa_dpo_b200.optim import B200AdamW
    from opa_dpo_b200.trainer import B200DPOTrainer, RESPONSES
    oc, ec = _cfgs()
    R = 24
    args = DPOArgs(beta=0.1, CoPO=True, CoPO_method="random", CoPO_mask_ratio=0.3, CoPO_coef=0.2, AncPO=True, response_len=R,
                   rollout_accumulation_steps=2, step_per_device_batch_size=2, gradient_accumulation_steps=2, noptepochs=2,
                   bf16=True)
    tok = SimpleNamespace(pad_token_id=0, eos_token_id=2)
    w = O.init_weights(oc)
    lp, lr = O.init_lora(oc, 11), O.init_lora(oc, 12)

    def make():
        model = B200LlavaModel(ec)
        sd = dict(w)
        sd[O.VT + "embeddings.patch_embedding.weight"] = w[O.VT + "embeddings.patch_embedding.weight"].reshape(oc.v_d, -1)
        model.engine.load_state_dict(sd)
        for name, lo in (("lora_policy", lp), ("lora_ref_policy", lr)):
            model.load_adapter_state(name, {f"base_model.model.{m}.lora_{x}.weight": t for m, ab in lo.items() for x, t in zip("AB", ab)})
        model.make_trainable("lora_policy")
        return (model, AutoregressivePolicy(args, model, tok, "lora_policy"), AutoregressivePolicy(args, model, tok, "lora_ref_policy"),
                B200AdamW(model, lr=1e-3))

    batches = []
    for s in range(2):
        images, q, qm, resp = O.synthetic_batch(oc, 2, 16, R, list(RESPONSES), seed=100 + s, left_pad=2)
        b = {"images": images, "queries": q, "queries_attention_mask": qm.long()}
        for n in RESPONSES:
            b[n] = resp[n]
            b[n + "_attention_mask"] = (resp[n] != 0).long()
        batches.append(b)

    # ---- the trainer
    model, policy, ref_policy, opt = make()
    tr = B200DPOTrainer(args, policy, ref_policy, opt)
    torch.manual_seed(7)
    tr.step(iter(batches))
    stats = tr.last_train_stats
    torch.cuda.synchronize()
    assert stats["loss/grad_norm"].shape == (2,) and torch.isfinite(stats["loss/grad_norm"]).all()   # 4 micro-batches / gas 2
    got = model.engine.buffers()["master"].clone()
    assert opt.step_count == 2
    model.engine.close()
    init = None

    # ---- by hand, same random stream
    model, policy, ref_policy, opt = make()
    dev = model.engine.device
    init = model.engine.buffers()["master"].clone()
    fused = FusedDPO(args)
    torch.manual_seed(7)
    rows = []
    with torch.no_grad():
        for b in batches:
            img = b["images"].to(dev).to(torch.bfloat16)
            masked = collate.mask_images(img, 0.3, "random")
            resp = {n: b[n].to(dev) for n in RESPONSES}
            sub = {"standard_response": resp["standard_response"], "AI_pseudo_response": resp["AI_pseudo_response"]}
            r = {"images": img, "masked_images": masked, "queries": b["queries"].to(dev), "qm": b["queries_attention_mask"].to(dev), **resp}
            r.update({"ref_base_" + k: v for k, v in ref_policy(img, r["queries"], r["qm"], **resp).items()})
            r.update({"ref_mask_" + k: v for k, v in ref_policy(masked, r["queries"], r["qm"], **sub).items()})
            rows.append(r)
    ro = {k: torch.cat([r[k] for r in rows]) for k in rows[0]}
    micro = 0
    for _ in range(2):
        for idx in DataLoader(range(4), batch_size=2, shuffle=True, drop_last=True):
            idx = idx.to(dev)
            rb = {k: v.index_select(0, idx) for k, v in ro.items()}
            policy.train()
            out = policy(rb["images"], rb["queries"], rb["qm"], mode="policy", **{n: rb[n] for n in RESPONSES})
            out_m = policy(rb["masked_images"], rb["queries"], rb["qm"], mode="policy",
                           mask_standard_response=rb["standard_response"], mask_AI_pseudo_response=rb["AI_pseudo_response"])
            loss, _ = fused.compute_policy_loss(out, rb, out_m)
            (loss / 2).backward()
            micro += 1
            if micro % 2 == 0:
                opt.step(1.0)
                opt.zero_grad()
    torch.cuda.synchronize()
    want = model.engine.buffers()["master"]
    n_train = model.engine.buffers()["n_train"]
    d_got, d_want = (got - init)[:n_train], (want - init)[:n_train]
    assert d_want.abs().max().item() > 1e-4, "the step must move the parameters"
    cos = torch.nn.functional.cosine_similarity(d_got, d_want, dim=0).item()
    close = ((d_got - d_want).abs() <= 1e-6).float().mean().item()
    assert cos >= 0.999 and close >= 0.995, (cos, close)
    assert torch.equal(got[n_train:], want[n_train:])          # the CLIP-tower LoRA tail is never trained
    model.engine.close()


def test_online_generator_mirror_greedy_responses_vs_oracle():
    """B200OnlineGenerator.generate_responses (online_generator.py:292-323): generate -> slice off the query -> truncate
    after EOS / the extra stop ids, all on the device; greedy ids and the truncated rows equal the oracle's."""
    from types import SimpleNamespace
    from oracle import llava_oracle as O
    from opa_dpo_b200.generation import B200OnlineGenerator
    from opa_dpo_b200.policy import B200LlavaModel
    oc, ec = _cfgs()
    model = B200LlavaModel(ec)
    w = O.init_weights(oc, peaked=40.0)
    lp = O.init_lora(oc, 11)
    sd = dict(w)
    sd[O.VT + "embeddings.patch_embedding.weight"] = w[O.VT + "embeddings.patch_embedding.weight"].reshape(oc.v_d, -1)
    model.engine.load_state_dict(sd)
    model.load_adapter_state("lora_policy", {f"base_model.model.{m}.lora_{x}.weight": t for m, ab in lp.items() for x, t in zip("AB", ab)})
    model.set_adapter("lora_policy")
    images, q, qm, _ = O.synthetic_batch(oc, 3, 16, 8, NAMES2, left_pad=3)
    ids_ref, margin = O.greedy_generate(oc, w, lp, images, q, qm, 12)
    assert margin > 0.05
    gen = B200OnlineGenerator(SimpleNamespace(response_len=12, top_p=1.0, top_k=0, temperature=1.0),
                              model, SimpleNamespace(pad_token_id=oc.pad_id, eos_token_id=oc.eos_id))
    dev = model.engine.device
    got = gen.generate_responses(images.to(dev), q.to(dev), qm.to(dev), do_sample=False)
    want = O.truncate_after_eos_with_padding(ids_ref[:, 16:], oc.eos_id, oc.pad_id, [1577, 29973])
    assert got.shape == (3, 12) and torch.equal(got.cpu(), want)
    model.engine.close()


def test_opa_entropy_regulariser_reference_alignment_vs_oracle():
    """OPAArgs.entropy_alignment="reference": the regulariser reads the entropy at the position of each labelled token
    (opa_trainer.py:63,82-84), one later than the cross-entropy. Loss vs the oracle's logits at those positions."""
    from types import SimpleNamespace
    import torch.nn.functional as F
    from oracle import llava_oracle as O
    from opa_dpo_b200.policy import AutoregressivePolicy, B200LlavaModel
    from opa_dpo_b200.opa import OPAArgs, OPAStepper
    from opa_dpo_b200.dpo import DPOArgs
    oc, ec = _cfgs()
    model = B200LlavaModel(ec)
    w = O.init_weights(oc, peaked=8.0)
    lp = O.init_lora(oc, 11)
    sd = dict(w)
    sd[O.VT + "embeddings.patch_embedding.weight"] = w[O.VT + "embeddings.patch_embedding.weight"].reshape(oc.v_d, -1)
    model.engine.load_state_dict(sd)
    model.load_adapter_state("lora_policy", {f"base_model.model.{m}.lora_{x}.weight": t for m, ab in lp.items() for x, t in zip("AB", ab)})
    model.make_trainable("lora_policy")
    R, coef = 24, 0.5
    policy = AutoregressivePolicy(DPOArgs(response_len=R), model, SimpleNamespace(pad_token_id=0, eos_token_id=2), "lora_policy")
    images, q, qm, resp = O.synthetic_batch(oc, 3, 16, R, ["standard_response"], left_pad=2)
    masked = images.clone()
    masked[:, :, ::3, ::2] = images.mean()
    r = resp["standard_response"]
    ids = torch.cat([q, r], 1)
    mask = torch.cat([qm.bool(), r != 0], 1)
    labels = r.clone()
    labels[labels == 0] = -100
    valid = (labels != -100).float()

    def ref_pass(img):
        logits = O.llava_logits(oc, w, lp, ids, mask, img)
        ls = torch.log_softmax(logits[:, -R:], -1)                      # positions OF the response tokens
        ce = F.cross_entropy(logits[:, -R - 1:-1].reshape(-1, oc.vocab), labels.reshape(-1), ignore_index=-100)
        return ce, -(ls.exp() * ls).sum(-1)
    with torch.no_grad():
        sft_ref, h = ref_pass(images)
        _, hm = ref_pass(masked)
    want = sft_ref + coef * (-((hm - h) * valid).sum(-1) / valid.sum(-1)).mean()
    stepper = OPAStepper(policy, None, args=OPAArgs(entropy_loss=True, entropy_loss_coef=coef, entropy_decay_coef=1.0,
                                                    entropy_alignment="reference"))
    policy.train()
    model.engine.zero_grad()
    loss, logs = stepper.compute_loss(images, q, qm, resp, masked_images=masked)
    assert policy.args.response_len == R
    assert abs(loss.item() - want.item()) <= 1e-2 * abs(want.item()), (loss.item(), want.item())
    loss.backward()
    torch.cuda.synchronize()
    assert model._slot_busy == [False, False] and torch.isfinite(model.engine.buffers()["grad"]).all()
    model.engine.close()


@pytest.mark.parametrize("eos", [298, 215, 159])
def test_greedy_rollout_early_eos_ids_exact(eos):
    """Rows that finish early at different steps: the EOS id is moved onto tokens the greedy path really emits; finished rows
    emit pad, the others go on — ids equal to the oracle's (which equal transformers' generate())."""
    import dataclasses
    from oracle import llava_oracle as O
    from opa_dpo_b200 import EngineConfig, POLICY
    oc = dataclasses.replace(O.OracleConfig(), eos_id=eos)
    ec = EngineConfig(**oc.to_dict())
    eng, w, lp, lr = _build(oc, ec, peaked=40.0, trainable=False)
    images, q, qm, _ = O.synthetic_batch(oc, 2, 16, 24, NAMES3, left_pad=2)
    ids_ref, margin = O.greedy_generate(oc, w, lp, images, q, qm, 12)
    got = eng.generate(POLICY, images, q, qm, 12).cpu()
    assert margin > 0.05 and torch.equal(got, ids_ref), (got[:, 16:], ids_ref[:, 16:])
    assert (ids_ref[:, 16:] == oc.pad_id).any()
    eng.close()


@pytest.mark.parametrize("share", [0, 1])
def test_fused_rope_epilogue_option_vs_oracle(share):
    """Option fuse_rope: RoPE applied in the q|k|v GEMM epilogue instead of a separate pass. Same tolerances vs the oracle as
    the default path (forward log-probs, DPO loss 1e-3, LoRA gradients), and close to the unfused path."""
    from oracle import llava_oracle as O
    from opa_dpo_b200 import POLICY, REF
    from opa_dpo_b200.dpo import DPOArgs, FusedDPO
    oc, ec = _cfgs()
    eng, w, lp, lr = _build(oc, ec, share_prefix=share)
    images, q, qm, resp = O.synthetic_batch(oc, 3, 16, 24, NAMES2, left_pad=3)
    dev = eng.device
    args = [images.to(dev), q.to(dev), qm.to(dev), [resp[n].to(dev) for n in NAMES2]]
    eng.set_option("fuse_rope", 0)
    base_lp, _ = eng.forward_logprobs(POLICY, *args)
    eng.set_option("fuse_rope", 1)
    lc = O.LossConfig(beta=0.1)
    loss_ref, lpg, ref_o, pol_o = _oracle_grads(oc, w, lp, lr, images, q, qm, resp, lc, NAMES2)
    ref_lp, _ = eng.forward_logprobs(REF, *args)
    pol_lp, _ = eng.forward_logprobs(POLICY, *args, save_slot=0)
    for i, n in enumerate(NAMES2):
        _cmp_lp(pol_lp[i], pol_o[n + "_logprobs"], f"fused rope policy {n}")
        _cmp_lp(ref_lp[i], ref_o[n + "_logprobs"], f"fused rope ref {n}")
    assert (pol_lp - base_lp).abs().max().item() <= 0.05
    pl = pol_lp.clone().requires_grad_(True)
    total, _ = FusedDPO(DPOArgs(beta=0.1)).pair(pl[0], pl[1], ref_lp[0], ref_lp[1])
    total.backward()
    assert abs(total.item() - loss_ref.item()) <= 1e-3 * abs(loss_ref.item())
    eng.zero_grad()
    eng.backward(0, pl.grad)
    torch.cuda.synchronize()
    grad = eng.buffers()["grad"]
    for m in ("model.layers.0.self_attn.q_proj", "model.layers.1.self_attn.k_proj", "model.layers.1.mlp.down_proj"):
        A, B = lpg[m]
        oa, ob = eng.lora_offsets(m)
        for t, off in ((A, oa), (B, ob)):
            gg = grad[off:off + t.numel()].view_as(t).cpu()
            cos = torch.nn.functional.cosine_similarity(gg.flatten(), t.grad.flatten(), dim=0).item()
            assert cos >= 0.99, (m, cos)
    ids = eng.generate(POLICY, images, q, qm, 4)          # the rollout prefill goes through the same projection
    assert ids.shape == (3, 20)
    eng.close()


def test_fused_swiglu_on_merged_reference_pass_is_bit_identical():
    """Option fuse_swiglu (with merge_frozen): the merged reference adapter's gate|up projection runs SwiGLU in its epilogue
    over the interleaved merged weight. The roundings of the unfused path are reproduced, so the reference log-probs are
    bit-identical with the option on and off; the rollout (which reads the non-interleaved copy) is unaffected."""
    from oracle import llava_oracle as O
    from opa_dpo_b200 import POLICY, REF
    oc, ec = _cfgs(d_ffn=512)
    eng, w, lp, lr = _build(oc, ec, trainable=True)
    images, q, qm, resp = O.synthetic_batch(oc, 3, 16, 24, NAMES2, left_pad=3)
    dev = eng.device
    args = [images.to(dev), q.to(dev), qm.to(dev), [resp[n].to(dev) for n in NAMES2]]
    eng.set_option("merge_frozen", 1)
    eng.set_option("fuse_swiglu", 0)
    a_lp, a_ent = eng.forward_logprobs(REF, *args)
    ids_a = eng.generate(REF, images, q, qm, 6)
    eng.set_option("fuse_swiglu", 1)
    b_lp, b_ent = eng.forward_logprobs(REF, *args)
    ids_b = eng.generate(REF, images, q, qm, 6)
    torch.cuda.synchronize()
    assert torch.equal(a_lp, b_lp) and torch.equal(a_ent, b_ent) and torch.equal(ids_a, ids_b)
    ref_o = O.policy_forward(oc, w, lr, images, q, qm, resp)
    for i, n in enumerate(NAMES2):
        _cmp_lp(b_lp[i], ref_o[n + "_logprobs"], f"fused swiglu ref {n}")
    pol_a, _ = eng.forward_logprobs(POLICY, *args)            # the trainable adapter never takes the merged path
    eng.set_option("fuse_swiglu", 0)
    pol_b, _ = eng.forward_logprobs(POLICY, *args)
    assert torch.equal(pol_a, pol_b)
    eng.close()
