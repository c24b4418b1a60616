"""Pins oracle/llava_oracle.py against tests/golden/opadpo_tiny.npz — vectors produced by the reference's OWN Python
(rl_models.AutoregressivePolicy.forward, DPOTrainer.dpo_loss / compute_policy_loss, compute_logprobs,
truncate_after_eos_with_padding) driven over HuggingFace LlamaForCausalLM / CLIPVisionModel; see oracle/make_golden.py.
CPU only. fp32 tolerances: 2e-4 absolute on log-probs (different summation orders), 1e-5 relative on losses."""
import os

import numpy as np
import pytest
import torch

from oracle import llava_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "opadpo_tiny.npz")
NAMES3 = ["standard_response", "original_generate_response", "AI_pseudo_response"]


@pytest.fixture(scope="module")
def gold():
    return {k: torch.from_numpy(v) for k, v in np.load(GOLD).items()}


@pytest.fixture(scope="module")
def setup(gold):
    cfg = O.OracleConfig()
    w = O.init_weights(cfg)
    lp, lr = O.init_lora(cfg, 11), O.init_lora(cfg, 12)
    resp = {n: gold["in." + n] for n in NAMES3}
    return cfg, w, lp, lr, gold["in.images"], gold["in.queries"], gold["in.qmask"], resp


def test_synthetic_batch_is_reproducible(gold, setup):
    cfg = setup[0]
    images, q, qm, resp = O.synthetic_batch(cfg, 2, 16, 24, NAMES3, left_pad=2)
    assert torch.equal(q, gold["in.queries"]) and torch.equal(qm, gold["in.qmask"])
    assert torch.allclose(images, gold["in.images"])
    for n in NAMES3:
        assert torch.equal(resp[n], gold["in." + n])


def test_policy_forward_matches_reference(gold, setup):
    cfg, w, lp, lr, images, q, qm, resp = setup
    masked_qm = torch.cat([gold["in.img_mask"], qm], 1)
    mresp = {"mask_standard_response": resp["standard_response"], "mask_AI_pseudo_response": resp["AI_pseudo_response"]}
    with torch.no_grad():
        runs = {"ref": O.policy_forward(cfg, w, lr, images, q, qm, resp, temperature=0.7),
                "pol": O.policy_forward(cfg, w, lp, images, q, qm, resp, temperature=0.7),
                "refm": O.policy_forward(cfg, w, lr, images, q, masked_qm, mresp, temperature=0.7),
                "polm": O.policy_forward(cfg, w, lp, images, q, masked_qm, mresp, temperature=0.7)}
    n = 0
    for tag, out in runs.items():
        for k, v in out.items():
            g = gold[f"{tag}.{k}"]
            assert v.shape == g.shape
            assert (v - g).abs().max().item() <= 2e-4, (tag, k, (v - g).abs().max().item())
            n += 1
    assert n == 20
    # pads are exactly zero (compute_logprobs ignore_index + response_mask)
    assert (runs["pol"]["standard_response_logprobs"][resp["standard_response"] == 0] == 0).all()


def test_compute_policy_loss_and_grads_match_reference(gold, setup):
    cfg, w, lp, lr, images, q, qm, resp = setup
    masked_qm = torch.cat([gold["in.img_mask"], qm], 1)
    mresp = {"mask_standard_response": resp["standard_response"], "mask_AI_pseudo_response": resp["AI_pseudo_response"]}
    scores = {k[3:]: v for k, v in gold.items() if k.startswith("in.") and ("_scores" in k or "_image_relations" in k)}
    lc = O.LossConfig(beta=0.1, CoPO=True, CoPO_coef=0.2, AncPO=True, mDPO_anchor=True, detailed_report=True,
                      response_score=True, response_image_relation=True)
    ref = {"ref_base_" + k[4:]: v for k, v in gold.items() if k.startswith("ref.")}
    ref.update({"ref_" + k[5:]: v for k, v in gold.items() if k.startswith("refm.")})
    lpg = {m: (A.clone().requires_grad_(True), B.clone().requires_grad_(True)) for m, (A, B) in lp.items()}
    pol = O.policy_forward(cfg, w, lpg, images, q, qm, resp, temperature=0.7)
    polm = O.policy_forward(cfg, w, lpg, images, q, masked_qm, mresp, temperature=0.7)
    loss, stats = O.compute_policy_loss(lc, pol, ref, polm, scores)
    assert abs(loss.item() - gold["cpl.loss"].item()) <= 1e-5 * abs(gold["cpl.loss"].item())
    for k, v in stats.items():
        g = gold["cpl.stats." + k]
        assert abs(v.item() - g.item()) <= 1e-5 + 1e-4 * abs(g.item()), (k, v.item(), g.item())
    loss.backward()
    for m in ("model.layers.0.self_attn.q_proj", "model.layers.1.mlp.down_proj", "model.layers.1.mlp.gate_proj",
              "model.mm_projector.0", "model.mm_projector.2"):
        for i, nm in enumerate("AB"):
            g = gold[f"grad.{m}.{nm}"]
            got = lpg[m][i].grad
            assert (got - g).norm().item() <= 1e-3 * g.norm().item() + 1e-9, (m, nm)


@pytest.mark.parametrize("div,ls", [("reverse_kl", 0.0), ("reverse_kl", 0.1), ("js_divergence", 0.0), ("alpha_divergence", 0.0)])
def test_dpo_loss_variants_match_reference(gold, div, ls):
    lc = O.LossConfig(beta=0.1, label_smoothing=ls, f_divergence_type=div)
    l, c, j = O.dpo_loss(lc, gold["pol.standard_response_logprobs"], gold["pol.original_generate_response_logprobs"],
                         gold["ref.standard_response_logprobs"], gold["ref.original_generate_response_logprobs"],
                         gold["in.AI_pseudo_response_scores"], gold["in.original_generate_response_scores"])
    for got, key in ((l, "losses"), (c, "chosen"), (j, "rejected")):
        assert torch.allclose(got, gold[f"dpo.{div}.{ls}.{key}"], rtol=1e-5, atol=1e-6), key


def test_greedy_rollout_matches_hf_cache_loop(gold, setup):
    cfg, _, lp, _, images, q, qm, _ = setup
    wp = O.init_weights(cfg, peaked=40.0)
    ids, margin = O.greedy_generate(cfg, wp, lp, images, q, qm, 12)
    assert margin > 1e-3
    assert torch.equal(ids, gold["gen.ids"]), (ids[:, 16:], gold["gen.ids"][:, 16:])


def test_truncate_and_compute_logprobs_match_reference(gold):
    comp = gold["trunc.in"]
    assert torch.equal(O.truncate_after_eos_with_padding(comp, 2, 0, [1577, 29973]), gold["trunc.out"])
    assert torch.equal(O.truncate_after_eos_with_padding(comp, 2, 0), gold["trunc.out_noextra"])
    got = O.compute_logprobs(gold["clp.logits"], gold["clp.labels"], 0)
    assert torch.allclose(got, gold["clp.out"], atol=1e-6)
    assert (got[gold["clp.labels"] == 0] == 0).all()


@pytest.mark.parametrize("eos", [2, 298, 215, 159])
def test_greedy_rollout_matches_hf_generate_with_early_eos(gold, setup, eos):
    """The oracle's greedy loop against transformers' own `generate()` (GenerationMixin greedy search [UPSTREAM], what
    online_generator.py:292-309 calls) over the HF stand-in model, with the EOS id moved onto tokens that the greedy path
    really emits, so that rows finish early at different steps: finished rows emit pad, the others go on."""
    import dataclasses
    from oracle.make_golden import HFLlava
    cfg, _, lp, _, images, q, qm, _ = setup
    wp = O.init_weights(cfg, peaked=40.0)
    model = HFLlava(cfg, wp, {"lora_policy": lp})
    model.set_adapter("lora_policy")
    with torch.no_grad():
        x, mask = model.splice(q, qm, images)
        want = model.llm.generate(inputs_embeds=x, attention_mask=mask.long(), max_new_tokens=12, do_sample=False,
                                  pad_token_id=cfg.pad_id, eos_token_id=eos)
    ids, margin = O.greedy_generate(dataclasses.replace(cfg, eos_id=eos), wp, lp, images, q, qm, 12)
    got = ids[:, q.size(1):]
    if want.size(1) < 12:        # generate() stops once every row has finished; the oracle (and the engine) pad to max_new_tokens
        want = torch.cat([want, torch.full((want.size(0), 12 - want.size(1)), cfg.pad_id)], 1)
    assert margin > 1e-3 and torch.equal(got, want), (got, want)
    if eos != 2:
        assert (got == cfg.pad_id).any(), "the case must make at least one row finish early"


def test_layer_checkpointing_option_is_bit_identical():
    """oracle.policy_forward(checkpoint_layers=True) — HF gradient checkpointing per decoder layer, which the reference
    trains with (opadpo_train.py:324-327) and which the reference arm's full-depth CPU timing uses to stay within the host's
    memory — gives the same loss and the same LoRA gradients, bit for bit."""
    from oracle import llava_oracle as O
    oc = O.OracleConfig()
    w, lp = O.init_weights(oc), O.init_lora(oc, 11)
    names = ["chosen_response", "rejected_response"]
    images, q, qm, resp = O.synthetic_batch(oc, 2, 16, 24, names, left_pad=3)

    def run(ck):
        lpg = {m: (A.clone().requires_grad_(not m.startswith(O.VT)), B.clone().requires_grad_(not m.startswith(O.VT)))
               for m, (A, B) in lp.items()}
        out = O.policy_forward(oc, w, lpg, images, q, qm, resp, checkpoint_layers=ck)
        loss = out["chosen_response_logprobs"].sum() - out["rejected_response_logprobs"].sum()
        loss.backward()
        return loss.detach(), {m: (A.grad, B.grad) for m, (A, B) in lpg.items() if A.grad is not None}
    l0, g0 = run(False)
    l1, g1 = run(True)
    assert torch.equal(l0, l1) and set(g0) == set(g1) and len(g0) > 0
    for m in g0:
        assert torch.equal(g0[m][0], g1[m][0]) and torch.equal(g0[m][1], g1[m][1]), m
