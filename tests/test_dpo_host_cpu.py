"""Host composition of the DPO loss: FusedDPO.compute_policy_loss — which pairs are formed, with which
coefficients, score / image-relation weights, CoPO pairs, the AncPO term (incl. its constant when CoPO is off), the
reward statistics — against the reference's DPOTrainer.compute_policy_loss run VERBATIM over a stub policy, for every
combination of the switches that reach it (hypothesis). The one-pair operation, which the product only has as a CUDA
kernel, is answered here by a torch restatement of the kernel's documented contract (include/opadpo_b200.h,
opadpo_dpo_pair_desc); the kernel itself is checked against the oracle on the GPU. Needs /root/reference (this container)."""
import os
import sys
from types import SimpleNamespace

import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
NAMES = ("standard_response", "original_generate_response", "AI_pseudo_response")


def torch_pair(pi_c, pi_r, ref_c, ref_r, s_c, s_r, p):
    """opadpo_op_dpo_pair's contract in torch: total = loss_coef*mean(loss) + anchor_coef*mean(anchor);
    out = [mean loss, mean anchor, masked-mean chosen reward, masked-mean rejected reward, #chosen, #rejected, 0, 0]."""
    rf = 0.0 if p["reference_free"] else 1.0
    dc, dr = pi_c - rf * ref_c, pi_r - rf * ref_r
    sc = s_c if s_c is not None else 1.0
    sr = s_r if s_r is not None else 1.0
    if p["divergence"] == 2:
        cap = 88.7228
        a = p["alpha_div"]
        u = (torch.exp(torch.clamp(-a * dr, max=cap)) - torch.exp(torch.clamp(-a * dc, max=cap))) / a
    else:
        u = sc * dc - sr * dr
        if p["divergence"] == 1:
            u = u - (F.softplus(dc) - F.softplus(dr))
    loss = -F.logsigmoid(p["beta"] * u) * (1 - p["label_smoothing"]) - F.logsigmoid(-p["beta"] * u) * p["label_smoothing"]
    cr, rr = p["beta"] * dc, p["beta"] * dr
    if p["anchor_mode"] == 1:
        anchor = -F.logsigmoid(p["anchor_sign"] * (cr - p["anchor_value"]))
    elif p["anchor_mode"] == 2:
        anchor = (cr - p["anchor_value"]) ** 2
    else:
        anchor = torch.zeros_like(cr)
    mc, mr = (ref_c != 0).float(), (ref_r != 0).float()
    out = torch.stack([loss.mean(), anchor.mean(), (cr * mc).sum() / mc.sum(), (rr * mr).sum() / mr.sum(), mc.sum(), mr.sum(),
                       torch.zeros(()), torch.zeros(())])
    return out[0] * p["loss_coef"] + out[1] * p["anchor_coef"], out.detach()


def test_compute_policy_loss_composition_vs_reference_all_switches():
    if not os.path.isdir("/root/reference"):
        pytest.skip("needs /root/reference")
    pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st
    from oracle.make_golden import install_stubs
    install_stubs()
    import opadpo.dpo_models.dpo_trainer as dpo_trainer
    from opa_dpo_b200.dpo import DPOArgs, FusedDPO
    b, R = 3, 7

    @settings(max_examples=300, deadline=None)
    @given(seed=st.integers(0, 10 ** 6), copo=st.booleans(), method=st.sampled_from(["random", "attention"]), ancpo=st.booleans(),
           mdpo=st.booleans(), anchor=st.sampled_from([0.0, 0.3]), detailed=st.booleans(), use_score=st.booleans(),
           use_rel=st.booleans(), ref_free=st.booleans(), ls=st.sampled_from([0.0, 0.1]),
           div=st.sampled_from(["reverse_kl", "js_divergence", "alpha_divergence"]), beta=st.sampled_from([0.1, 0.5]),
           coefs=st.tuples(st.sampled_from([1.0, 0.5]), st.sampled_from([1.0, 2.0]), st.sampled_from([0.2, 1.0]), st.sampled_from([1.0, 0.3])))
    def check(seed, copo, method, ancpo, mdpo, anchor, detailed, use_score, use_rel, ref_free, ls, div, beta, coefs):
        g = torch.Generator().manual_seed(seed)
        rnd = lambda: -torch.rand(b, R, generator=g) * 3
        valid = {n: (torch.rand(b, R, generator=g) > 0.25).float() for n in NAMES}
        for n in NAMES:
            valid[n][:, 0] = 1.0
        pol = {n + "_logprobs": rnd() * valid[n] for n in NAMES}
        ref = {"ref_base_" + n + "_logprobs": rnd() * valid[n] for n in NAMES}
        polm = {"mask_" + n + "_logprobs": rnd() * valid[n] for n in ("standard_response", "AI_pseudo_response")}
        refm = {"ref_mask_" + n + "_logprobs": rnd() * valid[n] for n in ("standard_response", "AI_pseudo_response")}
        scores = {}
        for n in ("original_generate_response", "AI_pseudo_response"):
            scores[n + "_scores"] = torch.tensor([1.0, 1.5, 2.0, 2.5])[torch.randint(0, 4, (b, R), generator=g)] * valid[n]
            scores[n + "_image_relations"] = torch.tensor([1.0, 3.0])[torch.randint(0, 2, (b, R), generator=g)] * valid[n]
        sp, ap, cc, ac = coefs
        args = SimpleNamespace(temperature=1.0, response_len=R, detailed_report=detailed, response_score=use_score,
                               response_image_relation=use_rel, CoPO=copo, CoPO_method=method, CoPO_coef=cc, AncPO=ancpo,
                               mDPO_anchor=mdpo, Anchor_value=anchor, Anchor_coef=ac, standard_pair_coef=sp, AI_pair_coef=ap)

        def stub_policy(**kw):      # AutoregressivePolicy.forward's contract: '<name>_logprobs' for every response kwarg
            table = {**pol, **polm}
            return {k + "_logprobs": table[k + "_logprobs"].clone().requires_grad_(True) for k in kw if "response" in k}
        stub_policy.train = lambda: None
        tr = object.__new__(dpo_trainer.DPOTrainer)
        tr.accelerator = SimpleNamespace(device=torch.device("cpu"), num_processes=1)
        tr.args, tr.tokenizer, tr.policy = args, SimpleNamespace(pad_token_id=0, eos_token_id=2), stub_policy
        tr.reference_free, tr.beta, tr.label_smoothing, tr.loss_type = ref_free, beta, ls, "sigmoid"
        tr.f_divergence_type = div
        tr.f_divergence_params = {"alpha_divergence_coef": 0.7}
        dummy = torch.zeros(b, 1)
        rollouts = dict(queries=dummy, queries_attn_masks=dummy, images=dummy, masked_query_attn_masks=dummy, masked_images=dummy,
                        **{n: valid[n] for n in NAMES}, **scores, **ref, **refm)
        loss_ref, stats_ref = tr.compute_policy_loss(rollouts)

        mine = DPOArgs(beta=beta, label_smoothing=ls, reference_free=ref_free, f_divergence_type=div, alpha_divergence_coef=0.7,
                       standard_pair_coef=sp, AI_pair_coef=ap, CoPO=copo, CoPO_coef=cc, AncPO=ancpo, mDPO_anchor=mdpo,
                       Anchor_value=anchor, Anchor_coef=ac, detailed_report=detailed, response_score=use_score,
                       response_image_relation=use_rel, response_len=R)
        loss, stats = FusedDPO(mine, pair_impl=torch_pair).compute_policy_loss(dict(pol), rollouts, dict(polm) if copo else None)
        assert abs(loss.item() - loss_ref.item()) <= 1e-5 * max(1.0, abs(loss_ref.item())), (loss.item(), loss_ref.item())
        assert set(stats) == set(stats_ref), set(stats) ^ set(stats_ref)     # loss/* 5, policy/* 12, logprobs/* 15
        for k, v in stats.items():
            assert abs(v.item() - stats_ref[k].item()) <= 1e-5 * max(1.0, abs(stats_ref[k].item())), (k, v.item(), stats_ref[k].item())
    check()
