"""Row a1, host half (SURVEY.md §8b "Python surface kept verbatim"): opa_dpo_b200.policy.AutoregressivePolicy.forward — the
discovery of response kwargs by name, the split of a CoPO 'attention' mask, the stacking order, the output names — against
the outputs of the REFERENCE's rl_models.AutoregressivePolicy.forward (tests/golden/opadpo_tiny.npz, produced by
oracle/make_golden.py). The engine call underneath (`base_model.policy_logprobs`), which the product only has on the GPU,
is answered here by the CPU oracle; the engine itself is compared with the oracle in the GPU tests."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import llava_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "opadpo_tiny.npz")
NAMES3 = ["standard_response", "original_generate_response", "AI_pseudo_response"]


class OracleBackedModel:
    """Stands where B200LlavaModel stands (`set_adapter`, `config.use_cache`, `policy_logprobs`)."""

    def __init__(self, cfg, w, adapters):
        self.cfg, self.w, self.adapters = cfg, w, adapters
        self.config = SimpleNamespace(use_cache=True)
        self.active, self.calls = None, []

    def set_adapter(self, name):
        self.active = name

    def policy_logprobs(self, images, queries, qmask, responses, temperature, img_mask=None):
        self.calls.append((self.active, len(responses), None if img_mask is None else tuple(img_mask.shape)))
        names = [f"response_{i}" for i in range(len(responses))]
        qam = qmask if img_mask is None else torch.cat([img_mask, qmask], 1)
        with torch.no_grad():
            out = O.policy_forward(self.cfg, self.w, self.adapters[self.active], images, queries, qam, dict(zip(names, responses)),
                                   temperature=temperature)
        return (torch.stack([out[n + "_logprobs"] for n in names]), torch.stack([out[n + "_entropies"] for n in names]))


def test_policy_forward_host_logic_vs_reference_outputs():
    from opa_dpo_b200.policy import AutoregressivePolicy, make_policy_with_base_model
    gold = {k: torch.from_numpy(v) for k, v in np.load(GOLD).items()}
    cfg = O.OracleConfig()
    w = O.init_weights(cfg)
    model = OracleBackedModel(cfg, w, {"lora_policy": O.init_lora(cfg, 11), "lora_ref_policy": O.init_lora(cfg, 12)})
    images, q, qm = gold["in.images"], gold["in.queries"], gold["in.qmask"]
    resp = {n: gold["in." + n] for n in NAMES3}
    R = resp[NAMES3[0]].shape[1]
    args = SimpleNamespace(temperature=0.7, response_len=R)
    tok = SimpleNamespace(pad_token_id=0, eos_token_id=2)
    policy = make_policy_with_base_model(args, model, tok, adapter_name="lora_policy")
    ref_policy = AutoregressivePolicy(args, model, tok, adapter_name="lora_ref_policy")
    masked_qm = torch.cat([gold["in.img_mask"], qm], 1)
    mresp = {"mask_standard_response": resp["standard_response"], "mask_AI_pseudo_response": resp["AI_pseudo_response"]}
    # what the reference's callers pass along and forward() must ignore: mode=, *_attention_mask, *_scores, *_image_relations
    noise = {"mode": "policy", "standard_response_attention_mask": resp["standard_response"] != 0,
             "AI_pseudo_response_scores": gold["in.AI_pseudo_response_scores"],
             "AI_pseudo_response_image_relations": gold["in.AI_pseudo_response_image_relations"]}
    runs = {"pol": policy(images, q, qm, **resp, **noise),                       # temperature defaults to args.temperature
            "ref": ref_policy(images=images, queries=q, queries_attn_masks=qm, temperature=0.7, **resp),
            "polm": policy(images, q, masked_qm, temperature=0.7, **mresp),
            "refm": ref_policy(images, q, masked_qm, **mresp)}
    n = 0
    for tag, out in runs.items():
        want = {k[len(tag) + 1:]: v for k, v in gold.items() if k.startswith(tag + ".")}
        assert set(out) == set(want), (tag, sorted(out), sorted(want))
        for k, v in out.items():
            assert v.shape == want[k].shape and (v - want[k]).abs().max().item() <= 2e-4, (tag, k)
            n += 1
    assert n == 20
    assert model.calls == [("lora_policy", 3, None), ("lora_ref_policy", 3, None), ("lora_policy", 2, (2, cfg.n_patches)),
                           ("lora_ref_policy", 2, (2, cfg.n_patches))]
    assert model.config.use_cache is False                                        # rl_models.py:86
    with pytest.raises(ValueError):
        policy(images, q, qm, standard_response=resp["standard_response"][:, :-1])    # not padded to args.response_len
    with pytest.raises(ValueError):
        policy(images, q, qm, mode="policy")                                        # no response tensors at all


def test_respond_arguments_match_reference_call():
    """Policy._respond (rl_models.py:146-191): the generate() call it makes — sampling at `temperature` with top_k = 0,
    top_p = 1, max_new_tokens = response_len — and the slice it returns."""
    from opa_dpo_b200.policy import AutoregressivePolicy
    seen = {}

    class Model:
        config = SimpleNamespace(use_cache=False, cache_shape=None)

        def set_adapter(self, name):
            seen["adapter"] = name

        def get_vision_tower(self):
            return SimpleNamespace(num_patches=576)

        def generate(self, **kw):
            seen.update(kw)
            return torch.cat([kw["inputs"], torch.full((kw["inputs"].size(0), kw["max_new_tokens"]), 7)], 1)

    m = Model()
    pol = AutoregressivePolicy(SimpleNamespace(temperature=0.9, response_len=5), m, SimpleNamespace(pad_token_id=0, eos_token_id=2), "lora_policy")
    pol.eval()
    q = torch.ones(2, 4, dtype=torch.long)
    out = pol.respond(q, q != 0, images=torch.zeros(2, 3, 2, 2))
    assert out["responses"].shape == (2, 5) and (out["responses"] == 7).all()
    assert seen["adapter"] == "lora_policy" and seen["do_sample"] is True and seen["top_k"] == 0 and seen["top_p"] == 1.0
    assert seen["temperature"] == 0.9 and seen["max_new_tokens"] == 5 and seen["pad_token_id"] == 0
    assert m.config.use_cache is True and m.config.cache_shape == (4 + 5 + 576,)
    pol.train()
    with pytest.raises(AssertionError):
        pol.respond(q, q != 0)


def test_generate_num_return_sequences_and_argument_errors():
    """B200LlavaModel.generate (the `model.generate` surface of online_generator.py:292-310 / rl_models.py:166-188):
    num_return_sequences repeats every prompt row in place (transformers' repeat_interleave expansion), greedy with more
    than one return sequence is rejected as transformers rejects it, a pad id other than the engine's is refused, and every
    call gets a new RNG seed. The engine is a recording stub here (the product's engine needs a GPU)."""
    from opa_dpo_b200.engine import EngineConfig
    from opa_dpo_b200.policy import AutoregressivePolicy, B200LlavaModel
    from opa_dpo_b200._lib import OpadpoError
    calls = []

    class Eng:
        def generate(self, aid, images, queries, qmask, max_new, **kw):
            calls.append(dict(aid=aid, images=images, queries=queries, qmask=qmask, max_new=max_new, **kw))
            return torch.cat([queries, torch.full((queries.size(0), max_new), 7)], 1)

    model = object.__new__(B200LlavaModel)
    torch.nn.Module.__init__(model)
    model.cfg = EngineConfig(v_image=28, v_patch=14)
    model.engine, model.active_adapter = Eng(), "lora_policy"
    q = torch.tensor([[0, 5, -200, 6], [7, 8, -200, 9]])
    img = torch.arange(2.0).view(2, 1, 1, 1).expand(2, 3, 28, 28)
    out = model.generate(inputs=q, images=img, attention_mask=q != 0, do_sample=True, max_new_tokens=3, pad_token_id=0,
                         top_k=30, top_p=0.95, temperature=1.2, num_return_sequences=3, synced_gpus=True)
    assert out.shape == (6, 7)
    c = calls[-1]
    assert torch.equal(c["queries"], q[[0, 0, 0, 1, 1, 1]]) and torch.equal(c["qmask"], (q != 0)[[0, 0, 0, 1, 1, 1]])
    assert torch.equal(c["images"][:, 0, 0, 0], torch.tensor([0.0, 0, 0, 1, 1, 1]))
    assert c["do_sample"] is True and c["top_k"] == 30 and c["top_p"] == 0.95 and c["temperature"] == 1.2 and c["max_new"] == 3
    model.generate(inputs=q, images=img, attention_mask=q != 0, do_sample=True, max_new_tokens=3)
    assert calls[-1]["seed"] != c["seed"] and calls[-1]["queries"].shape == (2, 4)
    with pytest.raises(ValueError, match="num_return_sequences"):
        model.generate(inputs=q, images=img, attention_mask=q != 0, do_sample=False, max_new_tokens=3, num_return_sequences=2)
    with pytest.raises(OpadpoError, match="pad_id"):
        model.generate(inputs=q, images=img, attention_mask=q != 0, max_new_tokens=3, pad_token_id=32000)

    # Policy._respond passes num_return_sequences through (rl_models.py:166-183) and returns [b * n, response_len]
    pol = AutoregressivePolicy(SimpleNamespace(temperature=1.0, response_len=3, suppress_eos_at_generation=True), model,
                               SimpleNamespace(pad_token_id=0, eos_token_id=2), "lora_policy")
    pol.eval()
    model.get_vision_tower = lambda: SimpleNamespace(num_patches=4)
    model.config = SimpleNamespace(use_cache=False, cache_shape=None)
    r = pol.respond(q, q != 0, images=img, num_return_sequences=2)["responses"]
    assert r.shape == (4, 3) and calls[-1]["suppress_tokens"] == [2] and calls[-1]["queries"].shape == (4, 4)


def test_call_shapes_are_checked_before_pointers_cross_the_abi():
    from opa_dpo_b200.engine import EngineConfig, check_call_shapes
    from opa_dpo_b200._lib import OpadpoError
    cfg = EngineConfig(v_image=28, v_patch=14)
    b, Q, R = 2, 5, 4
    img, q, qm = torch.zeros(b, 3, 28, 28), torch.zeros(b, Q, dtype=torch.long), torch.ones(b, Q, dtype=torch.bool)
    resp = [torch.zeros(b, R, dtype=torch.long)] * 2
    check_call_shapes(cfg, img, q, qm, resp, torch.ones(b, 4))
    check_call_shapes(cfg, img, q, qm, what="generate")
    bad = [dict(images=torch.zeros(b, 3, 14, 14)), dict(images=torch.zeros(1, 3, 28, 28)), dict(qmask=torch.ones(b, Q + 4)),
           dict(responses=[resp[0], torch.zeros(b, R + 1, dtype=torch.long)]), dict(responses=[torch.zeros(1, R, dtype=torch.long)]),
           dict(img_mask=torch.ones(b, 5)), dict(queries=torch.zeros(Q, dtype=torch.long)), dict(queries=torch.zeros(0, Q, dtype=torch.long))]
    for kw in bad:
        args = dict(images=img, queries=q, qmask=qm, responses=resp, img_mask=None)
        args.update(kw)
        with pytest.raises(OpadpoError):
            check_call_shapes(cfg, **args)
