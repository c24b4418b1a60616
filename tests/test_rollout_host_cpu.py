"""Row a5: B200DPOTrainer.rollout — which forwards run, on which inputs, under which names, the CoPO
masks and their random stream, the bf16 image cast, the score rows — against the reference's DPOTrainer.rollout
(opadpo/dpo_models/dpo_trainer.py:214-427) run VERBATIM over stub policies. The mask operations, which the product only has
as CUDA kernels, are answered by the CPU oracle (same host-side draws: one torch.randperm per image / row, in order).
What differs by design and is asserted as such: the rollouts stay on the device (no .cpu() parking) and the reference's
discarded policy forward (:336-338) is not run. Needs /root/reference."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
NAMES = ("standard_response", "original_generate_response", "AI_pseudo_response")


class OracleMaskOps:
    @staticmethod
    def mask_images(images, ratio, method, block=14):
        from oracle import collate_oracle as C
        n, Cc, H, W = images.shape
        total = H * W if method == "random" else (H // block) * (W // block)
        m = int(total * ratio)
        out = []
        for i in range(n):                                   # one permutation per image, in image order (dpo_trainer.py:295)
            idx = torch.randperm(total)[:m].numpy()
            x = C.mask_single_image(images[i].float().numpy(), idx, method, block, bf16_pixels=images.dtype == torch.bfloat16)
            out.append(torch.from_numpy(x).to(images.dtype))
        return torch.stack(out)

    @staticmethod
    def mask_percentage_per_row(n_rows, n_cols, ratio, device=None):
        from oracle import collate_oracle as C
        m = int(n_cols * ratio)
        idx = np.stack([torch.randperm(n_cols)[:m].numpy() for _ in range(n_rows)]) if n_rows else np.zeros((0, m), np.int64)
        return torch.from_numpy(C.attention_holes(n_cols, idx))


class StubPolicy:
    """AutoregressivePolicy.forward's contract with a deterministic function of its inputs standing in for the model, so
    that a forward on other images / masks / responses gives other numbers."""

    def __init__(self, salt):
        self.salt, self.calls = salt, []
        self.base_model = SimpleNamespace(engine=SimpleNamespace(device=torch.device("cpu")), flat_param=None)

    def eval(self):
        return self

    def train(self):
        return self

    def __call__(self, images, queries, queries_attn_masks, temperature=None, **kw):
        resp = {k: v for k, v in kw.items() if "response" in k and "_mask" not in k and "scores" not in k and "image_relations" not in k}
        self.calls.append(sorted(resp))
        base = images.float().mean(dim=(1, 2, 3)) + queries_attn_masks.float().mean(1) * 0.1 + self.salt
        out = {}
        for k, r in resp.items():
            valid = (r != 0).float()
            out[k + "_logprobs"] = -(r.float() * 1e-3 + base[:, None]) * valid
            out[k + "_entropies"] = (r.float() * 2e-3 + base[:, None]) * valid
        return out


@pytest.mark.parametrize("copo,method", [(False, "random"), (True, "random"), (True, "blockwise"), (True, "attention")])
@pytest.mark.parametrize("detailed", [False, True])
@pytest.mark.parametrize("bf16", [True, False])
def test_rollout_vs_reference_dpo_trainer(copo, method, detailed, bf16):
    if not os.path.isdir("/root/reference"):
        pytest.skip("needs /root/reference")
    from oracle.make_golden import install_stubs
    install_stubs()
    import opadpo.dpo_models.dpo_trainer as dpo_trainer
    from opa_dpo_b200.dpo import DPOArgs
    from opa_dpo_b200.trainer import B200DPOTrainer
    g = torch.Generator().manual_seed(1)
    b, Q, R = 2, 6, 5
    batches = []
    for _ in range(3):
        bt = {"images": torch.randn(b, 3, 28, 28, generator=g), "queries": torch.randint(3, 99, (b, Q), generator=g),
              "queries_attention_mask": torch.ones(b, Q, dtype=torch.long)}
        for n in NAMES:
            r = torch.randint(3, 99, (b, R), generator=g)
            r[0, 3:] = 0
            bt[n] = r
            bt[n + "_attention_mask"] = (r != 0).long()
        for n in ("original_generate_response", "AI_pseudo_response"):
            bt[n + "_scores"] = torch.rand(b, R, generator=g)
            bt[n + "_image_relations"] = torch.rand(b, R, generator=g)
        batches.append(bt)

    class Tok:
        pad_token_id, eos_token_id = 0, 2
        def encode(self, text, add_special_tokens=True):
            return [1577] if text == "?" else [1, 29973]

    ref_args = SimpleNamespace(bf16=bf16, fp16=False, CoPO=copo, CoPO_method=method, CoPO_mask_ratio=0.3, detailed_report=detailed,
                               response_score=True, response_image_relation=False, reward_model_per_device_batch_size=None,
                               temperature=0.8)
    tr = object.__new__(dpo_trainer.DPOTrainer)
    tr.args, tr.tokenizer = ref_args, Tok()
    tr.accelerator = SimpleNamespace(device=torch.device("cpu"), is_main_process=False)
    tr.policy, tr.ref_policy = StubPolicy(0.5), StubPolicy(0.25)
    torch.manual_seed(11)
    want = tr.rollout([dict(bt) for bt in batches])

    args = DPOArgs(CoPO=copo, CoPO_method=method, CoPO_mask_ratio=0.3, detailed_report=detailed, response_score=True,
                   response_image_relation=False, temperature=0.8, bf16=bf16, response_len=R)
    pol, ref = StubPolicy(0.5), StubPolicy(0.25)
    mine = B200DPOTrainer(args, pol, ref, optimizer=None, mask_ops=OracleMaskOps(), loss=object(), device=torch.device("cpu"))
    torch.manual_seed(11)
    got = mine.rollout([dict(bt) for bt in batches])

    assert set(got) == set(want), (sorted(set(got) ^ set(want)))
    for k in want:
        assert got[k].dtype == want[k].dtype and got[k].shape == want[k].shape, k
        if not bf16 and (k == "masked_images" or k.startswith("ref_mask_")):
            # float32 images: the CoPO fill value is a float mean whose summation order is not part of torch's contract
            # (tolerance 2^-20 * mean|x|, as in the collator tests); everything that does not depend on it is exact
            assert torch.allclose(got[k], want[k], atol=2e-6, rtol=0), k
            assert torch.equal(got[k] == 0, want[k] == 0), k
        else:
            assert torch.equal(got[k], want[k]), k
    # by design: the reference also runs (and throws away) a policy forward per batch; the mirror does not
    assert len(tr.policy.calls) == 3 and len(pol.calls) == 0
    assert ref.calls == tr.ref_policy.calls


@pytest.mark.parametrize("top_k,top_p,temperature,suppress", [(0, 1.0, 1.0, None), (30, 0.95, 1.2, None), (0, 0.8, 0.7, [2]),
                                                              (300, 1.0, 1.0, [2, 7]), (1, 0.5, 1.0, None), (1000, 0.3, 2.0, None)])
def test_oracle_warpers_equal_transformers_logits_processors(top_k, top_p, temperature, suppress):
    """oracle.hf_warp (what the sampling kernel is checked against) == the installed transformers' own processor classes in
    the order generate() applies them for the reference's arguments (online_generator.py:292-309): suppress_tokens, then
    temperature, top-k, top-p."""
    from transformers.generation.logits_process import (SuppressTokensLogitsProcessor, TemperatureLogitsWarper,
                                                        TopKLogitsWarper, TopPLogitsWarper)
    from oracle import llava_oracle as O
    g = torch.Generator().manual_seed(top_k * 7 + int(top_p * 100))
    logits = torch.randn(5, 512, generator=g) * 3.0
    logits[1, 10:14] = logits[1, 10]           # ties, also at the top-k boundary of some rows
    ids = torch.zeros(5, 1, dtype=torch.long)
    want = logits.clone()
    if suppress:
        want = SuppressTokensLogitsProcessor(suppress, device="cpu")(ids, want)
    if temperature != 1.0:
        want = TemperatureLogitsWarper(temperature)(ids, want)
    if top_k:
        want = TopKLogitsWarper(top_k)(ids, want)
    if top_p < 1.0:
        want = TopPLogitsWarper(top_p)(ids, want)
    got = O.hf_warp(logits, temperature, top_k, top_p, suppress)
    assert torch.equal(torch.isinf(got), torch.isinf(want))
    keep = ~torch.isinf(want)
    assert torch.allclose(got[keep], want[keep], rtol=1e-6, atol=0)
    assert keep.any(-1).all()


class LearnableStubPolicy(StubPolicy):
    """StubPolicy whose log-probs depend on one trainable scalar, so step_with_rollouts can differentiate the loss."""

    def __init__(self, salt):
        super().__init__(salt)
        self.theta = torch.nn.Parameter(torch.tensor(0.3))

    def __call__(self, *a, **kw):
        out = super().__call__(*a, **kw)
        return {k: (v * (1.0 + self.theta) if k.endswith("_logprobs") else v) for k, v in out.items()}


class StubOptimizer:
    def __init__(self, params, lr):
        self.param_groups, self.params, self.steps = [dict(lr=lr)], params, 0

    def step(self, scale=1.0):
        self.steps += 1
        return torch.stack([p.grad.detach().norm() for p in self.params]).norm()

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            p.grad = None


@pytest.mark.parametrize("copo", [False, True])
def test_step_returns_reference_record_step_stats(copo):
    """Row a6: B200DPOTrainer.step = rollout -> step_with_rollouts -> scheduler -> record_step_stats, and what it returns /
    logs is the reference's DPOTrainer.record_step_stats (dpo_trainer.py:804-835) run VERBATIM on the same train stats and
    rollouts: same keys ('dpo/loss-grad_norm', 'logprobs/…', 'objective/lr', 'objective/<entropies>'), python floats."""
    if not os.path.isdir("/root/reference"):
        pytest.skip("needs /root/reference")
    from oracle.make_golden import install_stubs
    install_stubs()
    import opadpo.dpo_models.dpo_trainer as dpo_trainer
    from opa_dpo_b200.dpo import DPOArgs, FusedDPO
    from opa_dpo_b200.trainer import B200DPOTrainer
    from test_dpo_host_cpu import torch_pair
    g = torch.Generator().manual_seed(5)
    b, Q, R = 2, 6, 5
    batches = []
    for _ in range(4):
        bt = {"images": torch.randn(b, 3, 28, 28, generator=g), "queries": torch.randint(3, 99, (b, Q), generator=g),
              "queries_attention_mask": torch.ones(b, Q, dtype=torch.long)}
        for n in NAMES:
            r = torch.randint(3, 99, (b, R), generator=g)
            r[0, 3:] = 0
            bt[n] = r
            bt[n + "_attention_mask"] = (r != 0).long()
        batches.append(bt)
    args = DPOArgs(CoPO=copo, CoPO_method="random", CoPO_mask_ratio=0.3, AncPO=True, temperature=1.0, bf16=False, response_len=R,
                   rollout_accumulation_steps=4, step_per_device_batch_size=2, gradient_accumulation_steps=2, noptepochs=1)
    pol, ref = LearnableStubPolicy(0.5), StubPolicy(0.25)
    opt = StubOptimizer([pol.theta], lr=3e-6)
    logged = []
    tr = B200DPOTrainer(args, pol, ref, opt, mask_ops=OracleMaskOps(), loss=FusedDPO(args, pair_impl=torch_pair),
                        device=torch.device("cpu"))
    tr.log_fn = lambda stats, step: logged.append((step, dict(stats)))
    torch.manual_seed(2)
    got = tr.step(iter(batches), step_idx=7)
    assert opt.steps == 2 and tr.last_train_stats["loss/grad_norm"].shape == (2,)     # 8 rows / 2 per micro-batch / gas 2
    assert logged == [(7, got)]
    assert all(isinstance(v, float) for v in got.values())

    torch.manual_seed(2)
    rollouts = tr.rollout(batches)           # the same rollouts again (the mask stream is re-seeded)
    rt = object.__new__(dpo_trainer.DPOTrainer)
    rt.optimizer = opt
    ref_logged = []
    rt.accelerator = SimpleNamespace(is_main_process=True, log=lambda stats, step: ref_logged.append((step, stats)))
    want = rt.record_step_stats(train_stats=tr.last_train_stats, rollouts=rollouts, step_idx=7)
    assert set(got) == set(want), sorted(set(got) ^ set(want))
    assert {"objective/lr", "dpo/loss-grad_norm", "dpo/loss-AncPO", "dpo/policy-AI_mask_gap_mean", "logprobs/AI_mask_logprobs_max",
            "objective/ref_base_standard_response_entropies"} <= set(got)
    assert len(got) == 1 + (5 if copo else 3) + 33      # lr, entropy keys of the rollouts, 32 policy-loss stats + grad_norm
    for k, v in want.items():
        assert got[k] == pytest.approx(v, rel=1e-6, abs=1e-9), k
    assert ref_logged[0][0] == 7 and set(ref_logged[0][1]) == set(got)


class _GenTok:
    """Tokenizer stub for the generation-stage tests: ids decode to fixed pieces, 0 / 1 / 2 are special."""
    pad_token_id, eos_token_id = 0, 2
    PIECES = {50: "USER:  \n", 51: " ASSISTANT:"}

    def encode(self, text, add_special_tokens=True):
        return [1577] if text == "?" else [1, 29973]

    def batch_decode(self, ids, skip_special_tokens=True, clean_up_tokenization_spaces=True):
        return ["".join(self.PIECES.get(int(t), f"<{int(t)}>") for t in row if int(t) not in (0, 1, 2)) for row in ids]


def _gen_batches(n_batches, b=2, Q=8, seed=3):
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(n_batches):
        q = torch.randint(3, 40, (b, Q), generator=g)
        q[:, 0], q[:, 1], q[:, 2], q[:, -1] = 0, 50, -200, 51
        out.append({"images": torch.randn(b, 3, 4, 4, generator=g), "images_path": [f"img{i}_{j}.jpg" for j in range(b)],
                    "images_url": [f"http://x/{i}/{j}" for j in range(b)], "images_bytes": [bytes([i, j]) for j in range(b)],
                    "queries": q, "query_attn_masks": q != 0, "standard_responses": torch.randint(3, 40, (b, 5), generator=g)})
    return out


def _stub_generate(inputs, images, max_new_tokens, **kw):
    """A deterministic stand-in for model.generate: ids from the prompt and the (possibly bf16-rounded) image, with stop
    tokens sprinkled in so the truncation rule matters."""
    base = (inputs.clamp(min=0).sum(1) + (images.float().sum((1, 2, 3)) * 64).round().long()) % 17
    new = (base[:, None] + torch.arange(max_new_tokens)[None, :] * 3) % 23 + 3
    new[0, 3] = 2
    new[1, 4] = 1577
    return torch.cat([inputs, new], 1)


def test_generation_stage_rollout_and_loop_vs_reference(tmp_path, monkeypatch):
    """Rows a10 / f-4: B200OnlineGenerator.rollout / step / generate against the reference's Online_Generator.rollout
    (online_generator.py:235-378) and Generator.generate / find_max_step / record_step_stats (generator.py:99-175) run
    VERBATIM over the same stub model: generate() arguments, bf16 image cast, truncation, decoded texts, the query cut, the
    record columns, the files written, the skipped / repeated steps of a resumed run."""
    if not os.path.isdir("/root/reference"):
        pytest.skip("needs /root/reference")
    from oracle.make_golden import install_stubs
    from oracle import llava_oracle as O
    install_stubs()
    import opadpo.generator_models.online_generator as OG
    from opa_dpo_b200.generation import B200OnlineGenerator, find_max_step
    monkeypatch.setenv("RANK", "0")
    tok = _GenTok()

    def make_args(out):
        return SimpleNamespace(bf16=True, fp16=False, response_len=6, suppress_eos_at_generation=True, top_p=0.95, top_k=30,
                               temperature=1.1, rollout_accumulation_steps=2, rollout_batch_size=4, total_epochs=1, sample_num=16,
                               output_dir=str(out))
    seen_ref, seen_mine = [], []

    def revise(responses, images_url, text_queries, text_responses, text_standard_responses, *_rest):
        return {"Pseudo_response": [r + "|p" for r in text_responses], "Pseudo_response_ids": None,
                "Generated_response": [r + "|g" for r in text_responses], "Generated_response_ids": None,
                "report_json": [{"url": u, "q": q} for u, q in zip(images_url, text_queries)]}

    class RefPolicy:
        def eval(self):
            return self

        def generate(self, **kw):
            seen_ref.append({k: v for k, v in kw.items() if k not in ("inputs", "images", "attention_mask")} | {"dtype": kw["images"].dtype})
            return _stub_generate(kw["inputs"], kw["images"], kw["max_new_tokens"])

    class MyModel:
        def generate(self, **kw):
            seen_mine.append({k: v for k, v in kw.items() if k not in ("inputs", "images", "attention_mask")} | {"dtype": kw["images"].dtype})
            return _stub_generate(kw["inputs"], kw["images"], kw["max_new_tokens"])

    # ---- one rollout call
    batches = _gen_batches(2)
    ref = object.__new__(OG.Online_Generator)
    ref.args, ref.tokenizer, ref.policy = make_args(tmp_path / "ref"), tok, RefPolicy()
    ref.accelerator = SimpleNamespace(device=torch.device("cpu"), is_main_process=False)
    ref.process_responses_coco_gpt4v = revise
    want = ref.rollout([dict(b) for b in batches], 1)
    mine = B200OnlineGenerator(make_args(tmp_path / "mine"), MyModel(), tok, revise=revise, truncate=O.truncate_after_eos_with_padding)
    got = mine.rollout([dict(b) for b in batches], 1)
    assert got == want and list(got) == list(want)          # same columns in the same order, same records
    assert "USER" not in got["query"][0] and got["query"][0].startswith("<image>\n<")
    for a, b in zip(seen_mine, seen_ref):
        assert a["dtype"] == b["dtype"] == torch.bfloat16
        for k in ("do_sample", "max_new_tokens", "pad_token_id", "suppress_tokens", "top_p", "top_k", "temperature", "num_return_sequences"):
            assert a[k] == b[k], k
        assert a["synced_gpus"] is False                     # by design: ranks decode independently
    # without a revision service the AI_* columns are None and the rest is unchanged
    bare = B200OnlineGenerator(make_args(tmp_path / "bare"), MyModel(), tok, truncate=O.truncate_after_eos_with_padding).rollout(batches, 1)
    assert bare["AI_json_report"] == [None] * 4 and bare["original_generate_response"] == want["original_generate_response"]

    # ---- the loop: 3 steps of 2 batches (max_step = 16 // 4 = 4 stops before step 4), then a restart that finds step 3 recorded
    import itertools

    class RefLoop(OG.Online_Generator):
        def get_train_dataloader(self):
            return itertools.cycle([dict(b) for b in _gen_batches(5, seed=9)])

    ref2 = object.__new__(RefLoop)
    ref2.args, ref2.tokenizer, ref2.policy = make_args(tmp_path / "ref"), tok, RefPolicy()
    ref2.accelerator = SimpleNamespace(device=torch.device("cpu"), is_main_process=False)
    ref2.process_responses_coco_gpt4v = revise
    ref2.train_dataset, ref2.log_history = list(range(20)), []
    hist_ref = ref2.generate()
    mine2 = B200OnlineGenerator(make_args(tmp_path / "mine"), MyModel(), tok, revise=revise, truncate=O.truncate_after_eos_with_padding)
    hist = mine2.generate(20, lambda: itertools.cycle([dict(b) for b in _gen_batches(5, seed=9)]))
    assert hist == hist_ref == [None, None, None]
    files = sorted(os.listdir(tmp_path / "ref" / "rollouts"))
    assert files == sorted(os.listdir(tmp_path / "mine" / "rollouts")) == ["step1_rank0.json", "step2_rank0.json", "step3_rank0.json"]
    for f in files:
        assert (tmp_path / "ref" / "rollouts" / f).read_bytes() == (tmp_path / "mine" / "rollouts" / f).read_bytes(), f
    assert find_max_step(str(tmp_path / "mine")) == ref2.find_max_step() == 3
    # restart: steps 1-2 are skipped with their batches consumed, step 3 is generated again from the same batches -> same file
    before = (tmp_path / "mine" / "rollouts" / "step3_rank0.json").read_bytes()
    n_calls = len(seen_mine)
    mine3 = B200OnlineGenerator(make_args(tmp_path / "mine"), MyModel(), tok, revise=revise, truncate=O.truncate_after_eos_with_padding)
    assert mine3.generate(20, lambda: itertools.cycle([dict(b) for b in _gen_batches(5, seed=9)])) == [None]
    assert len(seen_mine) == n_calls + 2 and (tmp_path / "mine" / "rollouts" / "step3_rank0.json").read_bytes() == before
    assert find_max_step(str(tmp_path / "fresh")) == 0 and os.path.isdir(tmp_path / "fresh" / "rollouts")
