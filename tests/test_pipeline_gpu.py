"""The drop-in chain end to end on the GPU, at tiny dimensions: decoded uint8 images + GPT-4V-style reports ->
B200DataCollator (preprocess, report assembly, score rows) -> B200DPOTrainer.step (reference-adapter rollout with CoPO
attention holes, compute_policy_loss with AncPO and score / image-relation weights, LoRA backward, clip + AdamW, cosine
schedule) -> save_checkpoint / load_checkpoint in the reference's layout. Every stage has its own parity test; this one
checks that they compose through the reference's interfaces."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _setup():
    from types import SimpleNamespace
    from oracle import collate_oracle as C
    from oracle import llava_oracle as O
    from opa_dpo_b200 import EngineConfig
    from opa_dpo_b200.collate import B200DataCollator
    from opa_dpo_b200.dpo import DPOArgs
    from opa_dpo_b200.optim import B200AdamW, cosine_schedule, get_last_checkpoint, load_checkpoint, save_checkpoint
    from opa_dpo_b200.policy import AutoregressivePolicy, B200LlavaModel
    from opa_dpo_b200.trainer import B200DPOTrainer
    oc = O.OracleConfig()
    ec = EngineConfig(**oc.to_dict())
    Q, R = 32, 48

    class Tok(C.StubTokenizer):            # ids inside the tiny vocabulary; "IMG" stays the placeholder id the collator maps to -200
        def encode_one(self, text):
            out = [self.bos_token_id]
            for wd in text.split():
                out.append(C.IMAGE_PLACEHOLDER_ID if wd == "IMG" else 3 + int(hashlib.md5(wd.encode()).hexdigest()[:8], 16) % (oc.vocab - 3))
            return out

    rs = np.random.RandomState(4)
    words = lambda n, p: " ".join(f"{p}{rs.randint(0, 999)}" for _ in range(n))
    instances = []
    for i in range(4):
        report, orig, ai = {}, [], []
        for k in range(3):
            cop, rew = words(3 + k, "c"), words(4 + k, "r")
            report[f"Sentence {k + 1}"] = {"copied content": cop, "rewritten content": rew, "score": 1 + (i + k) % 4,
                                           "error type": ["correct", "image_recognition_error"][(i + k) % 2]}
            orig.append(cop)
            ai.append(rew)
        report["Added"] = {"rewritten content": words(2, "a")}
        instances.append({"queries": "SYS " + words(4 + i, "q") + " IMG " + words(3, "q"),
                          "images": rs.randint(0, 256, (40 + 7 * i, 61 - 5 * i, 3)).astype(np.uint8),
                          "standard_response": words(9 + i, "s"), "original_generate_response": " ".join(orig),
                          "AI_pseudo_response": " ".join(ai), "AI_json_report": json.dumps(report)})
    collator = B200DataCollator(Tok(), Q, R, detailed_report=True, image_size=oc.v_image, pad_to_square=True)
    batches = [collator(instances[:2]), collator(instances[2:])]
    for b in batches:
        assert b["images"].shape == (2, 3, oc.v_image, oc.v_image) and b["images"].dtype == torch.float32
        assert ((b["queries"] == -200).sum(1) == 1).all()
        assert b["AI_pseudo_response_scores"].dtype == torch.float32 and (b["AI_pseudo_response_scores"] > 0).any()
        assert int(b["queries"].max()) < oc.vocab and int(b["AI_pseudo_response"].max()) < oc.vocab
    want = C.clip_preprocess(instances[1]["images"], oc.v_image, True)
    assert np.array_equal(batches[0]["images"][1].cpu().numpy(), want)

    args = DPOArgs(beta=0.1, CoPO=True, CoPO_method="attention", CoPO_mask_ratio=0.3, CoPO_coef=0.2, AncPO=True, mDPO_anchor=True,
                   detailed_report=True, response_score=True, response_image_relation=True, response_len=R,
                   rollout_accumulation_steps=2, step_per_device_batch_size=2, gradient_accumulation_steps=1, noptepochs=1)
    tok = SimpleNamespace(pad_token_id=0, eos_token_id=2)
    w = O.init_weights(oc)

    def make():
        model = B200LlavaModel(ec)
        sd = dict(w)
        sd[O.VT + "embeddings.patch_embedding.weight"] = w[O.VT + "embeddings.patch_embedding.weight"].reshape(oc.v_d, -1)
        model.engine.load_state_dict(sd)
        for name, seed in (("lora_policy", 11), ("lora_ref_policy", 12)):
            lo = O.init_lora(oc, seed)
            model.load_adapter_state(name, {f"base_model.model.{m}.lora_{x}.weight": t for m, ab in lo.items() for x, t in zip("AB", ab)})
        model.make_trainable("lora_policy")
        return model

    return SimpleNamespace(**locals())


def test_collator_trainer_checkpoint_pipeline(tmp_path):
    S = _setup()
    oc, C, args, tok, make, batches, instances = S.oc, S.C, S.args, S.tok, S.make, S.batches, S.instances
    AutoregressivePolicy, B200AdamW, B200DPOTrainer, cosine_schedule = S.AutoregressivePolicy, S.B200AdamW, S.B200DPOTrainer, S.cosine_schedule
    get_last_checkpoint, load_checkpoint, save_checkpoint = S.get_last_checkpoint, S.load_checkpoint, S.save_checkpoint
    model = make()
    model.engine.set_option("merge_frozen", 1)
    policy = AutoregressivePolicy(args, model, tok, "lora_policy")
    ref_policy = AutoregressivePolicy(args, model, tok, "lora_ref_policy")
    opt = B200AdamW(model, lr=1e-3)
    sch = cosine_schedule(opt, 0, 10)     # no warm-up: the first step already has lr > 0
    trainer = B200DPOTrainer(args, policy, ref_policy, opt, sch)
    before = model.engine.buffers()["master"].clone()
    torch.manual_seed(3)
    logged = trainer.step(iter(batches), step_idx=1)          # what the reference's step returns: the logged floats
    torch.cuda.synchronize()
    import math
    assert {"objective/lr", "dpo/loss-grad_norm", "dpo/policy-stand_gen_gap_mean", "logprobs/original_logprobs",
            "objective/ref_base_standard_response_entropies"} <= set(logged)
    assert all(isinstance(v, float) and math.isfinite(v) for v in logged.values()), logged
    stats = trainer.last_train_stats                           # per-optimizer-step tensors of step_with_rollouts
    assert abs(logged["dpo/loss-grad_norm"] - stats["loss/grad_norm"].mean().item()) <= 1e-6 * abs(logged["dpo/loss-grad_norm"])
    assert opt.step_count == 2 and stats["loss/grad_norm"].shape == (2,)
    for k, v in stats.items():
        assert torch.isfinite(v).all(), k
    assert (stats["loss/grad_norm"] > 0).all() and (stats["loss/stand_mask"] > 0).all() and (stats["loss/AncPO"] > 0).all()
    n_train = model.engine.buffers()["n_train"]
    moved = (model.engine.buffers()["master"][:n_train] - before[:n_train]).abs().max().item()
    assert 1e-5 < moved < 1e-2, moved           # two Adam steps at lr <= 1e-3
    assert sch.last_epoch == 1

    ckpt = save_checkpoint(model, opt, sch, str(tmp_path), 1)
    assert get_last_checkpoint(str(tmp_path)) == ckpt
    assert os.path.exists(os.path.join(ckpt, "adapter_model", "lora_policy", "adapter_model.bin"))
    saved = model.adapter_state_dict("lora_policy")
    model2 = make()
    opt2 = B200AdamW(model2, lr=1e-3)
    sch2 = cosine_schedule(opt2, 0, 10)
    assert load_checkpoint(model2, opt2, sch2, ckpt) == 1 and opt2.step_count == 2
    again = model2.adapter_state_dict("lora_policy")
    for k, v in saved.items():
        assert torch.equal(again[k], v), k
    model.engine.close()
    model2.engine.close()


def test_trainer_train_outer_loop_on_device(tmp_path):
    """B200DPOTrainer.train (RLTrainer.train's loop, pinned on the CPU in tests/test_train_loop_cpu.py) driving real steps:
    8 samples, rollout_batch_size 4 -> 2 outer steps of 2 batches each, a checkpoint at the start of each, and a resumed
    run that skips what was trained."""
    import dataclasses
    import itertools
    S = _setup()
    args = dataclasses.replace(S.args, total_epochs=1, rollout_batch_size=4, save_steps=1, max_step=10, warmup_steps=0,
                               output_dir=str(tmp_path))
    model = S.make()
    policy = S.AutoregressivePolicy(args, model, S.tok, "lora_policy")
    ref_policy = S.AutoregressivePolicy(args, model, S.tok, "lora_ref_policy")
    trainer = S.B200DPOTrainer(args, policy, ref_policy, S.B200AdamW(model, lr=1e-3))
    torch.manual_seed(3)
    hist = trainer.train(8, lambda: itertools.cycle(S.batches))
    torch.cuda.synchronize()
    assert len(hist) == 2 and trainer.optimizer.step_count == 4 and trainer.lr_scheduler.last_epoch == 2
    for step in (1, 2):
        assert os.path.exists(os.path.join(str(tmp_path), f"checkpoint-{step}", "adapter_model", "lora_policy", "adapter_model.bin"))
    assert S.get_last_checkpoint(str(tmp_path)) == os.path.join(str(tmp_path), "checkpoint-2")
    # resume from checkpoint-2 (written before step 2 ran): step 1 is skipped, step 2 runs again, nothing is saved at 2
    model2 = S.make()
    trainer2 = S.B200DPOTrainer(args, S.AutoregressivePolicy(args, model2, S.tok, "lora_policy"),
                                S.AutoregressivePolicy(args, model2, S.tok, "lora_ref_policy"), S.B200AdamW(model2, lr=1e-3))
    hist2 = trainer2.train(8, lambda: itertools.cycle(S.batches), resume_training_ckpt=os.path.join(str(tmp_path), "checkpoint-2"))
    assert len(hist2) == 1 and trainer2.optimizer.step_count == 2 + 2     # the optimizer state of checkpoint-2 (2 steps) + one outer step
    model.engine.close()
    model2.engine.close()
