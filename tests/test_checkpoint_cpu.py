"""Reading the reference's on-disk layouts: sharded HF base checkpoint + index, CLIP tower directory named
by config.json["image_checkpoint"], adapter directory resolution, config -> EngineConfig. Host I/O only: the tensors are
streamed into a recorder here; Engine.load_tensor itself is covered by the GPU tests."""
import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _fake_llava(tmp_path):
    base, clip = tmp_path / "llava-v1.5-tiny", tmp_path / "vision_tower-clip"
    base.mkdir()
    clip.mkdir()
    g = torch.Generator().manual_seed(0)
    rnd = lambda *s: torch.randn(*s, generator=g)
    shard1 = {"model.embed_tokens.weight": rnd(32, 8), "model.layers.0.self_attn.q_proj.weight": rnd(8, 8),
              "model.layers.0.self_attn.rotary_emb.inv_freq": rnd(2)}
    shard2 = {"model.norm.weight": rnd(8), "lm_head.weight": rnd(32, 8), "model.mm_projector.0.weight": rnd(8, 4),
              "model.vision_tower.vision_tower.vision_model.pre_layrnorm.weight": rnd(4)}
    torch.save(shard1, base / "pytorch_model-00001-of-00002.bin")
    torch.save(shard2, base / "pytorch_model-00002-of-00002.bin")
    wm = {k: "pytorch_model-00001-of-00002.bin" for k in shard1}
    wm.update({k: "pytorch_model-00002-of-00002.bin" for k in shard2})
    (base / "pytorch_model.bin.index.json").write_text(json.dumps({"metadata": {}, "weight_map": wm}))
    (base / "config.json").write_text(json.dumps({
        "hidden_size": 4096, "num_hidden_layers": 32, "num_attention_heads": 32, "intermediate_size": 11008, "vocab_size": 32000,
        "rms_norm_eps": 1e-5, "max_position_embeddings": 4096, "mm_vision_select_layer": -2, "mm_projector_type": "mlp2x_gelu",
        "image_checkpoint": {"image_model": str(clip)}}))
    clip_sd = {"vision_model.pre_layrnorm.weight": rnd(4), "vision_model.embeddings.class_embedding": rnd(4),
               "vision_model.post_layernorm.weight": rnd(4), "text_model.embeddings.token_embedding.weight": rnd(6, 4)}
    torch.save(clip_sd, clip / "pytorch_model.bin")
    (clip / "config.json").write_text(json.dumps({"vision_config": {
        "hidden_size": 1024, "num_hidden_layers": 24, "num_attention_heads": 16, "intermediate_size": 4096, "image_size": 336,
        "patch_size": 14, "layer_norm_eps": 1e-5}}))
    return str(base), str(clip), shard1, shard2, clip_sd


def test_sharded_base_and_clip_are_streamed_in_reference_order(tmp_path):
    from opa_dpo_b200 import checkpoint as K
    from opa_dpo_b200._lib import OpadpoError, OpadpoUnknownName
    base, clip, s1, s2, cs = _fake_llava(tmp_path)
    assert K.clip_dir_from_config(base) == clip
    files = K.shard_files(base)
    assert [f for f, _ in files] == ["pytorch_model-00001-of-00002.bin", "pytorch_model-00002-of-00002.bin"]
    got = dict(K.iter_checkpoint_tensors(base))
    assert set(got) == set(s1) | set(s2) and all(torch.equal(got[k], {**s1, **s2}[k]) for k in got)
    only = dict(K.iter_checkpoint_tensors(base, lambda n: "vision_tower" in n or "mm_projector" in n))
    assert set(only) == {"model.mm_projector.0.weight", "model.vision_tower.vision_tower.vision_model.pre_layrnorm.weight"}

    class Recorder:
        def __init__(self):
            self.calls = []

        def load_tensor(self, name, t):
            if "inv_freq" in name or "post_layernorm" in name:
                raise OpadpoUnknownName("unknown tensor")
            self.calls.append((name, t.clone()))

    rec = Recorder()
    logs = []
    stats = K.load_llava_base(rec, base, log=logs.append)
    names = [n for n, _ in rec.calls]
    vt = "model.vision_tower.vision_tower.vision_model.pre_layrnorm.weight"
    # the CLIP directory first (text tower ignored), then the LLaVA shards, whose vision_tower entry overrides the tower's
    assert names[:2] == [vt, "model.vision_tower.vision_tower.vision_model.embeddings.class_embedding"]
    assert names.count(vt) == 2 and torch.equal([t for n, t in rec.calls if n == vt][-1], s2[vt])
    assert not any("text_model" in n for n in names)
    assert stats == {"loaded": len(names), "skipped": 2} and len(logs) == 2

    # a shape mismatch on a tensor the engine holds is NOT a skip (from_pretrained raises on size mismatches) ...
    class Strict(Recorder):
        def load_tensor(self, name, t):
            if name.endswith("embed_tokens.weight"):
                raise OpadpoError("load_tensor(model.embed_tokens.weight): shape [32003,4096], expected [32000,4096]")
            super().load_tensor(name, t)
    with pytest.raises(OpadpoError, match="expected"):
        K.load_llava_base(Strict(), base)
    # ... and a checkpoint that does not cover what the engine executes raises with the missing names
    with pytest.raises(OpadpoError, match="layers.5.mlp.up_proj.weight"):
        K.load_llava_base(Recorder(), base, required=names + ["model.layers.5.mlp.up_proj.weight"])
    assert K.load_llava_base(Recorder(), base, required=names)["loaded"] == len(names)


def test_required_tensor_names_cover_the_llava_state_dict_layout():
    from opa_dpo_b200.engine import EngineConfig, required_tensor_names, VT
    cfg = EngineConfig(n_layers=2, v_layers=3)
    req = required_tensor_names(cfg)
    assert len(req) == len(set(req))
    for n in ("model.embed_tokens.weight", "lm_head.weight", "model.norm.weight", "model.layers.1.mlp.down_proj.weight",
              "model.layers.0.input_layernorm.weight", "model.mm_projector.2.bias", VT + "encoder.layers.2.mlp.fc1.bias",
              VT + "embeddings.position_embedding.weight", VT + "pre_layrnorm.bias", VT + "encoder.layers.0.layer_norm2.weight"):
        assert n in req, n
    assert not any("layers.2.self_attn" in n and "encoder" not in n for n in req) and not any("encoder.layers.3" in n for n in req)
    assert not any(n.endswith("q_proj.bias") and "vision_model" not in n for n in req)       # LLaMA linears carry no bias


def test_engine_config_from_hf_and_adapter_dir(tmp_path):
    from opa_dpo_b200 import checkpoint as K
    from opa_dpo_b200.engine import EngineConfig
    base, clip, *_ = _fake_llava(tmp_path)
    cfg = K.engine_config_from_hf(base)
    assert cfg == EngineConfig()                          # LLaVA-1.5-7B + CLIP-L/14-336, select layer -2 -> 23 executed layers
    c = json.load(open(os.path.join(base, "config.json")))
    c["mm_vision_select_layer"] = -1
    json.dump(c, open(os.path.join(base, "config.json"), "w"))
    assert K.engine_config_from_hf(base).v_layers == 24
    c.update(hidden_size=5120, num_hidden_layers=40, num_attention_heads=40, intermediate_size=13824, mm_vision_select_layer=-2)
    json.dump(c, open(os.path.join(base, "config.json"), "w"))
    assert K.engine_config_from_hf(base) == EngineConfig.llava_13b()
    c["num_key_value_heads"] = 8
    json.dump(c, open(os.path.join(base, "config.json"), "w"))
    with pytest.raises(ValueError):
        K.engine_config_from_hf(base)
    # adapter directory resolution (qlora_model.py:57-64): adapter_model/ then lora_default/
    ck = tmp_path / "checkpoint-10"
    (ck / "adapter_model" / "lora_default").mkdir(parents=True)
    assert K.resolve_adapter_dir(str(ck)) == str(ck / "adapter_model" / "lora_default")
    flat = tmp_path / "flat"
    flat.mkdir()
    assert K.resolve_adapter_dir(str(flat)) == str(flat)
    single = tmp_path / "single"
    single.mkdir()
    torch.save({"a": torch.ones(2)}, single / "pytorch_model.bin")
    assert dict(K.iter_checkpoint_tensors(str(single)))["a"].tolist() == [1.0, 1.0]
    with pytest.raises(FileNotFoundError):
        K.shard_files(str(flat))


def test_last_checkpoint_contract_vs_reference_function(tmp_path):
    """utils/lora_utils.get_last_checkpoint run verbatim against optim.last_checkpoint_and_completion on the same trees."""
    if not os.path.isdir("/root/reference"):
        pytest.skip("needs /root/reference")
    import contextlib
    import io
    from oracle.make_golden import install_stubs
    install_stubs()
    from utils import lora_utils as L
    from opa_dpo_b200.optim import last_checkpoint_and_completion as mine

    def ref(d):
        with contextlib.redirect_stdout(io.StringIO()):
            return L.get_last_checkpoint(d)
    d = tmp_path / "run"
    assert mine(str(d)) == ref(str(d)) == (None, False)                 # first training
    d.mkdir()
    assert mine(str(d)) == ref(str(d)) == (None, False)                 # started, nothing saved
    (d / "checkpoint-0").mkdir()
    assert mine(str(d)) == ref(str(d))
    for n in (75, 150, 9):
        (d / f"checkpoint-{n}").mkdir()
        assert mine(str(d)) == ref(str(d))
    assert mine(str(d))[0].endswith("checkpoint-150")
    (d / "checkpoint-300").write_text("a file, not a directory")
    assert mine(str(d)) == ref(str(d))
    (d / "completed").write_text("")
    assert mine(str(d)) == ref(str(d)) == (None, True)


def test_make_models_mirrors_reference_construction(tmp_path):
    """policy.make_models (dpo_trainer.py:958-1038 over qlora_model.py:44-127,297-356): which directories are read, which
    base the adapters sit on, the two adapter names on ONE shared base, trainable / frozen, the resumed policy adapter,
    the fresh-adapter branch (train_from_sft=False: A ~ kaiming_uniform, B = 0) and the vision-tower re-read from
    args.base_model_name. The engine is a recorder (the product's engine needs a GPU)."""
    import math
    from types import SimpleNamespace
    from opa_dpo_b200 import policy as P
    from opa_dpo_b200.engine import EngineConfig, linear_modules
    from opa_dpo_b200._lib import OpadpoError, OpadpoUnknownName
    base, clip, s1, s2, cs = _fake_llava(tmp_path)
    tiny = EngineConfig(d_model=8, n_layers=1, n_heads=1, d_ffn=16, vocab=32, v_d=4, v_layers=1, v_heads=1, v_ffn=8, v_image=28,
                        v_patch=14, lora_r=2, lora_alpha=4.0)
    mods = linear_modules(tiny)

    class RecEngine:
        device = torch.device("cpu")

        def __init__(self):
            self.cfg, self.tensors, self.lora, self.trainable = tiny, {}, {}, None

        def load_tensor(self, name, t):
            if "inv_freq" in name or "post_layernorm" in name:
                raise OpadpoUnknownName("unknown tensor")
            self.tensors[name] = t.clone()

        def lora_shape(self, module):
            if module not in mods:
                raise OpadpoError("not a LoRA target")
            return mods[module]

        def load_lora(self, aid, module, A, B):
            assert A.shape == (2, mods[module][1]) and B.shape == (mods[module][0], 2), module
            self.lora[(aid, module)] = (A.clone(), B.clone())

        def set_trainable(self, aid):
            self.trainable = aid

        def layer_grad_range(self, layer):
            assert layer == tiny.n_layers            # the projector's slice
            return 6, 2

        def buffers(self):
            return dict(master=torch.zeros(8), grad=torch.zeros(8), n_train=8)

    made = []

    def factory(cfg, device=None, base_model_name_or_path=""):
        made.append((cfg, device, base_model_name_or_path))
        return P.B200LlavaModel(tiny, base_model_name_or_path=base_model_name_or_path, engine=RecEngine())

    g = torch.Generator().manual_seed(1)
    sft = tmp_path / "sft" / "adapter_model" / "lora_default"
    sft.mkdir(parents=True)
    state = {}
    for m, (o, i) in mods.items():
        state[f"base_model.model.{m}.lora_A.weight"] = torch.randn(2, i, generator=g)
        state[f"base_model.model.{m}.lora_B.weight"] = torch.randn(o, 2, generator=g)
    torch.save(state, sft / "adapter_model.bin")
    (sft / "adapter_config.json").write_text(json.dumps({"r": 2, "lora_alpha": 4.0, "lora_dropout": 0.0, "target_modules": ["q_proj"],
                                                         "base_model_name_or_path": base, "peft_type": "LORA"}))
    tok = SimpleNamespace(pad_token_id=0, eos_token_id=2)
    args = SimpleNamespace(policy_model_name_or_path=str(tmp_path / "sft"), base_model_name=base, temperature=1.0, response_len=4)
    acc = SimpleNamespace(device=torch.device("cpu"))
    out = P.make_models(None, tok, args, acc, model_factory=factory)
    assert set(out) == {"policy", "ref_policy"} and len(made) == 1
    cfg_seen, dev_seen, name_seen = made[0]
    assert name_seen == base and dev_seen == acc.device
    assert cfg_seen == EngineConfig(lora_r=2, lora_alpha=4.0)      # dims from the base's config.json, LoRA from adapter_config.json
    pol, ref = out["policy"], out["ref_policy"]
    assert pol.base_model is ref.base_model and pol.adapter_name == "lora_policy" and ref.adapter_name == "lora_ref_policy"
    eng = pol.base_model.engine
    assert eng.trainable == P.POLICY and pol.base_model.flat_param is not None
    m0 = next(iter(mods))
    for aid in (P.POLICY, P.REF):
        assert set(k[1] for k in eng.lora if k[0] == aid) == set(mods)
        assert torch.equal(eng.lora[(aid, m0)][0], state[f"base_model.model.{m0}.lora_A.weight"])
    assert set(eng.tensors) == (set(s1) | set(s2) | {"model.vision_tower.vision_tower." + k for k in cs if k.startswith("vision_model.")}) \
        - {"model.layers.0.self_attn.rotary_emb.inv_freq", "model.vision_tower.vision_tower.vision_model.post_layernorm.weight"}
    assert pol.base_model.peft_config["lora_policy"]["r"] == 2
    # the DPO stage freezes the projector with its LoRA tensors (dpo_trainer.py:1016-1029): its gradient slice is declared frozen
    assert pol.base_model.frozen_grad_ranges == [(6, 2)]
    with pytest.raises(NotImplementedError):
        P.make_models(None, tok, SimpleNamespace(**{**vars(args), "vision_tower": "same"}), acc, model_factory=factory)

    # resumed run: the policy adapter comes from <ckpt>/adapter_model/lora_policy, the reference adapter from the SFT directory
    ck = tmp_path / "checkpoint-5"
    state2 = {k: v + 1.0 for k, v in state.items()}
    (ck / "adapter_model" / "lora_policy").mkdir(parents=True)
    torch.save(state2, ck / "adapter_model" / "lora_policy" / "adapter_model.bin")
    (ck / "adapter_model" / "lora_policy" / "adapter_config.json").write_text((sft / "adapter_config.json").read_text())
    out2 = P.make_models(None, tok, args, acc, resume_from_checkpoint=str(ck), model_factory=factory)
    e2 = out2["policy"].base_model.engine
    assert torch.equal(e2.lora[(P.POLICY, m0)][1], state2[f"base_model.model.{m0}.lora_B.weight"])
    assert torch.equal(e2.lora[(P.REF, m0)][1], state[f"base_model.model.{m0}.lora_B.weight"])

    # train_from_sft=False: fresh adapters, PEFT's init (B = 0, |A| <= 1/sqrt(fan_in)); an already-built base is reused
    args2 = SimpleNamespace(**{**vars(args), "train_from_sft": False})
    out3 = P.make_models(out2["policy"].base_model, tok, args2, acc, model_factory=factory)
    assert len(made) == 2 and out3["policy"].base_model is out2["policy"].base_model
    for aid in (P.POLICY, P.REF):
        A, B = e2.lora[(aid, m0)]
        assert (B == 0).all() and A.abs().max() <= 1.0 / math.sqrt(A.shape[1]) + 1e-6 and A.abs().max() > 0
    assert not torch.equal(e2.lora[(P.POLICY, m0)][0], e2.lora[(P.REF, m0)][0])

    # the vision tower / projector are re-read from args.base_model_name when it is another directory
    other = tmp_path / "other-base"
    other.mkdir()
    vt = "model.vision_tower.vision_tower.vision_model.pre_layrnorm.weight"
    torch.save({vt: torch.full((4,), 7.0), "model.embed_tokens.weight": torch.zeros(32, 8)}, other / "pytorch_model.bin")
    args3 = SimpleNamespace(**{**vars(args), "base_model_name": str(other)})
    out4 = P.make_models(None, tok, args3, acc, model_factory=factory)
    e4 = out4["policy"].base_model.engine
    assert (e4.tensors[vt] == 7.0).all() and torch.equal(e4.tensors["model.embed_tokens.weight"], s1["model.embed_tokens.weight"])
    with pytest.raises(OpadpoError):
        P.make_models(None, tok, args, acc, num_new_tokens=1, model_factory=factory)


def test_save_model_directory_rules(tmp_path, monkeypatch):
    """DPOTrainer.save_model's file rules (dpo_trainer.py:837-899): the directory it is given (any name), adapter under
    adapter_model/lora_policy, stray pytorch_model*.bin removed, optimizer.pt of every sibling directory deleted, nothing but
    the adapter for a best-checkpoint save, nothing at all off rank 0."""
    from types import SimpleNamespace
    from opa_dpo_b200.trainer import B200DPOTrainer

    class Model:
        flat_param = None
        engine = SimpleNamespace(device=torch.device("cpu"))

        def save_adapter(self, d, name):
            os.makedirs(os.path.join(d, name), exist_ok=True)
            open(os.path.join(d, name, "adapter_model.bin"), "w").write("x")

    sd = SimpleNamespace(state_dict=lambda: {"k": 1})
    tr = B200DPOTrainer(SimpleNamespace(output_dir=str(tmp_path / "run" / "final"), max_step=3), SimpleNamespace(base_model=Model()),
                        None, optimizer=sd, lr_scheduler=sd, loss=object())
    old = tmp_path / "run" / "checkpoint-3"
    old.mkdir(parents=True)
    (old / "optimizer.pt").write_text("old")
    (old / "scheduler.pt").write_text("old")
    out = tmp_path / "run" / "final"
    out.mkdir()
    (out / "pytorch_model-00001-of-00002.bin").write_text("stray")
    monkeypatch.setenv("RANK", "1")
    assert tr.save_model() is None and not (out / "adapter_model").exists()
    monkeypatch.setenv("RANK", "0")
    assert tr.save_model() == str(out)                       # output_dir defaults to args.output_dir
    assert (out / "adapter_model" / "lora_policy" / "adapter_model.bin").exists()
    assert (out / "optimizer.pt").exists() and (out / "scheduler.pt").exists() and not (out / "pytorch_model-00001-of-00002.bin").exists()
    assert not (old / "optimizer.pt").exists() and (old / "scheduler.pt").exists()
    best = tmp_path / "run" / "best"
    assert tr.save_model(str(best), for_best_ckpt=True) == str(best)
    assert (best / "adapter_model" / "lora_policy" / "adapter_model.bin").exists() and not (best / "optimizer.pt").exists()
    assert (out / "optimizer.pt").exists()                   # a best-checkpoint save leaves the others' optimizer state alone


def test_create_optimizer_and_scheduler_from_args():
    """RLTrainer.create_optimizer_and_scheduler (rl_trainer.py:197-215 over utils/trainer_utils.py:9-49) when neither was
    handed in: AdamW hyper-parameters and the schedule come from the args; max_grad_norm None = no clipping."""
    from types import SimpleNamespace
    from opa_dpo_b200.optim import B200AdamW, cosine_schedule
    from opa_dpo_b200.trainer import B200DPOTrainer
    calls = []
    eng = SimpleNamespace(device=torch.device("cpu"),
                          optimizer_step=lambda lr, betas, eps, wd, mg, scale: calls.append((lr, betas, eps, wd, mg, scale)) or torch.ones(1))
    model = SimpleNamespace(flat_param=torch.nn.Parameter(torch.zeros(4)), engine=eng)
    args = SimpleNamespace(learning_rate=2e-6, adam_beta1=0.8, adam_beta2=0.95, adam_epsilon=1e-6, weight_decay=0.1, max_grad_norm=None,
                           lr_scheduler_type="linear", warmup_steps=0, warmup_ratio=0.2, max_step=50)
    tr = B200DPOTrainer(args, SimpleNamespace(base_model=model), None, optimizer=None, loss=object())
    opt, sch = tr.create_optimizer_and_scheduler(10)
    assert isinstance(opt, B200AdamW) and isinstance(sch, cosine_schedule) and sch.kind == "linear" and sch.warmup == 2 and sch.total == 10
    g = opt.param_groups[0]
    assert g["initial_lr"] == 2e-6 and g["betas"] == (0.8, 0.95) and g["eps"] == 1e-6 and g["weight_decay"] == 0.1 and g["lr"] == 0.0
    sch.step()
    opt.step(0.5)
    assert calls == [(1e-6, (0.8, 0.95), 1e-6, 0.1, 0.0, 0.5)]          # lr after one warm-up step; clip disabled -> 0
    # `loss/grad_norm` is the norm AFTER clipping, as the reference logs it (rl_trainer.py:164-171); no clip -> the raw norm
    eng.zero_grad = lambda: None
    model.release_slots = lambda: None
    eng.optimizer_step = lambda *a: torch.full((1,), 4.0)
    assert tr._optimizer_step().item() == 4.0 and opt.last_grad_norm.item() == 4.0          # max_grad_norm None
    opt.max_grad_norm = 1.0
    assert abs(tr._optimizer_step().item() - 1.0) < 1e-5 and opt.last_grad_norm.item() == 4.0
    eng.optimizer_step = lambda *a: torch.full((1,), 0.5)
    assert tr._optimizer_step().item() == 0.5
    opt.max_grad_norm = None
    # frozen slices (the projector's LoRA in the DPO stage): gradient cleared before the fused step, so it is outside the norm
    # and untouched by Adam; with weight decay the weights are put back as well
    seen = {}

    def fused(lr, betas, eps, wd, mg, scale):
        seen["grad"] = model.flat_param.grad.clone()
        model.flat_param.data.mul_(0.5)                       # stands for the decoupled decay touching every weight
        return torch.ones(1)
    eng.optimizer_step, eng.sync_shadow = fused, lambda: seen.__setitem__("synced", True)
    model.flat_param.data.fill_(2.0)
    model.flat_param.grad = torch.ones(4)
    model.frozen_grad_ranges = [(1, 2)]
    opt.step(1.0)
    assert seen["grad"].tolist() == [1.0, 0.0, 0.0, 1.0] and model.flat_param.data.tolist() == [1.0, 2.0, 2.0, 1.0] and seen["synced"]
    opt.param_groups[0]["weight_decay"] = 0.0                 # no decay: nothing to put back, no extra copies
    seen.clear()
    model.flat_param.grad = torch.ones(4)
    opt.step(1.0)
    assert seen["grad"].tolist() == [1.0, 0.0, 0.0, 1.0] and "synced" not in seen
    model.frozen_grad_ranges = []
    opt.param_groups[0]["weight_decay"] = 0.1
    # handed-in objects are kept (create_optimizer / create_scheduler return them unchanged)
    tr2 = B200DPOTrainer(args, SimpleNamespace(base_model=model), None, optimizer=opt, lr_scheduler=sch, loss=object())
    assert tr2.create_optimizer_and_scheduler(99) == (opt, sch)


def test_train_opa_dpo_entry_flow(tmp_path, monkeypatch):
    """trainer.train_opa_dpo = the model / training half of opadpo_train.train (opadpo_train.py:583-595, 727-797): LoRA from
    scratch into <output_dir>/init_model when base == policy path, make_models on it, the training loop, checkpoint-final; a
    second launch finds init_model and the last checkpoint and resumes. Recorder engine, stub policies' arithmetic."""
    from types import SimpleNamespace
    from opa_dpo_b200 import policy as P, trainer as T
    from opa_dpo_b200.engine import EngineConfig, linear_modules
    from opa_dpo_b200._lib import OpadpoError, OpadpoUnknownName
    monkeypatch.setenv("RANK", "0")
    monkeypatch.setenv("WORLD_SIZE", "1")
    base, clip, *_ = _fake_llava(tmp_path)
    tiny = EngineConfig(d_model=8, n_layers=1, n_heads=1, d_ffn=16, vocab=32, v_d=4, v_layers=1, v_heads=1, v_ffn=8, v_image=28,
                        v_patch=14, lora_r=2, lora_alpha=4.0)
    mods = linear_modules(tiny)
    events = []

    class RecEngine:
        device = torch.device("cpu")

        def __init__(self):
            self.cfg, self.lora, self.master = tiny, {}, torch.zeros(8)

        def load_tensor(self, name, t):
            if "inv_freq" in name or "post_layernorm" in name:
                raise OpadpoUnknownName("unknown tensor")

        def lora_shape(self, module):
            if module not in mods:
                raise OpadpoError("not a LoRA target")
            return mods[module]

        def load_lora(self, aid, module, A, B):
            self.lora[(aid, module)] = (A.clone(), B.clone())

        def export_lora(self, aid, module):
            return self.lora[(aid, module)]

        def set_trainable(self, aid):
            events.append(("trainable", aid))

        def layer_grad_range(self, layer):
            assert layer == tiny.n_layers            # the projector's slice
            return 6, 2

        def buffers(self):
            return dict(master=self.master, grad=torch.zeros(8), exp_avg=torch.zeros(8), exp_avg_sq=torch.zeros(8), n_train=8, n_total=8)

        def optimizer_step(self, *a):
            events.append("opt_step")
            return torch.ones(1)

        def zero_grad(self):
            pass

        def set_option(self, *a):
            pass

        def sync_shadow(self):
            pass

    def factory(cfg, device=None, base_model_name_or_path=""):
        events.append(("build", cfg.lora_r, cfg.lora_alpha))
        return P.B200LlavaModel(tiny, base_model_name_or_path=base_model_name_or_path, engine=RecEngine())

    class Trainer(T.B200DPOTrainer):      # the step itself is covered elsewhere: record it, advance the optimizer / scheduler
        def step(self, loader, step_idx=0):
            events.append(("step", step_idx))
            for _ in range(self.args.rollout_accumulation_steps):
                next(loader)
            self.optimizer.step(1.0)
            self.lr_scheduler.step()
            return {}
    monkeypatch.setattr(T, "B200DPOTrainer", Trainer)
    out = tmp_path / "out"
    args = SimpleNamespace(output_dir=str(out), resume_dir=None, base_model_name=base, policy_model_name_or_path=base, lora_r=2,
                           lora_alpha=4.0, resume_from_training=True, total_epochs=1, rollout_batch_size=2,
                           rollout_accumulation_steps=1, save_steps=2, save_steps_extra_list=[], eval_steps=None, max_step=10,
                           learning_rate=1e-5, warmup_steps=1, temperature=1.0, response_len=4, loss_type="sigmoid",
                           f_divergence_type="reverse_kl")
    import itertools
    tr = T.train_opa_dpo(args, SimpleNamespace(pad_token_id=0, eos_token_id=2), 8, lambda: itertools.cycle([{}]),
                         accelerator=SimpleNamespace(device=torch.device("cpu")), model_factory=factory, trainer_kwargs=dict(loss=object()))
    init = out / "init_model"
    assert (init / "adapter_model.bin").exists() and args.policy_model_name_or_path == str(init)
    cfgj = json.load(open(init / "adapter_config.json"))
    assert cfgj["r"] == 2 and cfgj["lora_alpha"] == 4.0 and cfgj["base_model_name_or_path"] == base and cfgj["bias"] == "none"
    sd = torch.load(init / "adapter_model.bin")
    m0 = next(iter(mods))
    assert (sd[f"base_model.model.{m0}.lora_B.weight"] == 0).all() and sd[f"base_model.model.{m0}.lora_A.weight"].abs().max() > 0
    eng = tr.model.engine
    assert torch.equal(eng.lora[(P.POLICY, m0)][0], eng.lora[(P.REF, m0)][0])           # both adapters start from init_model
    assert [e for e in events if e[0] == "build"] == [("build", 2, 4.0)]                 # ONE base for everything
    assert [e for e in events if isinstance(e, tuple) and e[0] == "step"] == [("step", i) for i in (1, 2, 3, 4)]   # 8 / 2 steps
    assert sorted(os.listdir(out)) == ["checkpoint-2", "checkpoint-4", "checkpoint-final", "init_model"]
    assert (out / "checkpoint-final" / "adapter_model" / "lora_policy" / "adapter_model.bin").exists()
    assert (out / "checkpoint-final" / "optimizer.pt").exists() and not (out / "checkpoint-4" / "optimizer.pt").exists()

    # second launch: init_model is reused (no rebuild of a fresh adapter), the run resumes from checkpoint-4: steps 1-3 skipped
    events.clear()
    args2 = SimpleNamespace(**{**vars(args), "policy_model_name_or_path": base})
    (out / "checkpoint-4" / "optimizer.pt").write_bytes((out / "checkpoint-final" / "optimizer.pt").read_bytes())
    tr2 = T.train_opa_dpo(args2, SimpleNamespace(pad_token_id=0, eos_token_id=2), 8, lambda: itertools.cycle([{}]),
                          accelerator=SimpleNamespace(device=torch.device("cpu")), model_factory=factory, trainer_kwargs=dict(loss=object()))
    assert [e for e in events if isinstance(e, tuple) and e[0] == "step"] == [("step", 4)]
    assert args2.policy_model_name_or_path == str(init) and tr2.optimizer.step_count == 4 + 1


def test_opa_stage_final_checkpoint_feeds_the_dpo_stage(tmp_path):
    """opa.save_final_checkpoint writes the OPA stage's output layout (opa_train.py:553-561) and policy.make_models — the DPO
    stage's loader (qlora_model.py:318-328) — reads exactly that directory back as both adapters' starting point."""
    from types import SimpleNamespace
    from opa_dpo_b200 import opa, policy as P
    from opa_dpo_b200.engine import EngineConfig, linear_modules
    from opa_dpo_b200._lib import OpadpoError, OpadpoUnknownName
    base, clip, *_ = _fake_llava(tmp_path)
    tiny = EngineConfig(d_model=8, n_layers=1, n_heads=1, d_ffn=16, vocab=32, v_d=4, v_layers=1, v_heads=1, v_ffn=8, v_image=28,
                        v_patch=14, lora_r=2, lora_alpha=4.0)
    mods = linear_modules(tiny)

    class RecEngine:
        device = torch.device("cpu")

        def __init__(self):
            self.cfg, self.lora = tiny, {}

        def load_tensor(self, name, t):
            if "inv_freq" in name or "post_layernorm" in name:
                raise OpadpoUnknownName("unknown tensor")

        def lora_shape(self, module):
            if module not in mods:
                raise OpadpoError("not a LoRA target")
            return mods[module]

        def load_lora(self, aid, module, A, B):
            self.lora[(aid, module)] = (A.clone().float(), B.clone().float())

        def export_lora(self, aid, module):
            return self.lora[(aid, module)]

        def set_trainable(self, aid):
            pass

        def layer_grad_range(self, layer):
            assert layer == tiny.n_layers            # the projector's slice
            return 6, 2

        def buffers(self):
            return dict(master=torch.zeros(8), grad=torch.zeros(8), n_train=8)

    factory = lambda cfg, device=None, base_model_name_or_path="": P.B200LlavaModel(tiny, base_model_name_or_path=base_model_name_or_path,
                                                                                   engine=RecEngine())
    sft = factory(None, base_model_name_or_path=base)
    sft.add_adapter("default", is_trainable=True, generator=torch.Generator().manual_seed(3))     # PEFT's single OPA adapter
    m0 = next(iter(mods))
    A0, _ = sft.engine.lora[(P.POLICY, m0)]
    sft.engine.lora[(P.POLICY, m0)] = (A0, torch.full((mods[m0][0], 2), 0.25))                    # "trained"
    final = opa.save_final_checkpoint(sft, str(tmp_path / "opa_out"), adapter_name="default")
    assert sorted(os.listdir(final)) == ["adapter_config.json", "adapter_model.bin", "config.json", "non_lora_trainables.bin"]
    assert torch.load(os.path.join(final, "non_lora_trainables.bin")) == {}
    cfgj = json.load(open(os.path.join(final, "adapter_config.json")))
    assert cfgj["base_model_name_or_path"] == base and cfgj["r"] == 2 and cfgj["inference_mode"] is True
    assert opa.save_final_checkpoint(sft, str(tmp_path / "nope"), is_main_process=False) is None and not (tmp_path / "nope").exists()

    args = SimpleNamespace(policy_model_name_or_path=final, base_model_name=base, temperature=1.0, response_len=4)
    out = P.make_models(None, SimpleNamespace(pad_token_id=0, eos_token_id=2), args, None, model_factory=factory)
    eng = out["policy"].base_model.engine
    for aid in (P.POLICY, P.REF):
        A, B = eng.lora[(aid, m0)]
        assert torch.equal(A, A0.to(torch.bfloat16).float()) and (B == 0.25).all()                # adapter_model.bin is bf16, as PEFT's
    # the reference's default adapter name resolves (rl_models.py:193-204 `adapter_name="default"`)
    pol = P.make_policy_with_base_model(args, out["policy"].base_model, None)
    assert pol.adapter_name == "default"
    out["policy"].base_model.set_adapter("default")


def test_load_adapter_refuses_configs_the_engine_cannot_honour(tmp_path):
    """PeftModel.load_adapter's contract (qlora_model.py:84-88) for what the engine cannot express: another rank or alpha than
    it was built with, LoRA dropout on a trainable adapter, bias terms."""
    from opa_dpo_b200 import policy as P
    from opa_dpo_b200.engine import EngineConfig
    from opa_dpo_b200._lib import OpadpoError
    from types import SimpleNamespace
    model = P.B200LlavaModel(EngineConfig(lora_r=2, lora_alpha=4.0), engine=SimpleNamespace(device=torch.device("cpu")))
    d = tmp_path / "ad"
    d.mkdir()
    torch.save({}, d / "adapter_model.bin")

    def write(**kw):
        (d / "adapter_config.json").write_text(json.dumps({"r": 2, "lora_alpha": 4.0, "lora_dropout": 0.0, "bias": "none", **kw}))
    write()
    model.load_adapter(str(d), "lora_ref_policy")
    for bad in (dict(r=4), dict(lora_alpha=8.0), dict(bias="all")):
        write(**bad)
        with pytest.raises(OpadpoError):
            model.load_adapter(str(d), "lora_ref_policy")
    write(lora_dropout=0.05)
    model.load_adapter(str(d), "lora_ref_policy")              # frozen: dropout never applies in eval mode
    with pytest.raises(OpadpoError, match="dropout"):
        model.load_adapter(str(d), "lora_policy", is_trainable=True)
