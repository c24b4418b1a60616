"""Host-side mirror of the reference's policy boundary (opadpo/dpo_models/rl_models.py) on top of the C-ABI engine.

`AutoregressivePolicy` keeps the reference's constructor `(args, base_model, base_tokenizer, adapter_name)` and
`forward(images, queries, queries_attn_masks, temperature=None, **kwargs)` contract (rl_models.py:75-144): response
kwargs are discovered by name, K responses are stacked behind the shared query, and the result is
`{"<name>_logprobs", "<name>_entropies"}` of shape [b, response_len]. `B200LlavaModel` stands where the reference has
the PEFT-wrapped LlavaLlamaForCausalLM (`base_model`): `set_adapter`, `config.use_cache`, `generate(...)`,
`peft_config`, adapter save/load in PEFT layout. All arithmetic happens in libopadpo_b200.so.
"""
from __future__ import annotations

import json
import math
import os
import weakref
from types import SimpleNamespace
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor, nn

from .engine import BASE, POLICY, REF, Engine, EngineConfig, linear_modules
from ._lib import OpadpoError, OpadpoUnknownName

# "default" is PEFT's name for the single adapter of the OPA LoRA-SFT stage (opa_train.py:466-484) and the default of
# make_policy_with_base_model (rl_models.py:193-204); it occupies the trainable slot like lora_policy
ADAPTER_IDS = {"lora_policy": POLICY, "lora_ref_policy": REF, "lora_default": POLICY, "default": POLICY, None: BASE}


class _PolicyForward(torch.autograd.Function):
    """Autograd edge across the C ABI: forward keeps activations in an engine slot, backward feeds d(logprobs) to
    opadpo_backward, which accumulates LoRA gradients straight into the engine's flat bucket (exposed as
    `flat_param.grad`)."""

    @staticmethod
    def forward(ctx, anchor: Tensor, model: "B200LlavaModel", adapter: int, images, queries, qmask, img_mask, temperature,
                slot: int, *responses):
        lp, ent = model.engine.forward_logprobs(adapter, images, queries, qmask, list(responses), temperature,
                                                img_mask=img_mask, save_slot=slot, want_entropy=True)
        ctx.model, ctx.slot, ctx.n_in = model, slot, 9 + len(responses)
        ctx.gen = model._slot_gen[slot]
        # a graph that is dropped without ever running backward (an exception mid-step, a logging forward in train mode)
        # must not leak its slot: the slot is handed back when the autograd node dies
        weakref.finalize(ctx, model._drop_slot, slot, ctx.gen)
        ctx.out_shape = tuple(lp.shape)
        ctx.set_materialize_grads(False)   # an unused output arrives as None instead of a zero tensor
        return lp, ent

    @staticmethod
    def backward(ctx, dlp, dent):
        # the entropies are differentiable too (OPA's entropy regulariser, opa_models/opa_trainer.py:63,82-85)
        model = ctx.model
        if not model._slot_busy[ctx.slot] or model._slot_gen[ctx.slot] != ctx.gen:
            raise OpadpoError(f"backward of a forward whose activation slot {ctx.slot} was released or reused in the meantime "
                              "(release_slots() / zero_grad() between forward and backward, or a second backward)")
        if dlp is None and dent is None:   # nothing flows back: still run the backward so the slot is released
            dlp = torch.zeros(ctx.out_shape, device=model.engine.device)
        model.engine.backward(ctx.slot, dlp, dent)
        model._slot_busy[ctx.slot] = False
        return (None,) * ctx.n_in


class B200LlavaModel(nn.Module):
    """Stand-in for the reference's `base_model` (PeftModelForCausalLM over LlavaLlamaForCausalLM, built by
    qlora_model.get_accelerate_model :44-294): one shared base + adapters 'lora_policy' (trainable) and
    'lora_ref_policy' (frozen)."""

    def __init__(self, cfg: EngineConfig, device=None, base_model_name_or_path: str = "", engine=None):
        """`engine`: an already-built engine handle to wrap instead of creating one (the CPU tests pass a recorder to check
        the host logic of make_models / adapter loading; the package itself only ever builds `Engine`, which needs a GPU)."""
        super().__init__()
        self.engine = engine if engine is not None else Engine(cfg, device)
        self.cfg = cfg
        self.config = SimpleNamespace(use_cache=False, cache_shape=None, vocab_size=cfg.vocab, hidden_size=cfg.d_model)
        self.active_adapter: Optional[str] = "lora_policy"
        self.peft_config: Dict[str, dict] = {}
        self.base_model_name_or_path = base_model_name_or_path
        self.flat_param: Optional[nn.Parameter] = None
        self._slot_busy = [False, False]     # the engine keeps the activations of at most two undifferentiated forwards
        self._slot_gen = [0, 0]              # bumped at every hand-out, so a stale backward / finaliser is recognised
        self._passthrough: Dict[str, Dict[str, Tensor]] = {}   # adapter tensors the engine does not execute (CLIP layer 24)
        self.frozen_grad_ranges = []         # (offset, count) slices of the gradient bucket that must not train

    # ---- adapters
    def set_adapter(self, name: Optional[str]):
        if name not in ADAPTER_IDS:
            raise ValueError(f"unknown adapter {name!r}")
        self.active_adapter = name

    def get_vision_tower(self):
        return SimpleNamespace(num_patches=self.cfg.n_patches)

    def make_trainable(self, name: str = "lora_policy"):
        """Allocates fp32 masters / grad bucket / Adam moments and exposes them as ONE flat nn.Parameter whose .grad
        is a view of the engine's gradient bucket (what optimizer, grad-norm and all-reduce operate on)."""
        self.engine.set_trainable(ADAPTER_IDS[name])
        v = self.engine.buffers()
        self.flat_param = nn.Parameter(v["master"][: v["n_train"]], requires_grad=True)
        self.flat_param.grad = v["grad"]
        return self.flat_param

    def freeze_projector_lora(self):
        """The DPO stage freezes the projector — its LoRA tensors included — next to the vision tower
        (`policy_mm_projector.requires_grad_(False)`, dpo_trainer.py:1019-1029): the optimizer never sees them and they are
        saved unchanged. The engine differentiates the projector's LoRA like every other (the OPA stage trains it,
        opa_train.py:496-497), so here its slice of the gradient bucket is declared frozen: B200AdamW clears it before the
        fused clip + AdamW step, which keeps it out of the gradient norm and leaves weights and moments untouched."""
        rng = tuple(self.engine.layer_grad_range(self.cfg.n_layers))
        if rng[1] > 0 and rng not in self.frozen_grad_ranges:
            self.frozen_grad_ranges.append(rng)

    def default_peft_config(self) -> dict:
        mods = sorted({m.split(".")[-1] for m in linear_modules(self.cfg)})
        return dict(peft_type="LORA", task_type="CAUSAL_LM", r=self.cfg.lora_r, lora_alpha=self.cfg.lora_alpha,
                    lora_dropout=0.0, bias="none", target_modules=mods, inference_mode=False, fan_in_fan_out=False,
                    base_model_name_or_path=self.base_model_name_or_path, modules_to_save=None, init_lora_weights=True)

    def load_adapter_state(self, name: str, state: Dict[str, Tensor]):
        """`state` uses PEFT-0.5.0 saved key names: base_model.model.<module>.lora_A.weight / lora_B.weight
        (get_peft_model_state_dict strips the adapter name; dpo_trainer.py:1070-1095)."""
        aid = ADAPTER_IDS[name]
        pairs: Dict[str, Dict[str, Tensor]] = {}
        for k, t in state.items():
            k2 = k.replace(f".{name}.", ".")
            if ".lora_A" in k2:
                pairs.setdefault(k2.split(".lora_A")[0], {})["A"] = t
            elif ".lora_B" in k2:
                pairs.setdefault(k2.split(".lora_B")[0], {})["B"] = t
        for mod, ab in pairs.items():
            short = mod[len("base_model.model."):] if mod.startswith("base_model.model.") else mod
            try:
                self.engine.lora_shape(short)
            except OpadpoError:
                self._passthrough.setdefault(name, {})[mod + ".lora_A.weight"] = ab["A"].detach().cpu()
                self._passthrough[name][mod + ".lora_B.weight"] = ab["B"].detach().cpu()
                continue
            self.engine.load_lora(aid, short, ab["A"], ab["B"])
        self.peft_config.setdefault(name, self.default_peft_config())

    def add_adapter(self, name: str, is_trainable: bool = False, peft_config: Optional[dict] = None, generator=None):
        """A freshly initialised adapter on every LoRA target (qlora_model.py:84-113, the `train_from_sft=False` branch:
        `model.add_adapter(LoraConfig(r, lora_alpha, ...), adapter_name)`), with PEFT's reset_lora_parameters init
        [UPSTREAM peft 0.5.0 tuners/lora.py]: A ~ kaiming_uniform(a = sqrt(5)), B = 0."""
        aid = ADAPTER_IDS[name]
        r = self.cfg.lora_r
        for mod, (out_f, in_f) in linear_modules(self.cfg).items():
            A = torch.empty(r, in_f)
            torch.nn.init.kaiming_uniform_(A, a=math.sqrt(5), generator=generator)
            self.engine.load_lora(aid, mod, A, torch.zeros(out_f, r))
        self.peft_config[name] = dict(peft_config) if peft_config is not None else self.default_peft_config()
        if is_trainable:
            self.make_trainable(name)

    def adapter_state_dict(self, name: str) -> Dict[str, Tensor]:
        aid = ADAPTER_IDS[name]
        out: Dict[str, Tensor] = {}
        for mod in linear_modules(self.cfg):
            A, B = self.engine.export_lora(aid, mod)
            out[f"base_model.model.{mod}.lora_A.weight"] = A.to(torch.bfloat16).cpu()
            out[f"base_model.model.{mod}.lora_B.weight"] = B.to(torch.bfloat16).cpu()
        out.update(self._passthrough.get(name, {}))
        return out

    def save_adapter(self, save_directory: str, name: str):
        """PEFT layout written by save_adapters (dpo_trainer.py:1047-1095): <dir>/<name>/{adapter_model.bin,
        adapter_config.json} ('default' adapter goes to <dir> itself)."""
        out_dir = os.path.join(save_directory, name) if name != "default" else save_directory
        os.makedirs(out_dir, exist_ok=True)
        torch.save(self.adapter_state_dict(name), os.path.join(out_dir, "adapter_model.bin"))
        cfg = dict(self.peft_config.get(name) or self.default_peft_config())
        cfg["inference_mode"] = True
        with open(os.path.join(out_dir, "adapter_config.json"), "w") as f:
            json.dump(cfg, f, indent=2, sort_keys=True)

    def load_adapter(self, model_id: str, adapter_name: str, is_trainable: bool = False):
        """PeftModel.load_adapter contract used at qlora_model.py:84-88."""
        with open(os.path.join(model_id, "adapter_config.json")) as f:
            self.peft_config[adapter_name] = json.load(f)
        pc = self.peft_config[adapter_name]
        r = pc.get("r")
        if r != self.cfg.lora_r:
            raise OpadpoError(f"adapter rank {r} != engine lora_r {self.cfg.lora_r}")
        if "lora_alpha" in pc and float(pc["lora_alpha"]) != float(self.cfg.lora_alpha):
            raise OpadpoError(f"adapter lora_alpha {pc['lora_alpha']} != engine lora_alpha {self.cfg.lora_alpha} (the LoRA scale "
                              "alpha / r is one engine-wide constant)")
        if is_trainable and float(pc.get("lora_dropout") or 0.0) != 0.0:
            raise OpadpoError(f"adapter lora_dropout {pc['lora_dropout']}: the engine trains without LoRA dropout (the recipes use "
                              "0.0, llava_dpo.yaml:128, llava_opa.yaml:20)")
        if pc.get("bias", "none") != "none":
            raise OpadpoError(f"adapter bias={pc['bias']!r}: only bias='none' adapters exist in this path (qlora_model.py:244-251)")
        state = torch.load(os.path.join(model_id, "adapter_model.bin"), map_location="cpu")
        self.load_adapter_state(adapter_name, state)
        if is_trainable:
            self.make_trainable(adapter_name)

    # ---- activation slots
    def _take_slot(self) -> int:
        for slot in (0, 1):
            if not self._slot_busy[slot]:
                self._slot_busy[slot] = True
                self._slot_gen[slot] += 1
                return slot
        raise OpadpoError("both activation slots hold a grad-enabled forward that has not been differentiated yet (the engine "
                          "keeps two: CoPO's / the OPA regulariser's pair). Run backward on one of them, wrap forwards that "
                          "need no gradient in torch.no_grad(), or call release_slots()")

    def _drop_slot(self, slot: int, gen: int):
        if self._slot_busy[slot] and self._slot_gen[slot] == gen:
            self._slot_busy[slot] = False

    def release_slots(self):
        """Hands both slots back (after a failed step; B200AdamW.zero_grad calls it). A backward that arrives later for a
        forward made before this call raises instead of differentiating somebody else's activations."""
        for slot in (0, 1):
            if self._slot_busy[slot]:
                self._slot_busy[slot] = False
                self._slot_gen[slot] += 1

    # ---- forward used by the policy wrapper
    def policy_logprobs(self, images, queries, qmask, responses, temperature, img_mask=None) -> Tuple[Tensor, Tensor]:
        aid = ADAPTER_IDS[self.active_adapter]
        eng = self.engine
        dev = eng.device
        images, queries, qmask = images.to(dev), queries.to(dev), qmask.to(dev)
        responses = [r.to(dev) for r in responses]
        if img_mask is not None:
            img_mask = img_mask.to(dev)
        if ((queries == self.cfg.image_token).sum(1) != 1).any():
            raise OpadpoError("every query must hold exactly one image token (-200)")
        needs_grad = torch.is_grad_enabled() and self.flat_param is not None and aid == eng_trainable(self)
        if needs_grad:
            if len(responses) > 16:
                raise OpadpoError("at most 16 responses per call")
            slot = self._take_slot()
            return _PolicyForward.apply(self.flat_param, self, aid, images, queries, qmask, img_mask, float(temperature), slot,
                                        *responses)
        with torch.no_grad():
            return eng.forward_logprobs(aid, images, queries, qmask, responses, temperature, img_mask=img_mask)

    # ---- rollout
    @torch.no_grad()
    def generate(self, inputs: Tensor, images: Tensor, attention_mask: Tensor, do_sample: bool = False,
                 max_new_tokens: int = 896, pad_token_id: Optional[int] = None, suppress_tokens=None, top_p: float = 1.0,
                 top_k: int = 0, temperature: float = 1.0, num_return_sequences: int = 1, synced_gpus: bool = False,
                 **_unused) -> Tensor:
        """model.generate surface used by online_generator.py:292-310 and rl_models.py:166-188 -> LongTensor
        [b, Q + max_new_tokens] (finished rows padded). `synced_gpus` is ignored: ranks decode independently.
        Greedy ids reproduce the reference exactly; sampling follows HF's warper order with a counter-based RNG seeded from
        torch.initial_seed() (same distribution, not the same random stream as torch.multinomial)."""
        if num_return_sequences != 1:
            # [UPSTREAM transformers generate()]: every prompt row is repeated num_return_sequences times in place
            # (row 0's copies first); greedy decoding of identical copies is rejected there as well. The copies draw from
            # different random streams because the sampler's counter RNG is keyed by (seed, row, position).
            if not do_sample:
                raise ValueError("Greedy methods without beam search do not support `num_return_sequences` different than 1 "
                                 f"(got {num_return_sequences}).")
            n = int(num_return_sequences)
            inputs, images, attention_mask = (t.repeat_interleave(n, dim=0) for t in (inputs, images, attention_mask))
        if pad_token_id is not None and int(pad_token_id) != self.cfg.pad_id:
            raise OpadpoError(f"generate(pad_token_id={pad_token_id}): the engine was built with pad_id {self.cfg.pad_id} "
                              "(EngineConfig.pad_id; the reference pads with unk = 0, opadpo_train.py:696-698)")
        self._gen_calls = getattr(self, "_gen_calls", 0) + 1
        seed = (int(torch.initial_seed()) * 1000003 + self._gen_calls) & (2 ** 63 - 1)
        aid = ADAPTER_IDS[self.active_adapter]
        return self.engine.generate(aid, images, inputs, attention_mask, max_new_tokens, do_sample=do_sample, top_k=top_k,
                                    top_p=top_p, temperature=temperature, seed=seed, suppress_tokens=suppress_tokens)


def eng_trainable(model: B200LlavaModel) -> int:
    return POLICY if model.flat_param is not None else -2


class Policy(nn.Module):
    """rl_models.py:15-71."""

    def __init__(self, args, base_model: B200LlavaModel, base_tokenizer, adapter_name: Optional[str] = None):
        super().__init__()
        self.args = args
        self.base_model = base_model
        self.base_tokenizer = base_tokenizer
        self.adapter_name = adapter_name

    def respond(self, queries, query_attn_masks, images=None, temperature=None, num_return_sequences=1):
        assert not self.training, "Policy must be in eval model for generation."
        return self._post_respond(self._respond(queries, query_attn_masks, images, temperature, num_return_sequences))

    def _post_respond(self, respond_outputs):
        return respond_outputs


class AutoregressivePolicy(Policy):
    def forward(self, images: Tensor, queries: Tensor, queries_attn_masks: Tensor, temperature: Optional[float] = None,
                **kwargs) -> Dict[str, Tensor]:
        """rl_models.py:75-144."""
        if self.adapter_name is not None:
            self.base_model.set_adapter(self.adapter_name)
        self.base_model.config.use_cache = False
        if temperature is None:
            temperature = self.args.temperature
        keys = [k for k in kwargs if "response" in k and "_mask" not in k and "scores" not in k and "image_relations" not in k]
        if not keys:
            raise ValueError("no response tensors passed")
        R = kwargs[keys[0]].size(1)
        if R != self.args.response_len:
            raise ValueError(f"responses must be padded to args.response_len={self.args.response_len}, got {R}")
        Q = queries.size(1)
        img_mask = None
        qmask = queries_attn_masks
        if queries_attn_masks.size(1) != Q:   # CoPO 'attention': image-token mask prepended (rl_models.py:103-105)
            P = queries_attn_masks.size(1) - Q
            img_mask, qmask = queries_attn_masks[:, :P], queries_attn_masks[:, P:]
        lp, ent = self.base_model.policy_logprobs(images, queries, qmask, [kwargs[k] for k in keys], temperature, img_mask)
        out = {}
        for i, k in enumerate(keys):
            out[k + "_logprobs"] = lp[i]
            out[k + "_entropies"] = ent[i]
        return out

    def _respond(self, queries, query_attn_masks, images=None, temperature=None, num_return_sequences=1):
        """rl_models.py:146-191: sampling with top_k=0 / top_p=1 at `temperature`."""
        if self.adapter_name is not None:
            self.base_model.set_adapter(self.adapter_name)
        self.base_model.config.use_cache = True
        self.base_model.config.cache_shape = (queries.shape[-1] + self.args.response_len
                                              + self.base_model.get_vision_tower().num_patches,)   # rl_models.py:158-162
        if temperature is None:
            temperature = self.args.temperature
        seq = self.base_model.generate(inputs=queries, images=images, attention_mask=query_attn_masks, do_sample=True,
                                       max_new_tokens=self.args.response_len, top_p=1.0, top_k=0, temperature=temperature,
                                       pad_token_id=self.base_tokenizer.pad_token_id, num_return_sequences=num_return_sequences,
                                       suppress_tokens=([self.base_tokenizer.eos_token_id]
                                                        if getattr(self.args, "suppress_eos_at_generation", False) else None))
        return dict(responses=seq[:, queries.size(1):])     # [bsz * num_return_sequences, response_len], already right-padded


def make_policy_with_base_model(args, base_model, base_tokenizer, adapter_name: Optional[str] = "default"):
    """rl_models.py:193-204."""
    return AutoregressivePolicy(args, base_model, base_tokenizer, adapter_name=adapter_name)


def make_models(base_model, tokenizer, args, accelerator=None, num_new_tokens: int = 0,
                resume_from_checkpoint: Optional[str] = None, model_factory=None) -> dict:
    """DPOTrainer's model construction (dpo_trainer.py:958-1038 over load_nbit_model_for_inference /
    get_accelerate_model, qlora_model.py:44-127, 297-356), same signature and return dict `{policy, ref_policy}`:

    * the adapter directory is `args.policy_model_name_or_path` resolved through `[adapter_model/][lora_default/]`; the
      policy adapter of a resumed run comes from `<resume_from_checkpoint>/adapter_model/lora_policy` instead (:983-987);
    * the base checkpoint is the one `adapter_config.json` names (`base_model_name_or_path`, qlora_model.py:327-328), the
      LoRA rank / alpha are the ones it records; ONE base is shared by both adapters (REGISTERED_BASE_MODELS);
    * `args.train_from_sft` (default True): both adapters start from the checkpoint's weights, `lora_policy` trainable and
      `lora_ref_policy` frozen; False: both are freshly initialised with the recorded LoRA config (qlora_model.py:92-113);
    * the vision tower and the projector are frozen, LoRA tensors included (:1016-1029; `freeze_projector_lora`), then their
      base tensors are re-read from `args.base_model_name`
      (`loading_vision_tower_parameter`, opadpo_train.py:539-557), when that is another directory than the base.
    `base_model`: an already-loaded B200LlavaModel to reuse (None builds one on `accelerator.device`); there is no
    `accelerator.prepare` — data parallelism is dist.py's. New tokens (smart_tokenizer_and_embedding_resize, :934-955) are
    refused: the engine's vocabulary is the checkpoint's and its embeddings are frozen."""
    from . import checkpoint
    if num_new_tokens > 0:
        raise OpadpoError("make_models: num_new_tokens > 0 needs trainable, resized embeddings; the engine trains LoRA tensors only")
    if getattr(args, "vision_tower", "different") != "different":
        raise NotImplementedError("make_models: only vision_tower == 'different' exists (dpo_trainer.py:1016-1031)")
    adapter_dir = checkpoint.resolve_adapter_dir(args.policy_model_name_or_path)
    policy_dir = adapter_dir
    if resume_from_checkpoint:
        policy_dir = os.path.join(resume_from_checkpoint, "adapter_model", "lora_policy")
    with open(os.path.join(adapter_dir, "adapter_config.json")) as f:
        lora_cfg = json.load(f)
    base_path = lora_cfg["base_model_name_or_path"]
    if base_model is None:
        cfg = checkpoint.engine_config_from_hf(base_path, lora_r=int(lora_cfg["r"]), lora_alpha=float(lora_cfg["lora_alpha"]))
        factory = model_factory if model_factory is not None else B200LlavaModel
        base_model = factory(cfg, device=getattr(accelerator, "device", None), base_model_name_or_path=base_path)
        checkpoint.load_llava_base(base_model.engine, base_path)
    if getattr(args, "train_from_sft", True):
        base_model.load_adapter(adapter_dir, "lora_ref_policy", is_trainable=False)
        base_model.load_adapter(policy_dir, "lora_policy", is_trainable=True)
    else:
        base_model.add_adapter("lora_ref_policy", is_trainable=False, peft_config=lora_cfg)
        if resume_from_checkpoint:
            base_model.load_adapter(policy_dir, "lora_policy", is_trainable=True)
        else:
            base_model.add_adapter("lora_policy", is_trainable=True, peft_config=lora_cfg)
    base_model.freeze_projector_lora()       # :1019-1029; the tower itself runs without gradients in the engine
    tower_src = getattr(args, "base_model_name", None)
    if tower_src and os.path.abspath(tower_src.rstrip("/")) != os.path.abspath(base_path.rstrip("/")):
        for name, t in checkpoint.iter_checkpoint_tensors(tower_src, lambda n: "vision_tower" in n or "mm_projector" in n):
            try:
                base_model.engine.load_tensor(name, t)
            except OpadpoUnknownName:
                pass          # a tensor the engine does not execute (e.g. the tower's last layer): the reference skips unknown names too
    policy = make_policy_with_base_model(args, base_model, tokenizer, adapter_name="lora_policy")
    ref_policy = make_policy_with_base_model(args, base_model, tokenizer, adapter_name="lora_ref_policy")
    return dict(policy=policy, ref_policy=ref_policy)
