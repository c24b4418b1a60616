"""Generates tests/golden/opadpo_tiny.npz by running the REFERENCE's own Python for the hot path.

Run in the build container only (needs /root/reference; the GPU box never runs this):
    python oracle/make_golden.py

What is executed verbatim from /root/reference (imported with sys.modules stubs for the packages that are not
installed here: llava, peft, accelerate, bitsandbytes):
  * opadpo.dpo_models.rl_models.AutoregressivePolicy.forward          (rl_models.py:75-144)
  * utils.common_utils.compute_logprobs                                (common_utils.py:112-118)
  * opadpo.dpo_models.dpo_trainer.DPOTrainer.dpo_loss / compute_policy_loss  (dpo_trainer.py:429-802)
  * opadpo.generator_models.generator.truncate_after_eos_with_padding (generator.py:244-273)
What stands in for the un-vendored model stack: HuggingFace transformers' LlamaForCausalLM and CLIPVisionModel
(this image: transformers 5.5; reference pin 4.34.1) holding the oracle's seeded weights, a PEFT-0.5.0-equivalent
LoRA Linear wrapper, and the LLaVA multimodal splice restated from llava_setup/llava_modifications.patch:15022-15055.
The greedy rollout golden uses HF's own KV cache (DynamicCache) in a manual greedy loop.

The vectors pin oracle/llava_oracle.py (tests/test_oracle_golden.py) — an independent code path from the oracle's
hand-written LLaMA/CLIP arithmetic.
"""
from __future__ import annotations

import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn
import transformers  # must be imported before the stubs
import transformers.trainer_utils  # noqa: F401  (probes for peft/accelerate at import time)
from transformers import CLIPVisionConfig, CLIPVisionModel, LlamaConfig, LlamaForCausalLM  # noqa: F401
from transformers import Seq2SeqTrainingArguments, TrainingArguments  # noqa: F401  (resolve lazy imports pre-stub)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
from oracle import llava_oracle as O  # noqa: E402


def _stub(name, **attrs):
    import importlib.machinery
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)   # transformers probes find_spec() on these names
    m.__path__ = []
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs():
    class _Any:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, k):
            return _Any()

        def __call__(self, *a, **k):
            return _Any()

    _stub("llava", conversation=_Any(), mm_utils=_Any(), model=_Any())
    _stub("llava.conversation", default_conversation=_Any(), conv_templates={}, SeparatorStyle=_Any())
    _stub("llava.train")
    _stub("llava.train.train", DataArguments=_Any)
    _stub("llava.mm_utils", tokenizer_image_token=lambda *a, **k: None)
    _stub("llava.model", LlavaLlamaForCausalLM=_Any)
    _stub("llava.constants", IGNORE_INDEX=-100, IMAGE_TOKEN_INDEX=-200, DEFAULT_IMAGE_TOKEN="<image>")
    acc = _stub("accelerate", Accelerator=_Any, DistributedType=_Any())
    _stub("accelerate.optimizer", AcceleratedOptimizer=_Any)
    _stub("accelerate.utils", convert_outputs_to_fp32=lambda f: f, DistributedType=_Any(), set_seed=lambda *a, **k: None)
    _stub("accelerate.accelerator", AcceleratedOptimizer=_Any)
    acc.accelerator = sys.modules["accelerate.accelerator"]
    acc.optimizer = sys.modules["accelerate.optimizer"]
    acc.utils = sys.modules["accelerate.utils"]
    peft = _stub("peft", PeftModel=_Any, LoraConfig=_Any, get_peft_model=_Any(), PeftModelForCausalLM=_Any,
                 prepare_model_for_kbit_training=_Any(), prepare_model_for_int8_training=_Any())
    _stub("peft.utils", WEIGHTS_NAME="adapter_model.bin", get_peft_model_state_dict=_Any(), CONFIG_NAME="adapter_config.json")
    _stub("peft.tuners")
    _stub("peft.tuners.lora", LoraLayer=_Any)
    peft.utils = sys.modules["peft.utils"]
    _stub("bitsandbytes", optim=_Any(), nn=_Any())
    _stub("deepspeed")
    sys.path.insert(0, REF)


class LoRALinear(nn.Module):
    """peft 0.5.0 lora.Linear.forward [UPSTREAM]: result = F.linear(x, W, b) + lora_B(lora_A(dropout(x))) * scaling."""

    def __init__(self, base: nn.Linear, adapters, scaling, owner):
        super().__init__()
        self.base, self.adapters, self.scaling = base, adapters, scaling
        self._owner = [owner]   # in a list: must not register the parent as a submodule (cycle)

    def forward(self, x):
        y = self.base(x)
        ad = self.adapters.get(self._owner[0].active)
        if ad is not None:
            A, B = ad
            y = y + (x @ A.t()) @ B.t() * self.scaling
        return y


class HFLlava(nn.Module):
    """LlavaLlamaForCausalLM as the reference sees it (llava_modifications.patch:12127-12202), built on HF modules."""

    def __init__(self, cfg: O.OracleConfig, w, adapters):
        super().__init__()
        from transformers import CLIPVisionConfig, CLIPVisionModel, LlamaConfig, LlamaForCausalLM
        self.cfg, self.active = cfg, None
        self.config = SimpleNamespace(use_cache=False)
        lc = LlamaConfig(hidden_size=cfg.d_model, intermediate_size=cfg.d_ffn, num_hidden_layers=cfg.n_layers,
                         num_attention_heads=cfg.n_heads, num_key_value_heads=cfg.n_heads, vocab_size=cfg.vocab,
                         rms_norm_eps=cfg.rms_eps, rope_theta=cfg.rope_theta, max_position_embeddings=cfg.max_positions,
                         attention_bias=False, tie_word_embeddings=False, attn_implementation="eager")
        self.llm = LlamaForCausalLM(lc).float().eval()
        vc = CLIPVisionConfig(hidden_size=cfg.v_d, intermediate_size=cfg.v_ffn, num_hidden_layers=cfg.v_layers + 1,
                              num_attention_heads=cfg.v_heads, image_size=cfg.v_image, patch_size=cfg.v_patch,
                              hidden_act="quick_gelu", layer_norm_eps=cfg.v_eps, attn_implementation="eager")
        self.clip = CLIPVisionModel(vc).float().eval()
        self.proj = nn.Sequential(nn.Linear(cfg.v_d, cfg.d_model), nn.GELU(), nn.Linear(cfg.d_model, cfg.d_model))
        # ---- weights
        sd = {k: v for k, v in w.items() if k.startswith("model.layers") or k in ("model.embed_tokens.weight", "model.norm.weight", "lm_head.weight")}
        missing, unexpected = self.llm.load_state_dict(sd, strict=False)
        assert not unexpected and all("rotary" in m for m in missing), (missing, unexpected)
        csd = {k[len("model.vision_tower.vision_tower."):]: v for k, v in w.items() if k.startswith(O.VT)}
        missing, unexpected = self.clip.load_state_dict(csd, strict=False)
        assert not unexpected, unexpected   # the never-executed last layer + post_layernorm keep their random init
        self.proj[0].weight.data.copy_(w["model.mm_projector.0.weight"]); self.proj[0].bias.data.copy_(w["model.mm_projector.0.bias"])
        self.proj[2].weight.data.copy_(w["model.mm_projector.2.weight"]); self.proj[2].bias.data.copy_(w["model.mm_projector.2.bias"])
        # ---- LoRA on every Linear except lm_head (opa_train.py:177-190)
        scaling = cfg.lora_alpha / cfg.lora_r

        def wrap(root, prefix_map):
            for name, mod in list(root.named_modules()):
                if isinstance(mod, nn.Linear) and "lm_head" not in name:
                    key = prefix_map(name)
                    ads = {an: ad[key] for an, ad in adapters.items() if key in ad}
                    if not ads:
                        continue
                    parent = root.get_submodule(name.rsplit(".", 1)[0]) if "." in name else root
                    setattr(parent, name.rsplit(".", 1)[-1], LoRALinear(mod, ads, scaling, self))

        wrap(self.llm, lambda n: n)
        wrap(self.clip, lambda n: "model.vision_tower.vision_tower." + n)
        wrap(self.proj, lambda n: "model.mm_projector." + n)

    def set_adapter(self, name):
        self.active = name

    def prepare_inputs_for_generation(self, input_ids, attention_mask=None, images=None, use_cache=None, **kw):
        return dict(input_ids=input_ids, attention_mask=attention_mask, images=images, use_cache=use_cache)

    def encode_images(self, images):
        with torch.no_grad():   # CLIPVisionTower.forward is decorated with torch.no_grad (patch:15185-15217)
            hs = self.clip(pixel_values=images, output_hidden_states=True).hidden_states[-2][:, 1:]
        return self.proj(hs)

    def splice(self, input_ids, attention_mask, images):
        """prepare_inputs_labels_for_multimodal (patch:15022-15055): one image per row, features inserted at -200."""
        feats = self.encode_images(images)
        emb = self.llm.get_input_embeddings()
        xs, ms = [], []
        P = feats.shape[1]
        extra = attention_mask.shape[1] - input_ids.shape[1]   # CoPO 'attention': image-token mask prepended
        for n in range(input_ids.shape[0]):
            p = int((input_ids[n] == -200).nonzero()[0])
            te = emb(input_ids[n].clamp(min=0))
            xs.append(torch.cat([te[:p], feats[n], te[p + 1:]], 0))
            if extra:
                im, tm = attention_mask[n, :extra].bool(), attention_mask[n, extra:].bool()
            else:
                im, tm = torch.ones(P, dtype=torch.bool), attention_mask[n].bool()
            ms.append(torch.cat([tm[:p], im, tm[p + 1:]]))
        return torch.stack(xs), torch.stack(ms)

    def forward(self, input_ids=None, attention_mask=None, images=None, use_cache=None, output_hidden_states=False, **kw):
        x, m = self.splice(input_ids, attention_mask, images)
        out = self.llm(inputs_embeds=x, attention_mask=m.long(), use_cache=False, output_hidden_states=output_hidden_states)
        return out


def main():
    install_stubs()
    import opadpo.dpo_models.rl_models as rl_models
    import opadpo.dpo_models.dpo_trainer as dpo_trainer
    from opadpo.generator_models.generator import truncate_after_eos_with_padding
    import utils.common_utils as common_utils

    torch.manual_seed(0)
    torch.set_num_threads(4)
    cfg = O.OracleConfig()
    w = O.init_weights(cfg)
    lp, lr = O.init_lora(cfg, 11), O.init_lora(cfg, 12)
    model = HFLlava(cfg, w, {"lora_policy": lp, "lora_ref_policy": lr})
    Q, R, b = 16, 24, 2
    names3 = ["standard_response", "original_generate_response", "AI_pseudo_response"]
    images, q, qm, resp = O.synthetic_batch(cfg, b, Q, R, names3, left_pad=2)
    args = SimpleNamespace(temperature=0.7, response_len=R, detailed_report=True, response_score=True,
                           response_image_relation=True, CoPO=True, CoPO_method="attention", CoPO_coef=0.2, AncPO=True,
                           mDPO_anchor=True, Anchor_value=0.0, Anchor_coef=1.0, standard_pair_coef=1.0, AI_pair_coef=1.0)
    tok = SimpleNamespace(pad_token_id=0, eos_token_id=2)
    policy = rl_models.AutoregressivePolicy(args, model, tok, adapter_name="lora_policy")
    ref_policy = rl_models.AutoregressivePolicy(args, model, tok, adapter_name="lora_ref_policy")
    g = torch.Generator().manual_seed(9)
    img_mask = torch.rand(b, cfg.n_patches, generator=g) > 0.3
    masked_qm = torch.cat([img_mask, qm], 1)
    scores = {}
    for n in ("original_generate_response", "AI_pseudo_response"):
        scores[n + "_scores"] = torch.tensor([1.0, 1.5, 2.0, 2.5])[torch.randint(0, 4, (b, R), generator=g)] * (resp[n] != 0)
        scores[n + "_image_relations"] = torch.tensor([1.0, 3.0])[torch.randint(0, 2, (b, R), generator=g)] * (resp[n] != 0)
    mresp = {"mask_standard_response": resp["standard_response"], "mask_AI_pseudo_response": resp["AI_pseudo_response"]}

    out = {}
    with torch.no_grad():
        ref_out = ref_policy(images, q, qm, **resp)
        ref_m = ref_policy(images, q, masked_qm, **mresp)
        pol_out = policy(images, q, qm, **resp)
        pol_m = policy(images, q, masked_qm, **mresp)
    for tag, d in (("ref", ref_out), ("refm", ref_m), ("pol", pol_out), ("polm", pol_m)):
        for k, v in d.items():
            out[f"{tag}.{k}"] = v.numpy()

    # ---- DPOTrainer.compute_policy_loss, verbatim, on a trainer shell (no accelerate / dataloaders needed for the loss)
    tr = object.__new__(dpo_trainer.DPOTrainer)
    nn.Module.__init__(tr) if isinstance(tr, nn.Module) else None
    tr.accelerator = SimpleNamespace(device=torch.device("cpu"), num_processes=1)
    tr.args, tr.tokenizer, tr.policy = args, tok, policy
    tr.reference_free, tr.beta, tr.label_smoothing, tr.loss_type = False, 0.1, 0.0, "sigmoid"
    tr.f_divergence_type, tr.f_divergence_params = "reverse_kl", None
    rollouts = dict(queries=q, queries_attn_masks=qm, images=images, masked_query_attn_masks=masked_qm, **resp, **scores)
    rollouts.update({"ref_base_" + k: v for k, v in ref_out.items()})
    rollouts.update({"ref_" + k: v for k, v in ref_m.items()})
    loss, stats = tr.compute_policy_loss(rollouts)
    out["cpl.loss"] = loss.detach().numpy()
    for k, v in stats.items():
        out["cpl.stats." + k] = v.numpy()
    # LoRA gradients of that loss (policy adapter): the backward the engine must reproduce
    for t in list(lp.values()):
        for x in t:
            x.requires_grad_(True)
    loss2, _ = tr.compute_policy_loss(rollouts)
    loss2.backward()
    for m in ("model.layers.0.self_attn.q_proj", "model.layers.1.mlp.down_proj", "model.layers.1.mlp.gate_proj",
              "model.mm_projector.0", "model.mm_projector.2"):
        out[f"grad.{m}.A"] = lp[m][0].grad.numpy().copy()
        out[f"grad.{m}.B"] = lp[m][1].grad.numpy().copy()
    for t in list(lp.values()):
        for x in t:
            x.requires_grad_(False)
            x.grad = None

    # ---- dpo_loss variants, verbatim
    pc, pr = pol_out["standard_response_logprobs"], pol_out["original_generate_response_logprobs"]
    rc, rr = ref_out["standard_response_logprobs"], ref_out["original_generate_response_logprobs"]
    for div, ls in (("reverse_kl", 0.0), ("reverse_kl", 0.1), ("js_divergence", 0.0), ("alpha_divergence", 0.0)):
        tr.f_divergence_type, tr.label_smoothing = div, ls
        l, c, j = tr.dpo_loss(pc, pr, rc, rr, scores["AI_pseudo_response_scores"], scores["original_generate_response_scores"])
        out[f"dpo.{div}.{ls}.losses"], out[f"dpo.{div}.{ls}.chosen"], out[f"dpo.{div}.{ls}.rejected"] = l.numpy(), c.numpy(), j.numpy()

    # ---- greedy rollout with HF's own KV cache, peaked lm_head
    wp = O.init_weights(cfg, peaked=40.0)
    gen_model = HFLlava(cfg, wp, {"lora_policy": lp})
    gen_model.set_adapter("lora_policy")
    max_new = 12
    with torch.no_grad():
        x, m = gen_model.splice(q, qm, images)
        o = gen_model.llm(inputs_embeds=x, attention_mask=m.long(), use_cache=True)
        past, ids, unfinished = o.past_key_values, [q], torch.ones(b, dtype=torch.long)
        logits = o.logits[:, -1]
        for step in range(max_new):
            nxt = logits.argmax(-1) * unfinished + tok.pad_token_id * (1 - unfinished)
            ids.append(nxt[:, None])
            unfinished = unfinished * (nxt != tok.eos_token_id).long()
            if step == max_new - 1:
                break
            m = torch.cat([m, torch.ones(b, 1, dtype=torch.bool)], 1)
            o = gen_model.llm(inputs_embeds=gen_model.llm.get_input_embeddings()(nxt)[:, None], attention_mask=m.long(),
                              past_key_values=past, use_cache=True)
            past, logits = o.past_key_values, o.logits[:, -1]
    out["gen.ids"] = torch.cat(ids, 1).numpy()

    # ---- truncate_after_eos_with_padding, verbatim
    comp = torch.tensor([[5, 9, 2, 7, 7, 1577, 3, 0], [4, 4, 4, 4, 4, 4, 4, 4], [8, 29973, 6, 2, 1577, 9, 9, 9], [2, 0, 0, 0, 0, 0, 0, 0]])
    out["trunc.in"] = comp.numpy()
    out["trunc.out"] = truncate_after_eos_with_padding(comp, 2, 0, additional_tokens=[1577, 29973]).numpy()
    out["trunc.out_noextra"] = truncate_after_eos_with_padding(comp, 2, 0).numpy()
    # ---- compute_logprobs, verbatim
    gl = torch.Generator().manual_seed(4)
    lg_, lb_ = torch.randn(2, 5, 11, generator=gl), torch.randint(0, 11, (2, 5), generator=gl)
    lb_[0, 3:] = 0
    out["clp.logits"], out["clp.labels"] = lg_.numpy(), lb_.numpy()
    out["clp.out"] = common_utils.compute_logprobs(lg_, lb_, ignore_index=0).numpy()

    # inputs (so the test does not depend on RNG stream stability)
    out["in.images"], out["in.queries"], out["in.qmask"] = images.numpy(), q.numpy(), qm.numpy()
    out["in.img_mask"] = img_mask.numpy()
    for k, v in {**resp, **scores}.items():
        out["in." + k] = v.numpy()
    path = os.path.join(ROOT, "tests", "golden", "opadpo_tiny.npz")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays; transformers", transformers.__version__)


if __name__ == "__main__":
    main()
