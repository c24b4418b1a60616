"""CPU fp32 oracle of the OPA-DPO hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py and tools/bench_*.py may import
this module;
the product (opa_dpo_b200/) never does and has no CPU path.

Parity pinning: the reference repository ships NO tests, golden vectors or fixtures for this path (SURVEY.md §4,
§8c), and the arithmetic itself lives in un-vendored third-party packages (LLaVA@817a4af + the repo's patch,
transformers 4.34.1, peft 0.5.0).  This restatement is therefore pinned against outputs of the reference itself:
oracle/make_golden.py imports the reference's own rl_models.AutoregressivePolicy / DPOTrainer.dpo_loss /
compute_policy_loss / common_utils.compute_logprobs / generator.truncate_after_eos_with_padding verbatim from
/root/reference, drives them over HuggingFace transformers' LlamaForCausalLM + CLIPVisionModel holding the same
weights, and commits the resulting vectors under tests/golden/ (tests/test_oracle_golden.py checks them).

Every function cites the reference file:line (or the [UPSTREAM] module) whose behaviour it restates.
Plain torch on CPU in float32; no custom kernels, no dependency on transformers.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

IGNORE_INDEX = -100        # utils/constants.py:20
IMAGE_TOKEN_INDEX = -200   # utils/constants.py:21

VT = "model.vision_tower.vision_tower.vision_model."


@dataclass
class OracleConfig:
    d_model: int = 256
    n_layers: int = 2
    n_heads: int = 2
    d_ffn: int = 512
    vocab: int = 512
    rms_eps: float = 1e-5
    rope_theta: float = 10000.0
    max_positions: int = 4096
    v_d: int = 128
    v_layers: int = 2          # encoder layers EXECUTED (select_layer -2 of v_layers+1)
    v_heads: int = 2
    v_ffn: int = 256
    v_image: int = 56
    v_patch: int = 14
    v_eps: float = 1e-5
    lora_r: int = 8
    lora_alpha: float = 16.0
    pad_id: int = 0
    eos_id: int = 2
    image_token: int = IMAGE_TOKEN_INDEX

    @property
    def n_patches(self) -> int:
        return (self.v_image // self.v_patch) ** 2

    def to_dict(self):
        return asdict(self)


LLAVA_7B = dict(d_model=4096, n_layers=32, n_heads=32, d_ffn=11008, vocab=32000, v_d=1024, v_layers=23, v_heads=16,
                v_ffn=4096, v_image=336, v_patch=14, lora_r=256, lora_alpha=512.0)
LLAVA_13B = dict(d_model=5120, n_layers=40, n_heads=40, d_ffn=13824, vocab=32000, v_d=1024, v_layers=23, v_heads=16,
                 v_ffn=4096, v_image=336, v_patch=14, lora_r=256, lora_alpha=512.0)


def _bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


# ---------------------------------------------------------------------------------------------------- weights
def lora_target_modules(cfg: OracleConfig, vision: bool = True) -> List[str]:
    """Every nn.Linear except lm_head (opadpo/opa_train.py:177-190): LLaMA, projector and the CLIP tower."""
    mods = []
    for i in range(cfg.n_layers):
        for m in ("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.o_proj",
                  "mlp.gate_proj", "mlp.up_proj", "mlp.down_proj"):
            mods.append(f"model.layers.{i}.{m}")
    mods += ["model.mm_projector.0", "model.mm_projector.2"]
    if vision:
        for i in range(cfg.v_layers):
            for m in ("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.out_proj",
                      "mlp.fc1", "mlp.fc2"):
                mods.append(f"{VT}encoder.layers.{i}.{m}")
    return mods


def linear_shapes(cfg: OracleConfig) -> Dict[str, Tuple[int, int]]:
    """module -> (out_features, in_features)."""
    d, f, vd, vf = cfg.d_model, cfg.d_ffn, cfg.v_d, cfg.v_ffn
    s = {}
    for i in range(cfg.n_layers):
        p = f"model.layers.{i}."
        for m in ("q_proj", "k_proj", "v_proj", "o_proj"):
            s[p + "self_attn." + m] = (d, d)
        s[p + "mlp.gate_proj"] = (f, d)
        s[p + "mlp.up_proj"] = (f, d)
        s[p + "mlp.down_proj"] = (d, f)
    s["model.mm_projector.0"] = (d, vd)
    s["model.mm_projector.2"] = (d, d)
    for i in range(cfg.v_layers):
        p = f"{VT}encoder.layers.{i}."
        for m in ("q_proj", "k_proj", "v_proj", "out_proj"):
            s[p + "self_attn." + m] = (vd, vd)
        s[p + "mlp.fc1"] = (vf, vd)
        s[p + "mlp.fc2"] = (vd, vf)
    return s


def init_weights(cfg: OracleConfig, seed: int = 1234, std: float = 0.02, peaked: float = 1.0) -> Dict[str, torch.Tensor]:
    """Seeded random-init base weights: N(0, std^2) for linears/embeddings, norms = 1.
    Values are rounded to bf16 so the bf16 engine and this fp32 oracle hold IDENTICAL numbers.
    `peaked` scales lm_head so greedy top-1 margins exceed bf16 noise (rollout id parity)."""
    g = torch.Generator().manual_seed(seed)
    w: Dict[str, torch.Tensor] = {}

    def rnd(*shape, s=std):
        return _bf16_round(torch.randn(*shape, generator=g) * s)

    d = cfg.d_model
    w["model.embed_tokens.weight"] = rnd(cfg.vocab, d)
    for name, (o, i) in linear_shapes(cfg).items():
        w[name + ".weight"] = rnd(o, i)
        if "vision_model" in name or "mm_projector" in name:
            w[name + ".bias"] = rnd(o)
    for i in range(cfg.n_layers):
        w[f"model.layers.{i}.input_layernorm.weight"] = torch.ones(d)
        w[f"model.layers.{i}.post_attention_layernorm.weight"] = torch.ones(d)
    w["model.norm.weight"] = torch.ones(d)
    w["lm_head.weight"] = rnd(cfg.vocab, d, s=std * peaked)
    vd = cfg.v_d
    w[VT + "embeddings.class_embedding"] = rnd(vd)
    w[VT + "embeddings.patch_embedding.weight"] = rnd(vd, 3, cfg.v_patch, cfg.v_patch)
    w[VT + "embeddings.position_embedding.weight"] = rnd(cfg.n_patches + 1, vd)
    w[VT + "pre_layrnorm.weight"] = torch.ones(vd)
    w[VT + "pre_layrnorm.bias"] = torch.zeros(vd)
    for i in range(cfg.v_layers):
        for ln in ("layer_norm1", "layer_norm2"):
            w[f"{VT}encoder.layers.{i}.{ln}.weight"] = torch.ones(vd)
            w[f"{VT}encoder.layers.{i}.{ln}.bias"] = torch.zeros(vd)
    return w


def init_lora(cfg: OracleConfig, seed: int, std: float = 0.02, vision: bool = True) -> Dict[str, Tuple[torch.Tensor, torch.Tensor]]:
    """LoRA A [r, in] and B [out, r], BOTH N(0, std^2) (non-zero B so the adapter path is exercised)."""
    g = torch.Generator().manual_seed(seed)
    shapes = linear_shapes(cfg)
    out = {}
    for m in lora_target_modules(cfg, vision):
        o, i = shapes[m]
        A = _bf16_round(torch.randn(cfg.lora_r, i, generator=g) * std)
        B = _bf16_round(torch.randn(o, cfg.lora_r, generator=g) * std)
        out[m] = (A, B)
    return out


# ---------------------------------------------------------------------------------------------------- model
def _lin(x, w, name, lora, scale):
    """peft 0.5.0 lora.Linear.forward [UPSTREAM]: W x (+ b) + scaling * B(A x); dropout 0."""
    y = x @ w[name + ".weight"].t()
    b = w.get(name + ".bias")
    if b is not None:
        y = y + b
    if lora is not None and name in lora:
        A, B = lora[name]
        y = y + scale * ((x @ A.t()) @ B.t())
    return y


def clip_features(cfg: OracleConfig, w, lora, images: torch.Tensor) -> torch.Tensor:
    """CLIPVisionTower.forward (llava_modifications.patch:15185-15217): hidden_states[select_layer][:, 1:], run under
    no_grad. HF CLIPVisionModel arithmetic [UPSTREAM]: conv patch embed (no bias) + CLS + pos-emb, pre_layrnorm,
    pre-LN encoder layers with quick_gelu."""
    scale = cfg.lora_alpha / cfg.lora_r if cfg.lora_r else 0.0
    n = images.shape[0]
    vd, H = cfg.v_d, cfg.v_heads
    hd = vd // H
    with torch.no_grad():
        pe = F.conv2d(images.float(), w[VT + "embeddings.patch_embedding.weight"], stride=cfg.v_patch)
        pe = pe.flatten(2).transpose(1, 2)                                   # [n, P, vd]
        cls = w[VT + "embeddings.class_embedding"].expand(n, 1, vd)
        x = torch.cat([cls, pe], dim=1) + w[VT + "embeddings.position_embedding.weight"][None]
        x = F.layer_norm(x, (vd,), w[VT + "pre_layrnorm.weight"], w[VT + "pre_layrnorm.bias"], cfg.v_eps)
        for i in range(cfg.v_layers):
            p = f"{VT}encoder.layers.{i}."
            h = F.layer_norm(x, (vd,), w[p + "layer_norm1.weight"], w[p + "layer_norm1.bias"], cfg.v_eps)
            q = _lin(h, w, p + "self_attn.q_proj", lora, scale) * hd ** -0.5
            k = _lin(h, w, p + "self_attn.k_proj", lora, scale)
            v = _lin(h, w, p + "self_attn.v_proj", lora, scale)
            S = x.shape[1]
            q, k, v = (t.view(n, S, H, hd).transpose(1, 2) for t in (q, k, v))
            a = torch.softmax(q @ k.transpose(-1, -2), dim=-1) @ v
            a = a.transpose(1, 2).reshape(n, S, vd)
            x = x + _lin(a, w, p + "self_attn.out_proj", lora, scale)
            h = F.layer_norm(x, (vd,), w[p + "layer_norm2.weight"], w[p + "layer_norm2.bias"], cfg.v_eps)
            h = _lin(h, w, p + "mlp.fc1", lora, scale)
            h = h * torch.sigmoid(1.702 * h)
            x = x + _lin(h, w, p + "mlp.fc2", lora, scale)
    return x[:, 1:]


def projector(cfg: OracleConfig, w, lora, feats: torch.Tensor) -> torch.Tensor:
    """build_vision_projector 'mlp2x_gelu' (llava_modifications.patch:15354-15367)."""
    scale = cfg.lora_alpha / cfg.lora_r if cfg.lora_r else 0.0
    h = _lin(feats, w, "model.mm_projector.0", lora, scale)
    return _lin(F.gelu(h), w, "model.mm_projector.2", lora, scale)


def splice(cfg: OracleConfig, w, input_ids: torch.Tensor, attention_mask: torch.Tensor, img_feats: torch.Tensor,
           img_mask: Optional[torch.Tensor] = None):
    """prepare_inputs_labels_for_multimodal (llava_modifications.patch:15022-15055, helpers :14618-14736):
    token embeddings with the P projected patch rows inserted at the IMAGE_TOKEN_INDEX slot; image rows get mask 1
    (or the CoPO image-token mask, patch:14636-14643,14672-14676)."""
    N, L = input_ids.shape
    P = img_feats.shape[1]
    embeds, masks = [], []
    for n in range(N):
        pos = (input_ids[n] == cfg.image_token).nonzero()
        assert pos.numel() == 1, "exactly one image token per row on the hot path"
        p = int(pos[0])
        ids = input_ids[n].clamp(min=0)
        te = w["model.embed_tokens.weight"][ids]
        embeds.append(torch.cat([te[:p], img_feats[n], te[p + 1:]], dim=0))
        im = torch.ones(P, dtype=torch.bool) if img_mask is None else img_mask[n].bool()
        masks.append(torch.cat([attention_mask[n, :p].bool(), im, attention_mask[n, p + 1:].bool()]))
    return torch.stack(embeds), torch.stack(masks)


def _rope_tables(cfg: OracleConfig, S: int, hd: int):
    """HF LlamaRotaryEmbedding [UPSTREAM]: inv_freq = theta^(-2j/hd), emb = cat(freqs, freqs)."""
    inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd))
    fr = torch.outer(torch.arange(S, dtype=torch.float32), inv)
    emb = torch.cat([fr, fr], dim=-1)
    return emb.cos(), emb.sin()


def _rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat([-x[..., h:], x[..., :h]], dim=-1)


def llama_hidden(cfg: OracleConfig, w, lora, x: torch.Tensor, mask: torch.Tensor, past=None, pos0: int = 0,
                 checkpoint_layers: bool = False):
    """HF LlamaModel [UPSTREAM transformers 4.34.1] over inputs_embeds with a key-padding mask; position_ids are
    arange over the PADDED row (no position_ids are passed by rl_models.py:114-120). Optional KV cache for rollout:
    `past` = list of (k, v) per layer, `mask` then covers past + new keys. `checkpoint_layers`: HF gradient checkpointing
    per decoder layer, which the reference trains with (opadpo_train.py:324-327, qlora_model.py:204-205) — same values and
    gradients, each layer's internals recomputed in the backward instead of kept (used by the full-depth CPU timing)."""
    scale = cfg.lora_alpha / cfg.lora_r if cfg.lora_r else 0.0
    N, S, d = x.shape
    H = cfg.n_heads
    hd = d // H
    Sk = mask.shape[1]
    cos, sin = _rope_tables(cfg, pos0 + S, hd)
    cos, sin = cos[pos0:], sin[pos0:]
    qi = torch.arange(pos0, pos0 + S)[:, None]
    kj = torch.arange(Sk)[None, :]
    allowed = (kj <= qi)[None, None] & mask.bool()[:, None, None, :]           # [N,1,S,Sk]
    new_past = []

    def layer(i: int, x: torch.Tensor) -> torch.Tensor:
        p = f"model.layers.{i}."
        h = _rms(x, w[p + "input_layernorm.weight"], cfg.rms_eps)
        q = _lin(h, w, p + "self_attn.q_proj", lora, scale).view(N, S, H, hd).transpose(1, 2)
        k = _lin(h, w, p + "self_attn.k_proj", lora, scale).view(N, S, H, hd).transpose(1, 2)
        v = _lin(h, w, p + "self_attn.v_proj", lora, scale).view(N, S, H, hd).transpose(1, 2)
        q = q * cos + _rotate_half(q) * sin
        k = k * cos + _rotate_half(k) * sin
        if past is not None:
            if past[i] is not None:
                k = torch.cat([past[i][0], k], dim=2)
                v = torch.cat([past[i][1], v], dim=2)
            new_past.append((k, v))
        sc = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
        sc = sc.masked_fill(~allowed, float("-inf"))
        pr = torch.softmax(sc, dim=-1)
        pr = torch.nan_to_num(pr, nan=0.0)     # fully masked (padding) query rows: flash-attn never computes them
        a = (pr @ v).transpose(1, 2).reshape(N, S, d)
        x = x + _lin(a, w, p + "self_attn.o_proj", lora, scale)
        h = _rms(x, w[p + "post_attention_layernorm.weight"], cfg.rms_eps)
        g = _lin(h, w, p + "mlp.gate_proj", lora, scale)
        u = _lin(h, w, p + "mlp.up_proj", lora, scale)
        return x + _lin(F.silu(g) * u, w, p + "mlp.down_proj", lora, scale)

    recompute = checkpoint_layers and past is None and torch.is_grad_enabled() and x.requires_grad
    for i in range(cfg.n_layers):
        if recompute:
            from torch.utils.checkpoint import checkpoint
            x = checkpoint(layer, i, x, use_reentrant=False)
        else:
            x = layer(i, x)
    x = _rms(x, w["model.norm.weight"], cfg.rms_eps)
    return (x, new_past) if past is not None else x


def _rms(x, weight, eps):
    """HF LlamaRMSNorm [UPSTREAM]."""
    var = x.pow(2).mean(-1, keepdim=True)
    return weight * (x * torch.rsqrt(var + eps))


def llava_logits(cfg: OracleConfig, w, lora, input_ids, attention_mask, images, img_mask=None, img_of_row=None,
                 checkpoint_layers: bool = False):
    """LlavaLlamaForCausalLM.forward, multimodal branch (llava_modifications.patch:12127-12202)."""
    feats = projector(cfg, w, lora, clip_features(cfg, w, lora, images))
    if img_of_row is not None:
        feats = feats[img_of_row]
    x, mask = splice(cfg, w, input_ids, attention_mask, feats, img_mask)
    h = llama_hidden(cfg, w, lora, x, mask, checkpoint_layers=checkpoint_layers)
    return h @ w["lm_head.weight"].t()


# ---------------------------------------------------------------------------------------------------- policy
def compute_logprobs(logits: torch.Tensor, labels: torch.Tensor, ignore_index: int) -> torch.Tensor:
    """utils/common_utils.py:112-118."""
    return -F.cross_entropy(logits.permute(0, 2, 1), labels, reduction="none", ignore_index=ignore_index)


def policy_forward(cfg: OracleConfig, w, lora, images, queries, queries_attn_masks, responses: Dict[str, torch.Tensor],
                   temperature: float = 1.0, checkpoint_layers: bool = False) -> Dict[str, torch.Tensor]:
    """AutoregressivePolicy.forward (opadpo/dpo_models/rl_models.py:75-144)."""
    keys = [k for k in responses if "response" in k and "_mask" not in k and "scores" not in k and "image_relations" not in k]
    Q = queries.size(1)
    ids, masks = [], []
    img_mask = None
    for k in keys:
        row = torch.cat([queries, responses[k]], dim=1)
        if queries.size(1) == queries_attn_masks.size(1):
            m = row.ne(cfg.pad_id)
            m[:, :Q] = queries_attn_masks.bool()
        else:  # CoPO 'attention': mask is P longer, image-token part first (rl_models.py:103-105)
            P = queries_attn_masks.size(1) - Q
            img_mask = queries_attn_masks[:, :P].bool()
            m = torch.cat([queries_attn_masks[:, P:].bool(), responses[k].ne(cfg.pad_id)], dim=1)
        ids.append(row)
        masks.append(m)
    K, b = len(keys), queries.size(0)
    input_ids = torch.cat(ids, 0)
    attn = torch.cat(masks, 0)
    R = responses[keys[0]].size(1)
    response_mask = ~input_ids[:, Q:].eq(cfg.pad_id)
    img_of_row = torch.arange(K * b) % b
    logits = llava_logits(cfg, w, lora, input_ids, attn, images,
                          None if img_mask is None else img_mask.repeat(K, 1), img_of_row=img_of_row,
                          checkpoint_layers=checkpoint_layers)
    logits = logits[:, -R - 1:-1] / temperature
    labels = input_ids[:, -R:]
    logprobs = compute_logprobs(logits, labels, cfg.pad_id) * response_mask
    entropies = -(logits.softmax(-1) * logits.log_softmax(-1)).sum(-1) * response_mask
    out = {}
    for i, k in enumerate(keys):
        out[k + "_logprobs"] = logprobs[i * b:(i + 1) * b]
        out[k + "_entropies"] = entropies[i * b:(i + 1) * b]
    return out


# ---------------------------------------------------------------------------------------------------- DPO loss
@dataclass
class LossConfig:
    beta: float = 0.1
    label_smoothing: float = 0.0
    reference_free: bool = False
    f_divergence_type: str = "reverse_kl"
    alpha_divergence_coef: float = 1.0
    standard_pair_coef: float = 1.0
    AI_pair_coef: float = 1.0
    CoPO: bool = False
    CoPO_coef: float = 0.2
    AncPO: bool = False
    mDPO_anchor: bool = True
    Anchor_value: float = 0.0
    Anchor_coef: float = 1.0
    response_score: bool = False
    response_image_relation: bool = False
    detailed_report: bool = False


def dpo_loss(lc: LossConfig, pc, pr, rc, rr, chosen_scores=None, rejected_scores=None):
    """DPOTrainer.dpo_loss (opadpo/dpo_models/dpo_trainer.py:429-473)."""
    if chosen_scores is None:
        chosen_scores = torch.ones_like(pc)
    if rejected_scores is None:
        rejected_scores = torch.ones_like(pr)
    cl = pc - (not lc.reference_free) * rc
    rl = pr - (not lc.reference_free) * rr
    if lc.f_divergence_type == "alpha_divergence":
        a = lc.alpha_divergence_coef
        cap = math.floor(math.log(torch.finfo(torch.float32).max) * 1e4) / 1e4   # dpo_trainer.py:74-81
        logits = (torch.exp(torch.clamp(rl * -a, max=cap)) - torch.exp(torch.clamp(cl * -a, max=cap))) / a
    else:
        logits = chosen_scores * cl - rejected_scores * rl
        if lc.f_divergence_type == "js_divergence":
            logits = logits - (F.softplus(cl) - F.softplus(rl))
    losses = -F.logsigmoid(lc.beta * logits) * (1 - lc.label_smoothing) - F.logsigmoid(-lc.beta * logits) * lc.label_smoothing
    return losses, lc.beta * cl, lc.beta * rl


def masked_mean(values, mask):
    """dpo_trainer.py:1040-1045 (axis=None)."""
    return (values * mask).sum() / mask.sum()


def compute_policy_loss(lc: LossConfig, pol: Dict[str, torch.Tensor], ref: Dict[str, torch.Tensor],
                        pol_mask: Optional[Dict[str, torch.Tensor]] = None, scores: Optional[Dict[str, torch.Tensor]] = None):
    """DPOTrainer.compute_policy_loss (dpo_trainer.py:475-802), given the policy outputs.
    pol: '<name>_logprobs' for standard_response / original_generate_response / AI_pseudo_response (and, with CoPO,
    pol_mask: 'mask_standard_response_logprobs', 'mask_AI_pseudo_response_logprobs');
    ref: 'ref_base_<name>_logprobs' (+ 'ref_mask_*_logprobs'); scores: '<name>_scores', '<name>_image_relations'."""
    p_std, p_org, p_ai = (pol[k + "_logprobs"] for k in ("standard_response", "original_generate_response", "AI_pseudo_response"))
    r_std, r_org, r_ai = (ref["ref_base_" + k + "_logprobs"] for k in
                          ("standard_response", "original_generate_response", "AI_pseudo_response"))
    if lc.detailed_report and (lc.response_score or lc.response_image_relation):
        sc_org, sc_ai = scores["original_generate_response_scores"], scores["AI_pseudo_response_scores"]
        ir_ai = scores["AI_pseudo_response_image_relations"]
    else:
        sc_org = (r_org != 0).to(r_org.dtype)
        sc_ai = (r_ai != 0).to(r_ai.dtype)
        ir_ai = sc_ai
    use_sc = lc.detailed_report and lc.response_score
    use_ir = lc.detailed_report and lc.response_image_relation
    l1, c1, j1 = dpo_loss(lc, p_std, p_org, r_std, r_org)
    l2, c2, j2 = dpo_loss(lc, p_ai, p_org, r_ai, r_org, sc_ai if use_sc else None, sc_org if use_sc else None)
    m_c1, m_r1, m_c2 = r_std != 0, r_org != 0, r_ai != 0
    loss = l1.mean() * lc.standard_pair_coef + l2.mean() * lc.AI_pair_coef
    if lc.CoPO:
        pm_std, pm_ai = pol_mask["mask_standard_response_logprobs"], pol_mask["mask_AI_pseudo_response_logprobs"]
        rm_std, rm_ai = ref["ref_mask_standard_response_logprobs"], ref["ref_mask_AI_pseudo_response_logprobs"]
        l3, c3, j3 = dpo_loss(lc, p_std, pm_std, r_std, rm_std)
        l4, c4, j4 = dpo_loss(lc, p_ai, pm_ai, r_ai, rm_ai, ir_ai if use_ir else None, ir_ai if use_ir else None)
        loss = loss + l3.mean() * lc.standard_pair_coef * lc.CoPO_coef + l4.mean() * lc.AI_pair_coef * lc.CoPO_coef
    else:
        l3 = c3 = j3 = l4 = c4 = j4 = torch.zeros_like(loss)
    if lc.AncPO:
        a = lc.Anchor_value
        if lc.mDPO_anchor:
            anc = -F.logsigmoid(c1 - a) - F.logsigmoid(-c2 + a) - F.logsigmoid(c3 - a) - F.logsigmoid(-c4 + a)
        else:
            anc = (c1 - a) ** 2 + (c2 - a) ** 2 + (c3 - a) ** 2 + (c4 - a) ** 2
        anc = anc.mean()
        loss = loss + anc * lc.Anchor_coef
    else:
        anc = torch.zeros_like(loss)
    stats = {
        "loss/stand_gen": l1.mean(), "loss/AI_gen": l2.mean(), "loss/stand_mask": l3.mean(), "loss/AI_mask": l4.mean(),
        "loss/AncPO": anc,
        "policy/stand_gen_chosen_mean": masked_mean(c1, m_c1), "policy/stand_gen_reject_mean": masked_mean(j1, m_r1),
        "policy/AI_gen_chosen_mean": masked_mean(c2, m_c2), "policy/AI_gen_reject_mean": masked_mean(j2, m_r1),
    }
    stats["policy/stand_gen_gap_mean"] = stats["policy/stand_gen_chosen_mean"] - stats["policy/stand_gen_reject_mean"]
    stats["policy/AI_gen_gap_mean"] = stats["policy/AI_gen_chosen_mean"] - stats["policy/AI_gen_reject_mean"]
    if lc.CoPO:
        stats["policy/stand_mask_chosen_mean"] = masked_mean(c3, m_c1)
        stats["policy/stand_mask_reject_mean"] = masked_mean(j3, m_c1)
        stats["policy/AI_mask_chosen_mean"] = masked_mean(c4, m_c2)
        stats["policy/AI_mask_reject_mean"] = masked_mean(j4, m_c2)
    return loss, {k: v.detach() for k, v in stats.items()}


def pair_loss(lc: LossConfig, p_c, p_r, r_c, r_r):
    """The north-star reduced form: one chosen/rejected pair, loss = mean(dpo_loss) (dpo_trainer.py:631 with one pair)."""
    l, c, j = dpo_loss(lc, p_c, p_r, r_c, r_r)
    return l.mean(), c, j


# ---------------------------------------------------------------------------------------------------- rollout
def hf_warp(logits: torch.Tensor, temperature: float = 1.0, top_k: int = 0, top_p: float = 1.0,
            suppress_tokens: Optional[List[int]] = None) -> torch.Tensor:
    """The logits processors / warpers HF generate() applies for the arguments online_generator.py:292-309 and
    rl_models.py:166-188 pass [UPSTREAM transformers/generation/logits_process.py, pinned against the installed classes in
    tests/test_rollout_host_cpu.py]: SuppressTokensLogitsProcessor (score := -inf), TemperatureLogitsWarper (score / T),
    TopKLogitsWarper (drop scores < the k-th largest; k = min(top_k, vocab); 0 = off), TopPLogitsWarper (ascending sort,
    softmax, cumulative mass <= 1 - top_p dropped, the largest kept). Returns the filtered scores (-inf = removed); the
    sampler then draws from softmax of them."""
    s = logits.clone().float()
    if suppress_tokens:
        s[..., list(suppress_tokens)] = float("-inf")
    if temperature != 1.0:
        s = s / temperature
    if top_k and top_k > 0:
        k = min(int(top_k), s.shape[-1])
        kth = torch.topk(s, k, dim=-1).values[..., -1, None]
        s = s.masked_fill(s < kth, float("-inf"))
    if top_p < 1.0:
        sorted_s, sorted_i = torch.sort(s, descending=False, dim=-1)
        cum = sorted_s.softmax(-1).cumsum(-1)
        remove = cum <= (1.0 - top_p)
        remove[..., -1:] = False
        s = s.masked_fill(remove.scatter(-1, sorted_i, remove), float("-inf"))
    return s


def greedy_generate(cfg: OracleConfig, w, lora, images, queries, queries_attn_masks, max_new_tokens: int,
                    suppress_tokens: Optional[List[int]] = None):
    """model.generate(do_sample=False) as driven by online_generator.py:292-309 / rl_models.py:166-188:
    HF GenerationMixin.greedy_search [UPSTREAM] over LlavaLlamaForCausalLM.forward with a KV cache; the mask is
    extended by the image-token count on cached steps (llava_modifications.patch:12161-12175); finished rows emit
    pad. Returns (ids [b, Q+max_new] padded to full width, min top-1 margin seen)."""
    b, Q = queries.shape
    with torch.no_grad():
        feats = projector(cfg, w, lora, clip_features(cfg, w, lora, images))
        x, mask = splice(cfg, w, queries, queries_attn_masks, feats)
        past = [None] * cfg.n_layers
        h, past = llama_hidden(cfg, w, lora, x, mask, past=past, pos0=0)
        out = [queries]
        unfinished = torch.ones(b, dtype=torch.long)
        margin = float("inf")
        logits = hf_warp(h[:, -1] @ w["lm_head.weight"].t(), suppress_tokens=suppress_tokens)
        for step in range(max_new_tokens):
            top2 = logits.topk(2, dim=-1).values
            live = unfinished.bool()
            if live.any():
                margin = min(margin, float((top2[live, 0] - top2[live, 1]).min()))
            nxt = logits.argmax(-1)
            nxt = nxt * unfinished + cfg.pad_id * (1 - unfinished)
            out.append(nxt[:, None])
            unfinished = unfinished * (nxt != cfg.eos_id).long()
            if step == max_new_tokens - 1:
                break
            mask = torch.cat([mask, torch.ones(b, 1, dtype=torch.bool)], dim=1)
            xe = w["model.embed_tokens.weight"][nxt][:, None]
            h, past = llama_hidden(cfg, w, lora, xe, mask, past=past, pos0=mask.shape[1] - 1)
            logits = hf_warp(h[:, -1] @ w["lm_head.weight"].t(), suppress_tokens=suppress_tokens)
    return torch.cat(out, dim=1), margin


def truncate_after_eos_with_padding(completions: torch.Tensor, eos_token_id: int, pad_token_id: int,
                                    additional_tokens: Optional[List[int]] = None) -> torch.Tensor:
    """opadpo/generator_models/generator.py:244-273: keep up to and including the first EOS; each extra stop id, in
    order, overrides the cut with ITS first occurrence (even if later than EOS); rest := pad."""
    out = completions.clone()
    for i in range(out.size(0)):
        row = out[i]
        end = None
        hits = (row == eos_token_id).nonzero()
        if hits.numel() > 0:
            end = int(hits[0])
        for tok in additional_tokens or []:
            hits = (row == tok).nonzero()
            if hits.numel() > 0:
                end = int(hits[0])
        if end is not None:
            out[i, end + 1:] = pad_token_id
    return out


# ---------------------------------------------------------------------------------------------------- optimiser
def adamw_reference(params: List[torch.Tensor], grads: List[torch.Tensor], lr: float, max_grad_norm: float = 1.0,
                    betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
    """clip_grad_norm_ (rl_trainer.py:165-168) + torch.optim.AdamW (utils/trainer_utils.py:9-38), first step."""
    ps = [p.clone().requires_grad_(True) for p in params]
    for p, g in zip(ps, grads):
        p.grad = g.clone()
    norm = torch.nn.utils.clip_grad_norm_(ps, max_grad_norm) if max_grad_norm > 0 else None
    opt = torch.optim.AdamW(ps, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
    opt.step()
    return [p.detach() for p in ps], norm


# ---------------------------------------------------------------------------------------------------- synthetic inputs
def synthetic_batch(cfg: OracleConfig, b: int, Q: int, R: int, names: List[str], seed: int = 7, left_pad: int = 0,
                    ragged: bool = True, image_pos: int = 5):
    """Seeded synthetic preference batch in the reference's input format (data_utils_dpo.py:101-285): left-padded
    queries with one image token, right-padded responses ending in EOS."""
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(b, 3, cfg.v_image, cfg.v_image, generator=g)
    queries = torch.randint(3, cfg.vocab, (b, Q), generator=g)
    qmask = torch.ones(b, Q, dtype=torch.bool)
    for i in range(b):
        lp = (left_pad + i) % max(1, Q // 2) if left_pad else 0
        queries[i, :lp] = cfg.pad_id
        qmask[i, :lp] = False
        queries[i, lp + image_pos] = cfg.image_token
    resp = {}
    for k, name in enumerate(names):
        t = torch.randint(3, cfg.vocab, (b, R), generator=g)
        for i in range(b):
            n = int(torch.randint(max(2, R // 4), R + 1, (1,), generator=g)) if ragged else R
            t[i, n - 1] = cfg.eos_id
            t[i, n:] = cfg.pad_id
        resp[name] = t
    return images, queries, qmask, resp
