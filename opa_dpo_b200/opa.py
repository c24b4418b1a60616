"""OPA stage (LoRA-SFT) loss on the same engine — row f-1 of the scope table.

The reference trains OPA with HF Trainer over LlavaLlamaForCausalLM's own `loss` (opadpo/opa_train.py:543-548,
opa_models/opa_trainer.py:58-125): mean cross-entropy over every labelled (non IGNORE_INDEX) position of the batch,
labels = the response tokens (prompt masked, utils/data_utils_sft.py:20-69). With the policy forward returning
per-token log-probs of the response tokens, that loss is  -(sum of log-probs over valid tokens) / (number of valid
tokens); its backward is the engine's LoRA backward with d(logprob) = -1/count at valid tokens.

The optional entropy regulariser (opa_trainer.py:63-87,113-116) runs a second forward on masked images (or with holes in
the image attention mask) and adds  coef * mean_rows( -sum((H_masked - H) * mask) / sum(mask) ); both forwards stay in
the autograd graph, so the gradient flows through the entropies of both — the engine's fused lm_head backward takes
d(loss)/d(entropy) next to d(loss)/d(logprob) (`opadpo_backward_ent`), and the two forwards occupy its two save slots.

Known deviation (pinned by tests/test_opa_host_cpu.py, which runs the reference's compute_loss verbatim): the reference
reads the entropy at the position OF each labelled token — `entropies[t]` from `logits[t]`, masked by `labels[t] != -100`
(opa_trainer.py:63,82-84) — i.e. the distribution that predicts token t+1, one position later than its cross-entropy. The
engine's entropies are the ones rl_models.py:128 defines for the DPO path: of the distributions that predict the
labelled tokens themselves. `entropy_loss` below is the reference's formula exactly (equal value and gradient when fed
entropies at the reference's positions). `OPAArgs.entropy_alignment = "reference"` reproduces the reference's positions
with the engine as it is (one extra response slot with a placeholder label, see `OPAStepper._forward`); the default stays
"label" until that mode has run on the GPU.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor


_PLACEHOLDER = 1     # any non-pad id: the slot only has to carry a label so that its entropy (and its gradient) is emitted


def sft_loss(outputs: Dict[str, Tensor], responses: Dict[str, Tensor], pad_token_id: int = 0) -> Tensor:
    """HF causal-LM loss over the response tokens of every `<name>` in `responses` (token-level mean, as
    CrossEntropyLoss(ignore_index) over the flattened batch does)."""
    total = None
    count = None
    for name, ids in responses.items():
        lp = outputs[name + "_logprobs"]
        valid = (ids.to(lp.device) != pad_token_id)
        s = -(lp * valid).sum()
        n = valid.sum()
        total = s if total is None else total + s
        count = n if count is None else count + n
    return total / count.clamp(min=1)


def sft_batch_to_policy_inputs(input_ids: Tensor, labels: Tensor, pad_token_id: int, query_len: int, response_len: int,
                               ignore_index: int = -100):
    """The OPA stage's collated batch (utils/data_utils_sft.py:20-69: rows `prompt | response`, right-padded, labels =
    IGNORE_INDEX on the prompt and the padding) -> the policy boundary's inputs: queries [b, query_len] left-padded with
    their mask, responses [b, response_len] right-padded. Causal attention with RoPE depends on relative positions only, so
    moving the padding from the right of the row to the left of the prompt leaves every labelled token's distribution
    unchanged. Single-turn rows only (one labelled span); a row that does not fit raises instead of being truncated."""
    b, L = input_ids.shape
    labelled = labels != ignore_index
    n_resp = labelled.sum(1)
    n_tok = (input_ids != pad_token_id).sum(1)
    first = torch.where(n_resp > 0, labelled.int().argmax(1), n_tok)          # prompt length of every row
    pos = torch.arange(L, device=input_ids.device)[None, :]
    span = (pos >= first[:, None]) & (pos < (first + n_resp)[:, None])
    if not torch.equal(span, labelled):
        raise ValueError("sft_batch_to_policy_inputs: a row has more than one labelled span (multi-turn rows are not a "
                         "query + one response)")
    if int(first.max()) > query_len or int(n_resp.max()) > response_len:
        raise ValueError(f"sft_batch_to_policy_inputs: prompt {int(first.max())} / response {int(n_resp.max())} tokens do not "
                         f"fit query_len {query_len} / response_len {response_len}")
    q_pos = torch.arange(query_len, device=input_ids.device)[None, :] - (query_len - first)[:, None]   # source column or < 0
    q_valid = q_pos >= 0
    queries = torch.where(q_valid, input_ids.gather(1, q_pos.clamp(min=0)), torch.full_like(q_pos, pad_token_id))
    r_pos = torch.arange(response_len, device=input_ids.device)[None, :]
    r_valid = r_pos < n_resp[:, None]
    src = (first[:, None] + r_pos).clamp(max=L - 1)
    responses = torch.where(r_valid, input_ids.gather(1, src), torch.full_like(src, pad_token_id))
    return queries, q_valid, responses


def entropy_loss(outputs: Dict[str, Tensor], masked_outputs: Dict[str, Tensor], responses: Dict[str, Tensor],
                 pad_token_id: int = 0) -> Tensor:
    """opa_trainer.py:82-86: per row -sum((H_masked - H) * mask) / sum(mask) with mask = labelled positions, then the
    mean over rows (a row without labelled tokens gives nan there too). One response name per row set; with several
    names their rows are concatenated, which is what a batch of that many samples is to the reference."""
    rows = []
    for name, ids in responses.items():
        h, hm = outputs[name + "_entropies"], masked_outputs[name + "_entropies"]
        mask = (ids.to(h.device) != pad_token_id).to(h.dtype)
        rows.append(-((hm - h) * mask).sum(-1) / mask.sum(-1))
    return torch.cat(rows).mean()


@dataclass
class OPAArgs:
    """Loss-relevant TrainingArguments of the OPA stage (opa_train.py:144-163; run/train_opa.sh)."""
    entropy_loss: bool = False
    entropy_loss_coef: float = 0.1
    entropy_decay_coef: float = 1.0
    entropy_mask_method: str = "random"   # random | blockwise (masked images) | attention (holes in the image mask)
    entropy_mask_ratio: float = 0.3
    temperature: float = 1.0
    # "reference" (default): the reference's positions (opa_trainer.py:82-85), one later than the labels (see the module
    # docstring) — the forward runs over responses extended by one slot, with a placeholder token behind each row's last
    # token so that slot's entropy is emitted.
    # "label": entropies of the distributions that predict the labelled tokens (the engine's native output; one slot
    # cheaper, not what the reference computes).
    entropy_alignment: str = "reference"


class OPAStepper:
    """One LoRA-SFT optimisation step: policy forward over the given responses -> token-mean NLL [+ entropy regulariser
    over a second, image-masked forward] -> LoRA backward -> [all-reduce] -> clip + AdamW (what
    LLaVATrainer.compute_loss / training_step + optimizer.step do for the OPA stage)."""

    def __init__(self, policy, optimizer, temperature: float = 1.0, pad_token_id: int = 0, args: Optional[OPAArgs] = None):
        self.policy, self.opt, self.temperature, self.pad = policy, optimizer, temperature, pad_token_id
        self.args = args or OPAArgs(temperature=temperature)

    def _forward(self, images, queries, queries_attn_masks, responses: Dict[str, Tensor]) -> Dict[str, Tensor]:
        """Policy outputs [b, R] for `responses`. With entropy_alignment == "reference" the entropies are the ones the
        reference's regulariser reads: entropies[t] = H(distribution at the position of token t). The policy computes
        H(distribution predicting slot j) for every slot j whose label is not pad, so the forward runs on responses
        extended to R + 1 slots with a placeholder token right behind each row's last token (causal attention: nothing
        before it changes); slot t + 1 of that output is the wanted entropy, and the placeholder's own log-prob is cut off."""
        a = self.args
        if not (a.entropy_loss and a.entropy_alignment == "reference"):
            return self.policy(images, queries, queries_attn_masks, temperature=self.temperature, **responses)
        ext = {}
        for name, ids in responses.items():
            n = (ids != self.pad).sum(1, keepdim=True)                        # valid length of every row (right-padded)
            e = torch.cat([ids, torch.full_like(ids[:, :1], self.pad)], 1)
            ext[name] = e.scatter(1, n, torch.full_like(n, _PLACEHOLDER))
        pargs = self.policy.args
        R = pargs.response_len
        pargs.response_len = R + 1            # AutoregressivePolicy.forward insists on responses of args.response_len
        try:
            out = self.policy(images, queries, queries_attn_masks, temperature=self.temperature, **ext)
        finally:
            pargs.response_len = R
        res = {}
        for name, ids in responses.items():
            ent = out[name + "_entropies"]
            valid = (ids != self.pad).to(device=ent.device, dtype=ent.dtype)
            res[name + "_logprobs"] = out[name + "_logprobs"][:, :-1] * valid
            res[name + "_entropies"] = out[name + "_entropies"][:, 1:] * valid
        return res

    def compute_loss(self, images: Tensor, queries: Tensor, queries_attn_masks: Tensor, responses: Dict[str, Tensor],
                     masked_images: Optional[Tensor] = None, masked_attn_masks: Optional[Tensor] = None
                     ) -> Tuple[Tensor, Dict[str, float]]:
        """opa_trainer.py:58-125. `masked_images` (random / blockwise) or `masked_attn_masks` (the [b, P + Q] mask with the
        image part prepended, 'attention') is the second forward's input when args.entropy_loss is on."""
        a = self.args
        out = self._forward(images, queries, queries_attn_masks, responses)
        loss = sft_loss(out, responses, self.pad)
        logs = {"base_sft_loss": loss.detach(), "mask_sft_loss": 0.0, "entropy_loss": 0.0}
        if a.entropy_loss:
            if a.entropy_mask_method == "attention":
                if masked_attn_masks is None:
                    raise ValueError("entropy_mask_method='attention' needs masked_attn_masks")
                out_m = self._forward(images, queries, masked_attn_masks, responses)
            else:
                if masked_images is None:
                    raise ValueError("entropy_loss needs masked_images")
                out_m = self._forward(masked_images, queries, queries_attn_masks, responses)
            ent = entropy_loss(out, out_m, responses, self.pad)
            logs["mask_sft_loss"] = sft_loss(out_m, responses, self.pad).detach()
            logs["entropy_loss"] = ent.detach()
            a.entropy_loss_coef *= a.entropy_decay_coef           # decays before use (opa_trainer.py:114)
            loss = loss + ent * a.entropy_loss_coef
        return loss, logs

    def step(self, images: Tensor, queries: Tensor, queries_attn_masks: Tensor, responses: Dict[str, Tensor],
             grad_scale: float = 1.0, masked_images: Optional[Tensor] = None, masked_attn_masks: Optional[Tensor] = None):
        self.policy.train()
        loss, _ = self.compute_loss(images, queries, queries_attn_masks, responses, masked_images, masked_attn_masks)
        loss.backward()
        norm = self.opt.step(grad_scale)
        self.opt.zero_grad()
        return loss.detach(), norm


def save_final_checkpoint(model, output_dir: str, adapter_name: str = "lora_policy", is_main_process: bool = True) -> Optional[str]:
    """The OPA stage's final save (opa_train.py:553-561): `<output_dir>/checkpoint-final/{adapter_model.bin,
    adapter_config.json, non_lora_trainables.bin, config.json}` — PEFT's flat single-adapter layout, which is what the DPO
    stage's `policy_model_name_or_path` points at (qlora_model.py:318-328; `policy.make_models` reads it back).
    `non_lora_trainables.bin` holds the trainable tensors that are not LoRA (get_peft_state_non_lora_maybe_zero_3): none here,
    everything but the adapters is frozen, so it is the empty dict the reference writes in the same situation. `config.json`
    is the base model's, copied from `model.base_model_name_or_path` when that directory has one."""
    import json
    import os
    import shutil
    if not is_main_process:
        return None
    final = os.path.join(output_dir, "checkpoint-final")
    os.makedirs(final, exist_ok=True)
    torch.save(model.adapter_state_dict(adapter_name), os.path.join(final, "adapter_model.bin"))
    cfg = dict(model.peft_config.get(adapter_name) or model.default_peft_config())
    cfg["inference_mode"] = True
    with open(os.path.join(final, "adapter_config.json"), "w") as f:
        json.dump(cfg, f, indent=2, sort_keys=True)
    torch.save({}, os.path.join(final, "non_lora_trainables.bin"))
    src = os.path.join(getattr(model, "base_model_name_or_path", "") or "", "config.json")
    if os.path.isfile(src):
        shutil.copyfile(src, os.path.join(final, "config.json"))
    return final
