"""OPA stage (LoRA-SFT) loss on the same engine — row f-1 of the scope table.

The reference trains OPA with HF Trainer over LlavaLlamaForCausalLM's own `loss` (opadpo/opa_train.py:543-548,
opa_models/opa_trainer.py:58-125 with entropy_loss off): mean cross-entropy over every labelled (non IGNORE_INDEX)
position of the batch, labels = the response tokens (prompt masked, utils/data_utils_sft.py:20-69). With the policy
forward returning per-token log-probs of the response tokens, that loss is  -(sum of log-probs over valid tokens) /
(number of valid tokens); its backward is the engine's LoRA backward with d(logprob) = -1/count at valid tokens.
(The optional entropy regulariser needs d(entropy)/d(theta), which the fused lm_head epilogue does not produce: next.)
"""
from __future__ import annotations

from typing import Dict, Sequence

import torch
from torch import Tensor


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


class OPAStepper:
    """One LoRA-SFT optimisation step: policy forward over the given responses -> token-mean NLL -> LoRA backward ->
    [all-reduce] -> clip + AdamW (what LLaVATrainer.training_step + optimizer.step do for the OPA stage)."""

    def __init__(self, policy, optimizer, temperature: float = 1.0, pad_token_id: int = 0):
        self.policy, self.opt, self.temperature, self.pad = policy, optimizer, temperature, pad_token_id

    def step(self, images: Tensor, queries: Tensor, queries_attn_masks: Tensor, responses: Dict[str, Tensor],
             grad_scale: float = 1.0):
        self.policy.train()
        out = self.policy(images, queries, queries_attn_masks, temperature=self.temperature, **responses)
        loss = sft_loss(out, responses, self.pad)
        loss.backward()
        norm = self.opt.step(grad_scale)
        self.opt.zero_grad()
        return loss.detach(), norm
