"""Host-side mirror of the on-policy generation stage around the rollout kernels.

* `truncate_after_eos_with_padding` <- opadpo/generator_models/generator.py:244-273, same signature, on the device
  (the reference walks Python lists after a `.tolist()` round trip).
* `B200OnlineGenerator.generate_responses` <- the model half of `Online_Generator.rollout`
  (opadpo/generator_models/online_generator.py:283-323): `policy.generate(do_sample, top_k, top_p, temperature,
  max_new_tokens=response_len)`, slice off the query, truncate at EOS / the two extra stop ids. The GPT-4V revision that
  follows (:325-378) is a network service and stays with the caller.
* `record_step_stats` <- online_generator.py:380-396: `rollouts/step{i}_rank{r}.json`, one record per sample,
  `image_bytes` base64-encoded, `json.dump(..., indent=4)`.
"""
from __future__ import annotations

import base64
import ctypes
import json
import os
from typing import Dict, List, Optional, Sequence

import torch
from torch import Tensor

from . import _lib
from ._lib import OpadpoError

EXTRA_STOP_TOKENS = (1577, 29973)     # tokenizer.encode("?")[0], tokenizer.encode("\n?")[-1] (online_generator.py:311-315)


def truncate_after_eos_with_padding(completions: Tensor, eos_token_id: int, pad_token_id: int,
                                    additional_tokens: Optional[Sequence[int]] = None) -> Tensor:
    """generator.py:244-273 on a CUDA LongTensor [b, R]; returns a new tensor on the same device."""
    if completions.device.type != "cuda":
        raise OpadpoError("truncate_after_eos_with_padding: the rollout lives on the CUDA device (no CPU implementation)")
    lib = _lib.load()
    x = completions.to(torch.int64).contiguous()
    out = torch.empty_like(x)
    extra = list(additional_tokens or [])
    arr = (ctypes.c_int64 * max(1, len(extra)))(*extra)
    with torch.cuda.device(x.device):
        _lib.check(lib.opadpo_op_truncate_after_eos(x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1], int(eos_token_id),
                                                    int(pad_token_id), arr, len(extra),
                                                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
                   "opadpo_op_truncate_after_eos")
    return out


class B200OnlineGenerator:
    """The generation half of Online_Generator (online_generator.py:283-323) on the engine."""

    def __init__(self, args, policy_model, tokenizer):
        """args: response_len, top_p, top_k, temperature, suppress_eos_at_generation (online_generation.py flags);
        policy_model: B200LlavaModel (the `generate` surface); tokenizer: pad_token_id / eos_token_id."""
        self.args, self.model, self.tokenizer = args, policy_model, tokenizer

    @torch.no_grad()
    def generate_responses(self, images: Tensor, queries: Tensor, query_attn_masks: Tensor, do_sample: bool = True) -> Tensor:
        a, t = self.args, self.tokenizer
        out = self.model.generate(inputs=queries, images=images, attention_mask=query_attn_masks, do_sample=do_sample,
                                  max_new_tokens=a.response_len, pad_token_id=t.pad_token_id,
                                  suppress_tokens=[t.eos_token_id] if getattr(a, "suppress_eos_at_generation", False) else None,
                                  top_p=getattr(a, "top_p", 1.0), top_k=getattr(a, "top_k", 0),
                                  temperature=getattr(a, "temperature", 1.0), num_return_sequences=1, synced_gpus=False)
        responses = out[:, queries.size(1):]
        return truncate_after_eos_with_padding(responses, t.eos_token_id, t.pad_token_id, list(EXTRA_STOP_TOKENS))


def record_step_stats(output_dir: Optional[str], step_idx: int, response_dict: Dict[str, List]) -> Optional[str]:
    """online_generator.py:380-396. Returns the path written (None when output_dir is None)."""
    if output_dir is None:
        return None
    keys = list(response_dict.keys())
    n = len(response_dict[keys[0]]) if keys else 0
    for k in keys:
        if len(response_dict[k]) != n:
            raise ValueError("All arrays must be of the same length")     # what pandas raises for ragged columns
    records = [{k: response_dict[k][i] for k in keys} for i in range(n)]
    if "image_bytes" in response_dict:
        for item in records:
            item["image_bytes"] = base64.b64encode(item["image_bytes"]).decode("utf-8")
    log_dir = os.path.join(output_dir, "rollouts")
    os.makedirs(log_dir, exist_ok=True)
    path = os.path.join(log_dir, f"step{step_idx}_rank{int(os.environ.get('RANK', 0))}.json")
    with open(path, "w") as f:
        json.dump(records, f, indent=4)
    return path
