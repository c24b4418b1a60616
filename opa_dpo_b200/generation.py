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


FIRST_STEP_IDX = 1                      # generator.py:41
IMAGE_TOKEN_INDEX = -200                # utils/constants.py:21


def find_max_step(output_dir: str) -> int:
    """Generator.find_max_step (generator.py:126-140): the highest `step<N>` among the file names under
    <output_dir>/rollouts/ (created when missing), 0 when there is none — where a restarted generation run resumes."""
    import re
    directory = output_dir.rstrip("/") + "/rollouts/"
    os.makedirs(directory, exist_ok=True)
    best = 0
    for filename in os.listdir(directory):
        m = re.search(r"step(\d+)", filename)
        if m:
            best = max(best, int(m.group(1)))
    return best


class B200OnlineGenerator:
    """Online_Generator / Generator (online_generator.py:235-396, generator.py:115-175) on the engine: the generation half
    of `rollout` runs on the device, the loop around it (`step`, `generate`, resume through `find_max_step`,
    `record_step_stats`) keeps the reference's step numbering and file layout. The GPT-4V revision between generation and
    the record (`process_responses_coco_gpt4v`, online_generator.py:143-232) is a network service: it is the caller's
    `revise` callable here, and the AI_* columns are None without one."""

    def __init__(self, args, policy_model, tokenizer, revise=None, truncate=None):
        """args: response_len, top_p, top_k, temperature, suppress_eos_at_generation, bf16 / fp16, and for the loop
        rollout_accumulation_steps, rollout_batch_size, total_epochs, sample_num, output_dir (online_generation.py flags);
        policy_model: B200LlavaModel (the `generate` surface); tokenizer: pad_token_id / eos_token_id (+ batch_decode for
        `rollout`); revise(responses=, images_url=, text_queries=, text_responses=, text_standard_responses=) ->
        {"Pseudo_response", "Generated_response", "report_json"}; truncate: the stop-token operation, by default — and in
        this package only — the CUDA kernel (the CPU tests pass the oracle to compare the host logic with the reference)."""
        self.args, self.model, self.tokenizer = args, policy_model, tokenizer
        self.revise = revise
        self._truncate = truncate if truncate is not None else truncate_after_eos_with_padding
        self.log_history: List = []

    @torch.no_grad()
    def generate_responses(self, images: Tensor, queries: Tensor, query_attn_masks: Tensor, do_sample: bool = True) -> Tensor:
        a, t = self.args, self.tokenizer
        out = self.model.generate(inputs=queries, images=images, attention_mask=query_attn_masks, do_sample=do_sample,
                                  max_new_tokens=a.response_len, pad_token_id=t.pad_token_id,
                                  suppress_tokens=[t.eos_token_id] if getattr(a, "suppress_eos_at_generation", False) else None,
                                  top_p=getattr(a, "top_p", 1.0), top_k=getattr(a, "top_k", 0),
                                  temperature=getattr(a, "temperature", 1.0), num_return_sequences=1, synced_gpus=False)
        responses = out[:, queries.size(1):]
        return self._truncate(responses, t.eos_token_id, t.pad_token_id, list(EXTRA_STOP_TOKENS))

    @torch.no_grad()
    def rollout(self, queries_data, step_idx: Optional[int] = None) -> Dict[str, List]:
        """Online_Generator.rollout (online_generator.py:235-378): per batch — images cast to the compute dtype, sampled
        responses truncated at EOS / the two extra stop ids, responses / standard responses / queries decoded to text (the
        image placeholder turned into id 1, the question cut out between 'USER:  \\n' and ' ASSISTANT:'), the revision —
        and one flat record list over all batches under the reference's column names."""
        a, t = self.args, self.tokenizer
        if hasattr(self.model, "eval"):
            self.model.eval()
        dev = getattr(getattr(self.model, "engine", None), "device", None)
        cols: Dict[str, List] = {k: [] for k in ("query", "image_id", "standard_response", "original_generate_response",
                                                 "AI_generate_response", "AI_pseudo_response", "AI_json_report", "image_bytes")}
        decode = lambda ids: t.batch_decode(ids, skip_special_tokens=True, clean_up_tokenization_spaces=True)
        for batch in queries_data:
            images, queries, qmask = batch["images"], batch["queries"], batch["query_attn_masks"]
            if dev is not None:
                images, queries, qmask = images.to(dev), queries.to(dev), qmask.to(dev)
            if getattr(a, "bf16", False):
                images = images.to(torch.bfloat16)
            elif getattr(a, "fp16", False):
                images = images.half()
            responses = self.generate_responses(images, queries, qmask, do_sample=True)
            text_responses = decode(responses)
            text_standard = decode(batch["standard_responses"])
            q_ids = queries.clone()
            q_ids[q_ids == IMAGE_TOKEN_INDEX] = 1
            text_queries = [q[q.find("USER:  \n") + 8:q.find(" ASSISTANT:")] for q in decode(q_ids)]
            n = len(text_responses)
            if self.revise is not None:
                res = self.revise(responses=responses, images_url=batch["images_url"], text_queries=text_queries,
                                  text_responses=text_responses, text_standard_responses=text_standard)
                pseudo, generated, reports = res["Pseudo_response"], res["Generated_response"], res["report_json"]
            else:
                pseudo, generated, reports = [None] * n, [None] * n, [None] * n
            cols["query"] += ["<image>\n" + q for q in text_queries]
            cols["image_id"] += list(batch["images_path"])
            cols["standard_response"] += text_standard
            cols["original_generate_response"] += text_responses
            cols["AI_generate_response"] += list(generated)
            cols["AI_pseudo_response"] += list(pseudo)
            cols["AI_json_report"] += list(reports)
            cols["image_bytes"] += list(batch["images_bytes"])
        return cols

    def step(self, train_dataloader, step_idx: int):
        """Generator.step (generator.py:115-124)."""
        batches = [next(train_dataloader) for _ in range(self.args.rollout_accumulation_steps)]
        response_dict = self.rollout(batches, step_idx)
        record_step_stats(getattr(self.args, "output_dir", None), step_idx, response_dict)
        return None              # what the reference's record_step_stats returns

    def generate(self, dataset_len: int, next_batch_source) -> List:
        """Generator.generate (generator.py:142-175): total steps from the dataset length, `max_step = sample_num //
        rollout_batch_size`, steps below the highest recorded one are skipped with their batches consumed (so a restarted
        run continues in the same data order; the highest recorded step itself is generated again), stop at max_step."""
        a = self.args
        total_steps = dataset_len * a.total_epochs // a.rollout_batch_size
        self.max_step = a.sample_num // a.rollout_batch_size
        skipping_steps = find_max_step(a.output_dir)
        loader = next_batch_source()
        for step_idx in range(FIRST_STEP_IDX, total_steps + FIRST_STEP_IDX):
            if step_idx < skipping_steps:
                for _ in range(a.rollout_accumulation_steps):
                    next(loader)
                continue
            if step_idx >= self.max_step:
                break
            self.log_history.append(self.step(loader, step_idx))
        return self.log_history


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
