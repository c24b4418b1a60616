"""Data-parallel plumbing of the OPA-DPO step: preference pairs shard across ranks (one process per GPU, torchrun),
every rank holds a full replica, and the ONLY data-path collective is one all-reduce(sum) of the flat fp32 LoRA-grad
bucket per optimizer step; the 1/world factor is folded into the fused clip+AdamW kernel (grad_scale).

Deviation from the reference, on purpose: its DDP reducer is never armed (every backward runs under no_sync,
rl_trainer.py:155-162), so LoRA gradients are never synchronised there; this implements the intended DDP-mean
semantics. World size 1 is identical to the reference.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(global_pairs: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, equal shards (accelerate even_batches / drop_last semantics: opadpo_train.py:601, rl_trainer.py:314-320)."""
    if global_pairs % world != 0:
        raise ValueError(f"global batch {global_pairs} must be divisible by world size {world}")
    per = global_pairs // world
    return rank * per, (rank + 1) * per


def allreduce_grad_bucket(flat_grad: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> float:
    """Sum the flat LoRA-grad bucket over ranks in place; returns the scale (1/world) the optimizer must apply so the
    update uses the mean of per-rank gradients (= gradient of the global-mean loss for equal shards)."""
    if not (dist.is_available() and dist.is_initialized()):
        return 1.0
    world = dist.get_world_size(group)
    if world == 1:
        return 1.0
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world
