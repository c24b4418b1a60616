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


class OverlappedGradAllReduce:
    """All-reduce of the LoRA-grad bucket overlapped with the backward pass.

    The bucket is ordered layer by layer, and opadpo_backward reports (host callback) when the kernels that finish a
    layer's slice have been enqueued. Each slice is then all-reduced with async_op=True: ProcessGroupNCCL makes its own
    stream wait for everything enqueued so far on the compute stream, so the collective of layer l runs over NVLink while
    layers l-1..0 are still being differentiated. `finish()` makes the compute stream wait for all of them and returns
    the 1/world scale for the optimizer. Arm it only for the LAST micro-batch of an optimizer step."""

    def __init__(self, engine, flat_grad: torch.Tensor, group: Optional[dist.ProcessGroup] = None):
        self.engine, self.grad, self.group = engine, flat_grad, group
        self.works = []
        self.ranges = {}
        self.armed = False
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        if self.world > 1:
            engine.set_backward_callback(self._on_layer)

    def _on_layer(self, layer: int):
        if not self.armed:
            return
        if layer not in self.ranges:
            self.ranges[layer] = self.engine.layer_grad_range(layer)
        off, cnt = self.ranges[layer]
        if cnt > 0:
            self.works.append(dist.all_reduce(self.grad[off:off + cnt], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def arm(self):
        self.armed = self.world > 1

    def finish(self) -> float:
        if self.world == 1:
            return 1.0
        if not self.works:      # nothing was overlapped (not armed): one bucket-wide all-reduce
            return allreduce_grad_bucket(self.grad, self.group)
        for w in self.works:
            w.wait()            # stream-level wait: the compute stream waits for the NCCL stream
        self.works = []
        self.armed = False
        return 1.0 / self.world
