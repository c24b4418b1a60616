"""Data-parallel plumbing of the OPA-DPO step: preference pairs shard across ranks (one process per GPU, torchrun),
every rank holds a full replica, and the ONLY data-path collective is one all-reduce(sum) of the flat fp32 LoRA-grad
bucket per optimizer step; the 1/world factor is folded into the fused clip+AdamW kernel (grad_scale).

Deviation from the reference, on purpose: its DDP reducer is never armed (every backward runs under no_sync,
rl_trainer.py:155-162), so LoRA gradients are never synchronised there; this implements the intended DDP-mean
semantics. World size 1 is identical to the reference.
"""
from __future__ import annotations

import ctypes
import os
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
    """All-reduce of the LoRA-grad bucket overlapped with the backward pass (NCCL; the fallback of `make_grad_allreduce`).

    The bucket is ordered layer by layer, and opadpo_backward reports (host callback) when the kernels that finish a
    layer's slice have been enqueued. Each slice is then all-reduced with async_op=True: ProcessGroupNCCL makes its own
    stream wait for everything enqueued so far on the compute stream, so the collective of layer l runs over NVLink while
    layers l-1..0 are still being differentiated. `finish()` makes the compute stream wait for all of them and returns
    the 1/world scale for the optimizer. Arm it only for the LAST micro-batch of an optimizer step."""

    name = "NCCL all-reduce(sum), fp32, one slice per decoder layer"
    transport = "NCCL"

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


class NvlGradAllReduce:
    """The same overlap, with the collective done by this library's own kernel over NVLink / NVSwitch peer memory
    (csrc/collective.cu, `opadpo_op_allreduce_bf16wire`) instead of NCCL. The kernel takes no shared memory, so its few CTAs
    run beside the backward's persistent GEMM CTAs instead of displacing them. Two forms:

      wire="fp32" (default): the engine accumulates its gradients directly in a symmetric-memory bucket (torch maps it on
        every rank; `opadpo_engine_adopt_grad_buffer`), and each rank sums its 1/world share of a slice IN PLACE with
        multimem.ld_reduce (the addition happens in the NVSwitch) and multimem.st's it to all ranks: exact fp32 sums, no
        staging, the bucket is read once and written once per step.
      wire="bf16": the bucket stays where it is; the slice is packed to bf16 into a symmetric staging bucket, reduced with
        fp32 accumulation, broadcast and unpacked by the same launch (half the NVLink bytes, twice the HBM traffic).

    Without a multicast mapping both forms fall back to peer loads / stores inside the same kernel. Layers are grouped
    (`group_layers` decoder layers per launch; the last layers and the projector go on their own with a wide grid, since
    nothing is left to hide them behind). The reduced gradient is bit-identical on every rank. Construct through
    `make_grad_allreduce`, which falls back to NCCL on every rank if any rank cannot map the buffers."""

    def __init__(self, engine, flat_grad: torch.Tensor, group=None, group_layers: Optional[int] = None,
                 blocks: Optional[int] = None, tail_blocks: Optional[int] = None, wire: Optional[str] = None,
                 overlap: Optional[bool] = None):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self._lib = _lib
        self.lib = _lib.load()
        self.engine, self.group = engine, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.wire = wire or os.environ.get("OPADPO_DP_WIRE", "fp32")
        if self.wire not in ("fp32", "bf16"):
            raise ValueError(f"NvlGradAllReduce: wire must be 'fp32' or 'bf16', got {self.wire!r}")
        if self.world > _lib.MAX_RANKS:
            raise RuntimeError(f"NvlGradAllReduce: {self.world} ranks > {_lib.MAX_RANKS} (one NVLink domain)")
        n = flat_grad.numel()
        if n % 8 != 0 or flat_grad.data_ptr() % 32 != 0:
            raise RuntimeError("NvlGradAllReduce: the gradient bucket must be 32 B-aligned with a multiple of 8 elements")
        dev = flat_grad.device
        pg = group if group is not None else dist.group.WORLD
        # the symmetric buffer: the gradient bucket itself (fp32 in place) or its bf16 staging copy
        self.sym = symm_mem.empty(n, dtype=torch.float32 if self.wire == "fp32" else torch.bfloat16, device=dev)
        self.h_sym = symm_mem.rendezvous(self.sym, pg)
        self.signal = symm_mem.empty(_lib.AR_MAX_BLOCKS * _lib.MAX_RANKS, dtype=torch.int32, device=dev)
        self.signal.zero_()
        self.h_signal = symm_mem.rendezvous(self.signal, pg)
        torch.cuda.synchronize(dev)
        self.h_signal.barrier()          # every rank's signal words are zero before anybody signals
        torch.cuda.synchronize(dev)
        self.sym_ptrs = [int(p) for p in self.h_sym.buffer_ptrs]
        self.signal_ptrs = [int(p) for p in self.h_signal.buffer_ptrs]
        use_mc = os.environ.get("OPADPO_DP_MULTICAST", "1") != "0"
        self.mc_ptr = int(self.h_sym.multicast_ptr or 0) if use_mc else 0
        if self.wire == "fp32":
            self.sym.copy_(flat_grad)
            if engine is not None:       # from now on the wgrad epilogues accumulate straight into the symmetric bucket
                engine.adopt_grad_buffer(self.sym)
            self.grad = self.sym
        else:
            self.grad = flat_grad
        self.err = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.comm = torch.cuda.Stream(device=dev, priority=-1)
        self.epoch = 1
        env = os.environ.get
        self.group_layers = group_layers if group_layers is not None else int(env("OPADPO_DP_GROUP_LAYERS", "4"))
        self.blocks = blocks if blocks is not None else int(env("OPADPO_DP_BLOCKS", "16"))
        self.tail_blocks = tail_blocks if tail_blocks is not None else int(env("OPADPO_DP_TAIL_BLOCKS", "128"))
        # overlap=False: nothing runs under the backward; ONE wide launch over the whole bucket when the step's gradients are final
        self.overlap = overlap if overlap is not None else env("OPADPO_DP_OVERLAP", "1") != "0"
        self.n_layers = None
        self.ranges = {}
        self.pending = None      # (offset, count) of finished layers not yet sent
        self.launched = False
        self.armed = False
        self.done = torch.cuda.Event()
        if engine is not None:           # engine=None: a bare all-reduce of `flat_grad` (tests drive _launch / finish directly)
            engine.set_backward_callback(self._on_layer)

    @property
    def name(self) -> str:
        how = ("fp32 in place on a symmetric gradient bucket" if self.wire == "fp32"
               else "bf16 wire, fp32 accumulation, fused pack/reduce/broadcast/unpack")
        when = (f"{self.blocks} CTAs per launch under the backward, {self.group_layers} layers per launch" if self.overlap
                else f"one {self.tail_blocks}-CTA launch over the whole bucket after the backward")
        return f"own NVLink kernel (no shared memory; {when}): {how}"

    @property
    def transport(self) -> str:
        return "multimem (NVLS, sum taken in the NVSwitch)" if self.mc_ptr else "peer loads/stores"

    def _launch(self, off: int, cnt: int, blocks: int):
        if cnt <= 0:
            return
        if off % 8 != 0 or cnt % 8 != 0:
            raise RuntimeError(f"NvlGradAllReduce: slice [{off}, +{cnt}) is not a multiple of 8 elements")
        L = self._lib
        esz = 4 if self.wire == "fp32" else 2
        d = L.AllReduceDesc()
        d.rank, d.world = self.rank, self.world
        d.grad = self.grad.data_ptr() + 4 * off
        d.stage_local = self.sym_ptrs[self.rank] + esz * off
        d.stage_mc = (self.mc_ptr + esz * off) if self.mc_ptr else 0
        for r in range(self.world):
            d.stage_peer[r] = self.sym_ptrs[r] + esz * off
            d.signal_peer[r] = self.signal_ptrs[r]
        d.count, d.epoch, d.blocks = cnt, self.epoch & 0xFFFFFFFF, blocks
        d.err_flag = self.err.data_ptr()
        d.timeout_ms = int(os.environ.get("OPADPO_DP_TIMEOUT_MS", "60000"))
        d.fp32_in_place = 1 if self.wire == "fp32" else 0
        cur = torch.cuda.current_stream(self.grad.device)
        self.comm.wait_stream(cur)       # the slice is final in the compute stream's order
        L.check(self.lib.opadpo_op_allreduce_bf16wire(ctypes.byref(d), ctypes.c_void_p(self.comm.cuda_stream)),
                "allreduce_bf16wire")
        self.epoch += 2
        self.launched = True

    def _range(self, layer: int):
        if layer not in self.ranges:
            self.ranges[layer] = self.engine.layer_grad_range(layer)
        return self.ranges[layer]

    def _on_layer(self, layer: int):
        if not self.armed or not self.overlap:
            return
        if self.n_layers is None:
            self.n_layers = int(self.engine.cfg.n_layers)
        off, cnt = self._range(layer)
        if layer == self.n_layers:                       # projector: last in time, small
            self._launch(off, cnt, self.tail_blocks)
            return
        # decoder layers arrive top-down and are adjacent in the bucket: grow the pending run downwards
        if self.pending is None:
            self.pending = (off, cnt)
        else:
            p_off, p_cnt = self.pending
            if off + cnt != p_off:
                self._launch(p_off, p_cnt, self.blocks)
                self.pending = (off, cnt)
            else:
                self.pending = (off, p_cnt + cnt)
        tail = layer <= 1
        if tail or layer % self.group_layers == 0:
            self._launch(self.pending[0], self.pending[1], self.tail_blocks if layer == 0 else self.blocks)
            self.pending = None

    def arm(self):
        self.armed = True

    def check(self):
        """Raises if a cross-rank wait timed out in any launch so far (the kernel sets a host-visible flag instead of hanging)."""
        code = int(self.err[0])
        if code != 0:
            raise RuntimeError(f"NvlGradAllReduce: a peer did not arrive within the timeout (flag {code & 0xFFFFFFFF:#x}); "
                               "the reduced gradients of that step are invalid")

    def finish(self) -> float:
        self.check()
        if self.pending is not None:
            self._launch(self.pending[0], self.pending[1], self.tail_blocks)
            self.pending = None
        if not self.launched:            # not armed for this step: the whole bucket in one launch
            self._launch(0, self.grad.numel(), self.tail_blocks)
        self.done.record(self.comm)
        torch.cuda.current_stream(self.grad.device).wait_event(self.done)
        self.launched = False
        self.armed = False
        return 1.0 / self.world


def make_grad_allreduce(engine, flat_grad: torch.Tensor, group=None):
    """The overlapped gradient all-reduce of a data-parallel step: this library's NVLink kernel when every rank can map the
    symmetric buffers (CUDA bucket, one NVLink domain, <= 8 ranks), else NCCL slices. The choice is collective: one rank
    failing moves all of them to NCCL. OPADPO_DP_COLLECTIVE=nccl forces NCCL."""
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    if world == 1 or not flat_grad.is_cuda or os.environ.get("OPADPO_DP_COLLECTIVE", "nvl") == "nccl":
        return OverlappedGradAllReduce(engine, flat_grad, group)
    impl, why = None, ""
    try:
        impl = NvlGradAllReduce(engine, flat_grad, group)
    except Exception as e:   # noqa: BLE001 — any mapping failure means "use NCCL", decided together below
        why = f"{type(e).__name__}: {e}"
    ok = torch.tensor([1 if impl is not None else 0], device=flat_grad.device, dtype=torch.int32)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if int(ok.item()) == 1:
        return impl
    if impl is not None:             # this rank mapped its buffers but another could not: keep the (possibly adopted) bucket, use NCCL on it
        engine.set_backward_callback(None)
        flat_grad = impl.grad
    fb = OverlappedGradAllReduce(engine, flat_grad, group)
    fb._keep_alive = impl
    fb.fallback_reason = why or "another rank could not map the symmetric buffers"
    return fb


# ------------------------------------------------------------------------------------------------ sharded input pipeline
class InfiniteLoader:
    """utils/common_utils.py:174-187: `next()` forever, restarting the wrapped loader (a new shuffled epoch) when it ends."""

    def __init__(self, loader):
        self.loader = loader
        self.iterator = iter(loader)

    def __iter__(self):
        return self

    def __next__(self):
        try:
            return next(self.iterator)
        except StopIteration:
            self.iterator = iter(self.loader)
            return next(self.iterator)


class ShardedBatchSampler:
    """What `accelerator.prepare(DataLoader(shuffle=True, drop_last=True))` does to the batch sampler (rl_trainer.py:314-321,
    opadpo_train.py:598-603 `even_batches=True, split_batches=False`) [UPSTREAM accelerate 0.21 BatchSamplerShard]: all
    ranks walk the SAME shuffled sequence of batches, rank r keeps batch i when i % world == r, and a batch is only yielded
    once its whole round of `world` batches exists — so every rank sees the same number of equally sized batches and the
    pairs of one step are disjoint across ranks. The ranks agree on the permutation without communicating because each
    builds the sampler's generator from the same seed (accelerate broadcasts rank 0's generator state instead)."""

    def __init__(self, batch_sampler, rank: int, world: int):
        if not 0 <= rank < world:
            raise ValueError(f"rank {rank} outside world {world}")
        self.batch_sampler, self.rank, self.world = batch_sampler, rank, world

    def __len__(self):
        return len(self.batch_sampler) // self.world

    def __iter__(self):
        mine = None
        for i, batch in enumerate(self.batch_sampler):
            if i % self.world == self.rank:
                mine = batch
            if i % self.world == self.world - 1:
                yield mine
                mine = None


def get_train_dataloader(train_dataset, data_collator, per_device_batch_size: int, rank: int = 0, world: int = 1, seed: int = 0,
                         num_workers: int = 0) -> InfiniteLoader:
    """RLTrainer.get_train_dataloader (rl_trainer.py:309-323): shuffled, drop_last, collated batches of
    `rollout_per_device_batch_size` samples, sharded over the ranks as accelerate would, wrapped in an InfiniteLoader."""
    from torch.utils.data import BatchSampler, DataLoader, RandomSampler
    gen = torch.Generator().manual_seed(int(seed))
    sampler = BatchSampler(RandomSampler(train_dataset, generator=gen), batch_size=per_device_batch_size, drop_last=True)
    loader = DataLoader(train_dataset, batch_sampler=ShardedBatchSampler(sampler, rank, world), collate_fn=data_collator,
                        num_workers=num_workers)
    return InfiniteLoader(loader)
