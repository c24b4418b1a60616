"""world_size-2 gloo test of the data-parallel semantics (host logic in opa_dpo_b200/dist.py): per-rank LoRA gradients of
the local-mean loss, all-reduced (sum) and scaled by 1/world, equal the single-process gradient on the concatenated
batch. The per-rank compute stand-in is the CPU oracle (tests may use it); the product's GPU path is covered by -m gpu."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ["chosen_response", "rejected_response"]


def _flat_grads(cfg, w, lp, lr, images, q, qm, resp):
    from oracle import llava_oracle as O
    mods = [m for m in O.lora_target_modules(cfg) if "vision_model" not in m]
    lpg = {m: (A.clone().requires_grad_(m in mods), B.clone().requires_grad_(m in mods)) for m, (A, B) in lp.items()}
    with torch.no_grad():
        ref = O.policy_forward(cfg, w, lr, images, q, qm, resp)
    pol = O.policy_forward(cfg, w, lpg, images, q, qm, resp)
    loss, _, _ = O.pair_loss(O.LossConfig(beta=0.1), pol[NAMES[0] + "_logprobs"], pol[NAMES[1] + "_logprobs"],
                             ref[NAMES[0] + "_logprobs"], ref[NAMES[1] + "_logprobs"])
    loss.backward()
    return torch.cat([t.grad.flatten() for m in mods for t in lpg[m]]), loss.detach()


def _worker(rank, world, port, q_out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import llava_oracle as O
    from opa_dpo_b200.dist import allreduce_grad_bucket, shard_range
    cfg = O.OracleConfig(n_layers=1, v_layers=1)
    w = O.init_weights(cfg)
    lp, lr = O.init_lora(cfg, 11), O.init_lora(cfg, 12)
    images, q, qm, resp = O.synthetic_batch(cfg, 4, 16, 16, NAMES, left_pad=2)
    lo, hi = shard_range(4, rank, world)
    g, loss = _flat_grads(cfg, w, lp, lr, images[lo:hi], q[lo:hi], qm[lo:hi], {k: v[lo:hi] for k, v in resp.items()})
    scale = allreduce_grad_bucket(g)
    g = g * scale
    if rank == 0:
        g_full, loss_full = _flat_grads(cfg, w, lp, lr, images, q, qm, resp)
        q_out.put((float((g - g_full).norm() / g_full.norm()), scale))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_concatenated_batch():
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q_out)) for r in range(2)]
    for p in procs:
        p.start()
    rel, scale = q_out.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert scale == 0.5
    assert rel < 1e-4, f"DP gradient differs from the concatenated-batch gradient (rel {rel})"


def _overlap_worker(rank, world, port, q_out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from opa_dpo_b200.dist import OverlappedGradAllReduce

    class FakeEngine:  # stands in for the C-ABI engine: 3 "layers" + projector slices of a 100-element bucket
        ranges = {0: (0, 30), 1: (30, 30), 2: (60, 30), 3: (90, 10)}

        def set_backward_callback(self, fn):
            self.fn = fn

        def layer_grad_range(self, layer):
            return self.ranges[layer]

    eng = FakeEngine()
    g = torch.arange(100, dtype=torch.float32) * (rank + 1)
    ov = OverlappedGradAllReduce(eng, g)
    ov.arm()
    for layer in (2, 1, 0, 3):   # backward order, projector last
        eng.fn(layer)
    scale = ov.finish()
    expect = torch.arange(100, dtype=torch.float32) * sum(r + 1 for r in range(world))
    ok = torch.equal(g, expect) and scale == 1.0 / world
    # not armed -> falls back to one bucket-wide all-reduce
    g2 = torch.ones(100) * (rank + 1)
    ov2 = OverlappedGradAllReduce(eng, g2)
    scale2 = ov2.finish()
    ok = ok and torch.equal(g2, torch.ones(100) * 3) and scale2 == 0.5
    if rank == 0:
        q_out.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_overlapped_allreduce_slices_equal_bucket_allreduce():
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_overlap_worker, args=(r, 2, port, q_out)) for r in range(2)]
    for p in procs:
        p.start()
    assert q_out.get(timeout=120) is True
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0


def test_shard_range():
    from opa_dpo_b200.dist import shard_range
    assert [shard_range(64, r, 8) for r in (0, 7)] == [(0, 8), (56, 64)]
    with pytest.raises(ValueError):
        shard_range(10, 0, 4)


def _fake_nvl(n_layers, layer_elems, proj_elems, group_layers=4, blocks=16, tail_blocks=128, overlap=True):
    """NvlGradAllReduce with its launch replaced by a recorder (the launch plan is host logic; the kernel has its own
    hardware test in tests/dist_gpu_check.py)."""
    from types import SimpleNamespace
    from opa_dpo_b200.dist import NvlGradAllReduce
    total = n_layers * layer_elems + proj_elems

    def grad_range(layer):      # the engine's layout: decoder layers bottom-up, then the projector
        return (layer * layer_elems, layer_elems) if layer < n_layers else (n_layers * layer_elems, proj_elems)

    ar = object.__new__(NvlGradAllReduce)
    ar.engine = SimpleNamespace(cfg=SimpleNamespace(n_layers=n_layers), layer_grad_range=grad_range)
    ar.grad = torch.zeros(total)
    ar.world, ar.rank = 2, 0
    ar.group_layers, ar.blocks, ar.tail_blocks, ar.overlap = group_layers, blocks, tail_blocks, overlap
    ar.n_layers, ar.ranges, ar.pending, ar.launched, ar.armed = None, {}, None, False, False
    ar.err = torch.zeros(1, dtype=torch.int32)
    calls = []

    def launch(off, cnt, nblocks):
        calls.append((off, cnt, nblocks))
        ar.launched = True
    ar._launch = launch
    return ar, calls, total


@pytest.mark.parametrize("n_layers,group", [(32, 4), (32, 8), (6, 4), (2, 4), (1, 4), (5, 3)])
def test_nvl_allreduce_launch_plan_covers_the_bucket_exactly_once(n_layers, group):
    """The layer callback order of opadpo_backward (top layer first, projector last) turned into all-reduce launches: every
    element of the bucket is reduced exactly once, runs are contiguous and 8-aligned, grouped launches use the small grid,
    the last decoder layer and the projector the wide one; un-armed steps are one bucket-wide launch."""
    L, P = 8 * 1000, 8 * 70
    ar, calls, total = _fake_nvl(n_layers, L, P, group_layers=group)
    ar.arm()
    for layer in list(range(n_layers - 1, -1, -1)) + [n_layers]:
        ar._on_layer(layer)
    assert ar.pending is None        # layer 0 always flushes
    seen = torch.zeros(total, dtype=torch.int32)
    for off, cnt, nb in calls:
        assert off % 8 == 0 and cnt % 8 == 0 and cnt > 0
        seen[off:off + cnt] += 1
    assert bool((seen == 1).all())
    assert calls[-1] == (n_layers * L, P, 128)                      # the projector, wide grid
    assert calls[-2][0] == 0 and calls[-2][2] == 128                # the run that ends with layer 0, wide grid
    assert all(nb == 16 for _, _, nb in calls[:-2])
    assert len(calls) <= n_layers // group + 4
    # a step whose backward was not armed (gradient accumulation): one bucket-wide launch from finish()
    ar2, calls2, total2 = _fake_nvl(n_layers, L, P, group_layers=group)
    for layer in list(range(n_layers - 1, -1, -1)) + [n_layers]:
        ar2._on_layer(layer)
    assert calls2 == []
    ar2.comm = ar2.done = None
    ar2.finish = None    # (finish() needs CUDA streams; its un-armed branch is the single call below)
    if not ar2.launched:
        ar2._launch(0, ar2.grad.numel(), ar2.tail_blocks)
    assert calls2 == [(0, total2, 128)]
    # overlap switched off: the callbacks launch nothing
    ar3, calls3, _ = _fake_nvl(n_layers, L, P, group_layers=group, overlap=False)
    ar3.arm()
    for layer in list(range(n_layers - 1, -1, -1)) + [n_layers]:
        ar3._on_layer(layer)
    assert calls3 == []


def test_sharded_train_dataloader_partitions_each_round():
    """dist.get_train_dataloader: every rank walks the same shuffled batch sequence and keeps its share of each complete round
    (accelerate's BatchSamplerShard with even_batches / drop_last, rl_trainer.py:309-323): equal batch counts, disjoint
    samples inside a round, the un-sharded order preserved, a fresh permutation per epoch through the InfiniteLoader."""
    from torch.utils.data import BatchSampler, RandomSampler
    from opa_dpo_b200.dist import ShardedBatchSampler, get_train_dataloader
    data = list(range(23))
    collate = lambda items: torch.tensor(items)
    world, bs = 3, 2
    ref_gen = torch.Generator().manual_seed(5)
    plain = BatchSampler(RandomSampler(data, generator=ref_gen), batch_size=bs, drop_last=True)
    epoch1, epoch2 = list(plain), list(plain)             # 11 batches each -> 3 complete rounds of 3
    loaders = [get_train_dataloader(data, collate, bs, rank=r, world=world, seed=5) for r in range(world)]
    got = [[next(l).tolist() for _ in range(6)] for l in loaders]          # two epochs of 3 batches per rank
    for r in range(world):
        assert got[r][:3] == [epoch1[k * world + r] for k in range(3)]
        assert got[r][3:] == [epoch2[k * world + r] for k in range(3)]
    for k in range(6):
        round_items = [x for r in range(world) for x in got[r][k]]
        assert len(set(round_items)) == world * bs                          # one step's pairs are disjoint across ranks
    assert epoch1 != epoch2
    assert len(ShardedBatchSampler(plain, 0, world)) == 3
    single = get_train_dataloader(data, collate, bs, seed=5)
    assert [next(single).tolist() for _ in range(11)] == epoch1               # world 1 = the plain shuffled loader
    with pytest.raises(ValueError):
        ShardedBatchSampler(plain, 3, 3)


def test_sharded_batch_sampler_properties():
    """For any dataset size / batch size / world: every rank gets the same number of full batches, the ranks' batches of one
    round are the consecutive batches of the shared shuffled sequence, nothing is yielded from an incomplete round."""
    pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st
    from torch.utils.data import BatchSampler, SequentialSampler
    from opa_dpo_b200.dist import ShardedBatchSampler

    @settings(max_examples=200, deadline=None)
    @given(n=st.integers(0, 60), bs=st.integers(1, 7), world=st.integers(1, 8))
    def check(n, bs, world):
        base = BatchSampler(SequentialSampler(range(n)), batch_size=bs, drop_last=True)
        full = list(base)
        shards = [list(ShardedBatchSampler(base, r, world)) for r in range(world)]
        rounds = len(full) // world
        assert all(len(s) == rounds == len(ShardedBatchSampler(base, 0, world)) for s in shards)
        for k in range(rounds):
            assert [shards[r][k] for r in range(world)] == full[k * world:(k + 1) * world]
    check()
