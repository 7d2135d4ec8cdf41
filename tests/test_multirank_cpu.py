"""World-size-2 coverage of the multi-GPU host logic over gloo on CPU.

What runs here is the partitioning / exchange bookkeeping of legommenders_b200.sharding and trainer.FlatAdam's bucket
all-reduce.  The row gather of the product is a CUDA kernel with no CPU implementation; these tests inject a plain
indexing checker through `gather_fn` (test infrastructure, not a fallback of the product).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

WORLD = 2


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _run(rank, port, fn, args):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=WORLD)
    try:
        fn(rank, *args)
    finally:
        dist.destroy_process_group()


def spawn(fn, *args):
    mp.spawn(_run, args=(_free_port(), fn, args), nprocs=WORLD, join=True)


def _index(table, ids):
    return table[ids.reshape(-1)].reshape(*ids.shape, table.shape[1])


# ---------------------------------------------------------------------------------------------------------------
def _sharded_lookup(rank):
    from legommenders_b200 import sharding
    V, E = 1001, 12
    g = torch.Generator().manual_seed(5)
    table = torch.randn(V, E, generator=g)
    st = sharding.ShardedTable(sharding.shard_rows(table, rank, WORLD), V, gather_fn=_index)
    assert st.world == WORLD and st.local.shape[0] == len(range(rank, V, WORLD))
    g2 = torch.Generator().manual_seed(100 + rank)            # every rank asks for different ids
    ids = torch.randint(0, V, (7, 33), generator=g2)
    ids[torch.rand(7, 33, generator=g2) < 0.3] = -1           # unset positions (Env.UNSET)
    ids[0, :5] = ids[1, 5]                                     # duplicates
    out = st.lookup(ids)
    ref = torch.where((ids > -1).unsqueeze(-1), table[ids.clamp(min=0)], torch.zeros(()))
    assert torch.equal(out, ref)                               # moved bytes, not arithmetic: bit-exact
    rows, inverse = st.lookup_unique(ids)
    assert torch.equal(inverse > -1, ids > -1)
    assert torch.equal(rows[inverse[ids > -1]], table[ids[ids > -1]])
    assert rows.shape[0] == torch.unique(ids[ids > -1]).numel()
    # an empty request on one rank must not dead-lock the other
    empty = torch.full((4,), -1, dtype=torch.int64) if rank == 0 else torch.arange(4)
    o = st.lookup(empty)
    assert torch.equal(o, torch.zeros(4, E) if rank == 0 else table[:4])


def test_sharded_table_lookup_world2():
    spawn(_sharded_lookup)


def test_plan_lookup_buckets():
    from legommenders_b200 import sharding
    ids = torch.tensor([[5, -1, 2, 9], [2, 5, 4, -1]])
    p = sharding.plan_lookup(ids, 3)
    assert p.uniq.tolist() == [2, 4, 5, 9]
    assert p.inverse.tolist() == [2, -1, 0, 3, 0, 2, 1, -1]
    assert p.uniq[p.order].tolist() == [9, 4, 2, 5]           # owners 0,1,2,2 — ascending ids inside a bucket
    assert p.send_counts.tolist() == [1, 1, 2]


# ---------------------------------------------------------------------------------------------------------------
def _eval_partition(rank):
    from legommenders_b200 import sharding
    rng = np.random.default_rng(3)
    groups = torch.from_numpy(rng.integers(0, 50, size=400))
    rows = sharding.owned_rows(groups, rank, WORLD)
    assert torch.all(groups[rows] % WORLD == rank)
    counts = torch.tensor([rows.numel()])
    dist.all_reduce(counts)
    assert counts.item() == 400                                 # a partition: every row owned exactly once
    # replicated item cache from per-rank slices
    n_items, D = 11, 4
    full = torch.arange(n_items * D, dtype=torch.float32).reshape(n_items, D)
    a, b = sharding.item_slice(n_items, rank, WORLD)
    got = sharding.gather_item_cache(full[a:b].clone(), n_items)
    assert torch.equal(got, full)
    # global mean over groups from per-rank means
    vals = torch.from_numpy(rng.random(50))
    mine = vals[rank::WORLD]
    means, total = sharding.reduce_group_means(mine.mean().reshape(1), mine.numel())
    assert total == 50 and abs(means.item() - vals.mean().item()) < 1e-12


def test_eval_partition_world2():
    spawn(_eval_partition)


def test_item_slices_cover_everything():
    from legommenders_b200 import sharding
    for n in (0, 1, 7, 8, 65238):
        for w in (1, 2, 3, 8):
            sl = [sharding.item_slice(n, r, w) for r in range(w)]
            assert sl[0][0] == 0 and sl[-1][1] == n and all(sl[i][1] == sl[i + 1][0] for i in range(w - 1))


# ---------------------------------------------------------------------------------------------------------------
def _flat_bucket_allreduce(rank):
    from legommenders_b200.trainer import FlatAdam
    torch.manual_seed(0)                                        # identical replicas
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 3))
    opt = FlatAdam(model, lr=1e-3)
    assert opt.world == WORLD
    opt.zero_grad()
    x = torch.full((4, 6), float(rank + 1))
    model(x).sum().backward()                                   # autograd accumulates into the flat views
    local = opt.grad.clone()
    opt.allreduce()
    both = [torch.empty_like(local) for _ in range(WORLD)]
    dist.all_gather(both, local)
    assert torch.allclose(opt.grad, both[0] + both[1])
    for p in model.parameters():                                # parameters and grads are views of the flat buffers
        assert p.grad.data_ptr() >= opt.grad.data_ptr()


def test_flat_bucket_allreduce_world2():
    spawn(_flat_bucket_allreduce)


def test_balanced_partition_equalises_packed_rows():
    """batching.balanced_partition: every rank gets the same number of impressions and (within a few rows) the same packed-token cost; the
    global batch is only re-placed, never changed."""
    from legommenders_b200.batching import balanced_partition
    rng = np.random.default_rng(3)
    for world, per in ((2, 64), (8, 64), (4, 5)):
        cost = rng.integers(40, 1300, size=world * per).astype(np.float64)
        part = balanced_partition(cost, world)
        assert part.shape == (world, per)
        assert sorted(part.reshape(-1).tolist()) == list(range(world * per))
        sums = cost[part].sum(axis=1)
        naive = cost.reshape(world, per).sum(axis=1)
        assert sums.max() - sums.min() <= cost.max()                      # within one impression of each other
        assert sums.max() - sums.min() <= naive.max() - naive.min()
    with pytest.raises(ValueError):
        balanced_partition(np.ones(10), 4)
