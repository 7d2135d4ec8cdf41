"""World-size-2 coverage of the multi-GPU paths on real GPUs over NCCL; skipped on a single-GPU box.

 * batch data-parallel training: two ranks, each stepping half of a global batch through the native driver, end with the
   gradient (after the flat-bucket all-reduce and the 1/world mean) and the parameters (after Adam) of ONE rank stepping the
   whole batch;
 * the row-sharded table: all-to-all id exchange + lk_index_rows on the owner + all-to-all rows, bit-exact;
 * sharded cached evaluation: item slices all-gathered, users/impressions partitioned by group key, metric means all-reduced,
   equal to the single-rank values.
The CPU file tests/test_multirank_cpu.py covers the same host logic over gloo.
"""
import copy
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason='needs two GPUs')]

WORLD = 2
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _run(rank, port, fn, args):
    for p in (os.path.dirname(HERE), HERE, os.path.join(HERE, 'golden')):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=WORLD, device_id=torch.device('cuda', rank))
    try:
        fn(rank, *args)
        torch.cuda.synchronize()
    finally:
        dist.destroy_process_group()


def spawn(fn, *args):
    mp.spawn(_run, args=(_free_port(), fn, args), nprocs=WORLD, join=True)


# ---------------------------------------------------------------------------------------------------------------
def _sharded_lookup(rank):
    from legommenders_b200 import sharding
    dev = torch.device('cuda', rank)
    V, E = 100003, 300
    g = torch.Generator().manual_seed(5)
    table = torch.randn(V, E, generator=g)
    st = sharding.ShardedTable(sharding.shard_rows(table, rank, WORLD).to(dev), V)       # lk_index_rows on the owner
    g2 = torch.Generator().manual_seed(100 + rank)
    ids = torch.randint(0, V, (64, 33), generator=g2)
    ids[torch.rand(64, 33, generator=g2) < 0.3] = -1
    ids[0, :5] = ids[1, 5]
    out = st.lookup(ids.to(dev)).cpu()
    ref = torch.where((ids > -1).unsqueeze(-1), table[ids.clamp(min=0)], torch.zeros(()))
    assert torch.equal(out, ref)
    empty = torch.full((4,), -1, dtype=torch.int64) if rank == 0 else torch.arange(4)
    o = st.lookup(empty.to(dev)).cpu()
    assert torch.equal(o, torch.zeros(4, E) if rank == 0 else table[:4])


def test_sharded_table_lookup_nccl():
    spawn(_sharded_lookup)


# ---------------------------------------------------------------------------------------------------------------
def _build(rank):
    import cases
    import helpers
    from legommenders_b200 import builder
    c = cases.CASES['nrms_small']
    world, llm = cases.make_world(c)
    model, resampler, cfg = builder.build_model(world, c['kind'], hidden=c['hidden'], heads=c['heads'], additive=c['additive'],
                                                dropout=0.0, device_index=rank)
    np_state, _ = helpers.oracle_state(c, world, llm)
    builder.load_state(model, np_state)
    return c, world, model, resampler


def _dp_step(rank):
    from legommenders_b200 import Env
    from legommenders_b200.batching import BatchBuilder, tree_to_device
    from legommenders_b200.trainer import FlatAdam, NativeNRMSStep
    c, world, model, resampler = _build(rank)
    ref_model = copy.deepcopy(model)
    Env.train()
    rows = np.arange(32) % world.n_train
    bb = BatchBuilder(resampler, world, neg_count=4, seed=9, pin=False)
    full = bb.train_batch(rows)

    def half(tree, r):
        if isinstance(tree, dict):
            return type(tree)((k, half(v, r)) for k, v in tree.items())
        return tree[r * 16:(r + 1) * 16].contiguous()

    # reference: one rank, the whole batch (world forced to 1: no all-reduce, no 1/world)
    ropt = FlatAdam(ref_model, lr=1e-3)
    ropt.world = 1
    rnat = NativeNRMSStep(ref_model, ropt)
    rloss = rnat.fwd_bwd(tree_to_device(full, Env.device), training=False).item()
    rgrad = ropt.grad.clone()
    ropt.step()
    # data parallel: this rank's half
    opt = FlatAdam(model, lr=1e-3)
    assert opt.world == WORLD
    nat = NativeNRMSStep(model, opt)
    loss = nat.fwd_bwd(tree_to_device(half(full, rank), Env.device), training=False)
    lsum = loss.clone()
    dist.all_reduce(lsum)
    assert abs(lsum.item() / WORLD - rloss) <= 1e-5 * abs(rloss)                        # mean of equal-sized halves = global mean
    opt.allreduce()
    scale = rgrad.abs().max().item()
    assert (opt.grad / WORLD - rgrad).abs().max().item() <= 2e-5 * scale
    opt.world = 1                                                                       # the all-reduce above already ran
    opt.step_count += 1
    from legommenders_b200 import ops
    ops.adam_step(opt.flat, opt.grad, opt.m, opt.v, opt.step_count, lr=opt.lr, grad_scale=1.0 / WORLD)
    # Adam's first step moves every parameter by lr * sign(g) (up to eps): compare where the gradient is not round-off
    big = rgrad.abs() > 1e-3 * scale
    assert (opt.flat - ropt.flat)[big].abs().max().item() <= 1e-5
    both = [torch.empty_like(opt.flat) for _ in range(WORLD)]
    dist.all_gather(both, opt.flat)
    assert torch.equal(both[0], both[1])                                                # replicas stay bit-identical


def test_data_parallel_step_nccl():
    spawn(_dp_step)


def _sharded_table_step(rank):
    """Config 4: the native step fed from the row-sharded word table gives the loss and gradients of the replicated table."""
    from legommenders_b200 import Env, sharding
    from legommenders_b200.batching import BatchBuilder, tree_to_device
    from legommenders_b200.trainer import FlatAdam, NativeNRMSStep
    c, world, model, resampler = _build(rank)
    Env.train()
    rows = (np.arange(16) + 16 * rank) % world.n_train            # different impressions per rank
    batch = tree_to_device(BatchBuilder(resampler, world, neg_count=4, seed=3 + rank, pin=False).train_batch(rows), Env.device)
    opt = FlatAdam(model, lr=1e-3)
    nat = NativeNRMSStep(model, opt)
    l0 = nat.fwd_bwd(batch, training=False).item()
    g0 = opt.grad.clone()
    st = sharding.ShardedTable(sharding.shard_rows(nat.glove.detach(), rank, WORLD), nat.glove.shape[0])
    nat2 = NativeNRMSStep(model, opt, sharded_table=st)
    opt.grad.zero_()
    l1 = nat2.fwd_bwd(batch, training=False).item()
    assert l1 == l0                                               # the same rows reach the same kernels: bit-identical
    assert torch.equal(opt.grad, g0)


def _p2p_allreduce(rank):
    """FlatAdam's peer-memory all-reduce (lk_allreduce_p2p between two symmetric-memory barriers) == ncclAllReduce of the same buckets,
    bit-identical on all ranks, repeatedly (the barriers must order step k's reads before step k+1's writes)."""
    from legommenders_b200.trainer import FlatAdam
    dev = torch.device('cuda', rank)
    m = torch.nn.Sequential(torch.nn.Linear(301, 257), torch.nn.Linear(257, 33)).to(dev)
    opt = FlatAdam(m, lr=1e-3)
    assert opt.symm is not None, 'gradient bucket is not in symmetric memory (peer access unavailable?)'
    for it in range(20):
        g = torch.Generator(device=dev).manual_seed(1000 * it + rank)
        opt.grad.normal_(0, 1, generator=g)
        want = opt.grad.clone()
        dist.all_reduce(want)
        opt.allreduce()
        torch.testing.assert_close(opt.grad, want, rtol=1e-6, atol=1e-6)
        other = opt.grad.clone()
        dist.broadcast(other, src=0)
        assert torch.equal(other, opt.grad)                      # every element summed once, by one rank


def test_p2p_allreduce_matches_nccl():
    spawn(_p2p_allreduce)


def test_sharded_table_native_step_nccl():
    spawn(_sharded_table_step)


# ---------------------------------------------------------------------------------------------------------------
def _sharded_eval(rank):
    import cases
    from legommenders_b200 import DataSet, Env, evaluate
    c, world, model, resampler = _build(rank)
    g = cases.load('nrms_small')
    Env.test()
    model.eval()
    evaluate.build_caches(model, resampler.item_cache, DataSet(world.fast_table(), resampler))
    vals, scores, rows = evaluate.evaluate(model, torch.from_numpy(world.eval_users), torch.from_numpy(world.eval_items),
                                           torch.from_numpy(g['eval_labels']))
    assert rows is not None and torch.all(torch.from_numpy(world.eval_users)[rows] % WORLD == rank)
    ref = g['eval_scores'][rows.numpy()]
    assert np.abs(scores.cpu().numpy() - ref).max() <= 1e-4 * np.abs(g['eval_scores']).max()
    for (k, v), r in zip(vals.items(), g['metrics']):
        assert round(v, 4) == round(float(r), 4), k


def test_sharded_cached_eval_nccl():
    spawn(_sharded_eval)
