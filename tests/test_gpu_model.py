"""End-to-end parity of the CUDA path behind the reference's plugin surface: loss, scores, representations, every
parameter gradient, the two caches, cached scoring and the five default metrics — against the committed golden
vectors (minted from the live reference) and against the oracle on the same inputs."""
import copy
import random

import numpy as np
import pytest
import torch

import cases
import helpers
from oracle import lego_oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-4   # BASELINE.json north_star: 1e-4 relative on fp32 logits and losses (norm-wise)


def build(c, world, llm):
    from legommenders_b200 import builder
    model, resampler, cfg = builder.build_model(world, c['kind'], hidden=c['hidden'], heads=c['heads'], additive=c['additive'],
                                                dropout=0.0, use_neg_sampling=c.get('use_neg_sampling', True),
                                                llm_item_table=llm)
    np_state, _ = helpers.oracle_state(c, world, llm)
    builder.load_state(model, np_state)
    return model, resampler, cfg


@pytest.fixture(params=['packed', 'padded'])
def layout(request):
    """Every model-level parity test runs padding-free (default) and in the reference's padded layout."""
    from legommenders_b200 import Legommender
    old = Legommender.packed
    Legommender.packed = request.param == 'packed'
    yield request.param
    Legommender.packed = old


@pytest.mark.parametrize('name', list(cases.CASES))
def test_train_step_parity(name, layout):
    from legommenders_b200 import Env
    c = cases.CASES[name]
    g = cases.load(name)
    world, llm = cases.make_world(c)
    model, resampler, cfg = build(c, world, llm)
    batch = cases.unflatten_batch(g)
    ora = helpers.oracle_run(c, world, llm, batch)

    Env.train()
    model.train()
    loss = model(batch=copy.deepcopy(batch))
    loss.backward()
    for ref_loss in (float(g['loss']), ora['loss']):
        assert abs(loss.item() - ref_loss) <= TOL * abs(ref_loss)
    grads = {n: (p.grad.detach().cpu().numpy() if p.grad is not None else np.zeros(tuple(p.shape), dtype=np.float32))   # unused: CNNCat's linear
             for n, p in model.named_parameters() if p.requires_grad}
    helpers.check_grads(c, grads, ora['grads'], TOL, golden=g)

    Env.test()
    with torch.no_grad():
        scores = model(batch=copy.deepcopy(batch)).cpu().numpy()
        user = model.get_user_content(copy.deepcopy(batch)).cpu().numpy()
        assert helpers.normwise(scores, g['scores']) <= TOL and helpers.normwise(scores, ora['scores']) <= TOL
        assert helpers.normwise(user, g['user']) <= TOL
        if 'items' in g.files:
            items = model.get_item_content(copy.deepcopy(batch), 'item_id').cpu().numpy()
            assert helpers.normwise(items, g['items']) <= TOL
    Env.train()


@pytest.mark.parametrize('name', ['nrms_small', 'naml_small'])
def test_batch_builder_bit_exact(name):
    """Resampler + inputers + default_collate reproduce the reference batch bit for bit under the same `random` seed."""
    from torch.utils.data import DataLoader
    from legommenders_b200 import DataSet, Env
    c = cases.CASES[name]
    g = cases.load(name)
    world, llm = cases.make_world(c)
    model, resampler, cfg = build(c, world, llm)
    Env.train()
    random.seed(c['seed'])
    batch = next(iter(DataLoader(DataSet(world.train_table(), resampler), batch_size=c['batch'], num_workers=0, shuffle=False)))
    flat = cases.flatten_tree(batch)
    keys = [k for k in g.files if k.startswith('batch/')]
    assert sorted('batch/' + k for k in flat) == sorted(keys)
    for k, v in flat.items():
        assert v.dtype == torch.int64
        assert np.array_equal(v.numpy(), g['batch/' + k]), k


@pytest.mark.parametrize('name', [n for n, c in cases.CASES.items() if c.get('cached_eval')])
def test_cached_eval_parity(name, layout):
    from torch.utils.data import DataLoader
    from legommenders_b200 import DataSet, Env, ops
    c = cases.CASES[name]
    g = cases.load(name)
    world, llm = cases.make_world(c)
    model, resampler, cfg = build(c, world, llm)
    Env.test()
    model.eval()
    model.cacher.cache(item_contents=resampler.item_cache, user_contents=DataSet(world.fast_table(), resampler))
    assert Env.user_cache and model.cacher.user.cached
    if cfg.use_item_content:
        assert Env.item_cache and model.cacher.item.cached
        assert helpers.normwise(model.cacher.item.repr.cpu().numpy(), g['item_repr']) <= TOL
    assert helpers.normwise(model.cacher.user.repr.cpu().numpy(), g['user_repr']) <= TOL

    # the reference's own loop: 64-row batches of (user, item) ids through forward()
    sc = []
    with torch.no_grad():
        for eb in DataLoader(DataSet(world.eval_table(), resampler), batch_size=64, num_workers=0, shuffle=False):
            assert set(eb) == {'index', 'user_id', 'item_id', 'click'}
            sc.append(model(batch=eb).reshape(-1).cpu())
    sc = torch.cat(sc).numpy()
    assert helpers.normwise(sc, g['eval_scores']) <= TOL
    m = O.metric_pool(sc, g['eval_labels'], g['eval_groups'])
    for (k, v), ref in zip(m.items(), g['metrics']):
        assert round(v, 4) == round(float(ref), 4), k

    # one-launch scoring of every row (the cached-eval kernel proper)
    if cfg.use_item_content:
        allsc = ops.cached_scores(model.cacher.user.repr, model.cacher.item.repr,
                                  torch.from_numpy(world.eval_users).cuda(), torch.from_numpy(world.eval_items).cuda()).cpu().numpy()
        assert np.array_equal(allsc, sc)
    model.cacher.clean()
    assert not Env.item_cache and not Env.user_cache
    Env.train()


def test_state_dict_names_match_reference():
    c = cases.CASES['nrms_small']
    world, llm = cases.make_world(c)
    model, _, _ = build(c, world, llm)
    assert set(model.state_dict()) == set(helpers.state_shapes(c, world))
    c = cases.CASES['naml_small']
    world, llm = cases.make_world(c)
    model, _, _ = build(c, world, llm)
    assert set(model.state_dict()) == set(helpers.state_shapes(c, world))


def test_training_reduces_loss_with_dropout():
    """A few Adam steps with the reference's dropout rates on: loss is finite and goes down on a fixed batch."""
    from legommenders_b200 import Env, ops
    from legommenders_b200.trainer import FlatAdam
    c = cases.CASES['nrms_small']
    g = cases.load('nrms_small')
    world, llm = cases.make_world(c)
    from legommenders_b200 import builder
    torch.manual_seed(0)
    model, _, _ = builder.build_model(world, 'nrms', hidden=64, heads=8, additive=32, dropout=0.1)
    opt = FlatAdam(model, lr=1e-3)
    batch = cases.unflatten_batch(g)
    Env.train()
    model.train()
    losses = []
    for _ in range(30):
        opt.zero_grad()
        loss = model(batch=copy.deepcopy(batch))
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(np.isfinite(losses))
    assert np.mean(losses[-5:]) < np.mean(losses[:5])


@pytest.mark.parametrize('name', ['nrms_small', 'nrms_full', 'nrms_small_hot', 'nrms_full_hot', 'nrms_h100'])
def test_native_step_driver_parity(name):
    """lk_nrms_fwd_bwd (one C-ABI call for the whole forward+backward) against the oracle and the autograd path."""
    from legommenders_b200 import Env
    from legommenders_b200.trainer import FlatAdam, NativeNRMSStep
    c = cases.CASES[name]
    g = cases.load(name)
    world, llm = cases.make_world(c)
    model, resampler, cfg = build(c, world, llm)
    batch = cases.unflatten_batch(g)
    ora = helpers.oracle_run(c, world, llm, batch)
    Env.train()
    model.train()
    opt = FlatAdam(model, lr=1e-3)
    native = NativeNRMSStep(model, opt)
    opt.zero_grad()
    loss = native.fwd_bwd(copy.deepcopy(batch), training=False)          # dropout off: deterministic parity
    assert abs(loss.item() - ora['loss']) <= TOL * abs(ora['loss'])
    assert abs(loss.item() - float(g['loss'])) <= TOL * abs(float(g['loss']))
    grads = {n: p.grad.detach().cpu().numpy().copy() for n, p in model.named_parameters() if p.requires_grad}
    scale = max(np.abs(v).max() for v in ora['grads'].values())
    helpers.check_grads(c, grads, ora['grads'], TOL, golden=g)      # per tensor, no floor, on the "hot" cases: the benched path's own plumbing
    # and the autograd path on the same model gives the same numbers
    opt.zero_grad()
    loss2 = model(batch=copy.deepcopy(batch))
    loss2.backward()
    assert abs(loss2.item() - loss.item()) <= 2e-5 * abs(loss.item())
    for n, p in model.named_parameters():
        if p.requires_grad:
            assert np.abs(p.grad.cpu().numpy() - grads[n]).max() <= helpers.grad_bound(c, np.abs(grads[n]).max(), scale, 5e-5), n
    # a few optimiser steps through the native driver with dropout on: finite and decreasing (plain states only: the boosted w2 of the hot
    # states drives exp(score) of the reference's un-shifted additive softmax to overflow within a few Adam steps, in the reference too)
    if c.get('boost'):
        return
    losses = []
    for _ in range(25):
        losses.append(native.step(copy.deepcopy(batch)).item())
    assert all(np.isfinite(losses)) and np.mean(losses[-5:]) < np.mean(losses[:5])


def test_exact_fp32_mode_gradients(monkeypatch):
    """With the tensor-core contractions switched off (ops.USE_TC = False: the exact-fp32 SIMT GEMMs) every gradient of the hot state agrees
    with the fp64 oracle to 2e-5 of its own max — the split-bf16 planes, not the kernels' logic, set the 1e-4..3e-4 of the default mode."""
    from legommenders_b200 import Env, ops
    monkeypatch.setattr(ops, 'USE_TC', False)
    c = cases.CASES['nrms_small_hot']
    g = cases.load('nrms_small_hot')
    world, llm = cases.make_world(c)
    model, _, _ = build(c, world, llm)
    batch = cases.unflatten_batch(g)
    ref = helpers.oracle_run(c, world, llm, batch, dtype=torch.float64)
    Env.train()
    model.train()
    loss = model(batch=copy.deepcopy(batch))
    loss.backward()
    assert abs(loss.item() - ref['loss']) <= 2e-6 * abs(ref['loss'])
    for n, p in model.named_parameters():
        if p.requires_grad:
            r = ref['grads'][n]
            assert np.abs(p.grad.cpu().numpy() - r).max() <= 2e-5 * np.abs(r).max(), n


# ------------------------------------------------------------------------------------------------ evaluation on the device
@pytest.mark.parametrize('R,G,ties', [(2000, 60, True), (5000, 400, False), (64, 1, True), (70000, 2500, True)])
def test_group_metrics_kernel_matches_oracle(R, G, ties):
    """lk_group_metrics (GAUC / MRR / nDCG@k with sklearn tie semantics) against the oracle's restatement of
    utils/metrics.py; groups are interleaved (rows of a group are not contiguous) and scores contain ties."""
    from legommenders_b200 import Env
    from legommenders_b200.metrics import MetricPool
    Env.use_cuda(0)
    rng = np.random.default_rng(R + G)
    groups = rng.integers(0, G, size=R).astype(np.int64) * 7 - 3          # arbitrary (also negative) keys
    scores = rng.standard_normal(R).astype(np.float32)
    if ties:
        scores = np.round(scores * 2) / 2                                   # heavy ties
    labels = (rng.random(R) < 0.3).astype(np.int64)
    for g in np.unique(groups):                                             # both classes in every group
        idx = np.flatnonzero(groups == g)
        if len(idx) < 2:
            groups[idx] = groups[0]
    for g in np.unique(groups):
        idx = np.flatnonzero(groups == g)
        labels[idx[0]], labels[idx[-1]] = 1, 0
    names = ['GAUC', 'MRR', 'NDCG@1', 'NDCG@5', 'NDCG@10']
    ref = O.metric_pool(scores, labels, groups, names)
    pool = MetricPool.parse(names)
    got = pool.calculate(torch.from_numpy(scores), torch.from_numpy(labels), torch.from_numpy(groups), per_group=True)
    assert pool.n_groups == len(np.unique(groups))
    for k in names:
        assert abs(got[k] - ref[k]) <= 2e-6, (k, got[k], ref[k])
    # per-group values, group by group (ascending key order = pandas groupby order)
    keys = np.unique(groups)
    pg = pool.per_group.cpu().numpy()
    for gi in rng.choice(len(keys), size=min(20, len(keys)), replace=False):
        idx = np.flatnonzero(groups == keys[gi])
        assert abs(pg[0, gi] - O.auc(scores[idx], labels[idx])) <= 1e-6
        assert abs(pg[1, gi] - O.mrr(scores[idx], labels[idx])) <= 1e-6
        assert abs(pg[4, gi] - O.ndcg(scores[idx], labels[idx], 10)) <= 1e-6


@pytest.mark.parametrize('name', ['nrms_small', 'naml_small'])
def test_evaluate_matches_golden_metrics(name):
    """evaluate.py: caches -> one-kernel scoring -> one-kernel metrics; equal to the reference's metrics at 4 decimals."""
    from legommenders_b200 import DataSet, Env, evaluate
    c = cases.CASES[name]
    g = cases.load(name)
    world, llm = cases.make_world(c)
    model, resampler, cfg = build(c, world, llm)
    Env.test()
    model.eval()
    evaluate.build_caches(model, resampler.item_cache, DataSet(world.fast_table(), resampler))
    vals, scores, rows = evaluate.evaluate(model, torch.from_numpy(world.eval_users), torch.from_numpy(world.eval_items),
                                           torch.from_numpy(g['eval_labels']))
    assert rows is None
    assert helpers.normwise(scores.cpu().numpy(), g['eval_scores']) <= TOL
    for (k, v), ref in zip(vals.items(), g['metrics']):
        assert round(v, 4) == round(float(ref), 4), k
    model.cacher.clean()
    Env.train()


def test_device_batcher_packs_bit_exact():
    """Id-only batches (DeviceBatcher + lk_pack_item_tokens) give exactly the packed rows that packing.pack_tokens derives from the
    wire-format batch of the same candidates, and the native step computes the same loss from either."""
    from legommenders_b200 import Env
    from legommenders_b200.batching import BatchBuilder, DeviceBatcher, tree_to_device
    from legommenders_b200.trainer import FlatAdam, NativeNRMSStep
    c = cases.CASES['nrms_small']
    world, llm = cases.make_world(c)
    model, resampler, cfg = build(c, world, llm)
    Env.train()
    rows = np.arange(16) % world.n_train
    wire = BatchBuilder(resampler, world, neg_count=4, seed=5).train_batch(rows)
    dbat = DeviceBatcher(resampler, world, Env.device, neg_count=4, seed=5)
    hb = dbat.host_batch(rows)
    idb = dbat.to_device(hb)
    opt = FlatAdam(model, lr=1e-3)
    native = NativeNRMSStep(model, opt)
    pk_w, cu_u_w, max_u_w, B_w, C_w = native.pack(tree_to_device(wire, Env.device))
    pk_i, cu_u_i, max_u_i, B_i, C_i = native.pack(idb)
    assert (pk_w.n, pk_w.rows, pk_w.max_len, max_u_w, B_w, C_w) == (pk_i.n, pk_i.rows, pk_i.max_len, max_u_i, B_i, C_i)
    assert torch.equal(pk_w.cu.cpu(), pk_i.cu.cpu()) and torch.equal(cu_u_w.cpu(), cu_u_i.cpu())
    for col in pk_w.ids:
        assert torch.equal(pk_w.ids[col].cpu(), pk_i.ids[col].cpu()), col
    l1 = native.fwd_bwd(tree_to_device(wire, Env.device), training=False).item()
    l2 = native.fwd_bwd(idb, training=False).item()
    assert l1 == l2
    assert DeviceBatcher.h2d_bytes(hb) < 0.05 * sum(v.numel() * 8 for v in wire['history']['input_ids'].values())


@pytest.mark.parametrize('B', [16, 5])
def test_device_resampler_bit_exact(B):
    """batching.DeviceResampler (lk_resample_batch: negative sampling + history concat + offsets in one kernel, one step ahead on a side
    stream): (1) items / offsets equal the host replay of the same Philox draws bit for bit; (2) candidate order, negative provenance and
    history/mask semantics are the Resampler's (loader/resampler.py:139-259, oracle.candidates / pad_history); (3) the packed token rows equal
    packing.pack_tokens of the WIRE batch built for the same candidates, and the native step computes the same loss from either."""
    from legommenders_b200 import Env
    from legommenders_b200.batching import BatchBuilder, DeviceResampler, tree_to_device
    from legommenders_b200.trainer import FlatAdam, NativeNRMSStep
    c = cases.CASES['nrms_small']
    world, llm = cases.make_world(c)
    model, resampler, cfg = build(c, world, llm)
    Env.train()
    K = 4
    dres = DeviceResampler(resampler, world, Env.device, neg_count=K, seed=41, max_batch=16)
    rng = np.random.default_rng(B)
    all_rows = [rng.integers(0, world.n_train, size=B) for _ in range(5)]
    dres.submit(all_rows[0])
    opt = FlatAdam(model, lr=1e-3)
    native = NativeNRMSStep(model, opt)
    bb = BatchBuilder(resampler, world, neg_count=K, seed=0, pin=False)
    for i, rows in enumerate(all_rows):
        if i + 1 < len(all_rows):
            dres.submit(all_rows[i + 1])            # one step ahead, more slots than batches in flight
        batch = dres.take()
        pk, cu_u, max_u, Bb, C = batch['__lk_packed__']
        items, cu, cu_users = dres.replay_host(rows)
        assert (Bb, C, pk.n, pk.rows) == (B, K + 1, len(items), int(cu[-1]))
        assert np.array_equal(batch['__items__'].cpu().numpy(), items)
        assert np.array_equal(pk.cu.cpu().numpy(), cu) and np.array_equal(cu_u.cpu().numpy(), cu_users)
        assert np.array_equal(batch['user_id'].cpu().numpy(), world.train_users[rows])
        assert pk.max_len == int(np.diff(cu).max()) and max_u == int(np.diff(cu_users).max())
        cand = items[:B * C].reshape(B, C)
        for b, r in enumerate(rows):
            u = int(world.train_users[r])
            negs = list(world.negs[u])
            k = min(K, len(negs))
            assert cand[b, 0] == world.train_pos[r]
            assert all(int(x) in negs for x in cand[b, 1:1 + k])             # true negatives first ...
            assert all(0 <= int(x) < world.n_items for x in cand[b, 1 + k:])  # ... then uniform ids
            assert cand[b].tolist() == O.candidates(int(cand[b, 0]), cand[b, 1:1 + k], cand[b, 1 + k:]).tolist()
            hist_ids, hist_mask = O.pad_history(world.histories[u], world.hist_len)
            got = items[B * C + cu_users[b]: B * C + cu_users[b + 1]]
            assert np.array_equal(got, hist_ids[hist_mask > 0]) and len(got) == int(hist_mask.sum())
        wire = bb.train_batch(rows, cand=cand)
        pk_w, cu_u_w, max_u_w, B_w, C_w = native.pack(tree_to_device(wire, Env.device))
        assert (pk_w.n, pk_w.rows, pk_w.max_len, max_u_w, B_w, C_w) == (pk.n, pk.rows, pk.max_len, max_u, Bb, C)
        for col in pk_w.ids:
            assert torch.equal(pk_w.ids[col], pk.ids[col]), col
        l_dev = native.fwd_bwd(batch, training=False).item()
        l_wire = native.fwd_bwd(tree_to_device(wire, Env.device), training=False).item()
        assert l_dev == l_wire
    with pytest.raises(RuntimeError):
        dres.take()


def test_device_cache_build_matches_cacher():
    """evaluate.build_caches_device (id lists only, packed encoders, users from item-cache rows) leaves the same two caches as the
    cacher contract (ReprCacher.cache over the Resampler's per-item / per-user contents), and the same metrics."""
    from legommenders_b200 import DataSet, Env, evaluate
    from legommenders_b200.batching import DeviceBatcher
    c = cases.CASES['nrms_small']
    g = cases.load('nrms_small')
    world, llm = cases.make_world(c)
    model, resampler, cfg = build(c, world, llm)
    Env.test()
    model.eval()
    model.cacher.cache(item_contents=resampler.item_cache, user_contents=DataSet(world.fast_table(), resampler))
    ref_i, ref_u = model.cacher.item.repr.clone(), model.cacher.user.repr.clone()
    model.cacher.clean()
    assert not Env.item_cache and not Env.user_cache
    dbat = DeviceBatcher(resampler, world, Env.device)
    evaluate.build_caches_device(model, dbat, item_page=300, user_page=70)          # several ragged pages
    assert Env.item_cache and Env.user_cache and model.cacher.item.cached and model.cacher.user.cached
    assert helpers.normwise(model.cacher.item.repr.cpu().numpy(), ref_i.cpu().numpy()) <= 1e-5
    assert helpers.normwise(model.cacher.user.repr.cpu().numpy(), ref_u.cpu().numpy()) <= 1e-5
    assert helpers.normwise(model.cacher.user.repr.cpu().numpy(), g['user_repr']) <= TOL
    vals, scores, _ = evaluate.evaluate(model, torch.from_numpy(world.eval_users), torch.from_numpy(world.eval_items),
                                        torch.from_numpy(g['eval_labels']))
    for (k, v), ref in zip(vals.items(), g['metrics']):
        assert round(v, 4) == round(float(ref), 4), k
    model.cacher.clean()
    Env.train()
