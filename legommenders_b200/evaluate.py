"""Cached evaluation (mirror of base_lego.py:355-430 `base_evaluate` + `evaluate` on the fast-eval path).

The reference walks the test set in 64-row batches, calls the model (which indexes the two caches and takes a dot product,
model/legommender.py:153-157, 202-203), copies every batch of scores to the host and finally hands python lists to
MetricPool.  Here the whole test set is scored by ONE kernel per chunk of rows (`lk_cached_scores`) and the group metrics are
ONE more kernel (`lk_group_metrics`); with several ranks, users / impressions are sharded by group key and the item cache
is replicated.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional, Sequence

import torch
import torch.distributed as dist

from . import ops, sharding
from .env import Env
from .metrics import MetricPool

DEFAULT_METRICS = ('GAUC', 'MRR', 'NDCG@1', 'NDCG@5', 'NDCG@10')


def build_caches(model, item_contents, user_contents, group=None):
    """`model.cacher.cache(...)`, sharded when torch.distributed is initialised: every rank encodes a contiguous slice of the
    items (all-gathered into the replicated cache) and only the users it owns (`user_id % world == rank`; the fast-eval
    dataset is in user-id order, manager.py:209-227)."""
    rank, world = sharding._world(group)
    cacher = model.cacher
    if world == 1:
        cacher.cache(item_contents=item_contents, user_contents=user_contents)
        return
    if cacher.use_item_content:
        a, b = sharding.item_slice(len(item_contents), rank, world)
        cacher.item.cache(item_contents[a:b])
        cacher.item.repr = sharding.gather_item_cache(cacher.item.repr, len(item_contents), group)
    cacher.user.rows = range(rank, len(user_contents), world)
    try:
        cacher.user.cache(user_contents)
    finally:
        cacher.user.rows = None


def build_caches_device(model, batcher, item_page: int = 8192, user_page: int = 4096, group=None):
    """Both representation caches from id lists only (the fast pagers of the reference, loader/pager/fast_{item,user}_pager.py, taken
    to the device): the per-item token layouts already live on the device in `batcher` (batching.DeviceBatcher), so a page of items is
    one `lk_pack_item_tokens` launch + the packed item encoder, and a page of users is an `lk_index_rows` gather of its history items'
    rows from the fresh item cache (the reference's own short-circuit, legommender.py:153-157) + the packed user encoder.  No per-item or
    per-user Python, nothing but id lists and offsets crosses PCIe.  Leaves the cachers exactly as `ReprCacher.cache` does.
    Needs the packed item / user operators (NRMS / Ada configurations) and users with at least one click (config/data/mind.yaml:15-17).
    With several ranks: every rank encodes a contiguous slice of the items, one all-gather replicates the item cache;
    users are partitioned by `user_id % world` — rank r fills only its rows of the user cache, the rows `evaluate()` makes it score."""
    import ctypes
    from collections import OrderedDict
    import numpy as np
    from ._lib import call
    from .packing import Packed
    cacher, dev = model.cacher, Env.device
    if not (model._packed_items() and getattr(model.user_op, 'supports_packed', False) and cacher.use_item_content):
        raise ValueError('build_caches_device needs packed item and user operators over item content')
    rank, world = sharding._world(group)
    cacher.clean()
    was_training = model.training
    model.eval()                      # caches never carry dropout noise, whatever phase the caller is in (base_lego.py:417)
    n_items, n_users = len(batcher.item_len), len(batcher.hist)
    item_repr = model.item_op.get_full_placeholder(n_items).to(dev)
    tp = (ctypes.c_void_p * len(batcher.cols))(*[t.data_ptr() for t in batcher.tables])
    with torch.no_grad():
        i0, i1 = sharding.item_slice(n_items, rank, world)
        for s in range(i0, i1, item_page):
            e = min(s + item_page, i1)
            lens = batcher.item_len[s:e]
            cu = np.zeros(e - s + 1, dtype=np.int32)
            np.cumsum(lens, out=cu[1:])
            items = torch.arange(s, e, dtype=torch.int64, device=dev)
            cu_d = torch.from_numpy(cu).to(dev)
            outs = [torch.empty(int(cu[-1]), dtype=torch.int64, device=dev) for _ in batcher.cols]
            op = (ctypes.c_void_p * len(batcher.cols))(*[t.data_ptr() for t in outs])
            call('lk_pack_item_tokens', ctypes.addressof(tp), ctypes.addressof(op), len(batcher.cols), items.data_ptr(), cu_d.data_ptr(),
                 e - s, batcher.S)
            pk = Packed(OrderedDict(zip(batcher.cols, outs)), cu_d, e - s, int(cu[-1]), int(lens.max()))
            emb = model.item_op.inputer.get_embeddings({'input_ids': pk.ids}, training=False)
            item_repr[s:e] = model.item_op(emb, cu=pk.cu, max_len=pk.max_len)
        if world > 1:
            item_repr = sharding.gather_item_cache(item_repr[i0:i1].contiguous(), n_items, group)
        cacher.item.repr = item_repr
        cacher.item._set_cached(True)
        user_repr = torch.zeros_like(cacher.user.placeholder, device=dev)
        mine = np.arange(rank, n_users, world)                    # the users this rank owns (all of them on one GPU)
        for s in range(0, len(mine), user_page):
            page = mine[s:s + user_page]
            hists = [batcher.hist[int(u)] for u in page]
            hl = np.fromiter((len(h) for h in hists), dtype=np.int64, count=len(page))
            if hl.min() < 1:
                raise ValueError('build_caches_device: a user without clicks (the reference filters those out)')
            cu = np.zeros(len(page) + 1, dtype=np.int32)
            np.cumsum(hl, out=cu[1:])
            ids = torch.from_numpy(np.concatenate(hists)).to(dev)
            rows = ops.index_rows(item_repr, ids)
            rep = model.user_op(rows, cu=torch.from_numpy(cu).to(dev), max_len=int(hl.max()))
            if world == 1:
                user_repr[int(page[0]):int(page[0]) + len(page)] = rep
            else:
                user_repr[torch.from_numpy(page).to(dev)] = rep
        cacher.user.repr = user_repr
        cacher.user._set_cached(True)
    model.train(was_training)


def cached_scores(model, user_ids: torch.Tensor, item_ids: torch.Tensor, chunk_rows: int = 1 << 24) -> torch.Tensor:
    """score[r] = <user.repr[user_ids[r]], item.repr[item_ids[r]]> for every row (device tensor, fp32)."""
    cacher = model.cacher
    if not (cacher.user.cached and cacher.item.cached):
        raise RuntimeError('cached_scores needs both representation caches (call build_caches first)')
    dev = Env.device
    uid = user_ids.reshape(-1).to(dev, non_blocking=True)
    iid = item_ids.reshape(-1).to(dev, non_blocking=True)
    out = torch.empty(uid.numel(), dtype=torch.float32, device=dev)
    for s in range(0, uid.numel(), chunk_rows):
        e = min(s + chunk_rows, uid.numel())
        ops.cached_scores(cacher.user.repr, cacher.item.repr, uid[s:e], iid[s:e], out=out[s:e])
    return out


def evaluate(model, user_ids, item_ids, labels, groups=None, metrics: Sequence[str] = DEFAULT_METRICS, group=None, shard: bool = True,
             presharded: bool = False):
    """-> (OrderedDict metric -> float, scores of the rows this rank owns, row indices owned).
    `groups` defaults to `user_ids` (config/data/mind.yaml:24).  With several ranks each scores the rows whose group key it
    owns; metric means are combined with one all-reduce, so every rank returns the global values.
    presharded=True: the rows passed in are already this rank's share (`sharding.owned_rows` applied once to the evaluation set, SURVEY §8e
    "rows pre-partitioned by group"); only the metric reduction is collective then."""
    rank, world = sharding._world(group) if shard else (0, 1)      # shard=False: this rank evaluates every row itself (needs full caches)
    same_groups = groups is None or groups is user_ids          # config/data/mind.yaml:24: the group key IS the user id -> one copy, not two
    groups = user_ids if groups is None else groups
    user_ids, item_ids, labels, groups = (torch.as_tensor(t).reshape(-1) for t in (user_ids, item_ids, labels, groups))
    rows: Optional[torch.Tensor] = None
    if world > 1 and not presharded:
        # the user cache is partitioned by user id (build_caches: rank r holds the rows of users r, r+world, ...), so rows MUST be
        # partitioned by user id as well; a group key other than the user id could straddle ranks or hit un-encoded cache rows
        if groups is not user_ids and not torch.equal(groups, user_ids):
            raise ValueError('sharded evaluation groups by user id (config/data/mind.yaml:24); pass groups=None or groups == user_ids')
        rows = sharding.owned_rows(user_ids, rank, world)
        user_ids, item_ids, labels, groups = user_ids[rows], item_ids[rows], labels[rows], groups[rows]
    pool = MetricPool.parse(metrics)
    if user_ids.numel() == 0:        # a rank that owns no rows still takes part in the all-reduce below (0 sums, 0 groups)
        if world == 1:
            raise ValueError('no rows to evaluate')
        scores = torch.empty(0, dtype=torch.float32, device=Env.device)
        vals, pool.n_groups = OrderedDict((k, 0.0) for k in metrics), 0
    else:
        # every id array crosses PCIe once, in the integer width the host holds (int32 halves the bytes), and is widened on the device
        def up(t):
            return t.to(Env.device, non_blocking=True).to(torch.int64)
        uid, iid, lab = up(user_ids), up(item_ids), up(labels)
        scores = cached_scores(model, uid, iid)
        vals = pool.calculate(scores, lab, uid if same_groups else up(groups))
    if world > 1:
        local = torch.tensor(list(vals.values()), dtype=torch.float64, device=Env.device)
        means, _ = sharding.reduce_group_means(local, pool.n_groups, group)
        vals = OrderedDict((k, float(v)) for k, v in zip(vals, means.tolist()))
    return vals, scores, rows
