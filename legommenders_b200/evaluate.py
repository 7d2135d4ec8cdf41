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


def evaluate(model, user_ids, item_ids, labels, groups=None, metrics: Sequence[str] = DEFAULT_METRICS, group=None):
    """-> (OrderedDict metric -> float, scores of the rows this rank owns, row indices owned).
    `groups` defaults to `user_ids` (config/data/mind.yaml:24).  With several ranks each scores the rows whose group key it
    owns; metric means are combined with one all-reduce, so every rank returns the global values."""
    rank, world = sharding._world(group)
    groups = user_ids if groups is None else groups
    user_ids, item_ids, labels, groups = (torch.as_tensor(t).reshape(-1) for t in (user_ids, item_ids, labels, groups))
    rows: Optional[torch.Tensor] = None
    if world > 1:
        rows = sharding.owned_rows(groups, rank, world)
        user_ids, item_ids, labels, groups = user_ids[rows], item_ids[rows], labels[rows], groups[rows]
    scores = cached_scores(model, user_ids, item_ids)
    pool = MetricPool.parse(metrics)
    vals = pool.calculate(scores, labels, groups)
    if world > 1:
        local = torch.tensor(list(vals.values()), dtype=torch.float64, device=Env.device)
        means, _ = sharding.reduce_group_means(local, pool.n_groups, group)
        vals = OrderedDict((k, float(v)) for k, v in zip(vals, means.tolist()))
    return vals, scores, rows
