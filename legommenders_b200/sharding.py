"""Multi-GPU partitioning of the hot path — one process per GPU, torch.distributed for the plumbing.

Only the places where the path really shards are here:

* `ShardedTable`     — config 4: a large frozen word table row-sharded by `id % world`; lookup = dedup -> bucket by owner ->
                       all-to-all(ids) -> local row gather (CUDA kernel) -> all-to-all(rows) -> inverse index.
* `owned_rows`       — config 3: evaluation rows partitioned by group key so that a group never straddles ranks.
* `gather_item_cache`— config 3: every rank encodes a contiguous slice of the items, one all-gather replicates the cache.
* `reduce_group_means` — per-rank (metric sum, group count) -> global mean with one all-reduce.

Training data-parallelism (one flat-bucket all-reduce per step) lives in trainer.FlatAdam.

The integer bookkeeping (bucket plans, partitions) is plain tensor index work and runs wherever the ids live; it is covered on
CPU with gloo (tests/test_multirank_cpu.py).  The row gather itself is `ops.index_rows` (lk_index_rows) — there is no CPU
implementation of it in the product; tests inject a checker through `gather_fn`.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def _world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


# ----------------------------------------------------------------------------------------------------------------
# row-sharded table (config 4)
# ----------------------------------------------------------------------------------------------------------------
def shard_rows(table: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Rows owned by `rank` under the `id % world` rule, in local order (local index = id // world)."""
    return table[rank::world].contiguous()


class LookupPlan:
    """Integer plan of one sharded lookup (all tensors on the ids' device).

    uniq        sorted unique valid ids of this rank's request
    inverse     for every requested position: index into `uniq`, or -1 for an invalid (masked / unset) position
    order       permutation that groups `uniq` by owner rank (stable, so ids stay ascending inside a bucket)
    send_counts how many ids go to each owner
    """
    __slots__ = ('uniq', 'inverse', 'order', 'send_counts', 'shape')

    def __init__(self, uniq, inverse, order, send_counts, shape):
        self.uniq, self.inverse, self.order, self.send_counts, self.shape = uniq, inverse, order, send_counts, shape


def plan_lookup(ids: torch.Tensor, world: int) -> LookupPlan:
    flat = ids.reshape(-1)
    valid = flat > -1
    uniq, inv = torch.unique(flat[valid], sorted=True, return_inverse=True)
    inverse = torch.full_like(flat, -1)
    inverse[valid] = inv
    owner = uniq % world
    order = torch.sort(owner, stable=True).indices
    send_counts = torch.bincount(owner, minlength=world)
    return LookupPlan(uniq, inverse, order, send_counts, tuple(ids.shape))


class ShardedTable:
    """Frozen `[V, E]` fp32 table, row `i` stored on rank `i % world` at local index `i // world`."""

    def __init__(self, local_rows: torch.Tensor, num_rows: int, group=None,
                 gather_fn: Optional[Callable[[torch.Tensor, torch.Tensor], torch.Tensor]] = None):
        self.local = local_rows
        self.num_rows = num_rows
        self.group = group
        self.rank, self.world = _world(group)
        if gather_fn is None:
            from . import ops            # CUDA kernel (lk_index_rows); raises on CPU tensors — no fallback
            gather_fn = ops.index_rows
        self.gather = gather_fn

    def exchange(self, plan: LookupPlan) -> torch.Tensor:
        """-> rows of `plan.uniq` ([U, E], same order as uniq)."""
        W = self.world
        send_ids = plan.uniq[plan.order]
        if W == 1:
            rows = self.gather(self.local, send_ids)
        else:
            send_counts = plan.send_counts
            recv_counts = torch.empty_like(send_counts)
            dist.all_to_all_single(recv_counts, send_counts, group=self.group)
            sc, rc = send_counts.tolist(), recv_counts.tolist()       # split sizes must be host ints
            recv_ids = torch.empty(sum(rc), dtype=send_ids.dtype, device=send_ids.device)
            dist.all_to_all_single(recv_ids, send_ids, rc, sc, group=self.group)
            mine = self.gather(self.local, torch.div(recv_ids, W, rounding_mode='floor'))
            rows = torch.empty((sum(sc), self.local.shape[1]), dtype=self.local.dtype, device=self.local.device)
            dist.all_to_all_single(rows, mine.contiguous(), sc, rc, group=self.group)
        out = torch.empty_like(rows)
        out[plan.order] = rows            # undo the owner grouping -> uniq order
        return out

    def lookup_unique(self, ids: torch.Tensor):
        """-> (rows [U, E] of the unique valid ids, inverse [*ids.shape] with -1 at invalid positions).
        The NRMS step gathers straight from this compact table with `inverse` as the id tensor, so token rows cross
        NVLink once per distinct token, not once per occurrence."""
        plan = plan_lookup(ids, self.world)
        return self.exchange(plan), plan.inverse.reshape(plan.shape)

    def lookup(self, ids: torch.Tensor) -> torch.Tensor:
        """Dense result [*ids.shape, E]; invalid positions are zero rows (concat_inputer.py:108-112 semantics)."""
        rows, inverse = self.lookup_unique(ids)
        flat = inverse.reshape(-1)
        padded = torch.cat([rows, torch.zeros((1, rows.shape[1]), dtype=rows.dtype, device=rows.device)])
        idx = torch.where(flat < 0, torch.full_like(flat, rows.shape[0]), flat)
        return self.gather(padded, idx).reshape(*ids.shape, rows.shape[1])


# ----------------------------------------------------------------------------------------------------------------
# cached evaluation (config 3)
# ----------------------------------------------------------------------------------------------------------------
def owned_rows(group_keys: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Indices (ascending) of the evaluation rows this rank scores: all rows of a group go to `key % world`."""
    return torch.nonzero(group_keys % world == rank, as_tuple=False).reshape(-1)


def item_slice(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous item range encoded by `rank` (equal slices, remainder spread over the first ranks)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_item_cache(local_repr: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """All-gather the per-rank slices `[stop-start, D]` into the replicated `[n_items, D]` cache."""
    rank, world = _world(group)
    if world == 1:
        return local_repr
    D = local_repr.shape[1]
    sizes = [b - a for a, b in (item_slice(n_items, r, world) for r in range(world))]
    cap = max(sizes)                       # equal-sized contributions (slices differ by at most one row): pad, gather, trim
    mine = torch.zeros((cap, D), dtype=local_repr.dtype, device=local_repr.device)
    mine[:local_repr.shape[0]] = local_repr
    parts = [torch.empty((cap, D), dtype=local_repr.dtype, device=local_repr.device) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    return torch.cat([parts[r][:n] for r, n in enumerate(sizes)], dim=0)


def reduce_group_means(local_means: torch.Tensor, local_groups: int, group=None) -> Tuple[torch.Tensor, int]:
    """Per-rank means over `local_groups` groups -> global means over all groups (one all-reduce of sums + count)."""
    rank, world = _world(group)
    buf = torch.cat([local_means.to(torch.float64) * local_groups,
                     torch.tensor([float(local_groups)], dtype=torch.float64, device=local_means.device)])
    if world > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    total = int(round(buf[-1].item()))
    return buf[:-1] / max(total, 1), total
