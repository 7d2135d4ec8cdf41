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
                 gather_fn: Optional[Callable[[torch.Tensor, torch.Tensor], torch.Tensor]] = None, cap_factor: float = 1.5):
        self._device_gather = gather_fn is None
        self.cap_factor, self._cap, self._buf_key, self._buf = cap_factor, None, None, None
        self._overflow = torch.zeros(1, dtype=torch.int32, device=local_rows.device) if local_rows.is_cuda else None
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
        """-> (rows [U, E] of the distinct valid ids, inverse [*ids.shape] with -1 at invalid positions).
        The NRMS step gathers straight from this compact table with `inverse` as the id tensor, so token rows cross
        NVLink once per distinct token, not once per occurrence.
        CUDA ids take the device plan (`lookup_unique_device`: no sort, no split sizes, no host round trip); host ids (the gloo tests of the
        integer bookkeeping) take the index-arithmetic plan above."""
        if ids.is_cuda and self._device_gather:
            return self.lookup_unique_device(ids)
        plan = plan_lookup(ids, self.world)
        return self.exchange(plan), plan.inverse.reshape(plan.shape)

    # ---- device plan (csrc/lk_shard.cu) ------------------------------------------------------------------------------------------------
    def _buffers(self, P: int, cap: int):
        from ._lib import query
        slots = query('lk_shard_hash_slots', P)
        key = (slots, cap)
        if self._buf_key != key:
            dev, W, E = self.local.device, self.world, self.local.shape[1]
            neg = torch.empty(slots + W * cap, dtype=torch.int64, device=dev)           # [hash keys | send buckets]: one fill(-1)
            zero = torch.zeros(W + 1, dtype=torch.int32, device=dev)                    # [bucket counts | overflow]
            self._buf = dict(slots=slots, neg=neg, keys=neg[:slots], send=neg[slots:], vals=torch.empty(slots, dtype=torch.int32, device=dev),
                             zero=zero, recv=torch.empty(W * cap, dtype=torch.int64, device=dev),
                             rows_out=torch.empty((W * cap, E), dtype=torch.float32, device=dev))
            self._buf_key = key
        return self._buf

    def _plan(self, flat: torch.Tensor, cap: int):
        from ._lib import call
        W, P = self.world, flat.numel()
        b = self._buffers(P, cap)
        b['neg'].fill_(-1)
        b['zero'].zero_()
        call('lk_shard_plan', flat.data_ptr(), P, W, cap, b['keys'].data_ptr(), b['vals'].data_ptr(), b['slots'], b['zero'].data_ptr(),
             b['send'].data_ptr(), b['zero'][W:].data_ptr())
        return b

    def lookup_unique_device(self, ids: torch.Tensor):
        """Fixed-capacity all-to-all lookup: dedup + owner bucketing in one kernel (hash + atomics), equal-split NCCL all-to-all of the id
        buckets, owner-side gather, equal-split all-to-all of the row buckets.  -> (rows [W*cap, E], inverse).  The bucket capacity (the same
        on every rank) is measured once, on the first call: largest bucket over all ranks x `cap_factor`.  A later overflow is counted on
        the device and raised by `check()`, which the caller runs wherever it synchronises anyway."""
        from ._lib import call, id_violation_counter
        W = self.world
        flat = ids.reshape(-1).contiguous()
        P = flat.numel()
        if self._cap is None:                                      # first call only (synchronises): size the buckets
            b = self._plan(flat, max(1024, P))                     # a bucket can never hold more than P ids
            t = b['zero'][:W].max().reshape(1).to(torch.int64)
            if W > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            self._cap = max(1024, int(int(t.item()) * self.cap_factor) + 64)
        cap = self._cap
        b = self._plan(flat, cap)
        inverse = torch.empty(P, dtype=torch.int64, device=flat.device)
        call('lk_shard_inverse', flat.data_ptr(), P, b['keys'].data_ptr(), b['vals'].data_ptr(), b['slots'], inverse.data_ptr())
        self._overflow += b['zero'][W:]                            # device-side running count, read by check()
        if W == 1:
            recv = b['send']
        else:
            recv = b['recv']
            dist.all_to_all_single(recv, b['send'], group=self.group)
        id_violation_counter(self.local.device)
        rows_out = b['rows_out'] if W > 1 else torch.empty_like(b['rows_out'])
        call('lk_shard_gather', recv.data_ptr(), W * cap, W, self.local.data_ptr(), self.local.shape[0], self.local.shape[1], rows_out.data_ptr())
        if W == 1:
            rows = rows_out
        else:
            rows = torch.empty_like(rows_out)
            dist.all_to_all_single(rows, rows_out, group=self.group)
        return rows, inverse.reshape(ids.shape)

    def check(self):
        """Raise if any lookup since the last check overflowed a bucket (device->host read: call it where the host synchronises anyway)."""
        if self._overflow is None:
            return
        n = int(self._overflow.item())
        if n:
            self._overflow.zero_()
            raise RuntimeError(f'ShardedTable: {n} distinct ids did not fit the send buckets (capacity {self._cap}); construct with a larger cap_factor')

    def lookup(self, ids: torch.Tensor) -> torch.Tensor:
        """Dense result [*ids.shape, E]; invalid positions are zero rows (concat_inputer.py:108-112 semantics)."""
        rows, inverse = self.lookup_unique(ids)
        flat = inverse.reshape(-1)
        padded = torch.cat([rows, torch.zeros((1, rows.shape[1]), dtype=rows.dtype, device=rows.device)])
        idx = torch.where(flat < 0, torch.full_like(flat, rows.shape[0]), flat)
        return self.gather(padded, idx).reshape(*ids.shape, rows.shape[1])


# ----------------------------------------------------------------------------------------------------------------
# LLM-embedding catalog sweep (config 5)
# ----------------------------------------------------------------------------------------------------------------
def catalog_topk(user_planes, item_planes, k: int, item_offset: int = 0, group=None):
    """Per-user top-k of the catalog with the item table ROW-SHARDED over the ranks (contiguous blocks: this rank holds the projected rows
    [item_offset, item_offset + N_local) as split-bf16 planes; users are replicated).  The hot loop is local — `ops.sweep_topk`, scores never
    materialised — and the only collective is an all-gather of the [U, k] candidates followed by a k-way merge.
    -> (scores [U, k], global item ids [U, k]) identical on every rank."""
    from . import ops
    vals, idx = ops.sweep_topk(user_planes, item_planes, k, item_offset)
    rank, world = _world(group)
    if world > 1:
        gv = [torch.empty_like(vals) for _ in range(world)]
        gi = [torch.empty_like(idx) for _ in range(world)]
        dist.all_gather(gv, vals.contiguous(), group=group)
        dist.all_gather(gi, idx.contiguous(), group=group)
        vals, idx = ops.merge_topk(torch.cat(gv, dim=1), torch.cat(gi, dim=1), k)
    return vals, idx


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
