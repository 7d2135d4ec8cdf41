"""Vectorised host-side batch builder producing the reference's wire format.

`Resampler` (resampler.py) is the sample-at-a-time mirror of loader/resampler.py; it costs 10-20 ms of Python per
64-sample batch.  `BatchBuilder` produces the same nested int64 batch (`item_id.input_ids.<col> [B,C,S]`,
`history.input_ids.<col> [B,H,S]`, attention masks, `__clicks_mask__ [B,H]`) by fancy-indexing per-item token
tables `[N_items, S]` built once from the inputer layouts, so the host path no longer bounds the step.
Semantics kept (resampler.py:139-259 of the reference): candidate order [pos, negatives…], history right-padded with
item id 0, `__clicks_mask__ = 1^len 0^pad`.  Negative draws use a numpy generator instead of Python's `random`
(same distribution, different stream) — parity tests use `Resampler`, throughput runs use this.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

from .cacher import stack_trees


class BatchBuilder:
    def __init__(self, resampler, world, neg_count: int = 4, seed: int = 0, pin: bool = True):
        self.world = world
        self.neg_count = neg_count
        self.rng = np.random.default_rng(seed)
        self.pin = pin and torch.cuda.is_available()
        self.H = world.hist_len
        self.item_tokens = stack_trees(resampler.item_cache)       # nested dict of [N_items, S] int64 tensors
        hist = np.zeros((world.n_users, self.H), dtype=np.int64)
        hmask = np.zeros((world.n_users, self.H), dtype=np.int64)
        for u, h in enumerate(world.histories):
            hist[u, :len(h)] = h
            hmask[u, :len(h)] = 1
        self.hist, self.hmask = torch.from_numpy(hist), torch.from_numpy(hmask)
        self.neg_lists = world.negs

    def _index(self, tree, ids: torch.Tensor):
        if isinstance(tree, dict):
            return type(tree)((k, self._index(v, ids)) for k, v in tree.items())
        out = tree[ids.reshape(-1)].reshape(*ids.shape, tree.shape[-1])
        return out.pin_memory() if self.pin else out

    def sample_candidates(self, users: np.ndarray, pos: np.ndarray) -> np.ndarray:
        B, K = len(users), self.neg_count
        cand = np.empty((B, 1 + K), dtype=np.int64)
        cand[:, 0] = pos
        for b, u in enumerate(users):
            negs = self.neg_lists[int(u)]
            k = min(K, len(negs))
            if k:
                cand[b, 1:1 + k] = self.rng.choice(negs, size=k, replace=False)
            if k < K:
                cand[b, 1 + k:] = self.rng.integers(0, self.world.n_items, size=K - k)
        return cand

    def train_batch(self, rows: np.ndarray, cand: np.ndarray = None) -> dict:
        """rows: indices into the world's training impressions; cand [B, 1+K]: explicit candidate ids (default: sampled here)."""
        users, pos = self.world.train_users[rows], self.world.train_pos[rows]
        cand = torch.from_numpy(self.sample_candidates(users, pos) if cand is None else np.asarray(cand, dtype=np.int64))
        ut = torch.from_numpy(users)
        pin = (lambda t: t.pin_memory()) if self.pin else (lambda t: t)
        batch = OrderedDict()
        batch['index'] = pin(torch.from_numpy(np.asarray(rows, dtype=np.int64)))
        batch['user_id'] = pin(ut)
        batch['item_id'] = self._index(self.item_tokens, cand)
        batch['click'] = pin(torch.ones(len(rows), dtype=torch.int64))
        batch['history'] = self._index(self.item_tokens, self.hist[ut])
        batch['__clicks_mask__'] = pin(self.hmask[ut])
        return batch


def tree_to_device(tree, device, non_blocking=True):
    if isinstance(tree, dict):
        return type(tree)((k, tree_to_device(v, device, non_blocking)) for k, v in tree.items())
    return tree.to(device, non_blocking=non_blocking)


def tree_bytes(tree) -> int:
    if isinstance(tree, dict):
        return sum(tree_bytes(v) for v in tree.values())
    return tree.numel() * tree.element_size()


class DeviceBatcher:
    """Id-only training batches.  The per-item token layouts (`Resampler.item_cache`, built once by the inputer,
    loader/resampler.py:113-126) live on the DEVICE as `[N_items, S]` tables; a training batch crosses PCIe as the item-id list of its
    candidates + valid history items and two int32 offset vectors (~50 KB instead of the 3.7 MB of `[B, 55, S]` int64 trees), and
    `lk_pack_item_tokens` expands it straight into the packed rows the native step consumes.  The packed ids are bit-identical to
    `packing.pack_tokens` applied to the Resampler / BatchBuilder batch of the same candidates (tests/test_gpu_model.py).
    Layout of the item list: the B*C candidates first ([pos, negatives...] per impression), then the valid history items user by
    user (history right-padding with item 0 is never materialised)."""

    def __init__(self, resampler, world, device, neg_count: int = 4, seed: int = 0):
        from ._lib import load
        self.world, self.device, self.neg_count = world, device, neg_count
        self.sampler = BatchBuilder.__new__(BatchBuilder)          # reuse the candidate sampler only
        self.sampler.world, self.sampler.neg_count, self.sampler.neg_lists = world, neg_count, world.negs
        self.sampler.rng = np.random.default_rng(seed)
        trees = stack_trees(resampler.item_cache)
        self.cols = list(trees['input_ids'].keys())
        mask = trees['attention_mask']
        if isinstance(mask, dict):
            raise ValueError('DeviceBatcher needs a single-sequence inputer (ConcatInputer)')
        self.S = mask.shape[1]
        self.item_len = mask.sum(dim=1).numpy().astype(np.int64)                    # host: offsets are host arithmetic
        self.tables = [trees['input_ids'][c].to(device).contiguous() for c in self.cols]
        self.hist = [np.asarray(h, dtype=np.int64) for h in world.histories]
        self._lib = load()

    def host_batch(self, rows: np.ndarray) -> dict:
        """Pinned host buffers of one batch: item list, item offsets, user offsets (all that is copied per step)."""
        users, pos = self.world.train_users[rows], self.world.train_pos[rows]
        cand = self.sampler.sample_candidates(users, pos)
        hists = [self.hist[int(u)] for u in users]
        items = np.concatenate([cand.reshape(-1)] + hists)
        lens = self.item_len[items]
        cu = np.zeros(len(items) + 1, dtype=np.int32)
        np.cumsum(lens, out=cu[1:])
        cu_u = np.zeros(len(users) + 1, dtype=np.int32)
        np.cumsum([len(h) for h in hists], out=cu_u[1:])
        pin = (lambda t: t.pin_memory()) if torch.cuda.is_available() else (lambda t: t)
        return dict(items=pin(torch.from_numpy(items)), cu=pin(torch.from_numpy(cu)), cu_users=pin(torch.from_numpy(cu_u)),
                    n=len(items), rows=int(cu[-1]), max_len=int(lens.max()), max_hist=int(max(len(h) for h in hists)),
                    B=len(users), C=cand.shape[1], user_id=torch.from_numpy(users))

    def to_device(self, hb: dict) -> dict:
        """H2D of the id lists + one kernel -> a batch the native step accepts (carries its packed rows)."""
        import ctypes
        from ._lib import call
        from .packing import Packed
        dev = self.device
        items = hb['items'].to(dev, non_blocking=True)
        cu = hb['cu'].to(dev, non_blocking=True)
        cu_u = hb['cu_users'].to(dev, non_blocking=True)
        outs = [torch.empty(hb['rows'], dtype=torch.int64, device=dev) for _ in self.cols]
        tp = (ctypes.c_void_p * len(self.cols))(*[t.data_ptr() for t in self.tables])
        op = (ctypes.c_void_p * len(self.cols))(*[t.data_ptr() for t in outs])
        call('lk_pack_item_tokens', ctypes.addressof(tp), ctypes.addressof(op), len(self.cols), items.data_ptr(), cu.data_ptr(), hb['n'], self.S)
        pk = Packed(OrderedDict(zip(self.cols, outs)), cu, hb['n'], hb['rows'], hb['max_len'])
        return {'__lk_packed__': (pk, cu_u, hb['max_hist'], hb['B'], hb['C']), '__keep__': (items,)}

    @staticmethod
    def h2d_bytes(hb: dict) -> int:
        return sum(hb[k].numel() * hb[k].element_size() for k in ('items', 'cu', 'cu_users'))


class DeviceResampler:
    """The Resampler on the device, one step ahead (loader/resampler.py:139-259).

    Everything the reference's per-sample Python does for a training batch — negative sampling from the user's true-negative list plus
    uniform ids, candidate order [pos, negs...], the user's click history, offsets — is ONE kernel (`lk_resample_batch`) over device-resident
    tables; the only per-step host->device traffic is the B impression indices (8 B each).  The step needs four integers on the host to size
    its launches (token rows T, items n, longest item, longest history); they come back through pinned memory while the PREVIOUS step is still
    running: `submit(rows)` enqueues on a side stream, `take()` returns the oldest submitted batch (carrying its packed rows, as
    DeviceBatcher.to_device does).  Draws are Philox4x32-10 keyed by (seed, impression row): a batch is a pure function of (rows, seed).
    """

    def __init__(self, resampler, world, device, neg_count: int = 4, seed: int = 0, max_batch: int = 512, depth: int = 3):
        import ctypes
        from ._lib import load
        self.world, self.device, self.K, self.seed = world, device, neg_count, int(seed)
        trees = stack_trees(resampler.item_cache)
        self.cols = list(trees['input_ids'].keys())
        mask = trees['attention_mask']
        if isinstance(mask, dict):
            raise ValueError('DeviceResampler needs a single-sequence inputer (ConcatInputer)')
        self.S = mask.shape[1]
        dev = device
        self.tables = [trees['input_ids'][c].to(dev).contiguous() for c in self.cols]
        self.item_len = mask.sum(dim=1).to(torch.int32).to(dev)
        self.n_items = int(self.item_len.numel())

        def csr(lists):
            off = np.zeros(len(lists) + 1, dtype=np.int64)
            np.cumsum([len(x) for x in lists], out=off[1:])
            flat = np.concatenate([np.asarray(x, dtype=np.int64) for x in lists]) if off[-1] else np.zeros(0, dtype=np.int64)
            return torch.from_numpy(off).to(dev), torch.from_numpy(np.ascontiguousarray(flat)).to(dev)

        self.neg_off, self.neg_items = csr(world.negs)
        self.hist_off, self.hist_items = csr(world.histories)
        self.imp_user = torch.from_numpy(world.train_users).to(dev)
        self.imp_pos = torch.from_numpy(world.train_pos).to(dev)
        self.n_users, self.n_imps = len(world.histories), len(world.train_users)
        self.max_batch = max_batch
        cap = max_batch * (neg_count + 1 + world.hist_len)
        self.cap = cap
        self.stream = torch.cuda.Stream(device=dev)
        self.slots = []
        for _ in range(depth):
            self.slots.append(dict(rows_host=torch.empty(max_batch, dtype=torch.int64).pin_memory(),
                                   rows=torch.empty(max_batch, dtype=torch.int64, device=dev),
                                   items=torch.empty(cap, dtype=torch.int64, device=dev),
                                   cu=torch.empty(cap + 1, dtype=torch.int32, device=dev),
                                   cu_users=torch.empty(max_batch + 1, dtype=torch.int32, device=dev),
                                   user_ids=torch.empty(max_batch, dtype=torch.int64, device=dev),
                                   meta=torch.zeros(4, dtype=torch.int32, device=dev),
                                   meta_host=torch.zeros(4, dtype=torch.int32).pin_memory(),
                                   ready=torch.cuda.Event(), consumed=torch.cuda.Event(), B=0, busy=False))
        self._head = self._tail = 0
        self._last = None
        self._lib = load()
        self._tp = (ctypes.c_void_p * len(self.cols))(*[t.data_ptr() for t in self.tables])

    def submit(self, rows):
        """Enqueue the batch of impression `rows` on the side stream (returns immediately).  rows: numpy / host tensor (copied up through
        pinned memory, 8 B per impression) or an int64 CUDA tensor already resident on the device."""
        from ._lib import call, id_violation_counter
        slot = self.slots[self._head % len(self.slots)]
        if slot['busy']:
            raise RuntimeError('DeviceResampler: more batches in flight than slots (take() the oldest first)')
        B = len(rows)
        if B > self.max_batch:
            raise ValueError(f'batch of {B} impressions exceeds max_batch={self.max_batch}')
        resident = isinstance(rows, torch.Tensor) and rows.is_cuda
        if not resident:
            slot['rows_host'][:B] = torch.as_tensor(np.asarray(rows), dtype=torch.int64)
        id_violation_counter(self.device)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(slot['consumed'])            # the step that used this slot's buffers has been enqueued and will finish first
            if resident:
                src = rows                                       # must be complete already (the caller's responsibility, as for any input)
            else:
                slot['rows'][:B].copy_(slot['rows_host'][:B], non_blocking=True)
                src = slot['rows']
            slot['src'] = src                                    # keep a resident rows tensor alive until the kernel has run
            call('lk_resample_batch', src.data_ptr(), B, self.K, self.seed, self.imp_user.data_ptr(), self.imp_pos.data_ptr(),
                 self.neg_off.data_ptr(), self.neg_items.data_ptr(), self.hist_off.data_ptr(), self.hist_items.data_ptr(),
                 self.item_len.data_ptr(), self.n_items, self.n_imps, self.n_users, slot['items'].data_ptr(), slot['cu'].data_ptr(),
                 slot['cu_users'].data_ptr(), slot['user_ids'].data_ptr(), slot['meta'].data_ptr(), self.cap)
            slot['meta_host'].copy_(slot['meta'], non_blocking=True)
            slot['ready'].record(self.stream)
        slot['B'], slot['busy'] = B, True
        self._head += 1

    def take(self) -> dict:
        """The oldest submitted batch, expanded to packed token rows on the current stream -> a batch NativeNRMSStep accepts."""
        import ctypes
        from ._lib import call
        from .packing import Packed
        if self._tail == self._head:
            raise RuntimeError('DeviceResampler.take() without a submitted batch')
        main = torch.cuda.current_stream()
        if self._last is not None:
            self._last['consumed'].record(main)                 # everything that reads the previous slot's buffers is enqueued by now
        slot = self.slots[self._tail % len(self.slots)]
        self._tail += 1
        slot['ready'].synchronize()                             # normally complete long ago: it was submitted a whole step earlier
        T, n, max_len, max_hist = (int(v) for v in slot['meta_host'].tolist())
        if T < 0:
            raise RuntimeError(f'DeviceResampler: batch of {n} items exceeds the item buffer ({self.cap})')
        main.wait_event(slot['ready'])
        B, C = slot['B'], self.K + 1
        outs = [torch.empty(T, dtype=torch.int64, device=self.device) for _ in self.cols]
        op = (ctypes.c_void_p * len(self.cols))(*[t.data_ptr() for t in outs])
        call('lk_pack_item_tokens', ctypes.addressof(self._tp), ctypes.addressof(op), len(self.cols), slot['items'].data_ptr(),
             slot['cu'].data_ptr(), n, self.S)
        pk = Packed(OrderedDict(zip(self.cols, outs)), slot['cu'][:n + 1], n, T, max_len)
        slot['busy'] = False
        self._last = slot
        return {'__lk_packed__': (pk, slot['cu_users'][:B + 1], max_hist, B, C), '__keep__': (slot['items'],), 'user_id': slot['user_ids'][:B],
                '__items__': slot['items'][:n]}

    def replay_host(self, rows: np.ndarray):
        """The same batch from the host restatement (`lk_resample_reference`): (items, cu_items, cu_users) as numpy arrays."""
        import ctypes
        C = self.K + 1
        cand = np.zeros((len(rows), C), dtype=np.int64)
        buf = (ctypes.c_int64 * C)()
        for b, r in enumerate(rows):
            u = int(self.world.train_users[r])
            negs = np.ascontiguousarray(self.world.negs[u], dtype=np.int64)
            rc = self._lib.lk_resample_reference(self.seed, int(r), int(self.world.train_pos[r]), negs.ctypes.data, len(negs), self.K,
                                                 self.n_items, ctypes.cast(buf, ctypes.c_void_p))
            if rc:
                raise RuntimeError(self._lib.lk_last_error().decode())
            cand[b] = list(buf)
        hists = [np.asarray(self.world.histories[int(self.world.train_users[r])], dtype=np.int64) for r in rows]
        items = np.concatenate([cand.reshape(-1)] + hists)
        lens = self.item_len.cpu().numpy()[items]
        cu = np.zeros(len(items) + 1, dtype=np.int32)
        np.cumsum(lens, out=cu[1:])
        cu_u = np.zeros(len(rows) + 1, dtype=np.int32)
        np.cumsum([len(h) for h in hists], out=cu_u[1:])
        return items, cu, cu_u


def balanced_partition(costs: np.ndarray, world: int) -> np.ndarray:
    """Split a global batch into `world` equal-sized shares of (nearly) equal total cost -> index array [world, n // world].

    Data-parallel steps end with an all-reduce, so every step waits for the rank with the most packed token rows; with impressions drawn
    independently per rank the slowest of 8 is ≈ 8 % slower than the average (histories run from 1 to 50 clicks).  Sorting the GLOBAL batch by
    cost and dealing it out in serpentine order gives every rank the same number of impressions and within a few rows the same number of token
    rows.  Which impressions form the global batch — and therefore the averaged gradient — is unchanged; only their placement is."""
    n = len(costs)
    if n % world:
        raise ValueError(f'global batch of {n} does not divide over {world} ranks')
    order = np.argsort(-np.asarray(costs), kind='stable').reshape(n // world, world)
    order[1::2] = order[1::2, ::-1].copy()                    # serpentine: 0..W-1, W-1..0, ...
    return np.ascontiguousarray(order.T)


def impression_costs(world, item_len: np.ndarray, neg_count: int = 4) -> np.ndarray:
    """Expected packed token rows of every training impression: the tokens of its user's history items + the positive + an average
    item for each sampled negative (host arithmetic, once per data set)."""
    off = np.zeros(len(world.histories) + 1, dtype=np.int64)
    np.cumsum([len(h) for h in world.histories], out=off[1:])
    flat = np.concatenate([np.asarray(h, dtype=np.int64) for h in world.histories])
    per_user = np.add.reduceat(item_len[flat], off[:-1]) if len(flat) else np.zeros(len(world.histories))
    per_user[off[1:] == off[:-1]] = 0
    return per_user[world.train_users] + item_len[world.train_pos] + neg_count * float(item_len.mean())
