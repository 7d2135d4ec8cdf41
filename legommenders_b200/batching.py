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

    def train_batch(self, rows: np.ndarray) -> dict:
        """rows: indices into the world's training impressions."""
        users, pos = self.world.train_users[rows], self.world.train_pos[rows]
        cand = torch.from_numpy(self.sample_candidates(users, pos))
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
