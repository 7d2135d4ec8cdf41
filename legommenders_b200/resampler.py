"""Per-sample host pre-processing (mirror of loader/resampler.py:48-274 and loader/data_set.py:33-96).

Semantics that must be bit-exact: candidate order [pos, sampled true negatives…, uniform random ids…]
(label 0), history right-padded with item id 0 to `max_click_num`, `__clicks_mask__ = 1^len 0^pad`, content tensors
injected from a per-item cache unless Env.item_cache / Env.lm_cache say the model will index caches itself.
Python's `random` module is used exactly like the reference so that a shared seed gives identical draws.
"""
from __future__ import annotations

import copy
import random
from typing import Any, Callable, Dict, List, Optional

import torch
from torch.utils.data import Dataset as BaseDataset

from .cacher import stack_trees
from .env import Env


class DataSet(BaseDataset):
    """loader/data_set.py:33-96 — shallow-copies every column, then applies the resampler."""

    def __init__(self, ut, resampler: Optional[Callable[[Dict[str, Any]], Dict[str, Any]]] = None):
        self.ut = ut
        self.resampler = resampler

    def __getitem__(self, index):
        raw = self.ut[index]
        sample = {col: copy.copy(raw[col]) for col in raw}
        return self.resampler(sample) if self.resampler else sample

    def __len__(self):
        return len(self.ut)

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]


class Resampler:
    def __init__(self, lego_config):
        self.lego_config = lego_config
        self.use_item_content = lego_config.use_item_content
        cm = self.cm = lego_config.cm
        self.history_col, self.item_col, self.user_col = cm.history_col, cm.item_col, cm.user_col
        self.neg_col, self.mask_col = cm.neg_col, cm.mask_col

        self.item_dataset = self.item_inputer = self.item_cache = None
        if self.use_item_content:
            self.item_dataset = DataSet(ut=lego_config.item_ut)
            self.item_inputer = lego_config.item_operator.inputer
            self.item_cache = self._build_item_cache()

        self.user_cache: Dict[int, Any] = {}
        self.user_inputer = lego_config.user_operator.inputer
        self.max_click_num = self.user_inputer.ut.meta.features[self.history_col].max_len
        self.use_neg_sampling = lego_config.use_neg_sampling
        self.item_size = lego_config.item_ut.meta.features[self.item_col].tokenizer.vocab.size

    def _build_item_cache(self) -> List[dict]:
        return [self.item_inputer(sample) for sample in self.item_dataset]

    @staticmethod
    def pack_tensor(array):
        return torch.tensor(array, dtype=torch.long)

    def rebuild_candidates(self, sample: dict):
        if not isinstance(sample[self.item_col], list):
            sample[self.item_col] = [sample[self.item_col]]
        if self.use_neg_sampling and (Env.is_training or (Env.is_evaluating and Env.simple_dev)):
            true_negs = sample[self.neg_col] if self.neg_col else []
            k = self.lego_config.neg_count
            rand_count = max(k - len(true_negs), 0)
            negs = random.sample(true_negs, k=min(k, len(true_negs)))
            negs += [random.randint(0, self.item_size - 1) for _ in range(rand_count)]
            sample[self.item_col].extend(negs)
        if self.neg_col:
            sample.pop(self.neg_col, None)
        if not self.use_item_content or Env.lm_cache or Env.item_cache:
            sample[self.item_col] = self.pack_tensor(sample[self.item_col])
            return
        sample[self.item_col] = stack_trees([self.item_cache[i] for i in sample[self.item_col]])

    def rebuild_clicks(self, sample: dict):
        if Env.user_cache:
            sample.pop(self.history_col, None)
            return
        n = len(sample[self.history_col])
        sample[self.mask_col] = torch.tensor([1] * n + [0] * (self.max_click_num - n), dtype=torch.long)
        if self.use_item_content:
            sample[self.history_col].extend([0] * (self.max_click_num - n))
        if not self.use_item_content:
            sample[self.history_col] = self.user_inputer(sample)
            return
        if self.lego_config.user_operator_class.flatten_mode:
            sample[self.history_col] = self.user_inputer(sample)
            sample[self.mask_col] = self.user_inputer.get_mask(sample[self.history_col])
            return
        if Env.lm_cache or Env.item_cache:
            sample[self.history_col] = self.pack_tensor(sample[self.history_col])
            return
        uid = sample[self.user_col]
        if uid not in self.user_cache:
            self.user_cache[uid] = stack_trees([self.item_cache[i] for i in sample[self.history_col]])
        sample[self.history_col] = self.user_cache[uid]

    def rebuild(self, sample: dict) -> dict:
        self.rebuild_candidates(sample)
        self.rebuild_clicks(sample)
        return sample

    __call__ = rebuild
