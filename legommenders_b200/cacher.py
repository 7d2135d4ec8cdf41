"""Item / user representation caches for fast evaluation (mirror of loader/cacher/{base,item,user,repr}_cacher.py and
loader/pager/{base,fast_item,fast_user}_pager.py).

Contract kept: `cache(contents)` / `clean()` / `.repr` / `.cached`, positional slices (row i = i-th content), page size
`cache_page_size`, triggers that flip Env.item_cache / Env.user_cache (Resampler reads them), outputs detached.
Re-design: instead of one un-batched gather per item (fast_item_pager.py:101-104) a whole page of id rows is stacked
on the host once and runs through the batched gather + encoder kernels under no_grad.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Any, Callable, Optional, Sequence

import torch

from .env import Env


def stack_trees(items: Sequence[Any]):
    """utils/stacker.py:51-150 — stack a list of equally-shaped nested dicts of tensors along a new dim 0."""
    first = items[0]
    if isinstance(first, dict):
        return type(first)((k, stack_trees([it[k] for it in items])) for k in first)
    if isinstance(first, torch.Tensor):
        return torch.stack(list(items))
    return torch.tensor(list(items))


class BaseCacher:
    def __init__(self, operator, page_size: int, hidden_size: int, activate: bool = True,
                 trigger: Optional[Callable[[bool], None]] = None):
        self.operator = operator
        self.page_size = page_size
        self.hidden_size = hidden_size
        self.trigger = trigger or (lambda *_: None)
        self.cached = False
        self._set_cached(False)
        self._activate = activate
        self.repr = None

    def _set_cached(self, cached: bool):
        self.cached = cached
        self.trigger(cached)

    def _cache(self, contents):
        raise NotImplementedError

    def cache(self, contents):
        self.clean()
        if not self._activate:
            return
        self.repr = self._cache(contents)
        from ._lib import raise_on_bad_ids
        raise_on_bad_ids(type(self).__name__)      # building a cache is a synchronisation point: out-of-range ids -> IndexError
        self._set_cached(True)

    def clean(self):
        self.repr = None
        self._set_cached(False)


class ItemCacher(BaseCacher):
    """contents = list of per-item inputer outputs (Resampler.item_cache) in item-id order (resampler.py:113-126)."""

    encode_packed = None   # set by ReprCacher when the item encoder supports padding-free execution

    def _cache(self, contents):
        op = self.operator
        out = op.get_full_placeholder(len(contents)).to(Env.device)
        with torch.no_grad():
            for s in range(0, len(contents), self.page_size):
                page = stack_trees(contents[s:s + self.page_size])
                n = len(contents[s:s + self.page_size])
                if self.encode_packed is not None:
                    out[s:s + n] = self.encode_packed(page['input_ids'], op.inputer.get_mask(page), training=False)[0]
                else:
                    emb = op.inputer.get_embeddings(page, training=False)
                    out[s:s + n] = op(emb, mask=op.inputer.get_mask(page))
        return out


class UserCacher(BaseCacher):
    """contents = the fast-eval dataset (one resampled row per user, user-id order; manager.py:209-227)."""

    def __init__(self, placeholder, **kwargs):
        super().__init__(**kwargs)
        self.placeholder = placeholder
        self.rows = None    # user-sharded evaluation (sharding.py): only these positions of `contents` are encoded

    def _cache(self, contents):
        out = torch.zeros_like(self.placeholder, device=Env.device)   # a fresh buffer every time: un-encoded rows are zeros, never stale
        todo = list(range(len(contents))) if self.rows is None else [int(i) for i in self.rows]
        with torch.no_grad():
            for s in range(0, len(todo), self.page_size):
                page = todo[s:s + self.page_size]
                rep = self.operator(batch=stack_trees([contents[i] for i in page]))
                if self.rows is None:
                    out[page[0]:page[0] + len(page)] = rep
                else:
                    out[torch.as_tensor(page, device=out.device)] = rep
        return out


class ReprCacher:
    def __init__(self, legommender):
        config = legommender.config
        self.use_item_content = config.use_item_content
        self.user_size = config.user_ut.meta.features[legommender.cm.user_col].tokenizer.vocab.size
        self._activate = True
        self.item = ItemCacher(operator=legommender.item_op, page_size=config.cache_page_size, hidden_size=config.hidden_size,
                               activate=legommender.item_op is not None and legommender.item_op.allow_caching,
                               trigger=Env.set_item_cache)
        if legommender._packed_items():
            self.item.encode_packed = legommender.encode_items_packed
        self.user = UserCacher(operator=legommender.get_user_content, page_size=config.cache_page_size,
                               hidden_size=config.hidden_size, activate=legommender.user_op.allow_caching,
                               placeholder=legommender.user_op.get_full_placeholder(self.user_size),
                               trigger=Env.set_user_cache)

    def activate(self, activate: bool):
        self._activate = activate

    def cache(self, item_contents, user_contents):
        if not self._activate:
            return
        if self.use_item_content:
            self.item.cache(item_contents)
        self.user.cache(user_contents)

    def clean(self):
        self.item.clean()
        self.user.clean()
