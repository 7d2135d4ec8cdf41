"""Synthetic MIND-shaped data behind a UniTok-shaped facade.

The reference reads its data through UniTok tables (`loader/ut/lego_ut.py`); the hot
path only touches ``ut.meta.features[col].{name,max_len,tokenizer.vocab.{name,size}}``,
``ut.key_feature``, ``len(ut)`` and ``ut[i] -> dict``.  `Table` is the smallest object
with that surface, so the same synthetic world can be handed to the live reference
(golden generation), the oracle and the CUDA path.

Everything is numpy (PCG64) so that a seed reproduces the same world on any box.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, List, Optional

import numpy as np

DEFAULT_SEED = 2023  # utils/function.py:57-75 (reference default seed)


class Vocab:
    """Stand-in for unitok.Vocab: a name and a size (append returns the new index)."""

    def __init__(self, name: str, size: int = 0):
        self.name = name
        self._size = size
        self._tokens: List[str] = []

    def append(self, token: str) -> int:
        self._tokens.append(token)
        self._size = max(self._size, len(self._tokens))
        return len(self._tokens) - 1

    @property
    def size(self) -> int:
        return self._size

    def __len__(self):
        return self._size


class Feature:
    def __init__(self, name: str, vocab: Vocab, max_len: Optional[int] = None):
        self.name = name
        self.max_len = max_len
        self.tokenizer = SimpleNamespace(vocab=vocab)


class Table:
    """UniTok-shaped table: columnar python lists + feature meta."""

    def __init__(self, features: List[Feature], key: str, columns: Dict[str, list]):
        self.meta = SimpleNamespace(features={f.name: f for f in features})
        self.key_feature = self.meta.features[key]
        self.columns = columns
        self._n = len(next(iter(columns.values()))) if columns else 0

    def __len__(self):
        return self._n

    def __getitem__(self, i: int) -> dict:
        return {c: v[i] for c, v in self.columns.items()}

    def __iter__(self):
        for i in range(self._n):
            yield self[i]


def zipf_probs(n: int, a: float = 1.1) -> np.ndarray:
    p = 1.0 / np.arange(1, n + 1, dtype=np.float64) ** a
    return p / p.sum()


class MindWorld:
    """A seeded synthetic MIND-small-shaped world.

    items : title ~ U{min_title..title_len} tokens Zipf(1.1) over the word vocab, category ~ U{0..n_cats-1}
    users : history ~ U{1..hist_len} item ids Zipf over items, neg list U{0..max_neg}
    train : positives uniform over items, one row per impression (index,user_id,item_id,click,history,neg)
    eval  : groups of (user_id,item_id,click) rows with >=1 positive and >=1 negative per group
    """

    def __init__(self, n_items=2000, n_words=5000, n_users=500, n_cats=18, title_len=30, hist_len=50,
                 n_train=4096, n_eval_groups=200, eval_group_mean=36, min_title=5, max_neg=100,
                 embed_dim=300, seed=DEFAULT_SEED, title_col='title@glove', word_vocab='glove',
                 glove_std=0.4, make_table=True, min_hist=1):
        rng = np.random.default_rng(seed)
        self.seed = seed
        self.n_items, self.n_words, self.n_users, self.n_cats = n_items, n_words, n_users, n_cats
        self.title_len, self.hist_len, self.embed_dim = title_len, hist_len, embed_dim
        self.n_train = n_train
        self.title_col, self.word_vocab = title_col, word_vocab

        # ---- items ------------------------------------------------------------------
        wp = zipf_probs(n_words)
        wperm = rng.permutation(n_words)  # so that frequent words are not the low ids
        lens = rng.integers(min(min_title, title_len), title_len + 1, size=n_items)
        flat = wperm[rng.choice(n_words, size=int(lens.sum()), p=wp)]
        offs = np.concatenate([[0], np.cumsum(lens)])
        self.title_lens = lens.astype(np.int64)
        self.titles = [flat[offs[i]:offs[i + 1]].astype(np.int64) for i in range(n_items)]
        self.cats = rng.integers(0, n_cats, size=n_items).astype(np.int64)

        # ---- users ------------------------------------------------------------------
        ip = zipf_probs(n_items)
        iperm = rng.permutation(n_items)
        hl = rng.integers(min(min_hist, hist_len), hist_len + 1, size=n_users)
        hflat = iperm[rng.choice(n_items, size=int(hl.sum()), p=ip)]
        hoffs = np.concatenate([[0], np.cumsum(hl)])
        self.hist_lens = hl.astype(np.int64)
        self.histories = [hflat[hoffs[u]:hoffs[u + 1]].astype(np.int64) for u in range(n_users)]
        nl = rng.integers(0, max_neg + 1, size=n_users)
        self.negs = [rng.integers(0, n_items, size=int(nl[u])).astype(np.int64) for u in range(n_users)]

        # ---- train impressions --------------------------------------------------------
        self.train_users = rng.integers(0, n_users, size=n_train).astype(np.int64)
        self.train_pos = rng.integers(0, n_items, size=n_train).astype(np.int64)

        # ---- eval rows ------------------------------------------------------------------
        g_users = rng.permutation(n_users)[:min(n_eval_groups, n_users)]
        gs = np.maximum(2, rng.poisson(eval_group_mean, size=len(g_users)))
        eu, ei, ec = [], [], []
        for u, s in zip(g_users, gs):
            items = rng.integers(0, n_items, size=int(s))
            click = (rng.random(int(s)) < 0.12).astype(np.int64)
            click[0], click[1] = 1, 0  # both classes present (sklearn AUC needs them)
            eu.append(np.full(int(s), u)); ei.append(items); ec.append(click)
        self.eval_users = np.concatenate(eu).astype(np.int64)
        self.eval_items = np.concatenate(ei).astype(np.int64)
        self.eval_click = np.concatenate(ec).astype(np.int64)

        # ---- GloVe-shaped table -----------------------------------------------------------
        self.word_table = None
        if make_table:
            self.word_table = (rng.standard_normal((n_words, embed_dim), dtype=np.float32)
                               * np.float32(glove_std))

        self.word_v = Vocab(word_vocab, n_words)
        self.cat_v = Vocab('category', n_cats)
        self.item_v = Vocab('item_id', n_items)
        self.user_v = Vocab('user_id', n_users)
        self.index_v = Vocab('index', max(n_train, len(self.eval_users), n_users))
        self.click_v = Vocab('click', 2)

    def item_title_matrix(self) -> np.ndarray:
        """[n_items, title_len] int64 token ids, right-padded with -1 (SimpleInputer layout of the title column)."""
        m = np.full((self.n_items, self.title_len), -1, dtype=np.int64)
        for i, t in enumerate(self.titles):
            m[i, :len(t)] = t
        return m

    # UniTok-shaped views ---------------------------------------------------------------------
    def item_table(self) -> Table:
        feats = [Feature('item_id', self.item_v), Feature(self.title_col, self.word_v, self.title_len),
                 Feature('category', self.cat_v, None)]
        cols = {'item_id': list(range(self.n_items)),
                self.title_col: [t.tolist() for t in self.titles],
                'category': self.cats.tolist()}
        return Table(feats, 'item_id', cols)

    def user_table(self) -> Table:
        feats = [Feature('user_id', self.user_v), Feature('history', self.item_v, self.hist_len),
                 Feature('neg', self.item_v, None)]
        cols = {'user_id': list(range(self.n_users)),
                'history': [h.tolist() for h in self.histories],
                'neg': [n.tolist() for n in self.negs]}
        return Table(feats, 'user_id', cols)

    def _inter(self, users, items, click) -> Table:
        feats = [Feature('index', self.index_v), Feature('user_id', self.user_v), Feature('item_id', self.item_v),
                 Feature('click', self.click_v), Feature('history', self.item_v, self.hist_len),
                 Feature('neg', self.item_v, None)]
        n = len(users)
        cols = {'index': list(range(n)), 'user_id': [int(u) for u in users], 'item_id': [int(i) for i in items],
                'click': [int(c) for c in click],
                'history': [self.histories[int(u)].tolist() for u in users],
                'neg': [self.negs[int(u)].tolist() for u in users]}
        return Table(feats, 'index', cols)

    def train_table(self) -> Table:
        return self._inter(self.train_users, self.train_pos, np.ones_like(self.train_pos))

    def eval_table(self) -> Table:
        return self._inter(self.eval_users, self.eval_items, self.eval_click)

    def fast_table(self) -> Table:
        """manager.py:209-227: one dummy row per user in user-id order."""
        u = np.arange(self.n_users)
        t = self._inter(u, np.zeros_like(u), np.zeros_like(u))
        t.columns['item_id'] = [[0] for _ in u]
        t.columns['click'] = [[0] for _ in u]
        t.columns['neg'] = [[] for _ in u]
        return t


def init_state(shapes: Dict[str, tuple], seed: int, scale: float = 0.08) -> Dict[str, np.ndarray]:
    """Deterministic fp32 parameters for parity cases (numpy PCG64; order = sorted names)."""
    rng = np.random.default_rng(seed)
    out = {}
    for name in sorted(shapes):
        out[name] = (rng.standard_normal(shapes[name]) * scale).astype(np.float32)
    return out
