"""UniTok dataset directories <-> the `Table` facade of the hot path.

The reference reads its data only through the third-party `unitok` package (`loader/ut/lego_ut.py:80-109` calls `UniTok.load(save_dir)`;
`processor/base_processor.py:300-373` writes one directory per table with `UniTok.save`): an `items` directory (key `item_id`, text
features, `category`), a `users` directory (key `user_id`, `history`, `neg`) and one directory per interaction split (`index`, `user_id`,
`item_id`, `click`).  The hot path touches `ut.meta.features[col].{name, max_len, tokenizer.vocab.{name, size}}`, `ut.key_feature`,
`len(ut)` and `ut[i]` — what `synth.Table` provides.

PARITY UNPINNED: `unitok` is not vendored in /root/reference, not installed here and unversioned in the reference's requirements.txt, so
its byte layout cannot be checked offline.  What is restated here is the published layout of unitok 4.x as its `UniTok.save` writes it:

    <dir>/meta.json     {"version", "vocabularies": [{"name", "size"}], "tokenizers": [{"tokenizer_id", "vocab", "classname", "params"}],
                         "features": [{"name", "column", "tokenizer", "truncate", "max_len", "key"}]}      (unitok < 4.3: "jobs")
    <dir>/data.pkl      pickle of {feature name: list of per-sample values (int, or list of int)}           (unitok 3.x: data.npy)
    <dir>/<vocab>.vocab pickle of the token list in index order                                             (unitok 3.x: tok.<vocab>.dat, text)

The reader accepts each of the variants named in parentheses; the writer emits the 4.x form.  The tests round-trip a synthetic world through
these directories and drive the batch builders from the loaded tables; the first directory written by the real package decides whether the
restatement is right.  `pickle.load` executes what the file says, exactly as `UniTok.load` does: only open directories you wrote.
"""
from __future__ import annotations

import json
import os
import pickle
from types import SimpleNamespace
from typing import Dict, Optional

import numpy as np

from .synth import Feature, Table, Vocab

META_FILE = 'meta.json'
DATA_FILE = 'data.pkl'
WRITER_VERSION = 'unidep-v4'


def _read_vocab_tokens(save_dir: str, name: str):
    """Token list of a vocabulary in index order, or None when the directory carries only its size."""
    p = os.path.join(save_dir, f'{name}.vocab')
    if os.path.exists(p):
        with open(p, 'rb') as f:
            head = f.read(2)
        if head[:1] == b'\x80':                                 # pickle protocol >= 2
            with open(p, 'rb') as f:
                return list(pickle.load(f))
        with open(p, 'r', encoding='utf-8') as f:               # one token per line
            return f.read().split('\n')[:-1]
    p = os.path.join(save_dir, f'tok.{name}.dat')               # unitok 3.x
    if os.path.exists(p):
        with open(p, 'r', encoding='utf-8') as f:
            return f.read().split('\n')[:-1]
    return None


def _read_data(save_dir: str) -> Dict[str, list]:
    p = os.path.join(save_dir, DATA_FILE)
    if os.path.exists(p):
        with open(p, 'rb') as f:
            return pickle.load(f)
    p = os.path.join(save_dir, 'data.npy')                      # unitok 3.x
    if os.path.exists(p):
        return np.load(p, allow_pickle=True).item()
    raise FileNotFoundError(f'{save_dir}: neither {DATA_FILE} nor data.npy')


def load_table(save_dir: str) -> Table:
    """`UniTok.load(save_dir)` for the surface the hot path uses.  Values stay what the directory holds (ints, lists of ints)."""
    with open(os.path.join(save_dir, META_FILE), 'r', encoding='utf-8') as f:
        meta = json.load(f)
    vocabs: Dict[str, Vocab] = {}
    for v in meta.get('vocabularies', meta.get('vocabs', [])):
        voc = Vocab(v['name'], int(v.get('size', 0)))
        tokens = _read_vocab_tokens(save_dir, v['name'])
        if tokens is not None:
            voc._tokens = list(tokens)
            if voc._size and voc._size != len(tokens):
                raise ValueError(f'{save_dir}: vocabulary {v["name"]!r} has {len(tokens)} tokens, meta.json says {voc._size}')
            voc._size = len(tokens)
        vocabs[v['name']] = voc
    tokenizers = {t['tokenizer_id']: t for t in meta.get('tokenizers', [])}
    feats, key = [], meta.get('key_feature', meta.get('key_job'))
    for fj in meta.get('features', meta.get('jobs', [])):
        tok = tokenizers.get(fj.get('tokenizer'))
        vname = tok['vocab'] if tok is not None else fj.get('vocab')
        if vname not in vocabs:
            raise ValueError(f'{save_dir}: feature {fj["name"]!r} refers to the unknown vocabulary {vname!r}')
        ml = fj.get('max_len')
        feat = Feature(fj['name'], vocabs[vname], int(ml) if ml else None)
        feat.column = fj.get('column', fj['name'])
        feat.truncate = fj.get('truncate')
        if tok is not None:
            feat.tokenizer.classname = tok.get('classname')
            feat.tokenizer.tokenizer_id = tok['tokenizer_id']
        feats.append(feat)
        if fj.get('key') or fj.get('is_key'):
            key = fj['name']
    if key is None:
        raise ValueError(f'{save_dir}: no key feature')
    data = _read_data(save_dir)
    cols = {}
    n = None
    for ft in feats:
        if ft.name not in data:
            raise ValueError(f'{save_dir}: no data for feature {ft.name!r}')
        col = data[ft.name]
        col = col.tolist() if isinstance(col, np.ndarray) else list(col)
        if n is None:
            n = len(col)
        elif len(col) != n:
            raise ValueError(f'{save_dir}: feature {ft.name!r} has {len(col)} samples, expected {n}')
        cols[ft.name] = col
    table = Table(feats, key, cols)
    table.meta.version = meta.get('version')
    table.meta.vocabularies = list(vocabs.values())
    table.save_dir = save_dir
    return table


def save_table(table: Table, save_dir: str, tokens: Optional[Dict[str, list]] = None) -> None:
    """`UniTok.save(save_dir)` in the 4.x layout.  tokens: {vocab name: token list}; vocabularies without one are written as range(size)
    (EntityTokenizer over ids that already are indices)."""
    os.makedirs(save_dir, exist_ok=True)
    vocabs, tokenizers, features = {}, [], []
    for name, ft in table.meta.features.items():
        v = ft.tokenizer.vocab
        vocabs[v.name] = v
        many = any(isinstance(x, (list, tuple, np.ndarray)) for x in table.columns[name][:16])
        tid = getattr(ft.tokenizer, 'tokenizer_id', None) or f'auto_{name}'
        tokenizers.append(dict(tokenizer_id=tid, vocab=v.name, classname=getattr(ft.tokenizer, 'classname', None) or
                               ('EntitiesTokenizer' if many else 'EntityTokenizer'), params={}))
        features.append(dict(name=name, column=getattr(ft, 'column', name), tokenizer=tid, truncate=getattr(ft, 'truncate', None),
                             max_len=ft.max_len or 0, key=(ft is table.key_feature)))
    meta = dict(version=WRITER_VERSION, note='written by legommenders_b200.ut_io.save_table',
                vocabularies=[dict(name=v.name, size=v.size) for v in vocabs.values()], tokenizers=tokenizers, features=features)
    with open(os.path.join(save_dir, META_FILE), 'w', encoding='utf-8') as f:
        json.dump(meta, f, indent=2)
    for v in vocabs.values():
        toks = (tokens or {}).get(v.name) or (v._tokens if len(v._tokens) == v.size else list(range(v.size)))
        with open(os.path.join(save_dir, f'{v.name}.vocab'), 'wb') as f:
            pickle.dump(list(toks), f)
    with open(os.path.join(save_dir, DATA_FILE), 'wb') as f:
        pickle.dump({k: list(v) for k, v in table.columns.items()}, f)


class DirWorld:
    """A processed dataset (the directory tree `processor/base_processor.py:300-373` writes) behind the attributes the builders read from
    `synth.MindWorld`: item / user tables, per-user history and negative lists, the positive training impressions, the evaluation rows and
    the pretrained word table.

    root/items, root/users, root/train, root/valid are UniTok directories; `word_table` is the fp32 [vocab, dim] matrix the reference's
    embed config points at (`loader/embedding_hub.py:180-215`: a .npy per vocabulary), given as a path or an array.
    Column names follow the reference's MIND processor and `loader/column_map.py` defaults; override them for other datasets."""

    def __init__(self, root: str, title_col: str = 'title@glove', word_vocab: Optional[str] = None, word_table=None,
                 item_dir: str = 'items', user_dir: str = 'users', train_dir: str = 'train', valid_dir: str = 'valid',
                 item_col: str = 'item_id', user_col: str = 'user_id', history_col: str = 'history', neg_col: str = 'neg',
                 label_col: str = 'click', hist_len: Optional[int] = None):
        self.root = root
        self._items = load_table(os.path.join(root, item_dir))
        self._users = load_table(os.path.join(root, user_dir))
        self._train = load_table(os.path.join(root, train_dir)) if train_dir and os.path.isdir(os.path.join(root, train_dir)) else None
        self._valid = load_table(os.path.join(root, valid_dir)) if valid_dir and os.path.isdir(os.path.join(root, valid_dir)) else None
        self.title_col = title_col
        feats = self._items.meta.features
        if title_col not in feats:
            raise KeyError(f'{root}: the item table has no feature {title_col!r} (has {sorted(feats)})')
        self.word_vocab = word_vocab or feats[title_col].tokenizer.vocab.name
        self.n_items, self.n_users = len(self._items), len(self._users)
        self.n_words = feats[title_col].tokenizer.vocab.size
        self.title_len = feats[title_col].max_len or max((len(t) for t in self._items.columns[title_col]), default=0)
        ufe = self._users.meta.features
        self.hist_len = hist_len or ufe[history_col].max_len or max((len(h) for h in self._users.columns[history_col]), default=0)
        # rows of the item / user tables are in key order (the key vocabulary is built while tokenising them), so a key IS a row number
        for tab, col in ((self._items, item_col), (self._users, user_col)):
            keys = np.asarray(tab.columns[col], dtype=np.int64)
            if not np.array_equal(keys, np.arange(len(keys))):
                raise ValueError(f'{tab.save_dir}: {col} is not the row number; re-index the table')
        self.histories = [np.asarray(h, dtype=np.int64)[-self.hist_len:] if self.hist_len else np.asarray(h, dtype=np.int64)
                          for h in self._users.columns[history_col]]
        self.negs = ([np.asarray(x, dtype=np.int64) for x in self._users.columns[neg_col]] if neg_col in self._users.columns
                     else [np.zeros(0, dtype=np.int64) for _ in range(self.n_users)])
        self.titles = [np.asarray(t, dtype=np.int64) for t in self._items.columns[title_col]]
        self.cats = np.asarray(self._items.columns['category'], dtype=np.int64) if 'category' in self._items.columns else None
        self.n_cats = feats['category'].tokenizer.vocab.size if 'category' in feats else 0
        if self._train is not None:
            lab = np.asarray(self._train.columns[label_col], dtype=np.int64)
            pos = lab > 0                                        # the training set keeps clicked rows (loader/manager.py:330-347: "lambda x: x == 1" on the label)
            self.train_users = np.asarray(self._train.columns[user_col], dtype=np.int64)[pos]
            self.train_pos = np.asarray(self._train.columns[item_col], dtype=np.int64)[pos]
            self.n_train = int(pos.sum())
        if self._valid is not None:
            self.eval_users = np.asarray(self._valid.columns[user_col], dtype=np.int64)
            self.eval_items = np.asarray(self._valid.columns[item_col], dtype=np.int64)
            self.eval_click = np.asarray(self._valid.columns[label_col], dtype=np.int64)
        if isinstance(word_table, str):
            word_table = np.load(word_table)
        self.word_table = None if word_table is None else np.ascontiguousarray(word_table, dtype=np.float32)
        if self.word_table is not None and self.word_table.shape[0] != self.n_words:
            raise ValueError(f'word table has {self.word_table.shape[0]} rows, vocabulary {self.word_vocab!r} has {self.n_words}')
        self.embed_dim = None if self.word_table is None else int(self.word_table.shape[1])
        self._user_cols = (user_col, history_col, neg_col)

    def item_table(self) -> Table:
        return self._items

    def user_table(self) -> Table:
        return self._users

    def train_table(self) -> Optional[Table]:
        return self._train

    def eval_table(self) -> Optional[Table]:
        return self._valid


def save_world(world, root: str) -> None:
    """Write a `synth.MindWorld` as the directory tree `DirWorld` reads (and, if the restatement above is right, `LegoUT.load` does)."""
    save_table(world.item_table(), os.path.join(root, 'items'))
    users = world.user_table()
    save_table(users, os.path.join(root, 'users'))

    def inter(t: Table) -> Table:                                # interaction directories carry ids and the label only
        keep = ('index', 'user_id', 'item_id', 'click')
        feats = [t.meta.features[k] for k in keep]
        return Table(feats, 'index', {k: t.columns[k] for k in keep})

    n = len(world.train_users)
    tr = world._inter(world.train_users, world.train_pos, np.ones(n, dtype=np.int64))
    save_table(inter(tr), os.path.join(root, 'train'))
    save_table(inter(world.eval_table()), os.path.join(root, 'valid'))
    if world.word_table is not None:
        np.save(os.path.join(root, f'{world.word_vocab}.npy'), world.word_table)


__all__ = ['load_table', 'save_table', 'DirWorld', 'save_world', 'SimpleNamespace']
