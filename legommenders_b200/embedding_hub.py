"""EmbeddingHub on B200 (mirror of loader/embedding_hub.py — same constructor, registration calls, error
behaviour and state-dict names; the lookups are CUDA kernels).

Tables:
  * `Table`          — a plain [V, D] table (nn.Embedding in the reference; key `<name>.weight`).
  * `Transformation` — pretrained [V, E] table + Linear(E -> D) + dropout (embedding_hub.py:73-96; keys
                       `<name>.embedding.weight`, `<name>.linear.{weight,bias}`).
Both expose `lookup_add(base, ids, mask)` = base + valid·row(ids), the fused form of the reference's
`emb = table(ids*mask); emb *= mask; acc += emb` (concat_inputer.py:105-113), and `__call__(ids)` for the
reference's plain-module use (legommender.py:244-246).
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, Optional

import numpy as np
import torch
from torch import nn

from . import ops
from .env import Env


_module_index = [0]


def next_module_index() -> int:
    """Stable per-module index for the counter-based dropout streams: construction order, not id() (reproducible across runs)."""
    _module_index[0] += 1
    return _module_index[0]


class _Weight(nn.Module):
    def __init__(self, weight: torch.Tensor, requires_grad=True):
        super().__init__()
        self.weight = nn.Parameter(weight, requires_grad=requires_grad)


class _Affine(nn.Module):
    """Parameter holder with nn.Linear's default initialisation (kaiming-uniform(a=√5), bias U(±1/√fan_in))."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if bias:
            bound = 1.0 / math.sqrt(in_features) if in_features > 0 else 0.0
            self.bias = nn.Parameter(torch.empty(out_features).uniform_(-bound, bound))
        else:
            self.register_parameter('bias', None)


class Table(nn.Module):
    """Plain trainable table — nn.Embedding(num_embeddings, embedding_dim) in embedding_hub.py:330-337 (N(0,1) init)."""

    def __init__(self, num_embeddings: int, embedding_dim: int, weight: Optional[torch.Tensor] = None, requires_grad=True):
        super().__init__()
        if weight is None:
            weight = torch.empty(num_embeddings, embedding_dim).normal_()
        self.weight = nn.Parameter(weight, requires_grad=requires_grad)

    def lookup_add(self, base, ids, mask=None, training=None):
        return ops.gather_add(base, ids, mask, self.weight)

    def forward(self, indexes):
        return ops.gather_add(None, indexes, None, self.weight)


class Transformation(nn.Module):
    """y = Dropout(Linear(Embedding(idx))) — embedding_hub.py:73-96."""

    def __init__(self, embedding: Table, to_dimension: int, transformation_dropout: float):
        super().__init__()
        self.embedding = _Weight(embedding.weight.data, requires_grad=embedding.weight.requires_grad)
        self.linear = _Affine(self.embedding.weight.shape[1], to_dimension)
        self.p = float(transformation_dropout)
        self._calls = 0
        self._index = next_module_index()

    def _seed(self):
        self._calls += 1
        return (torch.initial_seed() * 1000003 + self._index * 8191 + self._calls) & ((1 << 62) - 1)

    def lookup_add(self, base, ids, mask=None, training=None):
        """`training=None` follows this module's own flag (model.eval() switches every dropout off, base_lego.py:417)."""
        training = self.training if training is None else training
        valid = mask if mask is not None else ops.valid_mask(ids)
        rows = ops.gather_add(None, ids, valid, self.embedding.weight)
        p = self.p if (training and self.p > 0) else 0.0
        y = ops.linear(rows, self.linear.weight, self.linear.bias, rowmask=valid, drop_p=p, seed=self._seed() if p else 0)
        return y if base is None else base + y

    def forward(self, indexes):
        rows = ops.gather_add(None, indexes, None, self.embedding.weight)
        p = self.p if (self.training and self.p > 0) else 0.0
        return ops.linear(rows, self.linear.weight, self.linear.bias, drop_p=p, seed=self._seed() if p else 0)


class PretrainedEmbedding:
    def __init__(self, embedder, transformation, transformation_dropout, frozen):
        self.embedder = embedder
        self.transformation = transformation
        self.transformation_dropout = transformation_dropout
        self.frozen = frozen


class EmbeddingHub:
    LINEAR, AUTO, DEFAULT = 'linear', 'auto', 'default'
    global_types = {LINEAR, AUTO}
    pretrained_types = {DEFAULT, LINEAR, AUTO}

    def __init__(self, embedding_dim: int, transformation: str, transformation_dropout: float):
        if transformation not in self.global_types:
            raise ValueError(f'invalid transformation type {transformation}, expected {self.global_types}')
        self.embedding_dim = embedding_dim
        self.transformation = transformation
        self.transformation_dropout = transformation_dropout
        self._vocab_size: Dict[str, int] = {}
        self.vocab_table = nn.ModuleDict()
        self.feature_table = nn.ModuleDict()
        self._pretrained_vocab_embeddings: Dict[str, PretrainedEmbedding] = {}
        self._pretrained_feature_embeddings: Dict[str, PretrainedEmbedding] = {}

    # -- registration of pretrained matrices (embedding_hub.py:163-234) --------------------------------
    def load_pretrained_embedding(self, path, *, vocab_name=None, col_name=None, transformation=DEFAULT,
                                  transformation_dropout=None, frozen=True):
        if vocab_name is None and col_name is None:
            raise ValueError('vocab_name or col_name must be specified')
        if vocab_name is not None and col_name is not None:
            raise ValueError('only one of vocab_name and col_name can be specified')
        name = vocab_name or col_name
        arr = path if isinstance(path, (np.ndarray, torch.Tensor)) else np.load(path)
        weight = torch.as_tensor(arr, dtype=torch.float32)
        if name == '<vocab_name>':
            raise ValueError('please specify the vocab name for the pretrained embedding in the config')
        if transformation not in self.pretrained_types:
            raise ValueError(f'invalid transformation type {transformation}, expected {self.pretrained_types}')
        if transformation == self.DEFAULT:
            transformation = self.transformation
        if transformation_dropout is None:
            transformation_dropout = self.transformation_dropout
        target = self._pretrained_vocab_embeddings if vocab_name is not None else self._pretrained_feature_embeddings
        # nn.Embedding.from_pretrained freezes by default; requires_grad is finalised in _process_pretrained_embedding
        target[name] = PretrainedEmbedding(Table(weight.shape[0], weight.shape[1], weight=weight, requires_grad=False),
                                           transformation, transformation_dropout, frozen)

    def _process_pretrained_embedding(self, name: str, size: int, pe: PretrainedEmbedding):
        """embedding_hub.py:239-281 — note the projection decision uses the GLOBAL policy (reference quirk)."""
        if int(pe.embedder.weight.shape[0]) != size:
            raise ValueError(f'{name} does not match the expected vocab size {size}')
        pe.embedder.weight.requires_grad = not pe.frozen
        width = int(pe.embedder.weight.shape[1])
        if width != self.embedding_dim or self.transformation == self.LINEAR:
            pe.embedder = Transformation(pe.embedder, self.embedding_dim, pe.transformation_dropout)

    def build_feature_embedding(self, feature) -> bool:
        if feature.name in self.feature_table or feature.name not in self._pretrained_feature_embeddings:
            return False
        pe = self._pretrained_feature_embeddings[feature.name]
        self._process_pretrained_embedding(feature.name, feature.tokenizer.vocab.size, pe)
        self.feature_table.add_module(feature.name, pe.embedder.to(Env.device))
        return True

    def build_vocab_embedding(self, vocab):
        if vocab.name in self.vocab_table:
            return
        if vocab.name not in self._pretrained_vocab_embeddings:
            self.vocab_table.add_module(vocab.name, Table(vocab.size, self.embedding_dim).to(Env.device))
            return
        pe = self._pretrained_vocab_embeddings[vocab.name]
        self._process_pretrained_embedding(vocab.name, vocab.size, pe)
        self.vocab_table.add_module(vocab.name, pe.embedder.to(Env.device))

    def register_vocab(self, vocab):
        if vocab.name in self._vocab_size:
            if self._vocab_size[vocab.name] != vocab.size:
                raise ValueError(f'conflict in vocab {vocab.name}: {self._vocab_size[vocab.name]} vs {vocab.size}')
            return
        self._vocab_size[vocab.name] = vocab.size
        self.build_vocab_embedding(vocab)

    def register_ut(self, ut, used_cols: Iterable[str]):
        for col in used_cols:
            feature = ut.meta.features[col]
            self.build_feature_embedding(feature)
            self.register_vocab(feature.tokenizer.vocab)

    def __call__(self, vocab_name: str, col_name: Optional[str] = None) -> nn.Module:
        """Feature table wins over vocab table (embedding_hub.py:378-385)."""
        if col_name and col_name in self.feature_table:
            return self.feature_table[col_name]
        return self.vocab_table[vocab_name]
