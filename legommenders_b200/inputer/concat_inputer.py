"""ConcatInputer (mirror of model/inputer/concat_inputer.py:24-114).

Host side (`sample_rebuilder`) produces the same int64 layout as the reference, bit for bit; device side
(`get_embeddings`) replaces the per-column embedding/mask/add chain with fused gather kernels.
"""
from collections import OrderedDict
from typing import Dict

import torch

from ..env import Env
from ..synth import Vocab
from .base_inputer import BaseInputer


class ConcatInputer(BaseInputer):
    output_single_sequence = True

    vocab = Vocab(name='__cat_inputer_special_ids')
    PAD = vocab.append('[PAD]')
    CLS = vocab.append('[CLS]')
    SEP = vocab.append('[SEP]')

    def __init__(self, use_cls_token, use_sep_token, **kwargs):
        super().__init__(**kwargs)
        self.use_cls_token = use_cls_token
        self.use_sep_token = use_sep_token
        self.vocab_activated = bool(use_sep_token or use_cls_token)
        self.max_content_len = sum((self.ut.meta.features[c].max_len or 1) for c in self.inputs)
        self.max_sequence_len = self.max_content_len + int(bool(use_cls_token)) + int(bool(use_sep_token)) * len(self.inputs)

    def get_max_content_len(self):
        return self.max_content_len

    def get_max_sequence_len(self):
        return self.max_sequence_len

    def get_vocabs(self):
        return [self.vocab] if self.vocab_activated else []

    def get_empty_input(self):
        return torch.full((self.max_sequence_len,), Env.UNSET, dtype=torch.long)

    def sample_rebuilder(self, sample):
        """concat_inputer.py:58-87: [CLS?] col0… [SEP?] col1… [SEP?], left-packed; one row per source column."""
        S = self.max_sequence_len
        pos = 0
        special = self.get_empty_input()
        ids = OrderedDict()
        if self.use_cls_token:
            special[pos] = self.CLS
            pos += 1
        for col in self.inputs:
            value = sample[col]
            if not isinstance(value, list):
                value = [value]
            row = self.get_empty_input()
            row[pos:pos + len(value)] = torch.tensor(value, dtype=torch.long)
            pos += len(value)
            ids[col] = row
            if self.use_sep_token:
                special[pos] = self.SEP
                pos += 1
        if self.vocab_activated:
            special[pos:] = self.PAD
            ids[self.vocab.name] = special
        mask = torch.zeros(S, dtype=torch.long)
        mask[:pos] = 1
        return dict(input_ids=ids, attention_mask=mask)

    def get_mask(self, batched_samples: Dict[str, torch.Tensor]):
        return batched_samples['attention_mask']

    def _table(self, col):
        vocab = col if col == self.vocab.name else self.ut.meta.features[col].tokenizer.vocab.name
        return self.eh(vocab)

    def get_embeddings(self, batched_samples, training=None):
        """Σ_cols [ids>-1]·table_c[ids] (concat_inputer.py:92-114).  The caller's ids are NOT mutated
        (the reference's in-place `seq *= mask` is an artefact).
        `training=None`: dropout follows each table module's own `.training` flag, as nn.Dropout does in the reference."""
        out = None
        for col, ids in batched_samples['input_ids'].items():
            ids = ids.to(Env.device, non_blocking=True)
            out = self._table(col).lookup_add(out, ids, None, training=training)
        return out
