"""SimpleInputer (mirror of model/inputer/simple_inputer.py:11-66): one padded id row + mask per column."""
from collections import OrderedDict

import torch

from ..env import Env
from .base_inputer import BaseInputer


class SimpleInputer(BaseInputer):
    output_single_sequence = False

    def get_vocabs(self):
        return []

    @classmethod
    def pad(cls, l: list, max_len: int):
        n = max_len - len(l)
        return l + [Env.UNSET] * n, [1] * len(l) + [0] * n

    def sample_rebuilder(self, sample: dict):
        input_ids, attention_mask = dict(), dict()
        for col in self.inputs:
            max_len = self.ut.meta.features[col].max_len
            value = sample[col]
            if not max_len:
                value, max_len = [value], 1
                sample[col] = value      # the reference rewrites the sample in place (simple_inputer.py:27)
            ids, mask = self.pad(list(value), max_len)
            input_ids[col] = torch.tensor(ids)
            attention_mask[col] = torch.tensor(mask)
        return dict(input_ids=input_ids, attention_mask=attention_mask)

    def get_mask(self, batched_samples):
        return OrderedDict(batched_samples['attention_mask'])

    def get_embeddings(self, batched_samples, training=None):
        """Per column mask·table[ids] with the stored attention mask, no sum (simple_inputer.py:43-66)."""
        out = OrderedDict()
        for col, ids in batched_samples['input_ids'].items():
            vocab = self.ut.meta.features[col].tokenizer.vocab.name
            ids = ids.to(Env.device, non_blocking=True)
            mask = batched_samples['attention_mask'][col].to(Env.device, non_blocking=True)
            out[col] = self.eh(vocab, col_name=col).lookup_add(None, ids, mask, training=training)
        return out
