"""LSTUR user encoder: GRU over the click history -> last hidden state -> Linear (mirror of model/operators/gru_operator.py:8-54)."""
import math

import torch
from torch import nn

from .. import ops
from ..embedding_hub import _Affine
from ..env import Env
from ..inputer.concat_inputer import ConcatInputer
from .base_operator import BaseOperator, BaseOperatorConfig


class GRUOperatorConfig(BaseOperatorConfig):
    def __init__(self, num_layers: int = 1, **kwargs):
        super().__init__(**kwargs)
        self.num_layers = num_layers


class _GRUParams(nn.Module):
    """nn.GRU's parameter names and default initialisation (U(-1/sqrt(H), 1/sqrt(H)) for every tensor), gate order (r, z, n)."""

    def __init__(self, input_size, hidden_size):
        super().__init__()
        k = 1.0 / math.sqrt(hidden_size)
        self.weight_ih_l0 = nn.Parameter(torch.empty(3 * hidden_size, input_size).uniform_(-k, k))
        self.weight_hh_l0 = nn.Parameter(torch.empty(3 * hidden_size, hidden_size).uniform_(-k, k))
        self.bias_ih_l0 = nn.Parameter(torch.empty(3 * hidden_size).uniform_(-k, k))
        self.bias_hh_l0 = nn.Parameter(torch.empty(3 * hidden_size).uniform_(-k, k))


class GRUOperator(BaseOperator):
    config_class = GRUOperatorConfig
    inputer_class = ConcatInputer
    config: GRUOperatorConfig

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        if self.config.num_layers != 1:
            raise ValueError('GRUOperator: only num_layers = 1 (config/model/lstur.yaml) runs on the fused recurrence kernel')
        self.gru = _GRUParams(self.config.input_dim, self.config.hidden_size)
        self.linear = _Affine(self.config.hidden_size, self.config.input_dim)

    def forward(self, embeddings, mask=None, **kwargs):
        lengths = mask.to(Env.device).sum(dim=1)            # pack_padded_sequence: the first `length` steps of every sequence
        g = self.gru
        last = ops.gru_last_hidden(embeddings, lengths, g.weight_ih_l0, g.weight_hh_l0, g.bias_ih_l0, g.bias_hh_l0)
        return ops.linear(last, self.linear.weight, self.linear.bias)

    @property
    def output_dim(self):
        return self.config.input_dim
