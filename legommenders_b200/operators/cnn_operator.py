"""NAML item encoder: per column Conv1d('same')+ReLU+mask (or Linear for length-1 columns), concat, additive attention
(mirror of model/operators/cnn_operator.py:10-67)."""
import math

import torch
from torch import nn

from .. import ops
from ..embedding_hub import _Affine
from ..env import Env
from ..inputer.simple_inputer import SimpleInputer
from .base_operator import BaseOperator, BaseOperatorConfig
from .common import AdditiveAttention


class CNNOperatorConfig(BaseOperatorConfig):
    def __init__(self, kernel_size: int = 3, dropout: float = 0.1, additive_hidden_size: int = 256, **kwargs):
        super().__init__(**kwargs)
        self.kernel_size = kernel_size
        self.dropout = dropout
        self.additive_hidden_size = additive_hidden_size


class _ConvParams(nn.Module):
    """nn.Conv1d parameter shapes and default init: weight [out, in, k] kaiming-uniform(a=√5), bias U(±1/√(in·k))."""

    def __init__(self, cin, cout, k):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin, k))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        bound = 1.0 / math.sqrt(cin * k)
        self.bias = nn.Parameter(torch.empty(cout).uniform_(-bound, bound))


class CNNOperator(BaseOperator):
    config_class = CNNOperatorConfig
    inputer_class = SimpleInputer
    config: CNNOperatorConfig

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        c = self.config
        if c.kernel_size % 2 == 0:
            raise ValueError("padding='same' with an even kernel size is not supported")
        self.cnn = _ConvParams(c.input_dim, c.hidden_size, c.kernel_size)
        self.linear = _Affine(c.input_dim, c.hidden_size)
        self.additive_attention = AdditiveAttention(embed_dim=c.hidden_size, hidden_size=c.additive_hidden_size)

    def forward(self, embeddings: dict, mask=None, **kwargs):
        outs, masks = [], []
        p = self.config.dropout if self.training else 0.0
        for col, e in embeddings.items():
            m = mask[col].to(Env.device)
            if e.shape[1] > 1:
                outs.append(ops.conv1d_relu_mask(e, self.cnn.weight, self.cnn.bias, m, drop_p=p,
                                                 seed=self._next_seed() if p else 0))
            else:
                outs.append(ops.linear(e, self.linear.weight, self.linear.bias))
            masks.append(m)
        # column concat along the sequence axis is pure data movement (cnn_operator.py:64-65)
        return self.additive_attention(torch.cat(outs, dim=1), torch.cat(masks, dim=1))
