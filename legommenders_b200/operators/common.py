"""Parameter holders that keep the reference's state-dict names and PyTorch's default initialisation."""
import math

import torch
from torch import nn

from .. import ops
from ..embedding_hub import _Affine


class AdditiveAttention(nn.Module):
    """model/common/attention.py:10-38; parameters live under `encoder.0.{weight,bias}` and `encoder.2.weight`."""

    def __init__(self, embed_dim, hidden_size):
        super().__init__()
        self.embed_dim, self.hidden_size = embed_dim, hidden_size
        self.encoder = nn.ModuleDict({'0': _Affine(embed_dim, hidden_size), '2': _Affine(hidden_size, 1, bias=False)})

    def forward(self, inputs, attention_mask=None, cu=None, max_len=None):
        e = self.encoder
        return ops.additive_attention(inputs, attention_mask, e['0'].weight, e['0'].bias, e['2'].weight, cu=cu, max_len=max_len)


class MultiheadAttentionParams(nn.Module):
    """Holds nn.MultiheadAttention's parameters (`in_proj_weight`, `in_proj_bias`, `out_proj.{weight,bias}`) with its
    initialisation (xavier-uniform in_proj, zero biases); the computation is ops.linear + ops.mha_core."""

    def __init__(self, embed_dim, num_heads, dropout):
        super().__init__()
        if embed_dim % num_heads:
            raise ValueError('embed_dim must be divisible by num_heads')
        self.embed_dim, self.num_heads, self.dropout = embed_dim, num_heads, float(dropout)
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dim))
        nn.init.xavier_uniform_(self.in_proj_weight)
        self.out_proj = _Affine(embed_dim, embed_dim)
        nn.init.zeros_(self.out_proj.bias)
