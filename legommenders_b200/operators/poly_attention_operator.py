"""MINER user encoder: poly attention over the click history -> `num_context_codes` interest vectors per user
(mirror of model/operators/poly_attention_operator.py:8-62)."""
import torch
from torch import nn

from .. import ops
from ..embedding_hub import _Affine
from ..env import Env
from ..inputer.concat_inputer import ConcatInputer
from .base_operator import BaseOperator, BaseOperatorConfig


class PolyAttentionOperatorConfig(BaseOperatorConfig):
    def __init__(self, num_context_codes: int = 32, context_code_dim: int = 200, **kwargs):
        super().__init__(**kwargs)
        self.num_context_codes = num_context_codes
        self.context_code_dim = context_code_dim


class PolyAttentionOperator(BaseOperator):
    config_class = PolyAttentionOperatorConfig
    inputer_class = ConcatInputer
    config: PolyAttentionOperatorConfig
    allow_caching = False            # [B, codes, D] per user does not fit the [users, hidden] cache (poly_attention_operator.py:24)

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        c = self.config
        self.linear = _Affine(c.input_dim, c.context_code_dim, bias=False)
        self.context_codes = nn.Parameter(nn.init.xavier_uniform_(torch.empty(c.num_context_codes, c.context_code_dim),
                                                                  gain=nn.init.calculate_gain('tanh')))

    def forward(self, embeddings, mask=None, **kwargs):
        """[B, S, D], [B, S] -> [B, codes, D]"""
        mask = mask.to(Env.device)
        proj = ops.linear(embeddings, self.linear.weight, None, act=ops.ACT_TANH)           # tanh(Linear(x))
        logits = ops.linear(proj, self.context_codes, None)                                  # proj · codesᵀ  [B, S, codes]
        return ops.poly_pool(logits, mask, embeddings)

    @property
    def output_dim(self):
        return self.config.input_dim
