"""NRMS encoder: multi-head self-attention -> Linear -> additive attention
(mirror of model/operators/attention_operator.py:9-59; used for items and for users)."""
from .. import ops
from ..embedding_hub import _Affine
from ..env import Env
from ..inputer.concat_inputer import ConcatInputer
from .base_operator import BaseOperator, BaseOperatorConfig
from .common import AdditiveAttention, MultiheadAttentionParams


class AttentionOperatorConfig(BaseOperatorConfig):
    def __init__(self, num_attention_heads: int = 8, attention_dropout: float = 0.1, additive_hidden_size: int = 256, **kwargs):
        super().__init__(**kwargs)
        self.num_attention_heads = num_attention_heads
        self.attention_dropout = attention_dropout
        self.additive_hidden_size = additive_hidden_size


class AttentionOperator(BaseOperator):
    config_class = AttentionOperatorConfig
    inputer_class = ConcatInputer
    config: AttentionOperatorConfig

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        c = self.config
        self.multi_head_attention = MultiheadAttentionParams(c.input_dim, c.num_attention_heads, c.attention_dropout)
        self.linear = _Affine(c.input_dim, c.hidden_size)
        self.additive_attention = AdditiveAttention(embed_dim=c.hidden_size, hidden_size=c.additive_hidden_size)

    supports_packed = True

    def forward(self, embeddings, mask=None, cu=None, max_len=None, **kwargs):
        """Dense: embeddings [N,S,D] + mask [N,S] (the reference's call).  Packed: embeddings [T,D] + cu int32 [N+1]."""
        if cu is None:
            mask = mask.to(Env.device)
        mha = self.multi_head_attention
        p = mha.dropout if self.training else 0.0
        qkv = ops.linear(embeddings, mha.in_proj_weight, mha.in_proj_bias)
        ctx = ops.mha_core(qkv, mask, mha.num_heads, drop_p=p, seed=self._next_seed() if p else 0, cu=cu, max_len=max_len)
        out = ops.linear(ctx, mha.out_proj.weight, mha.out_proj.bias)
        lin = ops.linear(out, self.linear.weight, self.linear.bias)
        return self.additive_attention(lin, mask, cu=cu, max_len=max_len)
