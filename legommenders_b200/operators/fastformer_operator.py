"""Fastformer encoder -> Linear (mirror of model/operators/fastformer_operator.py:10-49 and model/common/fastformer.py:32-226).

State-dict names are the reference's (`fastformer.encoders.<i>.attention.self.{query,query_att,key,key_att,transform}`, `...attention.output`,
`...intermediate`, `...output`, `fastformer.position_embeddings`, `fastformer.LayerNorm`, `fastformer.poolers.0.att_fc{1,2}`, `linear`)."""
import math

import torch
from torch import nn

from .. import ops
from ..embedding_hub import _Affine
from ..env import Env
from ..inputer.concat_inputer import ConcatInputer
from .attention_operator import AttentionOperatorConfig
from .base_operator import BaseOperator
from .transformer_operator import LN_EPS, _Holder, _LN, _Table, _dense, bert_output


class FastformerOperatorConfig(AttentionOperatorConfig):
    def __init__(self, num_hidden_layers: int = 3, num_attention_heads: int = 8, hidden_dropout_prob: float = 0.1, **kwargs):
        super().__init__(num_attention_heads=num_attention_heads, **kwargs)
        self.num_hidden_layers = num_hidden_layers
        self.hidden_dropout_prob = hidden_dropout_prob


class FastformerParams(nn.Module):
    def __init__(self, dim, heads, layers, max_pos=1024):
        super().__init__()
        if dim % heads:
            raise ValueError('The hidden size (%d) is not a multiple of the number of attention heads (%d)' % (dim, heads))
        self.encoders = nn.ModuleList([
            _Holder(attention=_Holder(self=_Holder(query=_dense(dim, dim), query_att=_dense(dim, heads), key=_dense(dim, dim),
                                                   key_att=_dense(dim, heads), transform=_dense(dim, dim)),
                                      output=_Holder(dense=_dense(dim, dim), LayerNorm=_LN(dim))),
                    intermediate=_Holder(dense=_dense(dim, dim * 4)),
                    output=_Holder(dense=_dense(dim * 4, dim), LayerNorm=_LN(dim)))
            for _ in range(layers)])
        self.position_embeddings = _Table(max_pos, dim)
        self.LayerNorm = _LN(dim)
        self.poolers = nn.ModuleList([_Holder(att_fc1=_dense(dim, dim), att_fc2=_dense(dim, 1))])


class FastformerOperator(BaseOperator):
    config_class = FastformerOperatorConfig
    inputer_class = ConcatInputer
    config: FastformerOperatorConfig

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        c = self.config
        self.fastformer = FastformerParams(c.input_dim, c.num_attention_heads, c.num_hidden_layers)
        self.linear = _Affine(c.input_dim, c.hidden_size)

    def _self_attention(self, sa, h, mask):
        """FastSelfAttention.forward (fastformer.py:96-143)."""
        heads = self.config.num_attention_heads
        scale = 1.0 / math.sqrt(h.shape[-1] // heads)
        mq = ops.linear(h, sa.query.weight, sa.query.bias)
        mk = ops.linear(h, sa.key.weight, sa.key.bias)
        pooled_q = ops.head_pool(ops.linear(mq, sa.query_att.weight, sa.query_att.bias), mask, mq, scale)        # [B, D]
        mixed = ops.bcast_mul(mk, pooled_q)                                                                      # key * pooled query
        pooled_k = ops.head_pool(ops.linear(mixed, sa.key_att.weight, sa.key_att.bias), mask, mixed, scale)
        wv = ops.bcast_mul(mq, pooled_k)                                                                         # pooled key * query (query = value)
        return ops.add(ops.linear(wv, sa.transform.weight, sa.transform.bias), mq)

    def forward(self, embeddings, mask=None, **kwargs):
        mask = mask.to(Env.device)
        f = self.fastformer
        B, S, D = embeddings.shape
        p = self.config.hidden_dropout_prob if self.training else 0.0
        pos_ids = torch.arange(S, dtype=torch.int64, device=embeddings.device).repeat(B)
        x = ops.gather_add(embeddings.reshape(B * S, D), pos_ids, None, f.position_embeddings.weight)
        x = ops.dropout(ops.layernorm(x, f.LayerNorm.weight, f.LayerNorm.bias, LN_EPS), p, self._next_seed() if p else 0).view(B, S, D)
        for layer in f.encoders:
            so = self._self_attention(layer.attention.self, x, mask)
            att = bert_output(layer.attention.output, so, x, p, self._next_seed)
            inter = ops.gelu(ops.linear(att, layer.intermediate.dense.weight, layer.intermediate.dense.bias))
            x = bert_output(layer.output, inter, att, p, self._next_seed)
        # AttentionPooling (fastformer.py:32-59) is the library's additive attention: exp(fc2(tanh(fc1 x))) * mask / (sum + eps).  The scalar
        # bias of att_fc2 multiplies every weight by the same factor and cancels in the normalisation (up to eps: < 1e-7 relative); it is kept
        # as a parameter for state-dict parity and takes no part in the computation.
        pool = f.poolers[0]
        pooled = ops.additive_attention(x, mask, pool.att_fc1.weight, pool.att_fc1.bias, pool.att_fc2.weight)
        return ops.linear(pooled, self.linear.weight, self.linear.bias)
