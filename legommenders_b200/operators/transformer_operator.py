"""BERT-style encoder -> Linear -> additive attention (mirror of model/operators/transformer_operator.py:10-61, whose `transformer` is
transformers' BertModel(vocab_size=1, type_vocab_size=1, max_position_embeddings=1024) fed with `inputs_embeds`).

State-dict names are BertModel's (`transformer.embeddings.*`, `transformer.encoder.layer.<i>.*`, `transformer.pooler.dense.*` — the pooler and the
word embedding exist but are never used, exactly as in the reference).  Every tensor op is a kernel of this library: position / token-type rows by
the gather kernel, LayerNorm (+ residual), QKV / output / feed-forward contractions, the attention core, erf-GELU, counter-based dropout."""
import torch
from torch import nn

from .. import ops
from ..embedding_hub import _Affine
from ..env import Env
from ..inputer.concat_inputer import ConcatInputer
from .attention_operator import AttentionOperatorConfig
from .base_operator import BaseOperator
from .common import AdditiveAttention

LN_EPS = 1e-12            # BertConfig.layer_norm_eps
HIDDEN_DROPOUT = 0.1      # BertConfig.hidden_dropout_prob: the reference leaves it at the default


class _LN(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.bias = nn.Parameter(torch.zeros(dim))


class _Table(nn.Module):
    def __init__(self, rows, dim):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(rows, dim).normal_(0, 0.02))


def _dense(i, o):
    m = _Affine(i, o)
    with torch.no_grad():
        m.weight.normal_(0, 0.02)          # BertPreTrainedModel._init_weights
        m.bias.zero_()
    return m


class _Holder(nn.Module):
    """Plain container: attribute names become state-dict path components."""

    def __init__(this, **children):          # `self` is a legitimate child name (BertAttention.self)
        super().__init__()
        for k, v in children.items():
            this.add_module(k, v)


class BertParams(nn.Module):
    def __init__(self, dim, heads, layers, intermediate, max_pos=1024):
        super().__init__()
        self.embeddings = _Holder(word_embeddings=_Table(1, dim), position_embeddings=_Table(max_pos, dim),
                                  token_type_embeddings=_Table(1, dim), LayerNorm=_LN(dim))
        self.encoder = _Holder(layer=nn.ModuleList([
            _Holder(attention=_Holder(self=_Holder(query=_dense(dim, dim), key=_dense(dim, dim), value=_dense(dim, dim)),
                                      output=_Holder(dense=_dense(dim, dim), LayerNorm=_LN(dim))),
                    intermediate=_Holder(dense=_dense(dim, intermediate)),
                    output=_Holder(dense=_dense(intermediate, dim), LayerNorm=_LN(dim)))
            for _ in range(layers)]))
        self.pooler = _Holder(dense=_dense(dim, dim))
        self.heads = heads


def bert_embeddings(emb: BertParams, x, p_hidden, seed_fn):
    """BertEmbeddings.forward with inputs_embeds: x + token_type[0] + position[0..S) -> LayerNorm -> dropout."""
    B, S, D = x.shape
    pos_ids = torch.arange(S, dtype=torch.int64, device=x.device).repeat(B)
    e = emb.embeddings
    x2 = ops.gather_add(x.reshape(B * S, D), torch.zeros(B * S, dtype=torch.int64, device=x.device), None, e.token_type_embeddings.weight)
    x2 = ops.gather_add(x2, pos_ids, None, e.position_embeddings.weight)
    h = ops.layernorm(x2, e.LayerNorm.weight, e.LayerNorm.bias, LN_EPS)
    return ops.dropout(h, p_hidden, seed_fn() if p_hidden else 0).view(B, S, D)


def bert_output(block, hidden, residual, p_hidden, seed_fn):
    """BertSelfOutput / BertOutput: LayerNorm(dropout(dense(hidden)) + residual)."""
    y = ops.linear(hidden, block.dense.weight, block.dense.bias, drop_p=p_hidden, seed=seed_fn() if p_hidden else 0)
    return ops.layernorm(y, block.LayerNorm.weight, block.LayerNorm.bias, LN_EPS, res=residual)


class TransformerOperatorConfig(AttentionOperatorConfig):
    def __init__(self, num_hidden_layers: int = 3, hidden_dropout_prob: float = HIDDEN_DROPOUT, **kwargs):
        super().__init__(**kwargs)
        self.num_hidden_layers = num_hidden_layers
        # BertConfig.hidden_dropout_prob: the reference cannot set it (always 0.1 in training); an extra key here, ignored by the reference's config class
        self.hidden_dropout_prob = hidden_dropout_prob


class TransformerOperator(BaseOperator):
    config_class = TransformerOperatorConfig
    inputer_class = ConcatInputer
    config: TransformerOperatorConfig

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        c = self.config
        self.transformer = BertParams(c.input_dim, c.num_attention_heads, c.num_hidden_layers, c.hidden_size * 4)
        self.linear = _Affine(c.input_dim, c.hidden_size)
        self.additive_attention = AdditiveAttention(embed_dim=c.hidden_size, hidden_size=c.hidden_size)

    def forward(self, embeddings, mask=None, **kwargs):
        mask = mask.to(Env.device)
        c = self.config
        p_hid = c.hidden_dropout_prob if self.training else 0.0
        p_att = c.attention_dropout if self.training else 0.0
        x = bert_embeddings(self.transformer, embeddings, p_hid, self._next_seed)
        for layer in self.transformer.encoder.layer:
            sa = layer.attention.self
            w = torch.cat([sa.query.weight, sa.key.weight, sa.value.weight], dim=0)      # [3D, D]: one contraction for Q, K, V (layout only)
            b = torch.cat([sa.query.bias, sa.key.bias, sa.value.bias], dim=0)
            qkv = ops.linear(x, w, b)
            ctx = ops.mha_core(qkv, mask, c.num_attention_heads, drop_p=p_att, seed=self._next_seed() if p_att else 0)
            att = bert_output(layer.attention.output, ctx, x, p_hid, self._next_seed)
            inter = ops.gelu(ops.linear(att, layer.intermediate.dense.weight, layer.intermediate.dense.bias))
            x = bert_output(layer.output, inter, att, p_hid, self._next_seed)
        out = ops.linear(x, self.linear.weight, self.linear.bias)
        return self.additive_attention(out, mask)
