"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.

A functional fp32 (optionally fp64) CPU restatement of the Legommenders hot path named by
BASELINE.json:north_star / SURVEY.md §8a.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` leg may import this module; the product package
`legommenders_b200` never does (it fails loudly if its CUDA library is missing).

Parity pin: the reference ships no tests or golden vectors of its own, so this
oracle is pinned against outputs of the LIVE reference modules run in the build container
(`tests/golden/make_golden.py` -> committed `tests/golden/*.npz`, checked by
`tests/test_oracle_golden.py`) and, when the reference tree is present, against the reference
directly (`tests/test_oracle_vs_reference.py`).

Every function cites the reference file:line it restates (paths relative to the reference root).
Floating point is torch on CPU — the reference's arithmetic *is* PyTorch (requirements.txt:1), so
the same library is the faithful substrate; index/mask work is integer-exact.

State dicts use the reference's parameter names, e.g.
  embedding_vocab_table.glove.embedding.weight / .linear.weight / .linear.bias
  item_op.multi_head_attention.in_proj_weight ... item_op.additive_attention.encoder.{0,2}.weight
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

DROPOUT = 0.0                   # parity runs keep 0; bench's CPU baseline sets the reference's 0.1 (training mode)
UNSET = -1                      # loader/env.py:7
EPS32 = 1.1920928955078125e-07  # torch.finfo(float32).eps, model/common/attention.py:36
PAD, CLS, SEP = 0, 1, 2         # model/inputer/concat_inputer.py:27-30
SPECIAL_VOCAB = '__cat_inputer_special_ids'


# --------------------------------------------------------------------------------------------
# a1 / a2 : per-item token layouts (integer exact)
# --------------------------------------------------------------------------------------------
def concat_layout(sample: dict, inputs: Sequence[str], max_lens: Dict[str, Optional[int]],
                  use_cls_token: bool, use_sep_token: bool):
    """model/inputer/concat_inputer.py:43-87 — left-packed [CLS?, col0…, SEP?, col1…, SEP?]; one int64
    vector of length S per source column (unset = -1); the special column is PAD(0) after the packed
    length; attention_mask = 1^pos 0^(S-pos)."""
    content = sum((max_lens[c] or 1) for c in inputs)
    S = content + int(use_cls_token) + int(use_sep_token) * len(inputs)
    ids = OrderedDict()
    special = np.full(S, UNSET, dtype=np.int64)
    pos = 0
    if use_cls_token:
        special[pos] = CLS
        pos += 1
    for col in inputs:
        v = sample[col]
        if not isinstance(v, (list, tuple, np.ndarray)):
            v = [v]
        v = np.asarray(v, dtype=np.int64)
        row = np.full(S, UNSET, dtype=np.int64)
        row[pos:pos + len(v)] = v
        pos += len(v)
        ids[col] = row
        if use_sep_token:
            special[pos] = SEP
            pos += 1
    if use_cls_token or use_sep_token:
        special[pos:] = PAD
        ids[SPECIAL_VOCAB] = special
    mask = np.zeros(S, dtype=np.int64)
    mask[:pos] = 1
    return dict(input_ids=ids, attention_mask=mask)


def simple_layout(sample: dict, inputs: Sequence[str], max_lens: Dict[str, Optional[int]]):
    """model/inputer/simple_inputer.py:17-38 — per column: ids right-padded with -1, mask 1^len 0^pad."""
    ids, mask = OrderedDict(), OrderedDict()
    for col in inputs:
        v = sample[col]
        L = max_lens[col]
        if not L:
            v, L = [v], 1
        v = list(v)
        ids[col] = np.asarray(v + [UNSET] * (L - len(v)), dtype=np.int64)
        mask[col] = np.asarray([1] * len(v) + [0] * (L - len(v)), dtype=np.int64)
    return dict(input_ids=ids, attention_mask=mask)


# --------------------------------------------------------------------------------------------
# a17 : Resampler semantics (integer exact; the RNG draws are supplied by the caller)
# --------------------------------------------------------------------------------------------
def pad_history(history: Sequence[int], max_click_num: int):
    """loader/resampler.py:209-218 — history right-padded with item id 0, mask 1^len 0^pad."""
    n = len(history)
    ids = np.asarray(list(history) + [0] * (max_click_num - n), dtype=np.int64)
    mask = np.asarray([1] * n + [0] * (max_click_num - n), dtype=np.int64)
    return ids, mask


def candidates(pos_item: int, sampled_true_negs: Sequence[int], random_negs: Sequence[int]):
    """loader/resampler.py:159-173 — candidate order [pos, sampled true negs…, uniform random ids…]."""
    return np.asarray([pos_item] + list(sampled_true_negs) + list(random_negs), dtype=np.int64)


# --------------------------------------------------------------------------------------------
# a5 : EmbeddingHub tables
# --------------------------------------------------------------------------------------------
def table_lookup(state: dict, vocab: str, ids: torch.Tensor, prefix='embedding_vocab_table.') -> torch.Tensor:
    """loader/embedding_hub.py:378-385 + Transformation.forward :95-96 (dropout = identity: eval / p=0).
    A vocab with `.embedding.weight` is a Transformation (Embedding -> Linear); otherwise a plain nn.Embedding."""
    k = prefix + vocab
    if k + '.embedding.weight' in state:
        e = F.embedding(ids, state[k + '.embedding.weight'])
        return F.dropout(F.linear(e, state[k + '.linear.weight'], state[k + '.linear.bias']), DROPOUT, training=DROPOUT > 0)
    return F.embedding(ids, state[k + '.weight'])


# --------------------------------------------------------------------------------------------
# a3 / a4 : inputer.get_embeddings
# --------------------------------------------------------------------------------------------
def concat_embeddings(state: dict, input_ids: Dict[str, torch.Tensor], col_vocab: Dict[str, str]) -> torch.Tensor:
    """model/inputer/concat_inputer.py:92-114 — per column mask=(ids>-1); ids*=mask; emb=table(ids);
    emb*=mask; sum over columns.  (Does not mutate the caller's ids.)"""
    out = None
    for col, ids in input_ids.items():
        vocab = col if col == SPECIAL_VOCAB else col_vocab[col]
        m = (ids > UNSET).long()
        e = table_lookup(state, vocab, ids * m) * m.unsqueeze(-1)
        out = e if out is None else out + e
    return out


def simple_embeddings(state: dict, input_ids: Dict[str, torch.Tensor], attention_mask: Dict[str, torch.Tensor],
                      col_vocab: Dict[str, str]) -> "OrderedDict[str, torch.Tensor]":
    """model/inputer/simple_inputer.py:43-66 — per column ids*=mask; emb=table(ids); emb*=mask; no sum."""
    out = OrderedDict()
    for col, ids in input_ids.items():
        m = attention_mask[col]
        out[col] = table_lookup(state, col_vocab[col], ids * m) * m.unsqueeze(-1)
    return out


# --------------------------------------------------------------------------------------------
# a6 : AdditiveAttention
# --------------------------------------------------------------------------------------------
def additive_attention(x: torch.Tensor, mask: Optional[torch.Tensor], w1, b1, w2) -> torch.Tensor:
    """model/common/attention.py:23-38 — a=exp(w2·tanh(W1x+b1))·mask; α=a/(Σa+eps32); out=Σαx.
    No max-subtraction; an all-zero mask gives an all-zero output."""
    s = F.linear(torch.tanh(F.linear(x, w1, b1)), w2).squeeze(-1)
    a = torch.exp(s)
    if mask is not None:
        a = a * mask
    alpha = a / (a.sum(dim=-1, keepdim=True) + EPS32)
    return (x * alpha.unsqueeze(-1)).sum(dim=1)


def _additive(state, prefix, x, mask):
    return additive_attention(x, mask, state[prefix + 'additive_attention.encoder.0.weight'],
                              state[prefix + 'additive_attention.encoder.0.bias'],
                              state[prefix + 'additive_attention.encoder.2.weight'])


# --------------------------------------------------------------------------------------------
# a7 : AttentionOperator (nn.MultiheadAttention semantics)
# --------------------------------------------------------------------------------------------
def multi_head_self_attention(x, mask, in_w, in_b, out_w, out_b, heads: int) -> torch.Tensor:
    """torch.nn.MultiheadAttention(batch_first) self-attention as called at
    model/operators/attention_operator.py:49-55: qkv = x·in_wᵀ+in_b; q *= dh^-0.5; logits=q·kᵀ with
    -inf on padded keys; softmax over keys; ctx=probs·v; out = ctx·out_wᵀ+out_b.  Attention dropout =
    identity (eval / p=0)."""
    N, S, D = x.shape
    dh = D // heads
    qkv = F.linear(x, in_w, in_b)
    q, k, v = qkv.split(D, dim=-1)
    q = q.view(N, S, heads, dh).transpose(1, 2) * (dh ** -0.5)
    k = k.view(N, S, heads, dh).transpose(1, 2)
    v = v.view(N, S, heads, dh).transpose(1, 2)
    logits = q @ k.transpose(-1, -2)                                   # [N,h,S,S]
    key_pad = (1 - mask).bool()                                         # attention_operator.py:53
    logits = logits.masked_fill(key_pad[:, None, None, :], float('-inf'))
    probs = F.dropout(torch.softmax(logits, dim=-1), DROPOUT, training=DROPOUT > 0)
    ctx = (probs @ v).transpose(1, 2).reshape(N, S, D)
    return F.linear(ctx, out_w, out_b)


def attention_operator(state: dict, prefix: str, x, mask, heads: int) -> torch.Tensor:
    """model/operators/attention_operator.py:46-59."""
    p = prefix + 'multi_head_attention.'
    o = multi_head_self_attention(x, mask, state[p + 'in_proj_weight'], state[p + 'in_proj_bias'],
                                  state[p + 'out_proj.weight'], state[p + 'out_proj.bias'], heads)
    lin = F.linear(o, state[prefix + 'linear.weight'], state[prefix + 'linear.bias'])
    return _additive(state, prefix, lin, mask)


# --------------------------------------------------------------------------------------------
# a8 : CNNOperator
# --------------------------------------------------------------------------------------------
def cnn_operator(state: dict, prefix: str, embeddings: "OrderedDict[str, torch.Tensor]",
                 mask: Dict[str, torch.Tensor]) -> torch.Tensor:
    """model/operators/cnn_operator.py:48-67 — columns longer than 1: Conv1d(k,'same') -> ReLU -> *mask
    (dropout = identity); length-1 columns: Linear; concat along sequence; additive attention."""
    outs, masks = [], []
    for col, e in embeddings.items():
        if e.shape[1] > 1:
            w = state[prefix + 'cnn.weight']
            pad = (w.shape[-1] - 1) // 2
            assert w.shape[-1] % 2 == 1
            y = F.conv1d(e.permute(0, 2, 1), w, state[prefix + 'cnn.bias'], padding=pad).permute(0, 2, 1)
            y = F.dropout(torch.relu(y) * mask[col].unsqueeze(-1), DROPOUT, training=DROPOUT > 0)
        else:
            y = F.linear(e, state[prefix + 'linear.weight'], state[prefix + 'linear.bias'])
        outs.append(y)
        masks.append(mask[col])
    return _additive(state, prefix, torch.cat(outs, dim=1), torch.cat(masks, dim=1))


# --------------------------------------------------------------------------------------------
# a9 / a10 : AdaOperator, PoolingOperator
# --------------------------------------------------------------------------------------------
def cnn_cat_operator(state: dict, prefix: str, embeddings: "OrderedDict[str, torch.Tensor]", mask: "OrderedDict[str, torch.Tensor]") -> torch.Tensor:
    """model/operators/cnn_cat_operator.py:23-38 (dropout 0): per column conv('same') -> ReLU -> mask -> additive attention, a single-token
    column is its embedding; the per-column vectors are concatenated on the feature axis."""
    outs = []
    for col, e in embeddings.items():
        if e.shape[1] > 1:
            h = F.conv1d(e.permute(0, 2, 1), state[prefix + 'cnn.weight'], state[prefix + 'cnn.bias'], padding='same').permute(0, 2, 1)
            h = F.relu(h) * mask[col].unsqueeze(-1)
            outs.append(_additive(state, prefix, h, mask[col]))
        else:
            outs.append(e.squeeze(1))
    return torch.cat(outs, dim=-1)


def gru_operator(state: dict, prefix: str, x: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """model/operators/gru_operator.py:40-52: nn.GRU(1 layer) over the first `len` steps of every sequence (pack_padded_sequence), last
    hidden state, Linear.  Gate order (r, z, n); n = tanh(W_in x + b_in + r * (W_hn h + b_hn))."""
    w_ih, w_hh = state[prefix + 'gru.weight_ih_l0'], state[prefix + 'gru.weight_hh_l0']
    b_ih, b_hh = state[prefix + 'gru.bias_ih_l0'], state[prefix + 'gru.bias_hh_l0']
    B, S, _ = x.shape
    H = w_hh.shape[1]
    lengths = mask.sum(dim=1)
    h = x.new_zeros((B, H))
    gi_all = x @ w_ih.t() + b_ih
    for t in range(S):
        gh = h @ w_hh.t() + b_hh
        gi = gi_all[:, t]
        r = torch.sigmoid(gi[:, :H] + gh[:, :H])
        z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
        hn = (1 - z) * n + z * h
        h = torch.where((lengths > t).unsqueeze(1), hn, h)       # sequences that have ended keep their last state
    return h @ state[prefix + 'linear.weight'].t() + state[prefix + 'linear.bias']


def _layer_norm(x, w, b, eps=1e-12):
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def _gelu(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def bert_encoder(state: dict, prefix: str, x: torch.Tensor, mask: torch.Tensor, heads: int, layers: int) -> torch.Tensor:
    """transformers BertModel.forward(inputs_embeds=x, attention_mask=mask).last_hidden_state with every dropout at 0
    (modeling_bert.py: BertEmbeddings, BertSelfAttention, BertSelfOutput, BertIntermediate, BertOutput), restated:
    h = LN(x + token_type[0] + position[0..S));  per layer: a = softmax(q k^T / sqrt(dh) + (1-mask)*(-inf)) v;
    h1 = LN(dense(a) + h);  h = LN(dense(gelu(dense(h1))) + h1)."""
    B, S, D = x.shape
    dh = D // heads
    p = prefix
    h = x + state[p + 'embeddings.token_type_embeddings.weight'][0] + state[p + 'embeddings.position_embeddings.weight'][:S]
    h = _layer_norm(h, state[p + 'embeddings.LayerNorm.weight'], state[p + 'embeddings.LayerNorm.bias'])
    neg = torch.zeros(mask.shape, dtype=x.dtype).masked_fill(mask == 0, float('-inf'))[:, None, None, :]
    for i in range(layers):
        lp = f'{p}encoder.layer.{i}.'
        def lin(name, t):
            return t @ state[lp + name + '.weight'].t() + state[lp + name + '.bias']
        q = lin('attention.self.query', h).view(B, S, heads, dh).transpose(1, 2)
        k = lin('attention.self.key', h).view(B, S, heads, dh).transpose(1, 2)
        v = lin('attention.self.value', h).view(B, S, heads, dh).transpose(1, 2)
        att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh) + neg, dim=-1) @ v
        att = att.transpose(1, 2).reshape(B, S, D)
        h1 = _layer_norm(lin('attention.output.dense', att) + h, state[lp + 'attention.output.LayerNorm.weight'], state[lp + 'attention.output.LayerNorm.bias'])
        ff = lin('output.dense', _gelu(lin('intermediate.dense', h1)))
        h = _layer_norm(ff + h1, state[lp + 'output.LayerNorm.weight'], state[lp + 'output.LayerNorm.bias'])
    return h


def transformer_operator(state: dict, prefix: str, x, mask, heads: int, layers: int) -> torch.Tensor:
    """model/operators/transformer_operator.py:46-61: BertModel -> Linear -> AdditiveAttention."""
    h = bert_encoder(state, prefix + 'transformer.', x, mask, heads, layers)
    out = h @ state[prefix + 'linear.weight'].t() + state[prefix + 'linear.bias']
    return _additive(state, prefix, out, mask)


def fastformer_operator(state: dict, prefix: str, x: torch.Tensor, mask: torch.Tensor, heads: int, layers: int) -> torch.Tensor:
    """model/operators/fastformer_operator.py:41-49 + model/common/fastformer.py:62-226 with every dropout at 0: position embeddings + LayerNorm,
    `layers` x (FastSelfAttention -> BertSelfOutput -> BertIntermediate -> BertOutput), AttentionPooling, Linear.  The mask enters the two softmaxes
    additively as (1 - mask) * (-10000)."""
    B, S, D = x.shape
    dh = D // heads
    p = prefix + 'fastformer.'
    m = mask.to(x.dtype)
    ext = ((1.0 - m) * -10000.0).unsqueeze(1)                                   # [B, 1, S]

    def lin(name, t):
        return t @ state[name + '.weight'].t() + state[name + '.bias']

    h = _layer_norm(x + state[p + 'position_embeddings.weight'][:S], state[p + 'LayerNorm.weight'], state[p + 'LayerNorm.bias'])
    for i in range(layers):
        lp = f'{p}encoders.{i}.'
        sa = lp + 'attention.self.'
        mq, mk = lin(sa + 'query', h), lin(sa + 'key', h)
        qw = torch.softmax(lin(sa + 'query_att', mq).transpose(1, 2) / dh ** 0.5 + ext, dim=-1).unsqueeze(2)      # [B, heads, 1, S]
        ql = mq.view(B, S, heads, dh).permute(0, 2, 1, 3)
        pooled_q = (qw @ ql).transpose(1, 2).reshape(B, 1, D)
        mixed = mk * pooled_q
        kw = torch.softmax((lin(sa + 'key_att', mixed) / dh ** 0.5).transpose(1, 2) + ext, dim=-1).unsqueeze(2)
        kl = mixed.view(B, S, heads, dh).permute(0, 2, 1, 3)
        pooled_k = kw @ kl                                                        # [B, heads, 1, dh]
        wv = (pooled_k * ql).transpose(1, 2).reshape(B, S, D)
        so = lin(sa + 'transform', wv) + mq
        att = _layer_norm(lin(lp + 'attention.output.dense', so) + h, state[lp + 'attention.output.LayerNorm.weight'], state[lp + 'attention.output.LayerNorm.bias'])
        ff = lin(lp + 'output.dense', _gelu(lin(lp + 'intermediate.dense', att)))
        h = _layer_norm(ff + att, state[lp + 'output.LayerNorm.weight'], state[lp + 'output.LayerNorm.bias'])
    e = torch.tanh(lin(p + 'poolers.0.att_fc1', h))
    alpha = torch.exp(lin(p + 'poolers.0.att_fc2', e)) * m.unsqueeze(2)
    alpha = alpha / (alpha.sum(dim=1, keepdim=True) + 1e-8)
    pooled = (h * alpha).sum(dim=1)
    return pooled @ state[prefix + 'linear.weight'].t() + state[prefix + 'linear.bias']


def poly_attention_operator(state: dict, prefix: str, x: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """model/operators/poly_attention_operator.py:45-58: weights = softmax_s(masked_fill(tanh(Linear(x))·codesᵀ, ~mask, 1e-30)) — the fill value
    is a logit of ~0, not -inf; out [B, codes, D] = weights · x."""
    proj = torch.tanh(x @ state[prefix + 'linear.weight'].t())
    w = (proj @ state[prefix + 'context_codes'].t()).permute(0, 2, 1)
    w = torch.where(mask.unsqueeze(1) > 0, w, torch.full_like(w, 1e-30))
    return torch.softmax(w, dim=2) @ x


def miner_predictor(state: dict, prefix: str, user: torch.Tensor, items: torch.Tensor, score_type: str = 'weighted') -> torch.Tensor:
    """model/predictors/miner_predictor.py:30-62: scores = items·userᵀ [B, K+1, codes]; weighted: softmax_codes(items · gelu(Linear(user))ᵀ) * scores, summed."""
    scores = items @ user.permute(0, 2, 1)
    if score_type == 'max':
        return scores.max(dim=2)[0]
    if score_type == 'mean':
        return scores.mean(dim=2)
    proj = _gelu(user @ state[prefix + 'target_aware_attention.linear.weight'].t())
    w = torch.softmax(items @ proj.permute(0, 2, 1), dim=2)
    return (w * scores).sum(dim=2)


def ada_operator(state: dict, prefix: str, x, mask) -> torch.Tensor:
    """model/operators/ada_operator.py:31-38."""
    return _additive(state, prefix, x, mask)


def pooling_operator(embeddings, mask, flatten=False, max_pooling=False) -> torch.Tensor:
    """model/operators/pooling_operator.py:31-61."""
    if isinstance(embeddings, torch.Tensor):
        embeddings, mask = dict(temp=embeddings), dict(temp=mask)
    pooled = OrderedDict()
    for col, e in embeddings.items():
        m = mask[col]
        e = e * m.unsqueeze(-1)
        if max_pooling:
            pooled[col] = e.max(dim=1)[0]
        else:
            pooled[col] = e.sum(dim=1) / (m.sum(dim=1).unsqueeze(-1) + 1e-8)
    if flatten:
        return torch.cat(list(pooled.values()), dim=-1)
    stack = torch.stack(list(pooled.values()), dim=1)
    return stack.max(dim=1)[0] if max_pooling else stack.mean(dim=1)


# --------------------------------------------------------------------------------------------
# a13 / a14 : scoring + loss
# --------------------------------------------------------------------------------------------
def dot_scores(user: torch.Tensor, items: torch.Tensor) -> torch.Tensor:
    """model/legommender.py:268-283 + model/predictors/dot_predictor.py:7-10 — z[b,c] = Σ_d u[b,d]·v[b,c,d]."""
    return (user.unsqueeze(1) * items).sum(dim=-1)


def ce_loss(scores: torch.Tensor) -> torch.Tensor:
    """model/legommender.py:114-118, 254, 263 — CrossEntropyLoss(scores, label 0), mean."""
    return F.cross_entropy(scores, torch.zeros(scores.shape[0], dtype=torch.long))


def bce_loss(scores: torch.Tensor, click: torch.Tensor) -> torch.Tensor:
    """model/legommender.py:256-257, 290 — BCEWithLogitsLoss(z, click.float()), mean."""
    return F.binary_cross_entropy_with_logits(scores, click.float())


# --------------------------------------------------------------------------------------------
# a11 / a12 : model-level item / user content and the full forward
# --------------------------------------------------------------------------------------------
class ModelSpec:
    """What the oracle needs to know about a model (the yaml-level configuration)."""

    def __init__(self, kind: str, heads: int = 8, col_vocab: Optional[Dict[str, str]] = None,
                 use_neg_sampling: bool = True, item_vocab: str = 'item_id', layers: int = 0, score_type: str = 'weighted'):
        assert kind in ('nrms', 'naml', 'llmid', 'pool', 'lstur', 'miner', 'fastformer')
        self.layers, self.score_type = layers, score_type
        self.kind, self.heads = kind, heads
        self.col_vocab = col_vocab or {}
        self.use_neg_sampling = use_neg_sampling
        self.item_vocab = item_vocab


def _flat(t: torch.Tensor) -> torch.Tensor:
    return t.reshape(-1, t.shape[-1])   # utils/shaper.py:41 (Reshaper.custom_worker)


def item_content(state: dict, spec: ModelSpec, tree: dict) -> torch.Tensor:
    """model/legommender.py:138-192 without the cache short-circuit: flatten [B,C,S]->[B·C,S], mask, gather,
    item_op, reshape back to [B,C,D]."""
    if spec.kind == 'nrms':
        ids = {c: _flat(v) for c, v in tree['input_ids'].items()}
        B = next(iter(tree['input_ids'].values())).shape[0]
        mask = _flat(tree['attention_mask'])
        x = concat_embeddings(state, ids, spec.col_vocab)
        r = attention_operator(state, 'item_op.', x, mask, spec.heads)
    elif spec.kind == 'naml':
        ids = OrderedDict((c, _flat(v)) for c, v in tree['input_ids'].items())
        B = next(iter(tree['input_ids'].values())).shape[0]
        am = OrderedDict((c, _flat(v)) for c, v in tree['attention_mask'].items())
        x = simple_embeddings(state, ids, am, spec.col_vocab)
        r = cnn_operator(state, 'item_op.', x, am)
    elif spec.kind == 'fastformer':
        ids = {c: _flat(v) for c, v in tree['input_ids'].items()}
        B = next(iter(tree['input_ids'].values())).shape[0]
        mask = _flat(tree['attention_mask'])
        r = fastformer_operator(state, 'item_op.', concat_embeddings(state, ids, spec.col_vocab), mask, spec.heads, spec.layers)
    elif spec.kind == 'miner':
        ids = {c: _flat(v) for c, v in tree['input_ids'].items()}
        B = next(iter(tree['input_ids'].values())).shape[0]
        mask = _flat(tree['attention_mask'])
        r = transformer_operator(state, 'item_op.', concat_embeddings(state, ids, spec.col_vocab), mask, spec.heads, spec.layers)
    elif spec.kind == 'lstur':
        ids = OrderedDict((c, _flat(v)) for c, v in tree['input_ids'].items())
        B = next(iter(tree['input_ids'].values())).shape[0]
        am = OrderedDict((c, _flat(v)) for c, v in tree['attention_mask'].items())
        r = cnn_cat_operator(state, 'item_op.', simple_embeddings(state, ids, am, spec.col_vocab), am)
    elif spec.kind == 'pool':
        ids = OrderedDict((c, _flat(v)) for c, v in tree['input_ids'].items())
        B = next(iter(tree['input_ids'].values())).shape[0]
        am = OrderedDict((c, _flat(v)) for c, v in tree['attention_mask'].items())
        r = pooling_operator(simple_embeddings(state, ids, am, spec.col_vocab), am)
    else:
        raise ValueError(spec.kind)
    return r.view(B, -1, r.shape[-1])


def user_content(state: dict, spec: ModelSpec, batch: dict, clicks: Optional[torch.Tensor] = None) -> torch.Tensor:
    """model/legommender.py:197-214."""
    if clicks is None:
        if spec.kind == 'llmid':
            # use_item_content=False: clicks = user_op.inputer.get_embeddings(batch[history]) (ConcatInputer, no specials)
            clicks = concat_embeddings(state, batch['history']['input_ids'], {'history': spec.item_vocab})
        else:
            clicks = item_content(state, spec, batch['history'])
    m = batch['__clicks_mask__']
    if spec.kind == 'nrms':
        return attention_operator(state, 'user_op.', clicks, m, spec.heads)
    if spec.kind == 'lstur':
        return gru_operator(state, 'user_op.', clicks, m)
    if spec.kind == 'miner':
        return poly_attention_operator(state, 'user_op.', clicks, m)
    if spec.kind == 'fastformer':
        return fastformer_operator(state, 'user_op.', clicks, m, spec.heads, spec.layers)
    return ada_operator(state, 'user_op.', clicks, m)


def forward(state: dict, spec: ModelSpec, batch: dict, return_scores: bool = False, want: Optional[dict] = None):
    """model/legommender.py:219-263."""
    if spec.kind == 'llmid':
        iid = batch['item_id']
        if iid.dim() == 1:
            iid = iid.unsqueeze(1)
        items = table_lookup(state, spec.item_vocab, iid)      # legommender.py:239-246
    else:
        items = item_content(state, spec, batch['item_id'])
    user = user_content(state, spec, batch)
    if want is not None:
        want['items'], want['user'] = items, user
    if spec.use_neg_sampling:
        scores = miner_predictor(state, 'predictor.', user, items, spec.score_type) if spec.kind == 'miner' else dot_scores(user, items)
        if return_scores:
            return scores
        return ce_loss(scores)
    scores = (user * items.squeeze(1)).sum(-1)                 # legommender.py:285-290
    if return_scores:
        return scores
    return bce_loss(scores, batch['click'])


# --------------------------------------------------------------------------------------------
# a15 / a16 : caches and cached evaluation
# --------------------------------------------------------------------------------------------
def build_item_cache(state: dict, spec: ModelSpec, item_trees: List[dict], page_size: int = 512) -> torch.Tensor:
    """loader/cacher/item_cacher.py:75-97 + loader/pager/{base,fast_item}_pager.py — per item get_embeddings/get_mask,
    stack per page of `page_size`, item_op, positional slice assignment, detach."""
    out = []
    with torch.no_grad():
        for s in range(0, len(item_trees), page_size):
            page = item_trees[s:s + page_size]
            if spec.kind in ('nrms', 'miner', 'fastformer'):          # ConcatInputer: one mask for the concatenated sequence
                tree = dict(input_ids={c: torch.stack([torch.as_tensor(p['input_ids'][c]) for p in page]).unsqueeze(1)
                                       for c in page[0]['input_ids']},
                            attention_mask=torch.stack([torch.as_tensor(p['attention_mask']) for p in page]).unsqueeze(1))
            else:
                tree = dict(input_ids=OrderedDict((c, torch.stack([torch.as_tensor(p['input_ids'][c]) for p in page]).unsqueeze(1))
                                                  for c in page[0]['input_ids']),
                            attention_mask=OrderedDict((c, torch.stack([torch.as_tensor(p['attention_mask'][c]) for p in page]).unsqueeze(1))
                                                       for c in page[0]['attention_mask']))
            out.append(item_content(state, spec, tree).squeeze(1))
    return torch.cat(out, dim=0)


def build_user_cache(state: dict, spec: ModelSpec, item_repr: Optional[torch.Tensor], histories: torch.Tensor,
                     clicks_mask: torch.Tensor, page_size: int = 512) -> torch.Tensor:
    """loader/cacher/user_cacher.py:84-97 + fast_user_pager.py + legommender.py:153-157 — clicks = item.repr[history]
    (history right-padded with id 0), user_op with __clicks_mask__, positional slices."""
    out = []
    with torch.no_grad():
        for s in range(0, histories.shape[0], page_size):
            h, m = histories[s:s + page_size], clicks_mask[s:s + page_size]
            if spec.kind == 'llmid':
                hm = (h > UNSET).long()
                clicks = table_lookup(state, spec.item_vocab, h * hm) * hm.unsqueeze(-1)
            else:
                clicks = item_repr[h.reshape(-1)].reshape(*h.shape, -1)
            out.append(user_content(state, spec, {'__clicks_mask__': m}, clicks=clicks))
    return torch.cat(out, dim=0)


def cached_scores(user_repr: torch.Tensor, item_repr: torch.Tensor, user_ids: torch.Tensor,
                  item_ids: torch.Tensor) -> torch.Tensor:
    """model/legommender.py:153-157, 202-203, 268-290 — score[r] = <U[uid[r]], I[iid[r]]>."""
    return (user_repr[user_ids] * item_repr[item_ids]).sum(-1)


# --------------------------------------------------------------------------------------------
# a18 : group metrics (numpy restatement of sklearn's roc_auc_score / ndcg_score and the custom MRR)
# --------------------------------------------------------------------------------------------
def _avg_ranks(x: np.ndarray) -> np.ndarray:
    order = np.argsort(x, kind='mergesort')
    xs = x[order]
    ranks = np.empty(len(x), dtype=np.float64)
    i = 0
    while i < len(xs):
        j = i
        while j + 1 < len(xs) and xs[j + 1] == xs[i]:
            j += 1
        ranks[order[i:j + 1]] = 0.5 * (i + j) + 1.0
        i = j + 1
    return ranks


def auc(scores, labels) -> float:
    """utils/metrics.py:88-108 (sklearn roc_auc_score, binary): Mann-Whitney U with tie-averaged ranks."""
    s = np.asarray(scores, dtype=np.float64)
    y = np.asarray(labels)
    npos = int((y == 1).sum())
    nneg = len(y) - npos
    if npos == 0 or nneg == 0:
        return float('nan')
    r = _avg_ranks(s)
    return float((r[y == 1].sum() - npos * (npos + 1) / 2.0) / (npos * nneg))


def mrr(scores, labels) -> float:
    """utils/metrics.py:144-160 — stable descending sort (python `sorted(reverse=True)` keeps the original
    order of ties), Σ y_i/(i+1) / Σ y."""
    s = np.asarray(scores, dtype=np.float64)
    y = np.asarray(labels, dtype=np.float64)
    order = np.argsort(-s, kind='mergesort')
    yt = y[order]
    return float((yt / np.arange(1, len(yt) + 1)).sum() / yt.sum())


def ndcg(scores, labels, k: int) -> float:
    """utils/metrics.py:223-235 — sklearn.metrics.ndcg_score([labels],[scores],k): linear gain, log2 discount
    truncated at k, tie-averaged DCG over score ties; ideal DCG from sorted labels; 0 when ideal is 0."""
    s = np.asarray(scores, dtype=np.float64)
    y = np.asarray(labels, dtype=np.float64)
    n = len(y)
    disc = 1.0 / np.log2(np.arange(n) + 2.0)
    if k is not None:
        disc[k:] = 0.0
    # tie-averaged DCG (sklearn _tie_averaged_dcg)
    _, inv, counts = np.unique(-s, return_inverse=True, return_counts=True)
    sums = np.zeros(len(counts))
    np.add.at(sums, inv, y)
    avg = sums / counts
    ends = np.cumsum(counts) - 1
    dcs = np.cumsum(disc)
    dsum = np.empty(len(counts))
    dsum[0] = dcs[ends[0]]
    dsum[1:] = np.diff(dcs[ends])
    dcg = float((avg * dsum).sum())
    ideal = float((np.sort(y)[::-1] * disc).sum())
    return dcg / ideal if ideal > 0 else 0.0


def metric_pool(scores, labels, groups, names=('GAUC', 'MRR', 'NDCG@1', 'NDCG@5', 'NDCG@10')) -> "OrderedDict[str, float]":
    """utils/metrics.py:313-369 — pandas groupby(groups) (sorted group keys, rows in original order), metric per
    group, mean over groups accumulated as a float32 tensor mean."""
    s = np.asarray(scores, dtype=np.float64)
    y = np.asarray(labels)
    g = np.asarray(groups)
    order = np.argsort(g, kind='mergesort')
    gs = g[order]
    bounds = np.flatnonzero(np.concatenate([[True], gs[1:] != gs[:-1], [True]]))
    out = OrderedDict()
    for name in names:
        vals = []
        for a, b in zip(bounds[:-1], bounds[1:]):
            idx = order[a:b]
            if name == 'GAUC':
                vals.append(auc(s[idx], y[idx]))
            elif name == 'MRR':
                vals.append(mrr(s[idx], y[idx]))
            elif name.startswith('NDCG@'):
                vals.append(ndcg(s[idx], y[idx], int(name.split('@')[1])))
            else:
                raise ValueError(name)
        out[name] = torch.tensor(vals, dtype=torch.float).mean().item()
    return out


# --------------------------------------------------------------------------------------------
# optimiser (part of a training step; base_lego.py:198-204 — torch.optim.Adam defaults)
# --------------------------------------------------------------------------------------------
def adam_step(p, g, m, v, step: int, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


# --------------------------------------------------------------------------------------------
# parameter shapes / default initialisation for the CPU timing port
# --------------------------------------------------------------------------------------------
def state_shapes(kind: str, D: int, A: int, E: int, n_words: int, n_cats: int, n_items: int = 0, layers: int = 0, codes: int = 0,
                 code_dim: int = 0) -> "OrderedDict[str, tuple]":
    s = OrderedDict()

    def additive(prefix):
        s[prefix + 'additive_attention.encoder.0.weight'] = (A, D)
        s[prefix + 'additive_attention.encoder.0.bias'] = (A,)
        s[prefix + 'additive_attention.encoder.2.weight'] = (1, A)

    def mha(prefix):
        s[prefix + 'multi_head_attention.in_proj_weight'] = (3 * D, D)
        s[prefix + 'multi_head_attention.in_proj_bias'] = (3 * D,)
        s[prefix + 'multi_head_attention.out_proj.weight'] = (D, D)
        s[prefix + 'multi_head_attention.out_proj.bias'] = (D,)
        s[prefix + 'linear.weight'] = (D, D)
        s[prefix + 'linear.bias'] = (D,)
        additive(prefix)

    if kind in ('nrms', 'naml', 'pool', 'lstur', 'miner', 'fastformer'):
        s['embedding_vocab_table.glove.embedding.weight'] = (n_words, E)
        s['embedding_vocab_table.glove.linear.weight'] = (D, E)
        s['embedding_vocab_table.glove.linear.bias'] = (D,)
        s['embedding_vocab_table.category.weight'] = (n_cats, D)
    if kind == 'nrms':
        s['embedding_vocab_table.' + SPECIAL_VOCAB + '.weight'] = (3, D)
        mha('item_op.')
        mha('user_op.')
    elif kind == 'naml':
        s['item_op.cnn.weight'] = (D, D, 3)
        s['item_op.cnn.bias'] = (D,)
        s['item_op.linear.weight'] = (D, D)
        s['item_op.linear.bias'] = (D,)
        additive('item_op.')
        additive('user_op.')
    elif kind == 'fastformer':
        heads = codes                                    # (the caller passes the head count through `codes`)
        for side in ('item_op.', 'user_op.'):
            f = side + 'fastformer.'
            for i in range(layers):
                lp = f'{f}encoders.{i}.'
                for n, shp in (('attention.self.query', (D, D)), ('attention.self.query_att', (heads, D)), ('attention.self.key', (D, D)),
                               ('attention.self.key_att', (heads, D)), ('attention.self.transform', (D, D)), ('attention.output.dense', (D, D)),
                               ('intermediate.dense', (4 * D, D)), ('output.dense', (D, 4 * D))):
                    s[lp + n + '.weight'] = shp
                    s[lp + n + '.bias'] = (shp[0],)
                for n in ('attention.output.LayerNorm', 'output.LayerNorm'):
                    s[lp + n + '.weight'] = (D,)
                    s[lp + n + '.bias'] = (D,)
            s[f + 'position_embeddings.weight'] = (1024, D)
            s[f + 'LayerNorm.weight'] = (D,)
            s[f + 'LayerNorm.bias'] = (D,)
            s[f + 'poolers.0.att_fc1.weight'] = (D, D)
            s[f + 'poolers.0.att_fc1.bias'] = (D,)
            s[f + 'poolers.0.att_fc2.weight'] = (1, D)
            s[f + 'poolers.0.att_fc2.bias'] = (1,)
            s[side + 'linear.weight'] = (D, D)
            s[side + 'linear.bias'] = (D,)
    elif kind == 'miner':
        s['embedding_vocab_table.' + SPECIAL_VOCAB + '.weight'] = (3, D)
        t = 'item_op.transformer.'
        s[t + 'embeddings.word_embeddings.weight'] = (1, D)
        s[t + 'embeddings.position_embeddings.weight'] = (1024, D)
        s[t + 'embeddings.token_type_embeddings.weight'] = (1, D)
        s[t + 'embeddings.LayerNorm.weight'] = (D,)
        s[t + 'embeddings.LayerNorm.bias'] = (D,)
        for i in range(layers):
            lp = f'{t}encoder.layer.{i}.'
            for n, shp in (('attention.self.query', (D, D)), ('attention.self.key', (D, D)), ('attention.self.value', (D, D)),
                           ('attention.output.dense', (D, D)), ('intermediate.dense', (4 * D, D)), ('output.dense', (D, 4 * D))):
                s[lp + n + '.weight'] = shp
                s[lp + n + '.bias'] = (shp[0],)
            for n in ('attention.output.LayerNorm', 'output.LayerNorm'):
                s[lp + n + '.weight'] = (D,)
                s[lp + n + '.bias'] = (D,)
        s[t + 'pooler.dense.weight'] = (D, D)
        s[t + 'pooler.dense.bias'] = (D,)
        s['item_op.linear.weight'] = (D, D)
        s['item_op.linear.bias'] = (D,)
        s['item_op.additive_attention.encoder.0.weight'] = (D, D)       # AdditiveAttention(hidden_size=hidden) in transformer_operator.py:40-43
        s['item_op.additive_attention.encoder.0.bias'] = (D,)
        s['item_op.additive_attention.encoder.2.weight'] = (1, D)
        s['user_op.linear.weight'] = (code_dim, D)
        s['user_op.context_codes'] = (codes, code_dim)
        s['predictor.target_aware_attention.linear.weight'] = (D, D)
    elif kind == 'lstur':
        s['item_op.cnn.weight'] = (D, D, 3)
        s['item_op.cnn.bias'] = (D,)
        s['item_op.linear.weight'] = (D, D)
        s['item_op.linear.bias'] = (D,)
        additive('item_op.')
        s['user_op.gru.weight_ih_l0'] = (3 * D, 2 * D)
        s['user_op.gru.weight_hh_l0'] = (3 * D, D)
        s['user_op.gru.bias_ih_l0'] = (3 * D,)
        s['user_op.gru.bias_hh_l0'] = (3 * D,)
        s['user_op.linear.weight'] = (2 * D, D)
        s['user_op.linear.bias'] = (2 * D,)
    elif kind == 'pool':
        additive('user_op.')
    else:
        s['embedding_vocab_table.item_id.embedding.weight'] = (n_items, E)
        s['embedding_vocab_table.item_id.linear.weight'] = (D, E)
        s['embedding_vocab_table.item_id.linear.bias'] = (D,)
        additive('user_op.')
    return s


def default_state(shapes: dict, pretrained: Dict[str, torch.Tensor], seed: int = 2023) -> Dict[str, torch.Tensor]:
    """Random parameters of the right shapes and scales (U(+-1/sqrt(fan_in))); pretrained tables are frozen."""
    g = torch.Generator().manual_seed(seed)
    state = {}
    for k, shp in shapes.items():
        if k in pretrained:
            state[k] = pretrained[k]
            continue
        fan_in = shp[1] * (shp[2] if len(shp) > 2 else 1) if len(shp) > 1 else shp[0]
        bound = 1.0 / math.sqrt(max(fan_in, 1))
        if k.endswith('.weight') and 'embedding_vocab_table' in k and 'linear' not in k:
            t = torch.randn(shp, generator=g)
        else:
            t = (torch.rand(shp, generator=g) * 2 - 1) * bound
        state[k] = t.requires_grad_(True)
    return state
