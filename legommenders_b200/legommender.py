"""Legommender on B200 (mirror of model/legommender.py:70-336): candidate + history item encoding, user encoding,
scoring, loss/score switch and the cache short-circuits — every tensor op is a kernel from liblegommenders_b200.so."""
from typing import List, Tuple

import torch
from torch import nn

from . import ops
from .cacher import ReprCacher
from .env import Env
from .lego_config import LegoConfig
from .packing import pack_offsets, pack_tokens


def _flatten_tree(x):
    """[B, C, S] -> [B·C, S] on every leaf (utils/shaper.py:92-104); returns (tree, B)."""
    if isinstance(x, torch.Tensor):
        return x.reshape(-1, x.shape[-1]), x.shape[0]
    out, b = type(x)(), None
    for k, v in x.items():
        out[k], b = _flatten_tree(v)
    return out, b


def _rows(x):
    if isinstance(x, torch.Tensor):
        return x.shape[0]
    return _rows(next(iter(x.values())))


def _slice_tree(x, s, e):
    if x is None:
        return None
    if isinstance(x, torch.Tensor):
        return x[s:e]
    return type(x)((k, _slice_tree(v, s, e)) for k, v in x.items())


class Legommender(nn.Module):
    def __init__(self, config: LegoConfig):
        super().__init__()
        self.config = config
        self.user_operator_class = config.user_operator_class
        self.predictor_class = config.predictor_class
        self.use_neg_sampling = config.use_neg_sampling
        self.neg_count = config.neg_count

        self.eh = config.eh
        self.embedding_vocab_table = self.eh.vocab_table        # same attribute names -> same state-dict keys
        self.embedding_feature_table = self.eh.feature_table

        self.user_hub = config.user_ut
        self.item_hub = config.item_ut
        self.cm = config.cm

        self.flatten_mode = self.user_operator_class.flatten_mode
        self.item_op = config.item_operator
        self.user_op = config.user_operator
        self.predictor = config.predictor

        Env.set_lm_cache(False)   # LM operators are out of scope on this path
        self.cacher = ReprCacher(self)
        self.cacher.activate(config.use_fast_eval)

    # -- padding-free execution --------------------------------------------------------------------------------
    packed = True    # class-level switch: False forces the reference's padded layout through the same kernels

    def _packed_items(self) -> bool:
        return (self.packed and self.item_op is not None and getattr(self.item_op, 'supports_packed', False)
                and getattr(self.item_op.inputer, 'output_single_sequence', False))

    def encode_items_packed(self, input_ids: dict, mask: torch.Tensor, item_valid=None, keep_empty=True, training=None):
        """Encode items given per-column ids [N,S] + mask [N,S] without touching pad tokens -> ([n,D], Packed)."""
        pk = pack_tokens(input_ids, mask, item_valid, keep_empty=keep_empty)
        emb = self.item_op.inputer.get_embeddings({'input_ids': pk.ids}, training=training)
        return self.item_op(emb, cu=pk.cu, max_len=pk.max_len), pk

    def _forward_packed(self, batch: dict):
        """Candidates + valid history items in ONE packed item-encoder pass; the packed history encodings are directly
        the user encoder's token rows.  -> (items [B,C,D], user [B,D])"""
        cm = self.cm
        cand, hist = batch[cm.item_col], batch[cm.history_col]
        meta = batch.get('__lk_packed__')
        if meta is None:
            B, C, S = cand['attention_mask'].shape
            H = hist['attention_mask'].shape[1]
            clicks = batch[cm.mask_col]
            ids = {c: torch.cat([cand['input_ids'][c].reshape(B * C, S), hist['input_ids'][c].reshape(B * H, S)])
                   for c in cand['input_ids']}
            mask = torch.cat([cand['attention_mask'].reshape(B * C, S), hist['attention_mask'].reshape(B * H, S)])
            valid = torch.cat([torch.ones(B * C, dtype=clicks.dtype, device=clicks.device), clicks.reshape(-1)])
            pk = pack_tokens(ids, mask, valid, keep_empty=False)
            cu_u, max_u = pack_offsets(clicks)
            meta = (pk, cu_u, max_u, B, C)
            if mask.is_cuda:          # device-resident batch: the bookkeeping cost a sync, keep it with the batch
                batch['__lk_packed__'] = meta
        pk, cu_u, max_u, B, C = meta
        emb = self.item_op.inputer.get_embeddings({'input_ids': pk.ids})
        rep = self.item_op(emb, cu=pk.cu, max_len=pk.max_len)
        items = rep[:B * C].view(B, C, -1)
        user = self.user_op(rep[B * C:], cu=cu_u, max_len=max_u)
        return items, user

    # -- item side (model/legommender.py:138-192) ----------------------------------------------------------
    def get_item_content(self, batch: dict, col: str):
        if self.cacher.item.cached:
            return ops.index_rows(self.cacher.item.repr, batch[col].to(Env.device))

        content, bsz = _flatten_tree(batch[col])
        inputer = self.item_op.inputer
        mask = inputer.get_mask(content)
        if self._packed_items():
            rep, _ = self.encode_items_packed(content['input_ids'], mask)
            return rep.view(bsz, -1, rep.shape[-1])
        emb = inputer.get_embeddings(content)

        n = _rows(emb)
        page = self.config.item_page_size or n
        if page >= n:
            rep = self.item_op(emb, mask=mask)
        else:   # paging bounds activation memory (legommender.py:174-184)
            rep = torch.cat([self.item_op(_slice_tree(emb, s, min(s + page, n)), mask=_slice_tree(mask, s, min(s + page, n)))
                             for s in range(0, n, page)], dim=0)
        return rep.view(bsz, -1, rep.shape[-1])

    # -- user side (model/legommender.py:197-214) ----------------------------------------------------------
    def get_user_content(self, batch: dict):
        if self.cacher.user.cached:
            return ops.index_rows(self.cacher.user.repr, batch[self.cm.user_col].to(Env.device))
        if self.config.use_item_content and not self.flatten_mode:
            clicks = self.get_item_content(batch, self.cm.history_col)
        else:
            clicks = self.user_op.inputer.get_embeddings(batch[self.cm.history_col])
        return self.user_op(clicks, mask=batch[self.cm.mask_col].to(Env.device))

    # -- forward (model/legommender.py:219-263) ----------------------------------------------------------------
    def forward(self, batch: dict):
        cm = self.cm
        if isinstance(batch[cm.item_col], torch.Tensor) and batch[cm.item_col].dim() == 1:
            batch[cm.item_col] = batch[cm.item_col].unsqueeze(1)

        if (self._packed_items() and getattr(self.user_op, 'supports_packed', False) and not self.flatten_mode
                and not self.cacher.item.cached and not self.cacher.user.cached and isinstance(batch[cm.item_col], dict)
                and cm.history_col in batch):
            items, user = self._forward_packed(batch)
        else:
            if self.config.use_item_content:
                items = self.get_item_content(batch, cm.item_col)
            else:
                vocab = self.config.user_ut.meta.features[cm.history_col].tokenizer.vocab.name
                items = self.eh(vocab, col_name=cm.history_col)(batch[cm.item_col].to(Env.device))
            user = self.get_user_content(batch)

        want_scores = Env.is_testing or (Env.is_evaluating and not Env.simple_dev)
        fused = getattr(self.predictor, 'fused_scoring', False)
        if self.use_neg_sampling:
            if fused and not want_scores:
                loss, _ = ops.dot_ce_loss(user, items)
                return loss
            scores = self._predict_for_neg_sampling(items, user)
            if want_scores:
                return scores
            return nn.functional.cross_entropy(scores, torch.zeros(scores.size(0), dtype=torch.long, device=Env.device))
        if fused and not want_scores:
            loss, _ = ops.dot_bce_loss(user, items.squeeze(1), batch[cm.label_col].to(Env.device).float())
            return loss
        scores = self._predict_for_ranking(items, user)
        if want_scores:
            return scores
        return nn.functional.binary_cross_entropy_with_logits(scores, batch[cm.label_col].float().to(Env.device))

    def _predict_for_neg_sampling(self, item_embeddings, user_embeddings):
        b, c, d = item_embeddings.shape
        if self.predictor.keep_input_dim or getattr(self.predictor, 'fused_scoring', False):
            return self.predictor(user_embeddings, item_embeddings)
        user = self.user_op.prepare_for_predictor(user_embeddings, c)
        return self.predictor(user, item_embeddings.reshape(-1, d)).view(b, -1)

    def _predict_for_ranking(self, item_embeddings, user_embeddings):
        return self.predictor(user_embeddings, item_embeddings.squeeze(1))

    def __str__(self):
        return self.__class__.__name__

    __repr__ = __str__

    # -- parameter groups (model/legommender.py:304-336) -----------------------------------------------------
    def get_parameters(self) -> Tuple[List[nn.Parameter], List[nn.Parameter]]:
        pretrained, other = [], []
        signals = self.item_op.get_pretrained_parameter_names() if self.item_op is not None else []
        for name, p in self.named_parameters():
            if not p.requires_grad:
                continue
            (pretrained if any(name.startswith(f'item_op.{s}') for s in signals) else other).append(p)
        return pretrained, other
