"""Wire a model the way loader/manager.py:139-153 does, from an in-memory UniTok-shaped world.

The reference's Manager needs UniTok directories + refconfig yaml on disk (out of scope);
this is the same build order over the same plugin surface, driven by the yaml-level meta/config values:
  meta.item / meta.user / meta.predictor  -> operators.get / predictors.get
  config.*                                -> LegoConfig(**config)
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import operators, predictors
from .column_map import ColumnMap
from .embedding_hub import EmbeddingHub
from .env import Env
from .lego_config import LegoConfig
from .legommender import Legommender
from .resampler import Resampler

# yaml-level model descriptions (config/model/nrms.yaml, naml.yaml + common/operators/user-*.yaml, predictors/dot.yaml)
MODEL_META = {
    'nrms': dict(item='Attention', user='Attention', predictor='Dot'),
    'naml': dict(item='CNN', user='Ada', predictor='Dot'),
    'llmid': dict(item=None, user='Ada', predictor='Dot'),          # id-based path with per-item LLM embeddings
    'pool': dict(item='Pooling', user='Ada', predictor='Dot'),
    'lstur': dict(item='CNNCat', user='GRU', predictor='Dot'),       # config/model/lstur.yaml
    'miner': dict(item='Transformer', user='PolyAttention', predictor='MINER'),   # config/model/miner.yaml
    'fastformer': dict(item='Fastformer', user='Fastformer', predictor='Dot'),   # config/model/fastformer.yaml      # masked-mean item encoder (pooling_operator.py) + Ada users
}


def model_config(kind: str, hidden: int, heads: int = 8, additive: int = 256, dropout: float = 0.1, neg_count: int = 4,
                 use_neg_sampling: bool = True) -> dict:
    if kind == 'nrms':
        return dict(use_item_content=True, hidden_size=hidden, item_hidden_size=hidden, neg_count=neg_count,
                    use_neg_sampling=use_neg_sampling,
                    item_config=dict(num_attention_heads=heads, attention_dropout=dropout, additive_hidden_size=additive,
                                     inputer_config=dict(use_cls_token=False, use_sep_token=True)),
                    user_config=dict(num_attention_heads=heads, attention_dropout=dropout, additive_hidden_size=additive,
                                     inputer_config=dict(use_cls_token=False, use_sep_token=False)))
    if kind == 'naml':
        return dict(use_item_content=True, hidden_size=hidden, item_hidden_size=hidden, neg_count=neg_count,
                    use_neg_sampling=use_neg_sampling,
                    item_config=dict(dropout=dropout, kernel_size=3, additive_hidden_size=additive),
                    user_config=dict(additive_hidden_size=additive,
                                     inputer_config=dict(use_cls_token=False, use_sep_token=False)))
    if kind == 'fastformer':
        return dict(use_item_content=True, hidden_size=hidden, item_hidden_size=hidden, neg_count=neg_count,
                    use_neg_sampling=use_neg_sampling,
                    item_config=dict(num_attention_heads=heads, num_hidden_layers=1, hidden_dropout_prob=dropout,
                                     inputer_config=dict(use_cls_token=False, use_sep_token=False)),
                    user_config=dict(num_attention_heads=heads, num_hidden_layers=1, hidden_dropout_prob=dropout,
                                     inputer_config=dict(use_cls_token=False, use_sep_token=False)))
    if kind == 'miner':
        return dict(use_item_content=True, hidden_size=hidden, item_hidden_size=hidden, neg_count=neg_count,
                    use_neg_sampling=use_neg_sampling,
                    item_config=dict(num_attention_heads=heads, attention_dropout=dropout, num_hidden_layers=2, hidden_dropout_prob=dropout,
                                     inputer_config=dict(use_cls_token=False, use_sep_token=True)),
                    user_config=dict(num_context_codes=8, context_code_dim=24,
                                     inputer_config=dict(use_cls_token=False, use_sep_token=False)),
                    predictor_config=dict(score_type='weighted'))
    if kind == 'lstur':
        return dict(use_item_content=True, hidden_size=hidden, item_hidden_size=hidden, neg_count=neg_count,
                    use_neg_sampling=use_neg_sampling,
                    item_config=dict(dropout=dropout, kernel_size=3, additive_hidden_size=additive),
                    user_config=dict(inputer_config=dict(use_cls_token=False, use_sep_token=False)))
    if kind == 'pool':
        return dict(use_item_content=True, hidden_size=hidden, item_hidden_size=hidden, neg_count=neg_count,
                    use_neg_sampling=use_neg_sampling, item_config=dict(flatten=False, max_pooling=False),
                    user_config=dict(additive_hidden_size=additive,
                                     inputer_config=dict(use_cls_token=False, use_sep_token=False)))
    if kind == 'llmid':
        return dict(use_item_content=False, hidden_size=hidden, item_hidden_size=hidden, neg_count=neg_count,
                    use_neg_sampling=use_neg_sampling,
                    user_config=dict(additive_hidden_size=additive,
                                     inputer_config=dict(use_cls_token=False, use_sep_token=False)))
    raise ValueError(kind)


def build_model(world, kind: str = 'nrms', hidden: int = 256, heads: int = 8, additive: int = 256, dropout: float = 0.1,
                neg_count: int = 4, use_neg_sampling: bool = True, llm_item_table: Optional[np.ndarray] = None,
                device_index: int = 0, build_resampler: bool = True):
    """Returns (legommender on cuda, resampler, lego_config)."""
    Env.use_cuda(device_index)
    Env.simple_dev = False
    Env.train()
    Env.set_item_cache(False)
    Env.set_user_cache(False)

    meta = MODEL_META[kind]
    cm = ColumnMap(neg_col='neg')
    item_ut, user_ut = world.item_table(), world.user_table()
    cfg = LegoConfig(**model_config(kind, hidden, heads, additive, dropout, neg_count, use_neg_sampling), item_page_size=0)
    cfg.set_component_classes(operators.get(meta['item']) if meta['item'] else None, operators.get(meta['user']),
                              predictors.get(meta['predictor']))
    item_inputs = [world.title_col, 'category']
    cfg.set_item_ut(item_ut, item_inputs)
    cfg.set_user_ut(user_ut, ['history'])
    cfg.set_column_map(cm)

    eh = EmbeddingHub(embedding_dim=cfg.item_hidden_size, transformation='auto', transformation_dropout=dropout)
    if kind != 'llmid':
        eh.load_pretrained_embedding(world.word_table, vocab_name=world.word_vocab, frozen=True)
        eh.register_ut(item_ut, item_inputs)
    else:
        eh.load_pretrained_embedding(llm_item_table, vocab_name='item_id', frozen=True)
        eh.register_vocab(item_ut.meta.features['item_id'].tokenizer.vocab)
    cfg.set_embedding_hub(eh)
    cfg.build_components()
    cfg.register_inputer_vocabs()
    model = Legommender(cfg).to(Env.device)
    resampler = Resampler(cfg) if build_resampler else None
    return model, resampler, cfg


def load_state(model: torch.nn.Module, state: dict):
    """Copy a reference-named state dict (numpy or torch values) into the model; raises on any name/shape mismatch."""
    own = model.state_dict()
    missing = set(own) - set(state)
    extra = set(state) - set(own)
    if missing or extra:
        raise KeyError(f'state-dict mismatch: missing {sorted(missing)}, unexpected {sorted(extra)}')
    with torch.no_grad():
        for k, v in own.items():
            src = torch.as_tensor(np.asarray(state[k]) if not isinstance(state[k], torch.Tensor) else state[k])
            if tuple(src.shape) != tuple(v.shape):
                raise ValueError(f'{k}: shape {tuple(src.shape)} != {tuple(v.shape)}')
            v.copy_(src.to(v.device, v.dtype))
