"""Import the LIVE reference (read-only tree) on CPU with two import stubs and a synthetic world.

Only usable where the reference tree exists (this container: /root/reference, or $LEGO_REF).
Used by tests/golden/make_golden.py to mint committed golden vectors and by tests that
cross-check the oracle against the reference when the tree is present.  Never imported by the
product package, bench.py or the -m gpu tests (the GPU box has no reference tree).

Recipe follows SURVEY.md Appendix B (build order mirrors loader/manager.py:139-153).
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get('LEGO_REF', '/root/reference')


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, 'model', 'operators'))


def activate():
    """Put stubs + the reference on sys.path (idempotent)."""
    if not available():
        raise RuntimeError(f'reference tree not found at {REF_ROOT}')
    sys.dont_write_bytecode = True
    stubs = os.path.join(HERE, '_stubs')
    for p in (REF_ROOT, stubs):
        if p not in sys.path:
            sys.path.insert(0, p)


def build_reference(world, model='nrms', hidden=256, neg_count=4, dropout=0.0, heads=8, additive=256,
                    use_neg_sampling=True, glove_npy=None, llm_item_table=None):
    """Return (legommender, resampler, lego_config, Env) built from the reference's own classes.

    model in {'nrms', 'naml', 'llmid'}; dropouts are forced to `dropout` everywhere so parity
    can be stated deterministically (SURVEY §7 'Dropout parity').
    """
    activate()
    import torch
    from loader.env import Env
    from loader.column_map import ColumnMap
    from loader.embedding_hub import EmbeddingHub
    from loader.resampler import Resampler
    from model.lego_config import LegoConfig
    from model.legommender import Legommender
    from model.operators.attention_operator import AttentionOperator
    from model.operators.cnn_operator import CNNOperator
    from model.operators.ada_operator import AdaOperator
    from model.predictors.dot_predictor import DotPredictor

    Env.device = torch.device('cpu')
    Env.simple_dev = False
    Env.train()
    Env.set_item_cache(False)
    Env.set_user_cache(False)
    Env.set_lm_cache(False)

    cm = ColumnMap(neg_col='neg')
    item_ut, user_ut = world.item_table(), world.user_table()
    predictor_cls, predictor_cfg = DotPredictor, None
    if model == 'nrms':
        item_cls, user_cls = AttentionOperator, AttentionOperator
        item_cfg = dict(num_attention_heads=heads, attention_dropout=dropout, additive_hidden_size=additive,
                        inputer_config=dict(use_cls_token=False, use_sep_token=True))
        user_cfg = dict(num_attention_heads=heads, attention_dropout=dropout, additive_hidden_size=additive,
                        inputer_config=dict(use_cls_token=False, use_sep_token=False))
        use_item_content = True
    elif model == 'naml':
        item_cls, user_cls = CNNOperator, AdaOperator
        item_cfg = dict(dropout=dropout, kernel_size=3, additive_hidden_size=additive)
        user_cfg = dict(additive_hidden_size=additive,
                        inputer_config=dict(use_cls_token=False, use_sep_token=False))
        use_item_content = True
    elif model == 'fastformer':
        from model.operators.fastformer_operator import FastformerOperator
        item_cls, user_cls = FastformerOperator, FastformerOperator
        item_cfg = dict(num_attention_heads=heads, num_hidden_layers=1, hidden_dropout_prob=dropout,
                        inputer_config=dict(use_cls_token=False, use_sep_token=False))
        user_cfg = dict(num_attention_heads=heads, num_hidden_layers=1, hidden_dropout_prob=dropout,
                        inputer_config=dict(use_cls_token=False, use_sep_token=False))
        use_item_content = True
    elif model == 'miner':
        from model.operators.transformer_operator import TransformerOperator
        from model.operators.poly_attention_operator import PolyAttentionOperator
        from model.predictors.miner_predictor import MINERPredictor
        item_cls, user_cls = TransformerOperator, PolyAttentionOperator
        item_cfg = dict(num_attention_heads=heads, attention_dropout=dropout, num_hidden_layers=2,
                        inputer_config=dict(use_cls_token=False, use_sep_token=True))
        user_cfg = dict(num_context_codes=8, context_code_dim=24, inputer_config=dict(use_cls_token=False, use_sep_token=False))
        use_item_content = True
        predictor_cls, predictor_cfg = MINERPredictor, dict(score_type='weighted')
    elif model == 'lstur':
        from model.operators.cnn_cat_operator import CNNCatOperator
        from model.operators.gru_operator import GRUOperator
        item_cls, user_cls = CNNCatOperator, GRUOperator
        item_cfg = dict(dropout=dropout, kernel_size=3, additive_hidden_size=additive)
        user_cfg = dict(inputer_config=dict(use_cls_token=False, use_sep_token=False))
        use_item_content = True
    elif model == 'pool':
        from model.operators.pooling_operator import PoolingOperator
        item_cls, user_cls = PoolingOperator, AdaOperator
        item_cfg = dict(flatten=False, max_pooling=False)
        user_cfg = dict(additive_hidden_size=additive,
                        inputer_config=dict(use_cls_token=False, use_sep_token=False))
        use_item_content = True
    elif model == 'llmid':
        item_cls, user_cls = None, AdaOperator
        item_cfg = None
        user_cfg = dict(additive_hidden_size=additive,
                        inputer_config=dict(use_cls_token=False, use_sep_token=False))
        use_item_content = False
    else:
        raise ValueError(model)

    cfg = LegoConfig(hidden_size=hidden, user_config=user_cfg, item_config=item_cfg, neg_count=neg_count,
                     use_neg_sampling=use_neg_sampling, use_item_content=use_item_content, item_page_size=0, predictor_config=predictor_cfg)
    cfg.set_component_classes(item_cls, user_cls, predictor_cls)
    item_inputs = [world.title_col, 'category']
    cfg.set_item_ut(item_ut, item_inputs)
    cfg.set_user_ut(user_ut, ['history'])
    cfg.set_column_map(cm)

    eh = EmbeddingHub(embedding_dim=cfg.item_hidden_size, transformation='auto', transformation_dropout=dropout)
    tmp = None
    if model != 'llmid':
        if glove_npy is None:
            tmp = tempfile.NamedTemporaryFile(suffix='.npy', delete=False)
            np.save(tmp.name, world.word_table)
            glove_npy = tmp.name
        eh.load_pretrained_embedding(glove_npy, vocab_name=world.word_vocab, frozen=True)
        eh.register_ut(item_ut, item_inputs)
    else:
        tmp = tempfile.NamedTemporaryFile(suffix='.npy', delete=False)
        np.save(tmp.name, llm_item_table)
        eh.load_pretrained_embedding(tmp.name, vocab_name='item_id', frozen=True)
        eh.register_vocab(item_ut.meta.features['item_id'].tokenizer.vocab)
    cfg.set_embedding_hub(eh)
    cfg.build_components()
    cfg.register_inputer_vocabs()
    model_ = Legommender(cfg)
    if dropout == 0.0:      # BertConfig.hidden_dropout_prob (0.1) is not reachable through the operator config: parity runs switch every dropout off
        for m in model_.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
    resampler = Resampler(cfg)
    if tmp is not None:
        os.unlink(tmp.name)
    return model_, resampler, cfg, Env
