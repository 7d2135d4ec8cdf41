"""Parity cases shared by the golden generator, the oracle tests and the GPU tests.

A case is fully determined by its seeds: the world (`synth.MindWorld`) and the parameters (`synth.init_state`) are
rebuilt on whichever box runs the test; only reference INPUT batches and OUTPUTS are stored in <case>.npz.
"""
from __future__ import annotations

import os
from collections import OrderedDict

import numpy as np

from legommenders_b200.synth import MindWorld, init_state

HERE = os.path.dirname(os.path.abspath(__file__))

_small_world = dict(n_items=150, n_words=400, n_users=48, n_train=64, n_eval_groups=24, eval_group_mean=8,
                    title_len=10, hist_len=12, min_title=3, max_neg=12, embed_dim=48)
_full_world = dict(n_items=400, n_words=1500, n_users=64, n_train=64, n_eval_groups=16, eval_group_mean=10,
                   title_len=30, hist_len=50, embed_dim=300)

_h100_world = dict(n_items=300, n_words=1200, n_users=48, n_train=48, n_eval_groups=8, eval_group_mean=8,
                   title_len=30, hist_len=100, embed_dim=300, min_hist=60)
_hot_small = {'glove.linear.weight': 2, 'item_op.multi_head_attention.in_proj_weight': 3, 'item_op.linear.weight': 2,
              'item_op.additive_attention.encoder.0.weight': 16, 'item_op.additive_attention.encoder.2.weight': 24,
              'user_op.multi_head_attention.in_proj_weight': 3, 'user_op.linear.weight': 3,
              'user_op.additive_attention.encoder.0.weight': 4, 'user_op.additive_attention.encoder.2.weight': 16}
_hot_full = {'item_op.additive_attention.encoder.0.weight': 2, 'item_op.additive_attention.encoder.2.weight': 24,
             'user_op.multi_head_attention.in_proj_weight': 1.5, 'user_op.linear.weight': 1.5,
             'user_op.additive_attention.encoder.0.weight': 4, 'user_op.additive_attention.encoder.2.weight': 12}

CASES = OrderedDict(
    nrms_small=dict(kind='nrms', hidden=64, heads=8, additive=32, batch=6, seed=11, world=_small_world, cached_eval=True),
    naml_small=dict(kind='naml', hidden=64, heads=8, additive=32, batch=6, seed=12, world=_small_world, cached_eval=True),
    llmid_small=dict(kind='llmid', hidden=64, heads=8, additive=32, batch=6, seed=13, world=_small_world, llm_dim=96,
                     cached_eval=True),
    nrms_bce_small=dict(kind='nrms', hidden=64, heads=4, additive=32, batch=8, seed=14, world=_small_world,
                        use_neg_sampling=False),
    nrms_full=dict(kind='nrms', hidden=256, heads=8, additive=256, batch=4, seed=21, world=_full_world, full_grads=False),
    naml_full=dict(kind='naml', hidden=256, heads=8, additive=256, batch=4, seed=22, world=_full_world, full_grads=False),
    # "hot" states: selected weights scaled up so that EVERY parameter gradient (the additive-attention ones included, which are
    # ~1e-7 of the largest gradient at the plain init scale) is a few percent of the largest or more -> per-tensor 1e-4 bars, no floor
    nrms_small_hot=dict(kind='nrms', hidden=64, heads=8, additive=32, batch=6, seed=11, world=_small_world, strict_grads=True,
                        boost=_hot_small),
    nrms_full_hot=dict(kind='nrms', hidden=256, heads=8, additive=256, batch=4, seed=21, world=_full_world, full_grads=False,
                       strict_grads=True, boost=_hot_full),
    # history 100 (BASELINE config 4): user sequences beyond the 64-token tensor-core attention kernels
    nrms_h100=dict(kind='nrms', hidden=256, heads=8, additive=256, batch=3, seed=23, world=_h100_world, full_grads=False),
    # 4096-d LLM item embeddings (BASELINE config 5) at model level
    llmid_4096=dict(kind='llmid', hidden=64, heads=8, additive=32, batch=6, seed=24, world=_small_world, llm_dim=4096,
                    llm_scale=0.05, cached_eval=True, full_grads=False),
    # masked-mean PoolingOperator item encoder (a10) + Ada users
    # MINER: BERT-style Transformer item encoder (2 layers) + poly-attention user encoder (8 codes) + target-aware MINER predictor
    miner_small=dict(kind='miner', hidden=64, heads=8, additive=64, batch=6, seed=27, world=_small_world, layers=2, codes=8, code_dim=24),
    # Fastformer item and user encoders (config/model/fastformer.yaml: one layer each, no SEP tokens)
    fastformer_small=dict(kind='fastformer', hidden=64, heads=8, additive=64, batch=6, seed=28, world=_small_world, layers=1, cached_eval=True),
    # LSTUR: CNNCat item encoder (feature-axis concat of the columns) + GRU user encoder (config/model/lstur.yaml)
    lstur_small=dict(kind='lstur', hidden=64, heads=8, additive=32, batch=6, seed=26, world=_small_world),   # no cached eval: the reference's UserCacher placeholder is [users, hidden] but GRUOperator returns input_dim = 2*hidden (base_operator.py:52-53 vs gru_operator.py:54)
    pool_small=dict(kind='pool', hidden=64, heads=8, additive=32, batch=6, seed=25, world=_small_world, cached_eval=True),
)


def make_world(c: dict):
    world = MindWorld(seed=c['seed'], **c['world'])
    llm = None
    if c['kind'] == 'llmid':
        rng = np.random.default_rng(c['seed'] + 1000)
        llm = (rng.standard_normal((world.n_items, c['llm_dim'])) * c.get('llm_scale', 1.0)).astype(np.float32)
    return world, llm


def make_state(c: dict, world, shapes: dict, llm=None) -> dict:
    """Deterministic parameters; pretrained tables come from the world, everything else from init_state."""
    state = init_state(shapes, seed=c['seed'] + 7)
    for k in shapes:
        if k.endswith('.embedding.weight'):
            state[k] = llm if c['kind'] == 'llmid' else world.word_table
    for pattern, factor in c.get('boost', {}).items():
        hit = [k for k in state if pattern in k]
        assert hit, pattern
        for k in hit:
            state[k] = (state[k] * np.float32(factor)).astype(np.float32)
    return state


def flatten_tree(tree, prefix='') -> dict:
    out = {}
    for k, v in tree.items():
        if isinstance(v, dict):
            out.update(flatten_tree(v, prefix + k + '/'))
        else:
            out[prefix + k] = v
    return out


def unflatten_batch(npz) -> dict:
    """Rebuild the nested batch dict (torch int64 tensors) from 'batch/...' keys, preserving column order."""
    import torch
    root = OrderedDict()
    for key in npz.files:
        if not key.startswith('batch/'):
            continue
        parts = key[len('batch/'):].split('/')
        d = root
        for p in parts[:-1]:
            d = d.setdefault(p, OrderedDict())
        d[parts[-1]] = torch.from_numpy(npz[key])
    return root


def sample_strided(g: np.ndarray, n: int = 4096) -> np.ndarray:
    flat = g.reshape(-1)
    step = max(1, flat.size // n)
    return flat[::step][:n].copy()


def load(name: str):
    return np.load(os.path.join(HERE, name + '.npz'))
