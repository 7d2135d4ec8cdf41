"""Glue for running this package INSIDE the reference tree (Jyonn/Legommenders), i.e. the reference's own trainer, LegoConfig, Legommender
and ClassHub around the B200 operators / predictor / EmbeddingHub.

Why a module: the reference's discovery (`loader/class_hub.py:131-151`) registers a class only if it is a subclass of the REFERENCE's
`BaseOperator` / `BasePredictor`, and its model code reads the reference's `loader.env.Env`.  This package has its own contract bases
(`contracts.py`) and its own `Env`, so two bridges are needed:

* `dual(cls, ref_base)`       — a class that is both `cls` (this package's operator / predictor: its `__init__`, `forward`, parameters, state
                                 dict) and a subclass of the reference base (so `issubclass` in ClassHub holds).  The reference base's
                                 `__init__` never runs: the contract bases initialise `nn.Module` directly.
* `bind_reference_env()`      — this package's `Env` reads/writes the reference's `Env` (device, phase, cache flags): one global state.

A maintainer adds one file per plugin to the reference tree, e.g. `model/operators/b200attention_operator.py`:

    from legommenders_b200 import integration
    B200AttentionOperator = integration.plugin('attention')      # -> yaml  meta: {item: B200Attention}

and swaps the EmbeddingHub where the reference constructs it (`loader/manager.py:139-153`): `integration.EmbeddingHub(...)`.
`tests/test_reference_boundary.py` runs exactly this against the live reference tree (its real ClassHub, LegoConfig and Legommender).
"""
from __future__ import annotations

from typing import Dict, Type

from .embedding_hub import EmbeddingHub  # noqa: F401  (re-exported: the hub the B200 inputers need)
from .env import Env

OPERATORS = {
    'attention': ('operators.attention_operator', 'AttentionOperator'),
    'cnn': ('operators.cnn_operator', 'CNNOperator'),
    'ada': ('operators.ada_operator', 'AdaOperator'),
    'pooling': ('operators.pooling_operator', 'PoolingOperator'),
    'cnncat': ('operators.cnn_cat_operator', 'CNNCatOperator'),
    'gru': ('operators.gru_operator', 'GRUOperator'),
    'fastformer': ('operators.fastformer_operator', 'FastformerOperator'),
    'transformer': ('operators.transformer_operator', 'TransformerOperator'),
    'polyattention': ('operators.poly_attention_operator', 'PolyAttentionOperator'),
}
PREDICTORS = {
    'dot': ('predictors.dot_predictor', 'DotPredictor'),
    'miner': ('predictors.miner_predictor', 'MINERPredictor'),
}


def dual(cls: Type, ref_base: Type, name: str | None = None) -> Type:
    """`cls` re-based so that `issubclass(result, ref_base)`; MRO = result -> cls -> ... this package's base ... -> ref_base -> nn.Module."""
    if issubclass(cls, ref_base):
        return cls
    return type(name or 'B200' + cls.__name__, (cls, ref_base), {'__module__': cls.__module__, '__doc__': cls.__doc__})


def _load(table: Dict[str, tuple], key: str) -> Type:
    import importlib
    mod, attr = table[key.lower()]
    return getattr(importlib.import_module('legommenders_b200.' + mod), attr)


def plugin(key: str, kind: str = 'operator') -> Type:
    """The discoverable class for a reference-side plugin file: `B200<Name>Operator` / `B200<Name>Predictor`."""
    if kind == 'operator':
        from model.operators.base_operator import BaseOperator as ref_base      # the reference tree must be importable here
        cls = _load(OPERATORS, key)
    elif kind == 'predictor':
        from model.predictors.base_predictor import BasePredictor as ref_base
        cls = _load(PREDICTORS, key)
    else:
        raise ValueError(f'unknown plugin kind {kind!r}')
    bind_reference_env()
    return dual(cls, ref_base)


def bind_reference_env():
    """Share the reference's global `Env` (loader/env.py) with this package's modules."""
    from loader.env import Env as RefEnv
    Env.bind(RefEnv)
    return RefEnv


def unbind_env():
    Env.bind(None)
