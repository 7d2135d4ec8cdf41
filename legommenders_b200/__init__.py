"""legommenders_b200 — B200-native hot path of Jyonn/Legommenders behind the reference's plugin surface.

Python host (this package) mirrors LegoConfig / operators / predictor / EmbeddingHub / cacher; all tensor work
runs in hand-written sm_100a CUDA kernels from `liblegommenders_b200.so` (C ABI: include/legommenders_b200.h).
There is no CPU or eager fallback: importing is free, but any op raises if the library is missing.
"""
from .env import Env
from .column_map import ColumnMap
from .embedding_hub import EmbeddingHub
from .lego_config import LegoConfig
from .legommender import Legommender
from .resampler import Resampler, DataSet
from . import operators, predictors, ops, evaluate, sharding
from .metrics import MetricPool

__all__ = ['Env', 'ColumnMap', 'EmbeddingHub', 'LegoConfig', 'Legommender', 'Resampler', 'DataSet', 'operators',
           'predictors', 'ops', 'evaluate', 'sharding', 'MetricPool']
