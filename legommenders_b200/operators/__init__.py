from .base_operator import BaseOperator, BaseOperatorConfig
from .attention_operator import AttentionOperator, AttentionOperatorConfig
from .cnn_operator import CNNOperator, CNNOperatorConfig
from .ada_operator import AdaOperator, AdaOperatorConfig
from .pooling_operator import PoolingOperator, PoolingOperatorConfig
from .cnn_cat_operator import CNNCatOperator, CNNCatOperatorConfig
from .gru_operator import GRUOperator, GRUOperatorConfig
from .transformer_operator import TransformerOperator, TransformerOperatorConfig
from .poly_attention_operator import PolyAttentionOperator, PolyAttentionOperatorConfig
from .fastformer_operator import FastformerOperator, FastformerOperatorConfig

REGISTRY = {'attention': AttentionOperator, 'cnn': CNNOperator, 'ada': AdaOperator, 'pooling': PoolingOperator,
            'cnncat': CNNCatOperator, 'gru': GRUOperator, 'transformer': TransformerOperator, 'polyattention': PolyAttentionOperator, 'fastformer': FastformerOperator}


def get(name: str):
    """Same keying as loader/class_hub.py:112-115: class name minus 'Operator', lower-cased."""
    return REGISTRY[name.lower()]
