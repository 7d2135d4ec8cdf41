"""Additive-attention user encoder (mirror of model/operators/ada_operator.py:7-38)."""
from ..env import Env
from ..inputer.concat_inputer import ConcatInputer
from .base_operator import BaseOperator, BaseOperatorConfig
from .common import AdditiveAttention


class AdaOperatorConfig(BaseOperatorConfig):
    def __init__(self, additive_hidden_size: int = 256, **kwargs):
        super().__init__(**kwargs)
        self.additive_hidden_size = additive_hidden_size


class AdaOperator(BaseOperator):
    config_class = AdaOperatorConfig
    inputer_class = ConcatInputer
    config: AdaOperatorConfig

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.additive_attention = AdditiveAttention(embed_dim=self.config.input_dim,
                                                    hidden_size=self.config.additive_hidden_size)

    supports_packed = True

    def forward(self, embeddings, mask=None, cu=None, max_len=None, **kwargs):
        if cu is not None:
            return self.additive_attention(embeddings, None, cu=cu, max_len=max_len)
        return self.additive_attention(embeddings, mask.to(Env.device))

    @property
    def output_dim(self):
        return self.config.input_dim
