"""Masked mean / max pooling encoder (mirror of model/operators/pooling_operator.py:10-61)."""
from collections import OrderedDict

import torch

from .. import ops
from ..env import Env
from ..inputer.simple_inputer import SimpleInputer
from .base_operator import BaseOperator, BaseOperatorConfig


class PoolingOperatorConfig(BaseOperatorConfig):
    def __init__(self, flatten: bool = False, max_pooling: bool = False, **kwargs):
        super().__init__(**kwargs)
        self.flatten = flatten
        self.max_pooling = max_pooling


class PoolingOperator(BaseOperator):
    inputer_class = SimpleInputer
    config_class = PoolingOperatorConfig
    config: PoolingOperatorConfig

    def forward(self, embeddings, mask=None, **kwargs):
        assert mask is not None, 'mask is required for pooling fusion'
        if isinstance(embeddings, torch.Tensor):
            assert isinstance(mask, torch.Tensor)
            embeddings, mask = dict(temp=embeddings), dict(temp=mask)
        elif isinstance(mask, torch.Tensor):
            assert len(embeddings) == 1
            mask = {next(iter(embeddings)): mask}
        mode = ops.POOL_MAX if self.config.max_pooling else ops.POOL_MEAN
        pooled = OrderedDict((col, ops.masked_pool(e, mask[col].to(Env.device), mode)) for col, e in embeddings.items())
        cols = list(pooled.values())
        if self.config.flatten:
            return torch.cat(cols, dim=-1)
        if len(cols) == 1:
            return cols[0]
        stack = torch.stack(cols, dim=1)     # column combine: K small [N,D] tensors (pooling_operator.py:57-61)
        return stack.max(dim=1)[0] if self.config.max_pooling else stack.mean(dim=1)
