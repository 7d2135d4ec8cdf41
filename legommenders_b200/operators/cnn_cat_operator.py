"""LSTUR item encoder: per column Conv1d('same')+ReLU+mask -> additive attention, single-token columns pass through, columns CONCATENATED on the
feature axis (mirror of model/operators/cnn_cat_operator.py:8-46)."""
import torch

from .. import ops
from ..env import Env
from .cnn_operator import CNNOperator, CNNOperatorConfig


class CNNCatOperatorConfig(CNNOperatorConfig):
    pass


class CNNCatOperator(CNNOperator):
    config_class = CNNCatOperatorConfig
    config: CNNCatOperatorConfig

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.num_columns = len(self.lego_config.item_inputs)

    def forward(self, embeddings: dict, mask=None, **kwargs):
        outs = []
        p = self.config.dropout if self.training else 0.0
        for col, e in embeddings.items():
            if e.shape[1] > 1:
                m = mask[col].to(Env.device)
                h = ops.conv1d_relu_mask(e, self.cnn.weight, self.cnn.bias, m, drop_p=p, seed=self._next_seed() if p else 0)
                outs.append(self.additive_attention(h, m))
            else:
                outs.append(e.squeeze(1))                      # a single-token column is its embedding (cnn_cat_operator.py:34-35)
        return torch.cat(outs, dim=-1)                         # feature-axis concat: data movement

    @property
    def output_dim(self):
        return self.config.hidden_size * self.num_columns

    def get_full_placeholder(self, sample_size):
        return torch.zeros((sample_size, self.output_dim), dtype=torch.float32)
