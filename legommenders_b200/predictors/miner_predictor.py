"""MINER target-aware predictor (mirror of model/predictors/miner_predictor.py:8-64)."""
from torch import nn

from .. import ops
from ..embedding_hub import _Affine
from .base_predictor import BasePredictor, BasePredictorConfig


class MINERPredictorConfig(BasePredictorConfig):
    def __init__(self, score_type: str = 'weighted', **kwargs):
        super().__init__(**kwargs)
        self.score_type = score_type


class TargetAwareAttention(nn.Module):
    """Holds `linear.weight` (no bias); the computation is fused into lk_miner_fwd."""

    def __init__(self, hidden_size: int):
        super().__init__()
        self.linear = _Affine(hidden_size, hidden_size, bias=False)


class MINERPredictor(BasePredictor):
    config_class = MINERPredictorConfig
    config: MINERPredictorConfig
    allow_ranking = False
    keep_input_dim = True

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        if self.config.score_type not in ops.MINER_MODES:
            raise ValueError(f'Unknown score type: {self.config.score_type}')
        self.target_aware_attention = TargetAwareAttention(self.config.hidden_size)

    def predict(self, user_embeddings, item_embeddings):
        """user [B, codes, D], items [B, K+1, D] -> [B, K+1]"""
        proj = None
        if self.config.score_type == 'weighted':
            proj = ops.gelu(ops.linear(user_embeddings, self.target_aware_attention.linear.weight, None))
        return ops.miner_score(user_embeddings, proj, item_embeddings, self.config.score_type)
