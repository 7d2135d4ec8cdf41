"""DotPredictor (mirror of model/predictors/dot_predictor.py:6-10) on the fused scoring kernels."""
from .. import ops
from .base_predictor import BasePredictor


class DotPredictor(BasePredictor):
    # the fused kernel takes user [B,D] and items [B,C,D] directly: no repeat is materialised
    fused_scoring = True

    def predict(self, user_embeddings, item_embeddings):
        """Accepts the reference's calling conventions: ([B·C,D],[B·C,D]) / ([B,D],[B,D]) -> [B·C] / [B];
        and the fused one ([B,D],[B,C,D]) -> [B,C]."""
        if item_embeddings.dim() == 3:
            return ops.dot_scores(user_embeddings, item_embeddings)
        return ops.dot_scores(user_embeddings, item_embeddings.unsqueeze(1)).squeeze(1)
