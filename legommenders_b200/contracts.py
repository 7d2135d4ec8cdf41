"""The three plugin contracts of the hot path in one place: what an inputer, an operator and a predictor must offer so that the
reference's config files (`config/model/*.yaml`, keys `meta.item | user | predictor`) can select the B200 classes.

The attribute and method NAMES below are the drop-in boundary (reference: model/inputer/base_inputer.py,
model/operators/base_operator.py, model/predictors/base_predictor.py) and therefore fixed; everything behind them is this
package's own: operators run on packed rows through the C ABI, predictors are fused with their loss, dropout is counter-based.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Type

import torch
from torch import nn


_index = [0]


def _next_index() -> int:
    _index[0] += 1
    return _index[0]


# ---------------------------------------------------------------------------------------------------------------- inputer
class BaseInputer:
    """Turns one item sample (dict column -> token ids) into the integer layout an operator consumes, and a collated batch of
    those layouts into (embeddings, mask).  `output_single_sequence` tells the model whether columns were concatenated."""

    output_single_sequence = True

    def __init__(self, ut, inputs, eh, **_unused):
        self.ut, self.eh = ut, eh
        self.inputs: list = inputs

    # -- per sample (host, runs once per item when the Resampler builds its cache)
    def sample_rebuilder(self, sample: dict):
        raise NotImplementedError(f'{type(self).__name__} must lay out a sample')

    def __call__(self, sample: dict):
        return self.sample_rebuilder(sample)

    # -- per batch (device)
    def get_mask(self, batched_samples: Dict[str, torch.Tensor]):
        raise NotImplementedError(f'{type(self).__name__} must derive the attention mask')

    def get_embeddings(self, batched_samples: Dict[str, torch.Tensor]):
        raise NotImplementedError(f'{type(self).__name__} must embed a batch')

    def get_vocabs(self) -> Optional[List]:
        """Extra vocabularies the inputer itself introduces (special tokens); registered with the EmbeddingHub by LegoConfig."""
        return []


# --------------------------------------------------------------------------------------------------------------- operator
class BaseOperatorConfig:
    """hidden_size / input_dim / inputer_config plus operator-specific keywords; keywords an operator does not know are accepted
    and ignored (the shipped yaml files pass e.g. `use_fast_eval` to every user operator)."""

    def __init__(self, hidden_size, input_dim, inputer_config=None, **_ignored):
        self.hidden_size, self.input_dim = hidden_size, input_dim
        self.inputer_config = dict(inputer_config) if inputer_config else {}


class BaseOperator(nn.Module):
    """An item or user encoder: `forward(embeddings, mask=None) -> [N, hidden_size]` fp32 on Env.device."""

    config_class = BaseOperatorConfig
    inputer_class: Type[BaseInputer]
    inputer: BaseInputer
    allow_caching = True       # may its outputs be held in the repr cache during evaluation
    flatten_mode = False

    def __init__(self, config: BaseOperatorConfig, lego_config, target_user=False):
        nn.Module.__init__(self)   # not super(): integration.dual() puts the reference's base behind this class in the MRO
        self.config, self.lego_config, self.target_user = config, lego_config, target_user
        side = 'user' if target_user else 'item'
        self.inputer = self.inputer_class(ut=getattr(lego_config, f'{side}_ut'), inputs=getattr(lego_config, f'{side}_inputs'),
                                          eh=lego_config.eh, **config.inputer_config)
        self._calls = 0
        self._index = _next_index()

    def forward(self, embeddings, mask=None, **kwargs):
        raise NotImplementedError(f'{self.classname} has no forward')

    # -- sizes and placeholders
    @property
    def output_dim(self) -> int:
        return self.config.hidden_size

    def get_full_placeholder(self, sample_size: int) -> torch.Tensor:
        """Host-side zero buffer the cacher / pager fills page by page."""
        return torch.zeros((sample_size, self.config.hidden_size), dtype=torch.float32)

    # -- bookkeeping the trainer asks for
    def get_pretrained_parameter_names(self) -> List[str]:
        return []

    @property
    def classname(self) -> str:
        return type(self).__name__

    def prepare_for_predictor(self, user_embeddings, candidate_size):
        """[B, D] -> [B*C, D] (each user row repeated per candidate).  Only here for API parity: the fused dot/CE kernels read the
        user row once per impression and never materialise the repeat."""
        if not self.target_user:
            raise AssertionError('repeat is only designed for user encoder')
        B, D = user_embeddings.shape[0], user_embeddings.shape[-1]
        return user_embeddings[:, None, :].expand(B, candidate_size, D).reshape(B * candidate_size, D)

    def _next_seed(self) -> int:
        """A fresh counter-based dropout stream per (module, call); forward and backward of a call share it."""
        self._calls += 1
        return (torch.initial_seed() * 1000003 + self._index * 8191 + self._calls) & ((1 << 62) - 1)


# -------------------------------------------------------------------------------------------------------------- predictor
class BasePredictorConfig:
    def __init__(self, hidden_size, **_ignored):
        self.hidden_size = hidden_size


class BasePredictor(nn.Module):
    """Scores (user, item) representation pairs.  Calling the module IS `predict` (no nn.Module hooks in between), as the model
    code expects."""

    config_class = BasePredictorConfig
    allow_ranking = allow_matching = True
    keep_input_dim = False

    def __init__(self, config: BasePredictorConfig, lego_config):
        nn.Module.__init__(self)
        self.config, self.lego_config = config, lego_config

    def predict(self, user_embeddings, item_embeddings):
        raise NotImplementedError(f'{type(self).__name__} has no predict')

    def __call__(self, *args: Any, **kwargs: Any):
        return self.predict(*args, **kwargs)
