from .base_predictor import BasePredictor, BasePredictorConfig
from .dot_predictor import DotPredictor

REGISTRY = {'dot': DotPredictor}


def get(name: str):
    return REGISTRY[name.lower()]
