from .base_predictor import BasePredictor, BasePredictorConfig
from .dot_predictor import DotPredictor
from .miner_predictor import MINERPredictor

REGISTRY = {'dot': DotPredictor, 'miner': MINERPredictor}


def get(name: str):
    return REGISTRY[name.lower()]
