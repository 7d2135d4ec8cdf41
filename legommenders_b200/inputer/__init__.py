from .base_inputer import BaseInputer
from .concat_inputer import ConcatInputer
from .simple_inputer import SimpleInputer
