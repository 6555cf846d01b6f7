"""Parameter initialisers with the reference's names (reference src/models/init.py)."""
import torch.nn as nn
from torch.nn.init import constant_, xavier_normal_, xavier_uniform_


def _init_with(fn, module):
    if isinstance(module, (nn.Embedding, nn.Linear)):
        fn(module.weight.data)
        if isinstance(module, nn.Linear) and module.bias is not None:
            constant_(module.bias.data, 0)


def xavier_normal_initialization(module):
    """init.py:13-29."""
    _init_with(xavier_normal_, module)


def xavier_uniform_initialization(module):
    """init.py:32-48."""
    _init_with(xavier_uniform_, module)
