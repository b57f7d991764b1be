"""
ghn3_b200 -- B200-native implementation of the GHN-3 parameter-prediction hot path.

Same top-level names as the reference package (ghn3/__init__.py:8-13) for the parts that are in scope:
`from_pretrained`, `GHN3`, `Graph`, `GraphBatch`. Heavy imports are lazy so that `import ghn3_b200.weights`
works without CUDA.
"""
__all__ = ['from_pretrained', 'GHN3', 'GHN', 'Graph', 'GraphBatch', 'param_norm', 'CONFIGS', 'Trainer']


def __getattr__(name):
    if name in ('from_pretrained', 'GHN3', 'GHN', 'param_norm', 'ConvDecoder3'):
        from . import nn as _nn
        return getattr(_nn, name)
    if name in ('Graph', 'GraphBatch'):
        from . import graph as _graph
        return getattr(_graph, name)
    if name == 'Trainer':
        from .trainer import Trainer
        return Trainer
    if name == 'CONFIGS':
        from .weights import CONFIGS
        return CONFIGS
    raise AttributeError(name)
