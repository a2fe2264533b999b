"""Quantization module (reference modules/quantization.py:5-12): the binarizer as an nn.Module, stochastic while
the module is in training mode and a deterministic sign otherwise.  `noise` (optional uniforms) makes the
stochastic draw an explicit input, which is how the parity tests reproduce the reference's samples."""
from torch import nn

import functions

__all__ = ['Quantization']


class Quantization(nn.Module):
    def forward(self, input, noise=None):
        stochastic = bool(self.training)
        return functions.Quantize.apply(input, stochastic, noise)
