"""Quantization module (reference modules/quantization.py:5-12): binarize with the module's training flag."""
import torch.nn as nn

from functions import Quantize as QuantizeFunction

__all__ = ['Quantization']


class Quantization(nn.Module):
    def __init__(self):
        super(Quantization, self).__init__()

    def forward(self, input, noise=None):
        return QuantizeFunction.apply(input, self.training, noise)
