"""nn.Module face of the depth<->space re-layouts (same class names, constructor argument and attribute names
as reference modules/shuffle.py:5-28).  Inside the fused coding loop these modules never run: the gate
kernel's epilogue stores straight into the next layer's layout (csrc/convlstm_fwd.cu, NEXT_DOWN / NEXT_UP);
they exist for stand-alone use of a PixelShuffleCell and for state_dict / module-tree compatibility."""
import torch.nn as nn

import functions

__all__ = ['PixelUnShuffle', 'PixelShuffle']


class _Relayout(nn.Module):
    """one integer factor, one functional; subclasses name the attribute the way the reference does"""
    _attr = None
    _fn = None

    def __init__(self, factor):
        nn.Module.__init__(self)
        setattr(self, self._attr, factor)

    def forward(self, input):
        return getattr(functions, self._fn)(input, getattr(self, self._attr))

    def extra_repr(self):
        return '{}={}'.format(self._attr, getattr(self, self._attr))


class PixelUnShuffle(_Relayout):
    _attr, _fn = 'downscale_factor', 'pixel_unshuffle'

    def __init__(self, downscale_factor):
        _Relayout.__init__(self, downscale_factor)


class PixelShuffle(_Relayout):
    _attr, _fn = 'upscale_factor', 'pixel_shuffle'

    def __init__(self, upscale_factor):
        _Relayout.__init__(self, upscale_factor)
