"""Depth <-> space re-layouts with the reference's channel order (functions/shuffle.py:4-25:
out channel = c*r*r + dy*r + dx).  Pure index permutations: when a PixelShuffleCell is called on its
own they are expressed as torch views + one copy; inside the fused loop they do not exist at all --
the gate kernel's epilogue stores straight into the shuffled layout."""
from utils import ntuple

__all__ = ['pixel_unshuffle', 'pixel_shuffle']


def pixel_unshuffle(input, downscale_factor):
    ry, rx = ntuple(2)(downscale_factor)
    n, c, h, w = input.size()
    v = input.reshape(n, c, h // ry, ry, w // rx, rx)
    return v.permute(0, 1, 3, 5, 2, 4).reshape(n, c * ry * rx, h // ry, w // rx)


def pixel_shuffle(input, upscale_factor):
    ry, rx = ntuple(2)(upscale_factor)
    n, c, h, w = input.size()
    oc = c // (ry * rx)
    v = input.reshape(n, oc, ry, rx, h, w)
    return v.permute(0, 1, 4, 2, 5, 3).reshape(n, oc, h * ry, w * rx)
