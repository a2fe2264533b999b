"""Spec -> module tree (reference models/utils.py:5-29): a dict with a 'cell' key is a Cell, any other
dict a ModuleDict, a list a ModuleList, a tuple an nn.Sequential.  Every element is built on its own,
so `[spec] * num_node` yields independent parameters per node, in the reference's creation order."""
import torch.nn as nn

from modules import Cell


def make_model(model):
    if isinstance(model, dict):
        if 'cell' in model:
            return Cell(model)
        if 'nn' in model:
            return eval(model['nn'])
        return nn.ModuleDict({k: make_model(v) for k, v in model.items()})
    if isinstance(model, list):
        return nn.ModuleList([make_model(m) for m in model])
    if isinstance(model, tuple):
        return nn.Sequential(*[make_model(m) for m in model])
    raise ValueError('Not valid model format')
