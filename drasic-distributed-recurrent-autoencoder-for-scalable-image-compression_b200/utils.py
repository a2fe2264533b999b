"""Helpers of the reference's src/utils.py that the hot path or its callers touch, restated for
torch 2.x: apply_fn (utils.py:68-74), ntuple (:57-64), process_control_name (:113-120), recur /
to_device (:52-54,88-101), collate (:186-189), save / load (:25-47; load passes weights_only=False
because the reference's checkpoints pickle Logger and RandomState objects), resume (:169-183)."""
import collections.abc as container_abcs
import errno
import os
from itertools import repeat

import numpy as np
import torch

import config


def makedir_exist_ok(dirpath):
    try:
        os.makedirs(dirpath)
    except OSError as e:
        if e.errno != errno.EEXIST:
            raise


def save(input, path, protocol=2, mode='torch'):
    makedir_exist_ok(os.path.dirname(path))
    if mode == 'torch':
        torch.save(input, path, pickle_protocol=protocol)
    elif mode == 'numpy':
        np.save(path, input)
    else:
        raise ValueError('Not valid save mode')


def load(path, mode='torch'):
    if mode == 'torch':
        return torch.load(path, map_location=lambda storage, loc: storage, weights_only=False)
    elif mode == 'numpy':
        return np.load(path)
    raise ValueError('Not valid save mode')


def ntuple(n):
    def parse(x):
        if isinstance(x, container_abcs.Iterable) and not isinstance(x, str):
            return x
        return tuple(repeat(x, n))
    return parse


def apply_fn(module, fn):
    """call `fn` (by name) on every descendant that has it -- used for free_hidden()"""
    for _, m in module.named_children():
        f = getattr(m, fn, None)
        if f is not None:
            f()
        if next(m.named_children(), None) is not None:
            apply_fn(m, fn)


def recur(fn, input, *args):
    if isinstance(input, (torch.Tensor, np.ndarray)):
        return fn(input, *args)
    if isinstance(input, list) or (isinstance(input, container_abcs.Sequence) and not isinstance(input, (str, bytes, tuple))):
        return [recur(fn, v, *args) for v in input]
    if isinstance(input, dict):
        return {k: recur(fn, v, *args) for k, v in input.items()}
    raise ValueError('Not valid input type')


def to_device(input, device):
    return recur(lambda x, y: x.to(y), input, device)


def collate(input):
    for k in input:
        input[k] = torch.stack(input[k], 0)
    return input


def process_control_name():
    c = config.PARAM['control']
    config.PARAM['normalization'] = c['normalization']
    config.PARAM['activation'] = c['activation']
    config.PARAM['num_iter'] = int(c['num_iter'])
    config.PARAM['code_size'] = int(c['code_size'])
    config.PARAM['num_node'] = int(c['num_node'])
    config.PARAM['num_channel'] = 1 if config.PARAM['data_name'] == 'MNIST' else 3


def resume(model, model_tag, optimizer=None, scheduler=None, load_tag='checkpoint'):
    path = './output/model/{}_{}.pt'.format(model_tag, load_tag)
    if not os.path.exists(path):
        raise ValueError('Not exists model tag')
    checkpoint = load(path)
    model.load_state_dict(checkpoint['model_dict'])
    if optimizer is not None:
        optimizer.load_state_dict(checkpoint['optimizer_dict'])
    if scheduler is not None:
        scheduler.load_state_dict(checkpoint['scheduler_dict'])
    print('Resume from {}'.format(checkpoint['epoch']))
    return checkpoint['epoch'], model, optimizer, scheduler, checkpoint['logger']
