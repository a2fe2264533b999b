"""Seeded codec parameters for the profiling / timing tools, built through the package's own model
constructors (the reference's construction sequence, train_model.py:39-58)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'drasic-distributed-recurrent-autoencoder-for-scalable-image-compression_b200')
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def make_state_dict(num_channel=3, code_size=8, num_node=8, separate=False, seed=0, device='cuda'):
    import config
    config.init()
    config.PARAM['device'] = device
    config.PARAM['data_name'] = 'MNIST' if num_channel == 1 else 'CIFAR10'
    config.PARAM['control'] = {'normalization': 'none', 'activation': 'tanh', 'num_iter': '16',
                               'code_size': str(code_size), 'num_node': str(num_node)}
    import utils
    utils.process_control_name()
    import models
    torch.manual_seed(seed)
    model = (models.sep_codec() if separate else models.dist_codec()).to(device)
    return {k: v.detach() for k, v in model.state_dict().items()}
