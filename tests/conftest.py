import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'drasic-distributed-recurrent-autoencoder-for-scalable-image-compression_b200')
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a real B200 (run with -m gpu on the GPU box)')
    # a fresh checkout has no libdrasic_b200.so (built artefacts are not tracked): compile it once (nvcc needs no GPU)
    if not os.path.exists(os.path.join(PKG, 'libdrasic_b200.so')):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


def build_model(name, data_name, num_node, num_iter, seed=0, device='cuda', precision='parity', code_size=8,
                loss_fn='mae', pack_mode='reference', cta_pair='auto', stream_k='auto'):
    """the reference's own construction sequence (train_model.py:39-58) against the B200 packages"""
    import torch
    import config
    config.init()
    config.PARAM['device'] = device
    config.PARAM['data_name'] = data_name
    config.PARAM['control'] = {'normalization': 'none', 'activation': 'tanh', 'num_iter': str(num_iter),
                               'code_size': str(code_size), 'num_node': str(num_node)}
    config.PARAM['precision'] = precision
    config.PARAM['loss_fn'] = loss_fn
    config.PARAM['pack_mode'] = pack_mode
    config.PARAM['cta_pair'] = cta_pair
    config.PARAM['stream_k'] = stream_k
    import utils
    utils.process_control_name()
    import models
    torch.manual_seed(seed)
    m = getattr(models, name)()
    return m.to(device)


def sd_digest(sd):
    import hashlib
    h = hashlib.sha256()
    for k in sorted(sd.keys()):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()
