"""Soak: N training steps through the public API on the synthetic data module (fused Adam), printing the loss and
PSNR trajectory; checks that everything stays finite and that the codec learns.  usage: soak_train.py [steps] [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'drasic-distributed-recurrent-autoencoder-for-scalable-image-compression_b200')
sys.path.insert(0, ROOT)
sys.path.insert(0, PKG)
import config

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
config.init()
config.PARAM.update(device='cuda', data_name='CIFAR10', precision=sys.argv[3] if len(sys.argv) > 3 else 'parity')
config.PARAM['control'] = {'normalization': 'none', 'activation': 'tanh', 'num_iter': '16', 'code_size': '8', 'num_node': '8'}
config.PARAM['data_size'] = {'train': steps * B, 'test': B}
config.PARAM['batch_size'] = {'train': B, 'test': B}
import utils
utils.process_control_name()
import models
import data
import metrics
torch.manual_seed(0)
model = models.dist_codec().cuda()
from engine.optim import FusedAdam
opt = FusedAdam(model.parameters(), lr=1e-3)
loader = data.make_data_loader(data.fetch_dataset('CIFAR10', 'label'))
model.train(True)
first = last = None
for i, batch in enumerate(loader['train']):
    inp = utils.to_device(utils.collate(batch), 'cuda')
    opt.zero_grad(set_to_none=True)
    out = model(inp)
    out['loss'].backward()
    gn = float(torch.sqrt(sum((p.grad.double() ** 2).sum() for p in model.parameters())))
    opt.step()
    loss = out['loss'].item()
    assert loss == loss and gn == gn and gn < 1e6, (i, loss, gn)
    first = loss if first is None else first
    last = loss
    if i % 10 == 0 or i == steps - 1:
        print('step %3d loss %.5f grad-norm %.4f' % (i, loss, gn), flush=True)
model.train(False)
with torch.no_grad():
    inp = utils.to_device(utils.collate(next(iter(loader['test']))), 'cuda')
    out = model(inp)
ps = metrics.psnr_per_iteration(out)
print('eval PSNR per rate: ' + ' '.join('%.2f' % p for p in ps))
assert last < first and ps[-1] > ps[0], (first, last, ps)
print('ok')
