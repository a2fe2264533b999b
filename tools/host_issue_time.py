"""Host time to issue one training step (all launches enqueued, no synchronize) next to the device time of the
step: tells whether a batch size is launch-bound.   usage: python tools/host_issue_time.py [batch] [mode]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'drasic-distributed-recurrent-autoencoder-for-scalable-image-compression_b200')
sys.path.insert(0, ROOT)
sys.path.insert(0, PKG)
import config

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
mode = sys.argv[2] if len(sys.argv) > 2 else 'train'
config.init()
config.PARAM.update(device='cuda', data_name='CIFAR10', precision='parity', pack_mode='reference', grad_precision='fp16')
config.PARAM['control'] = {'normalization': 'none', 'activation': 'tanh', 'num_iter': '16', 'code_size': '8', 'num_node': '8'}
import utils
utils.process_control_name()
import models
from engine.optim import FusedAdam
torch.manual_seed(0)
model = models.dist_codec().cuda()
model.train(mode == 'train')
g = torch.Generator().manual_seed(1)
inp = {'img': torch.rand(B, 3, 32, 32, generator=g).cuda(), 'label': torch.randint(0, 8, (B,), generator=g).cuda()}
opt = FusedAdam(model.parameters(), lr=1e-3)


def step():
    if mode == 'train':
        opt.zero_grad(set_to_none=True)
        out = model(inp)
        out['loss'].backward()
        opt.step()
    else:
        with torch.no_grad():
            out = model(inp)
    return out


for _ in range(4):
    step()
torch.cuda.synchronize()
issue, total = [], []
for _ in range(8):
    t0 = time.perf_counter()
    step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    issue.append(t1 - t0)
    total.append(t2 - t0)
issue.sort()
total.sort()
print('B=%d %s: host issue %.2f ms (median), step %.2f ms' % (B, mode, issue[4] * 1e3, total[4] * 1e3))
if '--profile' in sys.argv:
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(3):
        step()
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(35)
