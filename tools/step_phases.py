"""Device time of the phases of one steady-state training step through the public API (forward, backward-through-
time, optimizer), next to the sum of the per-kernel times of the same step.   usage: python tools/step_phases.py [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'drasic-distributed-recurrent-autoencoder-for-scalable-image-compression_b200')
sys.path.insert(0, ROOT)
sys.path.insert(0, PKG)
import config

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
config.init()
config.PARAM.update(device='cuda', data_name='CIFAR10', precision='parity', pack_mode='reference', grad_precision='fp16')
config.PARAM['control'] = {'normalization': 'none', 'activation': 'tanh', 'num_iter': '16', 'code_size': '8', 'num_node': '8'}
import utils
utils.process_control_name()
import models
torch.manual_seed(0)
model = models.dist_codec().cuda()
model.train(True)
g = torch.Generator().manual_seed(1)
inp = {'img': torch.rand(B, 3, 32, 32, generator=g).cuda(), 'label': torch.randint(0, 8, (B,), generator=g)}
if len(sys.argv) <= 2 or sys.argv[2] != 'cpu_label':
    inp['label'] = inp['label'].cuda()
from engine.optim import FusedAdam
opt = FusedAdam(model.parameters(), lr=1e-3)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]


def step(timed=False):
    if timed:
        ev[0].record()
    opt.zero_grad(set_to_none=True)
    out = model(inp)
    if timed:
        ev[1].record()
    out['loss'].backward()
    if timed:
        ev[2].record()
    opt.step()
    if timed:
        ev[3].record()
    return out


for _ in range(4):
    step()
torch.cuda.synchronize()
acc = [0.0, 0.0, 0.0]
n = 5
for _ in range(n):
    step(True)
    torch.cuda.synchronize()
    for i in range(3):
        acc[i] += ev[i].elapsed_time(ev[i + 1])
print('forward %.2f ms  backward %.2f ms  optimizer %.2f ms  total %.2f ms' % (acc[0] / n, acc[1] / n, acc[2] / n, sum(acc) / n))
eng = model._engine
eng.profile = []
step()
torch.cuda.synchronize()
prof = eng.profile_summary()
eng.profile = None
fwd = sum(v[1] for k, v in prof.items() if k.startswith('gates') or k in ('tail_head', 'bottleneck'))
bwd = sum(v[1] for k, v in prof.items()) - fwd
print('sum of per-kernel times: forward kernels %.2f ms  backward kernels %.2f ms' % (fwd, bwd))
