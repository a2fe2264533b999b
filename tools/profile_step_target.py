"""One steady-state training step through the public API (models.dist_codec(), loss.backward(), Adam) for an
ncu launch list: run under `ncu --profile-from-start off ...`; only the last step is inside the
cudaProfilerStart/Stop range.   usage: python tools/profile_step_target.py [batch] [precision] [T]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'drasic-distributed-recurrent-autoencoder-for-scalable-image-compression_b200')
sys.path.insert(0, ROOT)
sys.path.insert(0, PKG)
import config

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
prec = sys.argv[2] if len(sys.argv) > 2 else 'parity'
T = int(sys.argv[3]) if len(sys.argv) > 3 else 16
config.init()
config.PARAM.update(device='cuda', data_name='CIFAR10', precision=prec, pack_mode='reference', grad_precision='fp16')
config.PARAM['control'] = {'normalization': 'none', 'activation': 'tanh', 'num_iter': str(T), 'code_size': '8', 'num_node': '8'}
import utils
utils.process_control_name()
import models
torch.manual_seed(0)
model = models.dist_codec().cuda()
model.train(True)
g = torch.Generator().manual_seed(1)
inp = {'img': torch.rand(B, 3, 32, 32, generator=g).cuda(), 'label': torch.randint(0, 8, (B,), generator=g).cuda()}
from engine.optim import FusedAdam
opt = FusedAdam(model.parameters(), lr=1e-3)


def step():
    opt.zero_grad(set_to_none=True)
    out = model(inp)
    out['loss'].backward()
    opt.step()
    return out


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
e0.record()
out = step()
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('ok step ms', e0.elapsed_time(e1), 'loss', float(out['loss']))
