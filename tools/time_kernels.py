"""Per-kernel device times of one training step (forward with saved state + backward-through-time),
CUDA events around every launch on the launching stream.
usage: python tools/time_kernels.py [batch] [precision] [T] [grad_precision]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'drasic-distributed-recurrent-autoencoder-for-scalable-image-compression_b200')
sys.path.insert(0, ROOT)
sys.path.insert(0, PKG)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
from _params import make_state_dict
from engine.codec import CodecEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
prec = sys.argv[2] if len(sys.argv) > 2 else 'parity'
T = int(sys.argv[3]) if len(sys.argv) > 3 else 4
gprec = sys.argv[4] if len(sys.argv) > 4 else 'fp16'
sd = make_state_dict(3, 8, 8, seed=0)
g = torch.Generator().manual_seed(1)
img = torch.rand(B, 3, 32, 32, generator=g).cuda()
lab = torch.randint(0, 8, (B,), generator=g)
eng = CodecEngine(3, 8, 8, precision=prec, grad_precision=gprec)


def step():
    out = eng.forward(sd, img, lab.cuda(), T, training=True, label_host=lab.numpy(), save_for_backward=True)
    return out, eng.backward(out)


for _ in range(2):
    step()
torch.cuda.synchronize()
eng.profile = []
reps = 3
for _ in range(reps):
    out, grads = step()
torch.cuda.synchronize()
prof = eng.profile_summary()
eng.profile = None
tot = 0.0
print('B=%d precision=%s grad=%s T=%d dbg=%s' % (B, prec, gprec, T, os.environ.get('DRASIC_WGRAD_DBG', '0')))
for k, (n, ms, fl) in sorted(prof.items()):
    tot += ms
    print('%-16s n=%4d  %9.3f ms/step  %8.1f us/launch  %8.1f TFLOP/s' % (k, n // reps, ms / reps, 1e3 * ms / n, fl / ms / 1e9 if ms else 0))
print('sum of timed kernels: %.2f ms/step (T=%d)' % (tot / reps, T))
print('loss_acc', float(out['loss_acc'].sum()), 'grad abs-sum', float(sum(v.abs().sum() for v in grads.values())))
