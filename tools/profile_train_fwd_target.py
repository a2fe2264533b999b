"""Forward pass of a TRAINING step (node ranges padded for the backward, every iteration's state kept) for ncu:
run under `ncu --profile-from-start off -k regex:conv_gemm ...`; only the forward of the last step sits inside the
cudaProfilerStart/Stop range, so the capture holds its 6 x T gate launches (t = 0 skips the hidden convolution).
usage: python tools/profile_train_fwd_target.py [batch] [precision] [T]"""
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
T = int(sys.argv[3]) if len(sys.argv) > 3 else 2
sd = make_state_dict(3, 8, 8, seed=0)
g = torch.Generator().manual_seed(1)
img = torch.rand(B, 3, 32, 32, generator=g).cuda()
lab = torch.randint(0, 8, (B,), generator=g)
eng = CodecEngine(3, 8, 8, precision=prec)
out = eng.forward(sd, img, lab.cuda(), T, training=True, label_host=lab.numpy(), save_for_backward=True)
grads = eng.backward(out)
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = eng.forward(sd, img, lab.cuda(), T, training=True, label_host=lab.numpy(), save_for_backward=True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
grads = eng.backward(out)
torch.cuda.synchronize()
print('ok', float(out['loss_acc'].sum()), float(sum(v.abs().sum() for v in grads.values())))
