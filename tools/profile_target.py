"""Short single-GPU run for ncu: two eval passes of the fused loop (CIFAR-shaped, M sources, T iterations).
usage: python tools/profile_target.py [batch] [precision] [T] [M] [stream_k auto|on|off] [channels]"""
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
T = int(sys.argv[3]) if len(sys.argv) > 3 else 16
M = int(sys.argv[4]) if len(sys.argv) > 4 else 8
sk = sys.argv[5] if len(sys.argv) > 5 else 'auto'
C = int(sys.argv[6]) if len(sys.argv) > 6 else 3
sd = make_state_dict(C, 8, M, seed=0)
g = torch.Generator().manual_seed(1)
img = torch.rand(B, C, 32, 32, generator=g).cuda()
lab = torch.randint(0, M, (B,), generator=g)
eng = CodecEngine(C, 8, M, precision=prec, stream_k=sk)
for _ in range(2):
    out = eng.forward(sd, img, lab.cuda(), T, label_host=lab.numpy())
torch.cuda.synchronize()
print('ok', float(out['loss_acc'].sum()), [(ly.name, ly.pair, ly.block_n, ly.pos_imgs, ly.n_units) for ly in eng._last_layers])
