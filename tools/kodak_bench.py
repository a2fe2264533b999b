"""BASELINE config 5 on one GPU: Kodak-size 768x512 synthetic images, inference-only encode+decode, T=16,
(a) tiled into 32x32 patches at batch 4096 (10.67 images per batch) and (b) un-tiled, fully convolutional.
usage: python tools/kodak_bench.py [precision]"""
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

prec = sys.argv[1] if len(sys.argv) > 1 else 'parity'
sd = make_state_dict(3, 8, 1, seed=0)
eng = CodecEngine(3, 8, 1, precision=prec)
g = torch.Generator().manual_seed(1)


def run(img, reps=3):
    lab = torch.zeros(img.shape[0], dtype=torch.long)
    for _ in range(2):
        eng.forward(sd, img, lab.cuda(), 16, label_host=lab.numpy())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = eng.forward(sd, img, lab.cuda(), 16, label_host=lab.numpy())
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


ms, _ = run(torch.rand(4096, 3, 32, 32, generator=g).cuda())
print('tiled   : 4096 patches of 32x32 (10.67 Kodak images) %.1f ms -> %.0f patches/s = %.1f Kodak images/s (%s)' % (
    ms, 4096 / ms * 1e3, 4096 / 384.0 / ms * 1e3, prec))
ms, out = run(torch.rand(8, 3, 512, 768, generator=g).cuda())
print('untiled : 8 images of 768x512 %.1f ms -> %.1f Kodak images/s (%s), launches %d' % (ms, 8 / ms * 1e3, prec, eng.launches))
