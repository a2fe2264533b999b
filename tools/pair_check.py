"""CTA-pair (cta_group::2) GEMM path against the single-CTA path on the same inputs.
usage: python tools/pair_check.py [batch] [T] [precision]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'drasic-distributed-recurrent-autoencoder-for-scalable-image-compression_b200')
sys.path.insert(0, ROOT)
sys.path.insert(0, PKG)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
from _params import make_state_dict
from engine.codec import CodecEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 3
prec = sys.argv[3] if len(sys.argv) > 3 else 'parity'
train = len(sys.argv) > 4 and sys.argv[4] == 'train'
sd = make_state_dict(3, 8, 8, seed=0)
g = torch.Generator().manual_seed(1)
img = torch.rand(B, 3, 32, 32, generator=g).cuda()
lab = torch.randint(0, 8, (B,), generator=g)
noise = torch.rand(T, B, 8, 4, 4, generator=g).cuda()
res = {}
for mode in ('off', 'on'):
    eng = CodecEngine(3, 8, 8, precision=prec, cta_pair=mode)
    t0 = time.time()
    out = eng.forward(sd, img, lab.cuda(), T, training=train, noise=noise if train else None, label_host=lab.numpy(),
                      save_for_backward=train)
    grads = eng.backward(out) if train else {}
    torch.cuda.synchronize()
    res[mode] = (out['codes'].clone(), out['recon'].clone(), {k: v.clone() for k, v in grads.items()})
    print(mode, 'ok %.2fs' % (time.time() - t0), 'launches', eng.launches, 'pair layers', [ly.name for ly in out.get('_bwd', {}).get('layers', []) if ly.pair], flush=True)
c0, r0, g0 = res['off']
c1, r1, g1 = res['on']
print('code mismatches', int((c0 != c1).sum()), 'of', c0.numel(), ' recon max diff', float((r0 - r1).abs().max()))
for k in sorted(g0):
    d = float((g0[k] - g1[k]).abs().max())
    s = float(g0[k].abs().max())
    if d > 1e-3 * s + 1e-12:
        print('GRAD DIFF', k, d, s)
print('grads compared', len(g0))
