"""Differential check of oracle/drasic_oracle.py against the REFERENCE's own modules on fresh seeds and shapes
(not the committed fixtures).  Runs only where the reference tree is mounted (the build container):
      python tests/golden/live_check.py
It is its own process because the reference's top-level module names (config, models, modules, functions, utils)
are the ones this repo's package re-implements.  Exit status 0 = every case agrees."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_golden as G                      # the reference harness (SURVEY 8c recipe)
from oracle import drasic_oracle as O        # torch-only: no module-name clash with the reference


def check(cond, what):
    if not cond:
        print('MISMATCH:', what)
        sys.exit(1)


def main():
    config, models, modules, functions, utils = G.import_reference()
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    cases = [('dist_codec', 'CIFAR10', 3, 3, 7, 100), ('dist_codec', 'MNIST', 2, 2, 5, 101), ('sep_codec', 'CIFAR10', 2, 2, 6, 102)]
    for name, data, M, T, Bn, seed in cases:
        m = G.build_model(config, models, utils, name, data, M, T, seed=seed)
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        C = 1 if data == 'MNIST' else 3
        g = torch.Generator().manual_seed(seed + 1)
        img = torch.rand(Bn, C, 32, 32, generator=g)
        label = torch.randint(0, M, (Bn,), generator=g)
        label[0] = M - 1
        sep = name == 'sep_codec'
        with torch.no_grad():
            if sep:
                ref_out, ref_codes = G.drive_sep(config, m, img, label)
                zs = None
            else:
                ref_out, zs, ref_codes, _ = G.drive_dist(config, m, img, label, False)
            out = O.codec_forward(sd, img, label, T, M, separate=sep)
        utils.apply_fn(m, 'free_hidden')          # models/dist_codec.py:64 (the harness restates the loop without it)
        for t in range(T):
            check(torch.equal(out['code_pm1'][t], ref_codes[t]), '{} eval codes t={}'.format(name, t))
            check((out['img'][t] - ref_out['img'][t]).abs().max().item() <= 1e-6, '{} eval recon t={}'.format(name, t))
            # (the sep harness returns per-iteration arrays; sep_codec.py:63-64 concatenates them like dist_codec)
            want = sum(c.nbytes for c in ref_out['code'][:t + 1]) if sep else ref_out['code'][t].nbytes
            check(out['code'][t].nbytes == want, '{} packed size t={}'.format(name, t))
            if zs is not None:
                check((out['z'][t] - zs[t]).abs().max().item() <= 1e-6, '{} z t={}'.format(name, t))
        check(abs(float(out['loss']) - float(ref_out['loss'])) <= 1e-6, name + ' eval loss')
        if sep:
            print('ok', name, data, 'M={} T={} B={} (eval)'.format(M, T, Bn))
            continue
        # training: stochastic binarizer with the reference's own uniforms, loss and gradients through autograd
        m.zero_grad()
        ref_out, zs, ref_codes, noises = G.drive_dist(config, m, img, label, True, noise_seed=seed + 7)
        ref_out['loss'].backward()
        utils.apply_fn(m, 'free_hidden')
        leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        out = O.codec_forward(leaf, img, label, T, M, training=True, noise=noises)
        out['loss'].backward()
        for t in range(T):
            check(torch.equal(out['code_pm1'][t].detach(), ref_codes[t]), '{} train codes t={}'.format(name, t))
        check(abs(float(out['loss'].detach()) - float(ref_out['loss'].detach())) <= 1e-6, name + ' train loss')
        for k, p in m.named_parameters():
            if p.grad is None:
                check(leaf[k].grad is None or float(leaf[k].grad.abs().max()) == 0.0, 'grad of unused ' + k)
                continue
            den = float(p.grad.norm()) + 1e-20
            check(float((leaf[k].grad - p.grad).norm()) / den <= 1e-4, 'grad ' + k)
        print('ok', name, data, 'M={} T={} B={}'.format(M, T, Bn))
    print('live check passed')


if __name__ == '__main__':
    main()
