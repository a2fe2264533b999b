"""GPU diagnostics (development aid, not part of the product path): runs one stage per process so a
faulting kernel cannot poison later stages.  usage: python tools/gpu_diag.py <stage> ; stages: see STAGES."""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = os.path.join(ROOT, 'drasic-distributed-recurrent-autoencoder-for-scalable-image-compression_b200')
sys.path.insert(0, ROOT)
sys.path.insert(0, PKG)

from oracle import drasic_oracle as O  # noqa: E402
from engine import binding as B        # noqa: E402
from engine.cell import CellRunner     # noqa: E402
from engine.codec import CodecEngine   # noqa: E402


def stage_info():
    L = B.lib()
    print('abi', L.drasic_version(), 'sms', L.drasic_device_sms(), torch.cuda.get_device_name(0))


def cell_case(cin, co, hw, n, precision, steps=3, wscale=1.0, seed=0):
    import torch.nn as nn
    torch.manual_seed(seed)
    w_in = nn.Conv2d(cin, 4 * co, 3, 1, 1, bias=False).weight.detach() * wscale
    w_h = nn.Conv2d(co, 4 * co, 3, 1, 1, bias=False).weight.detach() * wscale
    g = torch.Generator().manual_seed(seed + 1)
    run = CellRunner(precision)
    state64 = None
    state32 = None
    res = []
    for t in range(steps):
        x = torch.rand(n, cin, hw, hw, generator=g) * 2 - 1
        h64, state64 = O.convlstm_cell(x.double(), state64, w_in.double(), w_h.double())
        h32, state32 = O.convlstm_cell(x, state32, w_in, w_h)
        h = run.step(x.cuda(), w_in.cuda(), w_h.cuda())
        torch.cuda.synchronize()
        hid = run.hidden_nchw()
        eh = (h.cpu().double() - h64).abs().max().item()
        ec = (hid[1][0].cpu().double() - state64[1]).abs().max().item()
        eh32 = (h32.double() - h64).abs().max().item()
        bias = (h.cpu().double() - h64).mean().item()
        res.append((t, eh, ec, eh32, bias))
    return res


def stage_cell_small():
    for prec in ('parity', 'fast'):
        r = cell_case(128, 128, 4, 8, prec)
        print('cell 128->128 4x4 n=8', prec, ['t%d eh=%.2e ec=%.2e (fp32 ref err %.1e) bias=%.1e' % x for x in r])


def stage_cell_shapes():
    for (cin, co, hw, n) in [(128, 128, 16, 4), (512, 128, 8, 4), (512, 128, 4, 16), (128, 512, 4, 16),
                             (128, 512, 8, 4), (128, 128, 16, 3), (64, 64, 8, 2), (128, 128, 32, 1)]:
        for prec in ('parity', 'fast'):
            try:
                r = cell_case(cin, co, hw, n, prec)
                print('cell %d->%d %dx%d n=%d' % (cin, co, hw, hw, n), prec,
                      ['t%d eh=%.2e ec=%.2e (fp32 %.1e) bias=%.1e' % x for x in r])
            except Exception:
                traceback.print_exc()
    r = cell_case(128, 128, 8, 4, 'parity', wscale=4.0)
    print('cell stress x4 weights', ['t%d eh=%.2e ec=%.2e (fp32 %.1e) bias=%.1e' % x for x in r])


def loop_case(tag, num_channel, num_node, T, B_, seed, precision, separate=False, img=None, label=None, wscale=1.0):
    sd = O.init_state_dict(num_channel=num_channel, code_size=8, num_node=num_node, separate=separate, seed=seed)
    if wscale != 1.0:
        sd = {k: v * wscale for k, v in sd.items()}
    if img is None:
        g = torch.Generator().manual_seed(seed + 100)
        img = torch.rand(B_, num_channel, 32, 32, generator=g)
        label = torch.randint(0, num_node, (B_,), generator=g)
    with torch.no_grad():
        ref = O.codec_forward(sd, img, label, T, num_node, separate=separate)
    eng = CodecEngine(num_channel, 8, num_node, separate=separate, precision=precision)
    sdc = {k: v.cuda() for k, v in sd.items()}
    out = eng.forward(sdc, img.cuda(), label.cuda(), T, want_z=True)
    torch.cuda.synchronize()
    codes = out['codes'].cpu()
    refc = torch.stack(ref['code_pm1'])
    refz = torch.stack(ref['z'])
    mism = (codes != refc).flatten(1).sum(1).tolist()
    zerr = (out['z'].cpu() - refz).abs().flatten(1).max(1)[0].tolist()
    rerr = (out['recon'].cpu() - torch.stack(ref['img'])).abs().flatten(1).max(1)[0].tolist()
    loss = CodecEngine.loss_from_acc(out['loss_acc'], img.numel()).item()
    print(tag, precision, 'mismatching bits per iter', mism)
    print('   z err per iter', ['%.1e' % e for e in zerr])
    print('   recon err per iter', ['%.1e' % e for e in rerr])
    print('   loss %.8f ref %.8f  launches %d' % (loss, ref['loss'].item(), eng.launches))
    # margins of mismatching bits
    bad = (codes != refc)
    if bad.any():
        print('   |z_ref| at mismatches:', refz[bad].abs()[:20].tolist())
    # real bit pack check
    import numpy as _np
    pb = _np.stack([O.pack_bits(c) for c in ref['code_pm1']])
    print('   bits equal to packbits(code>0):', bool(_np.array_equal(out['bits'].cpu().numpy()[~bad.flatten(2).any(2).numpy()],
                                                                    pb[~bad.flatten(2).any(2).numpy()])))
    return out


def stage_loop_small():
    g = np.load(os.path.join(ROOT, 'tests/golden/loop_mnist_m1.npz'))
    for prec in ('parity', 'fast'):
        loop_case('mnist_m1 golden', 1, 1, 16, 4, 0, prec, img=torch.from_numpy(g['img']), label=torch.from_numpy(g['label']))
    g = np.load(os.path.join(ROOT, 'tests/golden/loop_cifar_m2.npz'))
    loop_case('cifar_m2 golden', 3, 2, 4, 6, 3, 'parity', img=torch.from_numpy(g['img']), label=torch.from_numpy(g['label']))
    g = np.load(os.path.join(ROOT, 'tests/golden/loop_sep_m2.npz'))
    loop_case('sep_m2 golden', 3, 2, 3, 5, 6, 'parity', separate=True, img=torch.from_numpy(g['img']), label=torch.from_numpy(g['label']))


def stage_loop_medium():
    loop_case('cifar_m8 B=64 T=16', 3, 8, 16, 64, 0, 'parity')
    loop_case('cifar_m8 B=64 T=16 weights x3', 3, 8, 16, 64, 0, 'parity', wscale=3.0)


def stage_bench():
    for prec in ('parity', 'fast'):
        for (M, Bn) in [(8, 128), (8, 1024), (1, 1024), (1, 4096)]:
            sd = {k: v.cuda() for k, v in O.init_state_dict(3, 8, M, seed=0).items()}
            eng = CodecEngine(3, 8, M, precision=prec)
            g = torch.Generator().manual_seed(1)
            img = torch.rand(Bn, 3, 32, 32, generator=g).cuda()
            lab_h = torch.randint(0, M, (Bn,), generator=g)
            lab = lab_h.cuda()
            for _ in range(2):
                eng.forward(sd, img, lab, 16, label_host=lab_h.numpy())
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            t0 = time.perf_counter()
            e0.record()
            for _ in range(reps):
                eng.forward(sd, img, lab, 16, label_host=lab_h.numpy())
            e1.record()
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) / reps
            ms = e0.elapsed_time(e1) / reps
            print('bench %s M=%d B=%d: %.2f ms/forward (wall %.2f ms) -> %.0f img/s, %.1f TFLOP/s algorithmic' %
                  (prec, M, Bn, ms, wall * 1e3, Bn / ms * 1e3, 57.083e9 * Bn / ms / 1e9))


def train_case(tag, num_channel, num_node, T, B_, seed, precision, separate=False, img=None, label=None, noise=None, gp='fp16'):
    sd = O.init_state_dict(num_channel=num_channel, code_size=8, num_node=num_node, separate=separate, seed=seed)
    if img is None:
        g = torch.Generator().manual_seed(seed + 100)
        img = torch.rand(B_, num_channel, 32, 32, generator=g)
        label = torch.randint(0, num_node, (B_,), generator=g)
        noise = torch.rand(T, B_, 8, 4, 4, generator=g)
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    t0 = time.time()
    ref = O.codec_forward(sdg, img, label, T, num_node, training=True, noise=[n for n in noise], separate=separate)
    ref['loss'].backward()
    t_ref = time.time() - t0
    eng = CodecEngine(num_channel, 8, num_node, separate=separate, precision=precision, grad_precision=gp)
    sdc = {k: v.cuda() for k, v in sd.items()}
    out = eng.forward(sdc, img.cuda(), label.cuda(), T, training=True, noise=noise.cuda(), save_for_backward=True)
    grads = eng.backward(out)
    torch.cuda.synchronize()
    loss = CodecEngine.loss_from_acc(out['loss_acc'], img.numel()).item()
    mism = int((out['codes'].cpu() != torch.stack(ref['code_pm1'])).sum())
    print(tag, precision, gp, 'loss %.8f ref %.8f code mismatches %d  (oracle fwd+bwd %.1fs)' % (loss, ref['loss'].item(), mism, t_ref))
    worst = []
    for k, v in sdg.items():
        gr = v.grad.double() if v.grad is not None else torch.zeros_like(v).double()
        gg = grads[k].cpu().double()
        rel = ((gg - gr).norm() / max(gr.norm().item(), 1e-30)).item()
        cos = (torch.dot(gg.flatten(), gr.flatten()) / max(gg.norm().item() * gr.norm().item(), 1e-30)).item()
        worst.append((rel, cos, k, gr.norm().item(), gg.norm().item()))
    worst.sort(reverse=True)
    for rel, cos, k, nr, ng in worst[:10]:
        print('   rel_err %.2e cos %.6f |ref| %.3e |ours| %.3e  %s' % (rel, cos, nr, ng, k))
    print('   median rel_err %.2e over %d tensors' % (sorted(w[0] for w in worst)[len(worst) // 2], len(worst)))


def stage_train_small():
    g = np.load(os.path.join(ROOT, 'tests/golden/loop_cifar_m2.npz'))
    for prec in ('parity', 'fast'):
        train_case('cifar_m2 golden train', 3, 2, 4, 6, 3, prec, img=torch.from_numpy(g['img']),
                   label=torch.from_numpy(g['label']), noise=torch.from_numpy(g['train_noise']))
    train_case('cifar_m2 golden train', 3, 2, 4, 6, 3, 'parity', img=torch.from_numpy(g['img']),
               label=torch.from_numpy(g['label']), noise=torch.from_numpy(g['train_noise']), gp='parity')
    train_case('mnist_m1 T=3', 1, 1, 3, 8, 1, 'parity')
    train_case('sep_m2 T=3', 3, 2, 3, 10, 2, 'parity', separate=True)


def stage_train_medium():
    for prec, gp in (('parity', 'fp16'), ('parity', 'parity'), ('fast', 'fp16')):
        train_case('cifar_m8 B=32 T=8', 3, 8, 8, 32, 5, prec, gp=gp)


def stage_train_bench():
    for prec, gp in (('parity', 'fp16'), ('parity', 'parity'), ('fast', 'fp16')):
        for (M, Bn) in [(8, 1024)]:
            sd = {k: v.cuda() for k, v in O.init_state_dict(3, 8, M, seed=0).items()}
            eng = CodecEngine(3, 8, M, precision=prec, grad_precision=gp)
            g = torch.Generator().manual_seed(1)
            img = torch.rand(Bn, 3, 32, 32, generator=g).cuda()
            lab_h = torch.randint(0, M, (Bn,), generator=g)
            lab = lab_h.cuda()
            def step():
                out = eng.forward(sd, img, lab, 16, training=True, label_host=lab_h.numpy(), save_for_backward=True)
                return eng.backward(out)
            step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            eng.profile = []
            e0.record()
            step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            prof = eng.profile_summary()
            eng.profile = None
            print('train bench %s/%s M=%d B=%d: %.1f ms/step -> %.0f img/s (%.1f algorithmic TFLOP/s) mem %.1f GB' %
                  (prec, gp, M, Bn, ms, Bn / ms * 1e3, 3 * 57.083e9 * Bn / ms / 1e9, torch.cuda.max_memory_allocated() / 2**30))
            grp = {}
            for k, (n, t, fl) in prof.items():
                kk = k.split('_')[0]
                a = grp.setdefault(kk, [0, 0.0, 0.0]); a[0] += n; a[1] += t; a[2] += fl
            print('   ', {k: '%d launches %.1f ms %.0f TF/s' % (v[0], v[1], v[2] / max(v[1], 1e-9) / 1e9) for k, v in grp.items()})
            del eng
            torch.cuda.empty_cache()


STAGES = {'info': stage_info, 'train_small': stage_train_small, 'train_medium': stage_train_medium,
          'train_bench': stage_train_bench, 'cell_small': stage_cell_small, 'cell_shapes': stage_cell_shapes,
          'loop_small': stage_loop_small, 'loop_medium': stage_loop_medium, 'bench': stage_bench}

if __name__ == '__main__':
    name = sys.argv[1]
    t0 = time.time()
    print('==== stage', name, flush=True)
    try:
        STAGES[name]()
    except Exception:
        traceback.print_exc()
        print('==== stage', name, 'FAILED after %.1fs' % (time.time() - t0), flush=True)
        sys.exit(1)
    print('==== stage', name, 'ok %.1fs' % (time.time() - t0), flush=True)
