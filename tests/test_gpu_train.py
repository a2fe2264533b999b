"""GPU parity of the training step (forward with the stochastic binarizer + backward-through-time)
through the reference-facing API: `output = model(input); output['loss'].backward()`
(train_model.py:113-115), against gradients the reference produced (golden fixture) and the CPU oracle."""
import os

import numpy as np
import pytest
import torch

from conftest import build_model
from oracle import drasic_oracle as O

pytestmark = pytest.mark.gpu

# stated tolerances: relative L2 error of every parameter's gradient against the fp32 reference
# (fp16 backward: fp16 operands / saved gates, errors compound over 6 layers x T iterations of BPTT and
# do not average out for a node that saw only a handful of samples; median over tensors is ~6e-4)
GRAD_TOL = {'fp16': 1e-2, 'parity': 5e-4}
GRAD_TOL_MEDIAN = 2e-3


def rel_err(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / max(b.norm().item(), 1e-30)).item()


def train_step(model, img, label, noise):
    model.train(True)
    model.zero_grad()
    out = model({'img': img.cuda(), 'label': label.cuda()}, noise=noise.cuda())
    out['loss'].backward()
    return out


@pytest.mark.parametrize('precision,grad_precision', [('parity', 'fp16'), ('parity', 'parity'), ('fast', 'fp16')])
def test_training_step_matches_reference_gradients(golden_dir, precision, grad_precision):
    import config
    g = np.load(os.path.join(golden_dir, 'loop_cifar_m2.npz'))
    m = build_model('dist_codec', 'CIFAR10', 2, 4, seed=3, precision=precision)
    config.PARAM['grad_precision'] = grad_precision
    img, label = torch.from_numpy(g['img']), torch.from_numpy(g['label'])
    out = train_step(m, img, label, torch.from_numpy(g['train_noise']))
    assert out['loss'].requires_grad is False or out['loss'].grad_fn is not None
    if precision == 'parity':
        assert np.array_equal(out['code_pm1'].cpu().numpy().astype(np.int8), g['train_codes'])
    assert abs(out['loss'].item() - float(g['train_loss'])) < 1e-5
    tol = GRAD_TOL[grad_precision]
    norms = {str(k): v for k, v in zip(g['grad_keys'], g['grad_norms'])}
    named = dict(m.named_parameters())
    assert set(named) == set(norms)
    for k, p in named.items():
        assert p.grad is not None and p.grad.shape == p.shape, k
        assert abs(p.grad.double().norm().item() - norms[k]) <= tol * norms[k], k
    full = {'g_e0_w': 'model.encoder.0.0.cell.cell.main.weight', 'g_e0_b': 'model.encoder.1.0.cell.cell.main.bias',
            'g_e4_w': 'model.encoder.1.7.cell.cell.main.weight', 'g_d0_w': 'model.decoder.0.cell.cell.main.weight',
            'g_d4_w': 'model.decoder.7.cell.cell.main.weight'}
    for short, k in full.items():
        assert rel_err(named[k].grad.cpu(), torch.from_numpy(g[short])) < tol, k
    assert rel_err(named['model.decoder.5.cell.cell.0.in.cell.cell.main.weight'].grad[:8, :8].cpu(),
                   torch.from_numpy(g['g_d3_in_slice'])) < tol
    assert rel_err(named['model.encoder.0.2.cell.cell.0.hidden.cell.cell.main.weight'].grad[:8, :8].cpu(),
                   torch.from_numpy(g['g_e1_hidden_slice'])) < tol


@pytest.mark.parametrize('name,M,B,T,separate,cta_pair', [('dist_codec', 8, 27, 5, False, 'off'), ('sep_codec', 3, 13, 3, True, 'off'),
                                                           ('dist_codec', 1, 9, 16, False, 'off'), ('dist_codec', 8, 27, 5, False, 'on'),
                                                           ('sep_codec', 3, 13, 3, True, 'on')])
def test_training_step_vs_oracle_autograd(name, M, B, T, separate, cta_pair):
    """ragged node groups (one node empty), per-source decoders (sep: detached residual), T=16 single source"""
    data = 'MNIST' if M == 1 else 'CIFAR10'
    C = 1 if M == 1 else 3
    m = build_model(name, data, M, T, seed=11, cta_pair=cta_pair)
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(12)
    img = torch.rand(B, C, 32, 32, generator=g)
    label = torch.randint(0, M, (B,), generator=g)
    if M > 2:
        label[label == 1] = 0                                       # node 1 sees no data in this batch
    noise = torch.rand(T, B, 8, 4, 4, generator=g)
    ref = O.codec_forward(sd, img, label, T, M, training=True, noise=[n for n in noise], separate=separate)
    ref['loss'].backward()
    out = train_step(m, img, label, noise)
    assert np.array_equal(out['code_pm1'].cpu().numpy(), torch.stack(ref['code_pm1']).numpy())
    assert abs(out['loss'].item() - ref['loss'].item()) < 1e-5
    errs = []
    for k, p in m.named_parameters():
        if sd[k].grad is None:                                      # parameters of the empty node
            assert p.grad is not None and float(p.grad.abs().max()) == 0.0, k
        else:
            errs.append(rel_err(p.grad.cpu(), sd[k].grad))
            assert errs[-1] < GRAD_TOL['fp16'], k
    assert sorted(errs)[len(errs) // 2] < GRAD_TOL_MEDIAN


def test_optimizer_trajectory_follows_oracle_and_weights_are_repacked():
    """SGD updates the fp32 master parameters in place (train_model.py:116); the packed fp16 weight planes
    are a cache keyed on the parameter version and must follow.  Three training steps with given binarizer
    noise must track the same three steps done by the CPU oracle."""
    m = build_model('dist_codec', 'CIFAR10', 2, 3, seed=5)
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    opt = torch.optim.SGD(m.parameters(), lr=0.5)
    opt_ref = torch.optim.SGD(list(sd.values()), lr=0.5)
    g = torch.Generator().manual_seed(6)
    img = torch.rand(8, 3, 32, 32, generator=g)
    label = torch.randint(0, 2, (8,), generator=g)
    for it in range(3):
        noise = torch.rand(3, 8, 8, 4, 4, generator=g)
        opt_ref.zero_grad()
        ref = O.codec_forward(sd, img, label, 3, 2, training=True, noise=[n for n in noise])
        ref['loss'].backward()
        opt_ref.step()
        m.train(True)
        opt.zero_grad()
        out = m({'img': img.cuda(), 'label': label.cuda()}, noise=noise.cuda())
        out['loss'].backward()
        opt.step()
        assert abs(out['loss'].item() - ref['loss'].item()) < 2e-5, it
        assert (out['code_pm1'].cpu() != torch.stack(ref['code_pm1'])).float().mean().item() < 2e-3
    for k, p in m.named_parameters():
        assert rel_err(p.detach().cpu(), sd[k].detach()) < 1e-4, k
    # eval after training equals the oracle on the *updated* weights (cache was refreshed)
    sdn = {k: v.detach() for k, v in sd.items()}
    with torch.no_grad():
        ref = O.codec_forward(sdn, img, label, 3, 2)
    m.train(False)
    with torch.no_grad():
        out = m({'img': img.cuda(), 'label': label.cuda()})
    assert abs(out['loss'].item() - ref['loss'].item()) < 1e-4
    assert (out['code_pm1'].cpu() != torch.stack(ref['code_pm1'])).float().mean().item() < 5e-3


def test_fused_adam_updates_reach_the_packed_weights():
    """torch.optim.Adam(fused=True) (train_model.py:161 on a CUDA box) updates parameters in place WITHOUT
    bumping Tensor._version; the packed fp16 weight planes must follow anyway.  After two steps the model
    must behave exactly like a fresh model (fresh engine, fresh packing) loaded with its state_dict, and
    its loss trajectory must track the CPU oracle's Adam."""
    m = build_model('dist_codec', 'CIFAR10', 2, 3, seed=15)
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, fused=True)
    opt_ref = torch.optim.Adam(list(sd.values()), lr=1e-3)
    g = torch.Generator().manual_seed(16)
    img = torch.rand(8, 3, 32, 32, generator=g)
    label = torch.randint(0, 2, (8,), generator=g)
    inp = {'img': img.cuda(), 'label': label.cuda()}
    v0 = [p._version for p in m.parameters()]
    losses = []
    for it in range(3):
        noise = torch.rand(3, 8, 8, 4, 4, generator=g)
        opt_ref.zero_grad()
        ref = O.codec_forward(sd, img, label, 3, 2, training=True, noise=[n for n in noise])
        ref['loss'].backward()
        opt_ref.step()
        m.train(True)
        opt.zero_grad()
        out = m(inp, noise=noise.cuda())
        out['loss'].backward()
        opt.step()
        losses.append((out['loss'].item(), ref['loss'].item()))
    assert [p._version for p in m.parameters()] == v0          # the premise: versions did not move
    assert losses[0][0] != losses[2][0]                         # ... while the weights did
    for a, b in losses:
        assert abs(a - b) < 5e-4, losses
    m.train(False)
    with torch.no_grad():
        out = m(inp)
    m2 = build_model('dist_codec', 'CIFAR10', 2, 3, seed=99)
    m2.load_state_dict(m.state_dict())
    m2.train(False)
    with torch.no_grad():
        out2 = m2(inp)
    assert torch.equal(out['code_pm1'], out2['code_pm1']) and torch.equal(out['img'][-1], out2['img'][-1])
    # the decoder-only path sees the updated weights too
    rec = m.decode(out['bits'], inp['label'])
    assert torch.equal(rec[-1], out['img'][-1])


def test_release_frees_arenas_and_next_call_rebuilds():
    m = build_model('dist_codec', 'CIFAR10', 2, 3, seed=17)
    g = torch.Generator().manual_seed(18)
    img, label = torch.rand(8, 3, 32, 32, generator=g), torch.randint(0, 2, (8,), generator=g)
    noise = torch.rand(3, 8, 8, 4, 4, generator=g)
    out1 = train_step(m, img, label, noise)
    g1 = {k: p.grad.clone() for k, p in m.named_parameters()}
    used = torch.cuda.memory_allocated()
    m.release()
    assert m._engine._arena is None and torch.cuda.memory_allocated() < used
    out2 = train_step(m, img, label, noise)
    assert torch.equal(out1['code_pm1'], out2['code_pm1']) and out1['loss'].item() == out2['loss'].item()
    for k, p in m.named_parameters():
        assert float((p.grad - g1[k]).abs().max()) <= 1e-3 * float(g1[k].abs().max()) + 1e-12, k


def test_backward_requires_matching_forward():
    m = build_model('dist_codec', 'MNIST', 1, 2, seed=0)
    inp = {'img': torch.rand(4, 1, 32, 32).cuda(), 'label': torch.zeros(4, dtype=torch.long).cuda()}
    m.train(True)
    out1 = m(inp)
    out2 = m(inp)                                                   # overwrites the saved per-iteration state
    with pytest.raises(RuntimeError, match='must follow the forward'):
        out1['loss'].backward()
    out2['loss'].backward()


def test_cuda_graph_matches_eager_eval_and_train():
    """config.PARAM['cuda_graph']: the whole loop (and backward-through-time) replayed as one CUDA graph
    must reproduce the eager launches bit for bit, across batches with different node assignments and
    across optimizer updates of the weights."""
    import config
    m = build_model('dist_codec', 'CIFAR10', 3, 4, seed=7)
    g = torch.Generator().manual_seed(8)
    batches = [(torch.rand(24, 3, 32, 32, generator=g), torch.randint(0, 3, (24,), generator=g)) for _ in range(3)]
    batches[2] = (batches[2][0], torch.zeros(24, dtype=torch.long))          # all samples on one node
    noise = torch.rand(4, 24, 8, 4, 4, generator=g)
    # ---- eval
    eager = []
    m.train(False)
    for img, label in batches:
        with torch.no_grad():
            eager.append(m({'img': img.cuda(), 'label': label.cuda()}))
    config.PARAM['cuda_graph'] = True
    graphed = []
    for (img, label), ref in zip(batches, eager):
        with torch.no_grad():
            out = m({'img': img.cuda(), 'label': label.cuda()})
        assert torch.equal(out['code_pm1'], ref['code_pm1'])
        assert torch.equal(torch.stack(out['img']), torch.stack(ref['img']))
        assert out['loss'].item() == ref['loss'].item()
        graphed.append(out)
    # the code lists are built lazily from copies of the graph's output buffers: later replays must not leak into them
    for out, ref in zip(graphed, eager):
        assert torch.equal(out['code_pm1'], ref['code_pm1']) and torch.equal(out['bits'], ref['bits'])
        assert len(out['code']) == len(ref['code'])
        assert all(np.array_equal(a, b) for a, b in zip(out['code'], ref['code']))
    # ---- train: gradients from the graph == eager gradients; then an optimizer step and again
    opt = torch.optim.SGD(m.parameters(), lr=0.05)
    for step in range(2):
        grads = {}
        for mode in (False, True):
            config.PARAM['cuda_graph'] = mode
            m.train(True)
            m.zero_grad()
            img, label = batches[step]
            out = m({'img': img.cuda(), 'label': label.cuda()}, noise=noise.cuda())
            out['loss'].backward()
            grads[mode] = ({k: p.grad.clone() for k, p in m.named_parameters()}, out['loss'].item(), out['code_pm1'].clone())
        assert grads[True][1] == grads[False][1]
        assert torch.equal(grads[True][2], grads[False][2])
        for k in grads[True][0]:
            # fp32 atomics in the weight-gradient accumulation make the sums order dependent at the 1e-6 level
            a, b = grads[True][0][k], grads[False][0][k]
            # ... and the graph's static plan has another slot count, hence other K cuts in the stream-K GEMMs
            assert (a - b).norm().item() <= 1e-3 * max(b.norm().item(), 1e-30), k
        opt.step()
    config.PARAM['cuda_graph'] = False


def test_config2_full_size_training_step_properties():
    """BASELINE configs[1] at the bench size (B=1024, 8 sources by random subsets, T=16): too big for the CPU
    oracle, checked through size-independent properties of one training step."""
    import config
    M, T, B = 8, 16, 1024
    m = build_model('dist_codec', 'CIFAR10', M, T, seed=40)
    g = torch.Generator().manual_seed(41)
    img = torch.rand(B, 3, 32, 32, generator=g)
    label = torch.randint(0, M - 1, (B,), generator=g)            # node M-1 sees no data in this batch
    noise = torch.rand(T, B, 8, 4, 4, generator=g)
    out = train_step(m, img, label, noise)
    codes = out['code_pm1']
    assert bool(((codes == 1) | (codes == -1)).all()) and torch.isfinite(out['loss'])
    # (1) the loss is the mean over iterations of the L1 error of the returned reconstructions (dist_codec.py:58,61)
    l1 = torch.stack([(r - img.cuda()).abs().mean() for r in out['img']]).mean().item()
    assert abs(l1 - out['loss'].item()) < 1e-6
    # (2) every parameter has a finite gradient; the encoder nobody used has exactly zero gradient
    for k, p in m.named_parameters():
        assert p.grad is not None and bool(torch.isfinite(p.grad).all()), k
        if k.startswith('model.encoder.{}.'.format(M - 1)):
            assert float(p.grad.abs().max()) == 0.0, k
        else:
            assert float(p.grad.abs().max()) > 0.0, k
    # (3) samples are independent: the gradient of the global-mean loss is the sample-weighted sum of the gradients
    #     of two halves of the batch (the property the multi-GPU all-reduce relies on)
    ref = {k: p.grad.clone() for k, p in m.named_parameters()}
    acc = {k: torch.zeros_like(v) for k, v in ref.items()}
    for lo, hi in ((0, 384), (384, 1024)):
        o = train_step(m, img[lo:hi], label[lo:hi], noise[:, lo:hi])
        assert torch.equal(o['code_pm1'], codes[:, lo:hi])       # same symbols as inside the big batch
        for k, p in m.named_parameters():
            acc[k] += p.grad * ((hi - lo) / float(B))
    worst = max(float((acc[k] - ref[k]).norm() / (ref[k].norm() + 1e-20)) for k in ref if float(ref[k].abs().max()) > 0)
    assert worst < 2e-3, worst
    # (4) the decoder-only path reproduces the training-mode reconstructions from the emitted bits
    rec = m.decode(out['bits'][:4], label.cuda())
    # (another slot count -> other K cuts in the stream-K GEMMs: equal up to fp32 summation order)
    assert (rec[3] - out['img'][3].detach()).abs().max().item() < 1e-5


@pytest.mark.parametrize('M,code,B', [(10, 16, 21), (2, 4, 6)])
def test_training_other_controls_vs_oracle(M, code, B):
    """training away from the BASELINE point: the reference's default num_node=10, other code sizes"""
    T = 2
    m = build_model('dist_codec', 'CIFAR10', M, T, seed=60, code_size=code)
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(61)
    img = torch.rand(B, 3, 32, 32, generator=g)
    label = torch.randint(0, M, (B,), generator=g)
    noise = torch.rand(T, B, code, 4, 4, generator=g)
    ref = O.codec_forward(sd, img, label, T, M, training=True, noise=[n for n in noise])
    ref['loss'].backward()
    out = train_step(m, img, label, noise)
    assert np.array_equal(out['code_pm1'].cpu().numpy(), torch.stack(ref['code_pm1']).numpy())
    assert abs(out['loss'].item() - ref['loss'].item()) < 1e-5
    for k, p in m.named_parameters():
        if sd[k].grad is None:
            assert float(p.grad.abs().max()) == 0.0, k
        else:
            assert rel_err(p.grad.cpu(), sd[k].grad) < GRAD_TOL['fp16'], k


def test_single_source_pairs_on_the_4x4_layers():
    """one source: node slots are padded to 16 images, so E3 / D1 (one M tile per 8 images) run as CTA pairs as
    well; forward codes and training gradients against the oracle"""
    M, T, B = 1, 3, 48
    m = build_model('dist_codec', 'CIFAR10', M, T, seed=90, cta_pair='on')
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(91)
    img = torch.rand(B, 3, 32, 32, generator=g)
    label = torch.zeros(B, dtype=torch.long)
    noise = torch.rand(T, B, 8, 4, 4, generator=g)
    ref = O.codec_forward(sd, img, label, T, M, training=True, noise=[n for n in noise])
    ref['loss'].backward()
    out = train_step(m, img, label, noise)
    layers = {ly.name: ly.pair for ly in m._engine._last_layers} if hasattr(m._engine, '_last_layers') else None
    assert layers is None or all(layers.values()), layers
    assert np.array_equal(out['code_pm1'].cpu().numpy(), torch.stack(ref['code_pm1']).numpy())
    assert abs(out['loss'].item() - ref['loss'].item()) < 1e-5
    for k, p in m.named_parameters():
        assert rel_err(p.grad.cpu(), sd[k].grad) < GRAD_TOL['fp16'], k
    # eval, deterministic sign
    with torch.no_grad():
        ref_e = O.codec_forward({k: v.detach() for k, v in sd.items()}, img, label, T, M)
    m.train(False)
    with torch.no_grad():
        out_e = m({'img': img.cuda(), 'label': label.cuda()})
    assert np.array_equal(out_e['code_pm1'].cpu().numpy(), torch.stack(ref_e['code_pm1']).numpy())


@pytest.mark.parametrize('prev_amax,redo', [(0.0, True), (1e-9, True), (3e4, True), (0.02, False)])
def test_cell_backward_scale_prediction_and_verify_pass(prev_amax, redo):
    """drasic_lstm_bwd_fused through the C ABI: cell backward (cell.py:133-139 differentiated by autograd on the CPU)
    into fp16 planes whose power-of-two scale is predicted from another tensor's amax.  Whatever the prediction
    (unknown, far too small, far too large, right), after the verify pass planes * inv_scale are the fp32 gate
    gradients to fp16-split precision, and the carry and the true amax are exact."""
    from engine import binding as B
    L = B.lib()
    P, Co = 8192, 128                                   # 512 blocks of work: the verify launch runs on a capped grid
    g = torch.Generator().manual_seed(90)
    pre = (torch.randn(P, 4, Co, generator=g) * 1.5).requires_grad_(True)
    c_prev = torch.randn(P, Co, generator=g).requires_grad_(True)
    i, f, o = torch.sigmoid(pre[:, 0]), torch.sigmoid(pre[:, 1]), torch.sigmoid(pre[:, 3])
    gg = torch.tanh(pre[:, 2])
    c_t = f * c_prev + i * gg
    h = o * torch.tanh(c_t)
    dh_above, dh_rec = torch.randn(P, Co, generator=g) * 1e-2, torch.randn(P, Co, generator=g) * 1e-2
    dc_in = torch.randn(P, Co, generator=g) * 1e-2
    (h * (dh_above + dh_rec)).sum().backward(retain_graph=True, inputs=[pre, c_prev])
    d_pre, d_cprev = pre.grad.clone(), c_prev.grad.clone()
    pre.grad = c_prev.grad = None
    (c_t * dc_in).sum().backward(inputs=[pre, c_prev])
    d_pre += pre.grad
    d_cprev += c_prev.grad
    gates = torch.stack([i, f, gg, o], 1).detach().cuda().contiguous()       # activated gates, fp32
    dev = dict(device='cuda')
    hi, lo = torch.zeros(P, 4, Co, dtype=torch.float16, **dev), torch.zeros(P, 4, Co, dtype=torch.float16, **dev)
    dc_out = torch.zeros(P, Co, **dev)
    prev = torch.tensor([prev_amax], dtype=torch.float32, **dev).view(torch.int32)
    amax = torch.zeros(1, dtype=torch.int32, **dev)
    inv = torch.zeros(1, **dev)
    t = [x.detach().cuda().contiguous() for x in (c_t, c_prev, dh_above, dh_rec, dc_in)]
    a = B.LstmBwdFusedArgs()
    a.gates, a.c_t, a.c_prev, a.dh_above, a.dh_rec, a.dc_in = [B.ptr(x) for x in [gates] + t]
    a.dc_out, a.dg_hi, a.dg_lo = B.ptr(dc_out), B.ptr(hi), B.ptr(lo)
    a.prev_amax_bits, a.amax_bits, a.inv_scale_out = B.ptr(prev), B.ptr(amax), B.ptr(inv)
    a.n_pix, a.Co, a.gates_f32 = P, Co, 1
    stream = torch.cuda.current_stream().cuda_stream
    a.verify = 0
    B.check(L.drasic_lstm_bwd_fused(a, stream), 'lstm_bwd_fused')
    inv_first = float(inv)
    a.verify = 1
    B.check(L.drasic_lstm_bwd_fused(a, stream), 'lstm_bwd_fused')
    torch.cuda.synchronize()
    true_amax = float(d_pre.abs().max())
    assert abs(float(amax.view(torch.float32)) - true_amax) <= 1e-5 * true_amax
    assert (float(inv) != inv_first) == redo
    scaled = true_amax / float(inv)
    assert 16.0 <= scaled <= 60000.0                     # the planes sit inside fp16's normal range
    got = (hi.float() + lo.float()) * inv
    assert (got.cpu() - d_pre).abs().max().item() <= 2e-6 * true_amax + 1e-9
    assert (dc_out.cpu() - d_cprev).abs().max().item() <= 1e-6 * float(d_cprev.abs().max())


def test_wgrad_wide_pair_tiles_agree_with_the_default_tiles():
    """the opt-in 512-column wgrad pair tiles (DRASIC_WGRAD_PAIR=2; used where a node has >= 256 K blocks) give the
    same weight gradients as the default 256-column tiles, and both match the oracle's autograd"""
    M, B_, T = 1, 72, 2
    m = build_model('dist_codec', 'CIFAR10', M, T, seed=21)
    g = torch.Generator().manual_seed(22)
    img = torch.rand(B_, 3, 32, 32, generator=g)
    label = torch.zeros(B_, dtype=torch.long)
    noise = torch.rand(T, B_, 8, 4, 4, generator=g)
    grads = {}
    for mode in (1, 2):
        m.zero_grad()
        m.train(True)
        if m._engine is not None:
            m._engine.wgrad_pair_mode = mode
        out = m({'img': img.cuda(), 'label': label.cuda()}, noise=noise.cuda())
        m._engine.wgrad_pair_mode = mode          # (the engine exists after the first forward)
        out['loss'].backward()
        grads[mode] = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    ref = O.codec_forward(sd, img, label, T, M, training=True, noise=[n for n in noise])
    ref['loss'].backward()
    for k in grads[1]:
        assert rel_err(grads[2][k].cpu(), grads[1][k].cpu()) < 2e-4, k       # fp32 accumulation order only
        assert rel_err(grads[2][k].cpu(), sd[k].grad) < GRAD_TOL['fp16'], k
