"""GPU parity at the BASELINE configurations' own sizes (round-1 VERDICT "untested-config gaps"): config 1 at
B=64 / T=16, config 4 at M=8 with T=16 / 8, config 5 with an oracle spot check of a 64-patch subset of the 4096,
config 2 training at M=8 / T=16 against the oracle's autograd, the mse / bce loss gradients, 16x16 inputs whose
smallest layer has 4 pixels per image, and the captured-graph buffer ownership."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, build_model
from oracle import drasic_oracle as O
from test_gpu_loop import MARGIN_TOL, PSNR_TOL, RECON_TOL, first_mismatch_margins, run
from test_gpu_train import GRAD_TOL, GRAD_TOL_MEDIAN, rel_err, train_step

pytestmark = pytest.mark.gpu


def _note_flips(name, symbols, margins):
    """the observed flip rate of every oracle-sized comparison goes to the test output and, when the directory
    exists, to gpurun_out/parity_flips.jsonl (copied to profiles/ per round)"""
    rec = {'test': name, 'symbols': int(symbols), 'images_with_a_flip': len(margins), 'margins': [float(m) for m in margins]}
    print('parity flips:', json.dumps(rec))
    d = os.path.join(ROOT, 'gpurun_out')
    if os.path.isdir(d):
        with open(os.path.join(d, 'parity_flips.jsonl'), 'a') as f:
            f.write(json.dumps(rec) + '\n')


def _compare_eval(m, img, label, T, M, name, node_iters=None):
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    with torch.no_grad():
        ref = O.codec_forward(sd, img, label, T, M, node_iters=node_iters)
    out = run(m, img, label)
    codes, refc, refz = out['code_pm1'].cpu(), torch.stack(ref['code_pm1']), torch.stack(ref['z'])
    if node_iters is not None:       # symbols a node emits after its budget is spent are not part of the stream
        active = torch.tensor([[t < node_iters[int(l)] for l in label] for t in range(T)])
        codes = torch.where(active[:, :, None, None, None], codes, refc)
    B = img.shape[0]
    margins, clean = first_mismatch_margins(codes, refc, refz)
    _note_flips(name, refc.numel(), margins)
    assert all(mg < MARGIN_TOL for mg in margins), margins
    assert clean >= B - 2
    ok = torch.tensor([bool((codes[:, b] == refc[:, b]).all()) for b in range(B)])
    err = (torch.stack(out['img']).cpu()[:, ok] - torch.stack(ref['img'])[:, ok]).abs().max().item()
    assert err < RECON_TOL
    ps = np.array([O.psnr(r.cpu(), img) for r in out['img']])
    ps_ref = np.array([O.psnr(r, img) for r in ref['img']])
    assert np.abs(ps - ps_ref).max() < PSNR_TOL
    assert abs(out['loss'].item() - ref['loss'].item()) < 1e-4
    return out, ref


def test_config1_mnist_single_source_b64_t16_vs_oracle():
    """BASELINE configs[0] at its own size: MNIST-shaped (1x32x32 after the reference's 28->32 resize, data.py:21-24),
    single source, eval, batch 64, T=16"""
    m = build_model('dist_codec', 'MNIST', 1, 16, seed=0)
    g = torch.Generator().manual_seed(0)
    img = torch.rand(64, 1, 32, 32, generator=g)
    _compare_eval(m, img, torch.zeros(64, dtype=torch.long), 16, 1, 'config1_b64_t16')


def test_config4_m8_t16_t8_vs_oracle():
    """BASELINE configs[3] at its own size: 8 sources, T=16 for the first half, T=8 for the second half"""
    import config
    M, T, B = 8, 16, 64
    m = build_model('dist_codec', 'CIFAR10', M, T, seed=7)
    iters = [16] * 4 + [8] * 4
    config.PARAM['node_iters'] = iters
    try:
        g = torch.Generator().manual_seed(70)
        img = torch.rand(B, 3, 32, 32, generator=g)
        label = torch.randint(0, M, (B,), generator=g)
        out, _ = _compare_eval(m, img, label, T, M, 'config4_m8_t16_t8', node_iters=iters)
        rec = torch.stack(out['img']).cpu()
        frozen = label >= 4
        assert torch.equal(rec[7][frozen], rec[15][frozen]) and not torch.equal(rec[7][~frozen], rec[15][~frozen])
    finally:
        config.PARAM.pop('node_iters', None)


def test_config5_4096_patches_oracle_spot_check():
    """BASELINE configs[4]: 4096 patches of 768x512 images, single source, T=16, eval -- the whole batch on the GPU,
    the oracle on 64 of the patches (samples are independent through the loop)"""
    m = build_model('dist_codec', 'CIFAR10', 1, 16, seed=0)
    g = torch.Generator().manual_seed(5)
    big = torch.rand(11, 3, 512, 768, generator=g)
    patches = big.unfold(2, 32, 32).unfold(3, 32, 32).permute(0, 2, 3, 1, 4, 5).reshape(-1, 3, 32, 32)[:4096].contiguous()
    label = torch.zeros(4096, dtype=torch.long)
    out = run(m, patches, label)
    sub = torch.arange(17, 4096, 64)[:64]                               # spread over every 128-image tile
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    with torch.no_grad():
        ref = O.codec_forward(sd, patches[sub], label[sub], 16, 1)
    codes, refc, refz = out['code_pm1'].cpu()[:, sub], torch.stack(ref['code_pm1']), torch.stack(ref['z'])
    margins, clean = first_mismatch_margins(codes, refc, refz)
    _note_flips('config5_4096_subset64', refc.numel(), margins)
    assert all(mg < MARGIN_TOL for mg in margins), margins
    assert clean >= 62
    ok = torch.tensor([bool((codes[:, b] == refc[:, b]).all()) for b in range(64)])
    err = (torch.stack(out['img']).cpu()[:, sub][:, ok] - torch.stack(ref['img'])[:, ok]).abs().max().item()
    assert err < RECON_TOL


def test_config2_training_m8_t16_vs_oracle_autograd():
    """BASELINE configs[1] at its own depth: 8 sources, T=16, stochastic binarizer (given noise), full BPTT"""
    M, T, B = 8, 16, 32
    m = build_model('dist_codec', 'CIFAR10', M, T, seed=13)
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(14)
    img = torch.rand(B, 3, 32, 32, generator=g)
    label = torch.randint(0, M, (B,), generator=g)
    noise = torch.rand(T, B, 8, 4, 4, generator=g)
    ref = O.codec_forward(sd, img, label, T, M, training=True, noise=[n for n in noise])
    ref['loss'].backward()
    out = train_step(m, img, label, noise)
    bad = (out['code_pm1'].cpu() != torch.stack(ref['code_pm1']))
    assert int(bad.sum()) == 0
    assert abs(out['loss'].item() - ref['loss'].item()) < 1e-5
    errs = []
    for k, p in m.named_parameters():
        if sd[k].grad is None:
            assert float(p.grad.abs().max()) == 0.0, k
            continue
        errs.append(rel_err(p.grad.cpu(), sd[k].grad))
        assert errs[-1] < GRAD_TOL['fp16'], (k, errs[-1])
    assert sorted(errs)[len(errs) // 2] < GRAD_TOL_MEDIAN
    print('config2 training M=8 T=16 B=32: worst gradient rel err {:.2e}, median {:.2e}'.format(max(errs), sorted(errs)[len(errs) // 2]))


@pytest.mark.parametrize('kind,T', [('mse', 3), ('bce', 1), ('bce', 3)])
def test_loss_kind_gradients_vs_oracle_autograd(kind, T):
    """dist_codec.py:24-34: the mse and bce branches of loss_fn, forward value and every parameter gradient"""
    M, B = 2, 8
    m = build_model('dist_codec', 'CIFAR10', M, T, seed=17, loss_fn=kind)
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(18)
    img = torch.rand(B, 3, 32, 32, generator=g)
    label = torch.randint(0, M, (B,), generator=g)
    noise = torch.rand(T, B, 8, 4, 4, generator=g)
    ref = O.codec_forward(sd, img, label, T, M, training=True, noise=[n for n in noise], loss_kind=kind)
    ref['loss'].backward()
    out = train_step(m, img, label, noise)
    assert abs(out['loss'].item() - ref['loss'].item()) < 1e-5 * max(1.0, abs(ref['loss'].item()))
    for k, p in m.named_parameters():
        assert rel_err(p.grad.cpu(), sd[k].grad) < GRAD_TOL['fp16'], k


def test_bce_outside_unit_interval_raises_like_torch():
    """F.binary_cross_entropy (dist_codec.py:26) refuses inputs outside [0, 1]; the cumulative reconstruction leaves
    it once the decoder's refinements are large.  Both the oracle (torch) and the B200 path raise RuntimeError."""
    M, T, B = 1, 4, 4
    m = build_model('dist_codec', 'CIFAR10', M, T, seed=19, loss_fn='bce')
    with torch.no_grad():
        for k, p in m.named_parameters():
            if k.startswith('model.decoder'):
                p.mul_(6.0)
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(20)
    img = torch.rand(B, 3, 32, 32, generator=g)
    label = torch.zeros(B, dtype=torch.long)
    with pytest.raises(RuntimeError):
        with torch.no_grad():
            O.codec_forward(sd, img, label, T, M, loss_kind='bce')
    with pytest.raises(RuntimeError, match='between 0 and 1'):
        run(m, img, label)
    # the model is usable afterwards (hidden state released)
    import config
    config.PARAM['loss_fn'] = 'mae'
    out = run(m, img, label)
    assert np.isfinite(out['loss'].item())


def test_16x16_images_two_nodes_training_vs_oracle_autograd():
    """at 16x16 the smallest layer has 4 pixels per image, so a 64-pixel wgrad K block spans 16 images;
    node ranges are padded to 16 images there (a 28 / 36 split used to mix two nodes' samples in one K block)"""
    M, T, B = 2, 2, 64
    m = build_model('dist_codec', 'CIFAR10', M, T, seed=23)
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(24)
    img = torch.rand(B, 3, 16, 16, generator=g)
    label = torch.tensor([0] * 28 + [1] * 36)[torch.randperm(B, generator=g)]
    noise = torch.rand(T, B, 8, 2, 2, generator=g)
    ref = O.codec_forward(sd, img, label, T, M, training=True, noise=[n for n in noise])
    ref['loss'].backward()
    out = train_step(m, img, label, noise)
    assert int((out['code_pm1'].cpu() != torch.stack(ref['code_pm1'])).sum()) == 0
    assert abs(out['loss'].item() - ref['loss'].item()) < 1e-5
    for k, p in m.named_parameters():
        assert rel_err(p.grad.cpu(), sd[k].grad) < GRAD_TOL['fp16'], k


def test_captured_graphs_own_their_buffers():
    """a captured graph bakes device addresses in.  Interleave a training graph, an eval graph, eager
    calls and a second batch size, then replay the FIRST graphs: results must equal the eager path's."""
    import config
    M, T = 2, 3
    m = build_model('dist_codec', 'CIFAR10', M, T, seed=29)
    g = torch.Generator().manual_seed(30)

    def batch(B):
        return (torch.rand(B, 3, 32, 32, generator=g).cuda(), torch.randint(0, M, (B,), generator=g).cuda(),
                torch.rand(T, B, 8, 4, 4, generator=g).cuda())

    def fwd(img, label, noise, train, graph):
        config.PARAM['cuda_graph'] = graph
        try:
            m.train(train)
            if train:
                m.zero_grad(set_to_none=True)
                out = m({'img': img, 'label': label}, noise=noise)
                out['loss'].backward()
                grads = torch.cat([p.grad.flatten() for p in m.parameters()]).clone()
                return out['code_pm1'].clone(), out['img'][-1].clone(), out['loss'].detach().clone(), grads
            with torch.no_grad():
                out = m({'img': img, 'label': label})
            return out['code_pm1'].clone(), out['img'][-1].clone(), out['loss'].detach().clone(), None
        finally:
            config.PARAM['cuda_graph'] = False
    a16, b16, a24 = batch(16), batch(16), batch(24)
    t_graph_1 = fwd(*a16, True, True)          # captures the training graph (B=16)
    e_graph_1 = fwd(*a16, False, True)         # captures the eval graph (B=16)
    fwd(*a24, True, False)                     # eager training at another batch size (replaces the eager arenas)
    fwd(*a24, False, False)
    fwd(*a24, True, True)                      # a second training graph (B=24)
    fwd(*b16, False, False)                    # eager eval at B=16
    junk = [torch.randn(1 << 22, device='cuda') for _ in range(8)]     # reuse whatever the allocator has freed
    t_graph_2 = fwd(*a16, True, True)          # replays the FIRST training graph
    e_graph_2 = fwd(*a16, False, True)         # replays the FIRST eval graph
    del junk
    t_eager, e_eager = fwd(*a16, True, False), fwd(*a16, False, False)
    for got in (t_graph_1, t_graph_2):
        assert torch.equal(got[0], t_eager[0]) and torch.equal(got[1], t_eager[1])
        assert abs(float(got[2]) - float(t_eager[2])) < 1e-6
        assert rel_err(got[3], t_eager[3]) < 1e-3      # (wgrad accumulates with atomics: run-to-run noise only)
    for got in (e_graph_1, e_graph_2):
        assert torch.equal(got[0], e_eager[0]) and torch.equal(got[1], e_eager[1])


@pytest.mark.parametrize('M,B,cta_pair', [(1, 256, 'on'), (1, 264, 'on'), (1, 136, 'off'), (4, 272, 'on'), (1, 136, 'on'),
                                          (4, 200, 'auto')])
def test_position_major_tiles_are_bit_identical_to_pixel_major(M, B, cta_pair):
    """The position-major M tiles (one pixel position x 128 images; padding taps never issued) skip only MMAs whose A
    operand is TMA zero fill, so every output bit equals the pixel-major walk's: forward (codes, reconstructions)
    and -- up to the atomics' order in wgrad -- the gradients.  M=1: encoder and decoder layers; M=4: the joint
    decoder only.  B=264 / 136 / 272 leave a pixel-major remainder behind the position-major region; with CTA pairs on
    fewer than 256 slots (B=136 'on', B=200) a pair walks two positions of the first 128 slots."""
    T = 3
    # (static schedule: under stream-K the two walks cut the K sequence at different places, which changes the order
    # of the fp32 partial sums -- test_stream_k_schedule_vs_oracle covers that combination against the oracle)
    m = build_model('dist_codec', 'CIFAR10', M, T, seed=37, cta_pair=cta_pair, stream_k='off')
    g = torch.Generator().manual_seed(38)
    img = torch.rand(B, 3, 32, 32, generator=g)
    label = torch.randint(0, M, (B,), generator=g)
    noise = torch.rand(T, B, 8, 4, 4, generator=g)
    res = {}
    for pos in (True, False):
        out = run(m, img, label)                                     # (creates the engine on the first pass)
        m._engine.pos_major = pos
        out = run(m, img, label)
        assert any(ly.pos_imgs > 0 for ly in m._engine._last_layers) == pos
        tr = train_step(m, img, label, noise)
        grads = torch.cat([p.grad.flatten() for p in m.parameters()]).clone()
        res[pos] = (out['code_pm1'].clone(), torch.stack(out['img']).clone(), tr['code_pm1'].clone(), tr['loss'].item(), grads)
    assert torch.equal(res[True][0], res[False][0]) and torch.equal(res[True][1], res[False][1])
    assert torch.equal(res[True][2], res[False][2]) and res[True][3] == res[False][3]
    assert rel_err(res[True][4], res[False][4]) < 1e-4
    # and against the oracle on a slice of the batch
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    sub = torch.arange(0, B, B // 24)[:24]
    with torch.no_grad():
        ref = O.codec_forward(sd, img[sub], label[sub], T, M)
    codes, refc = res[True][0].cpu()[:, sub], torch.stack(ref['code_pm1'])
    margins, clean = first_mismatch_margins(codes, refc, torch.stack(ref['z']))
    _note_flips('position_major_M{}_B{}_{}'.format(M, B, cta_pair), refc.numel(), margins)
    assert all(mg < MARGIN_TOL for mg in margins) and clean >= len(sub) - 1, margins
    ok = torch.tensor([bool((codes[:, b] == refc[:, b]).all()) for b in range(len(sub))])
    assert (res[True][1].cpu()[:, sub][:, ok] - torch.stack(ref['img'])[:, ok]).abs().max().item() < RECON_TOL


@pytest.mark.parametrize('M,B,T,cta_pair', [(8, 61, 6, 'off'), (8, 61, 6, 'on'), (1, 264, 3, 'on'), (3, 20, 16, 'auto'),
                                            (1, 136, 3, 'off'), (1, 136, 3, 'on')])
def test_stream_k_schedule_vs_oracle(M, B, T, cta_pair):
    """stream_k='on': every gate / dgrad GEMM cuts its K-block sequence into one range per CTA (pair) and exchanges
    partial accumulators.  Same fp32-grade numerics (the partial sums are added in another order): codes equal the
    oracle's, reconstructions within tolerance, training gradients within the fp16-backward tolerance -- for
    grouped layers (M=8, ragged, one empty node), position-major tiles with and without a pixel-major remainder."""
    m = build_model('dist_codec', 'CIFAR10', M, T, seed=43, cta_pair=cta_pair, stream_k='on')
    g = torch.Generator().manual_seed(44)
    img = torch.rand(B, 3, 32, 32, generator=g)
    label = torch.randint(0, M, (B,), generator=g)
    if M > 2:
        label[label == 1] = 0
    sub = torch.arange(0, B, max(1, B // 24))[:24]
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    with torch.no_grad():
        ref = O.codec_forward(sd, img[sub], label[sub], T, M)
    for _ in range(2):                                               # twice: the flags must come back to zero
        out = run(m, img, label)
        codes, refc, refz = out['code_pm1'].cpu()[:, sub], torch.stack(ref['code_pm1']), torch.stack(ref['z'])
        margins, clean = first_mismatch_margins(codes, refc, refz)
        assert all(mg < MARGIN_TOL for mg in margins), margins
        assert clean >= len(sub) - 1
        ok = torch.tensor([bool((codes[:, b] == refc[:, b]).all()) for b in range(len(sub))])
        err = (torch.stack(out['img']).cpu()[:, sub][:, ok] - torch.stack(ref['img'])[:, ok]).abs().max().item()
        assert err < RECON_TOL
    assert int(m._engine._sk_buffers(torch.device('cuda', 0))[1].abs().sum()) == 0
    # training step against the oracle's autograd on a small batch
    Bt = 16
    sdg = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    noise = torch.rand(T, Bt, 8, 4, 4, generator=g)
    rt = O.codec_forward(sdg, img[:Bt], label[:Bt], T, M, training=True, noise=[n for n in noise])
    rt['loss'].backward()
    m._engine.overlap_wgrad = False      # (the backward uses stream-K only when wgrad shares its stream)
    tr = train_step(m, img[:Bt], label[:Bt], noise)
    assert int((tr['code_pm1'].cpu() != torch.stack(rt['code_pm1'])).sum()) == 0
    assert abs(tr['loss'].item() - rt['loss'].item()) < 1e-5
    for k, p in m.named_parameters():
        if sdg[k].grad is None:
            assert float(p.grad.abs().max()) == 0.0, k
        else:
            assert rel_err(p.grad.cpu(), sdg[k].grad) < GRAD_TOL['fp16'], k
    assert int(m._engine._sk_buffers(torch.device('cuda', 0))[1].abs().sum()) == 0


@pytest.mark.parametrize('name,B', [('dist_codec', 40), ('dist_codec', 264), ('sep_codec', 24)])
def test_single_active_node_runs_as_one_group(name, B):
    """All samples of a batch on ONE node of a multi-node model (what a rank of a source-sharded run sees):
    the per-node layers run as a single group on that node's weight planes -- position-major tiles included at
    B >= 256 -- and only that node's weights are packed.  Eval and training step against the oracle."""
    M, T, node = 4, 3, 2
    m = build_model(name, 'CIFAR10', M, T, seed=47)
    sep = name == 'sep_codec'
    g = torch.Generator().manual_seed(48)
    img = torch.rand(B, 3, 32, 32, generator=g)
    label = torch.full((B,), node, dtype=torch.long)
    sub = torch.arange(0, B, max(1, B // 20))[:20]
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    with torch.no_grad():
        ref = O.codec_forward(sd, img[sub], label[sub], T, M, separate=sep)
    out = run(m, img, label)
    enc_layers = [ly for ly in m._engine._last_layers if ly.is_enc]
    assert all(ly.gemm_groups == 1 and ly.g0 == node and ly.active == [node] for ly in enc_layers)
    if B >= 256:
        assert all(ly.pos_imgs == 256 for ly in enc_layers)
    codes, refc = out['code_pm1'].cpu()[:, sub], torch.stack(ref['code_pm1'])
    assert torch.equal(codes, refc)
    assert (torch.stack(out['img']).cpu()[:, sub] - torch.stack(ref['img'])).abs().max().item() < RECON_TOL
    # training step: the other nodes' gradients are exactly zero, the active node's match autograd
    Bt = 12
    sdg = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    noise = torch.rand(T, Bt, 8, 4, 4, generator=g)
    rt = O.codec_forward(sdg, img[:Bt], label[:Bt], T, M, training=True, noise=[n for n in noise], separate=sep)
    rt['loss'].backward()
    tr = train_step(m, img[:Bt], label[:Bt], noise)
    assert abs(tr['loss'].item() - rt['loss'].item()) < 1e-5
    for k, p in m.named_parameters():
        if sdg[k].grad is None:
            assert float(p.grad.abs().max()) == 0.0, k
        else:
            assert rel_err(p.grad.cpu(), sdg[k].grad) < GRAD_TOL['fp16'], k
    # a later batch that touches another node packs that node's weights then
    label2 = label.clone()
    label2[::2] = 0
    with torch.no_grad():
        ref2 = O.codec_forward(sd, img[sub], label2[sub], T, M, separate=sep)
    out2 = run(m, img, label2)
    assert torch.equal(out2['code_pm1'].cpu()[:, sub], torch.stack(ref2['code_pm1']))


def test_deterministic_mode_gives_bit_identical_gradients():
    """config.PARAM['deterministic']: the same step twice gives the same gradient bits (single writer per wgrad element,
    fixed-point accumulation in the pointwise backward kernels), like the reference's CPU path; the default (float atomics)
    only agrees to ~1e-6.  The gradients still match the oracle's autograd."""
    import config
    M, T, B = 3, 4, 40
    m = build_model('dist_codec', 'CIFAR10', M, T, seed=53)
    config.PARAM['deterministic'] = True
    try:
        g = torch.Generator().manual_seed(54)
        img = torch.rand(B, 3, 32, 32, generator=g)
        label = torch.randint(0, M, (B,), generator=g)
        noise = torch.rand(T, B, 8, 4, 4, generator=g)
        runs = []
        for _ in range(3):
            tr = train_step(m, img, label, noise)
            runs.append((torch.cat([p.grad.flatten() for p in m.parameters()]).clone(), tr['loss'].item()))
        assert m._engine.deterministic
        assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[1][0], runs[2][0])
        assert runs[0][1] == runs[1][1] == runs[2][1]
        sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()}
        ref = O.codec_forward(sd, img, label, T, M, training=True, noise=[n for n in noise])
        ref['loss'].backward()
        for k, p in m.named_parameters():
            assert rel_err(p.grad.cpu(), sd[k].grad) < GRAD_TOL['fp16'], k
    finally:
        config.PARAM['deterministic'] = False
