"""GPU parity of the rows SURVEY.md 8(f) marks "next": heterogeneous per-node rates (BASELINE config 4),
decoder-only reconstruction from a bitstream prefix (N1) and on-device rate/distortion metrics (N3)."""
import numpy as np
import pytest
import torch

from conftest import build_model
from oracle import drasic_oracle as O

pytestmark = pytest.mark.gpu


def _batch(B, C, M, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, C, 32, 32, generator=g), torch.randint(0, M, (B,), generator=g), g


def test_config4_heterogeneous_rates_eval_vs_oracle():
    """T=6 for the first half of the sources, T=3 for the second half (config 4 scaled down; inferred
    semantics restated in oracle.codec_forward): frozen nodes keep their reconstruction, the loss keeps
    counting it, active nodes are untouched."""
    import config
    M, T, B = 4, 6, 26
    m = build_model('dist_codec', 'CIFAR10', M, T, seed=5)
    iters = [T, T, T // 2, T // 2]
    config.PARAM['node_iters'] = iters
    try:
        sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
        img, label, _ = _batch(B, 3, M, 50)
        with torch.no_grad():
            ref = O.codec_forward(sd, img, label, T, M, node_iters=iters)
            m.train(False)
            out = m({'img': img.cuda(), 'label': label.cuda()})
        codes, refc = out['code_pm1'].cpu(), torch.stack(ref['code_pm1'])
        active = torch.tensor([[t < iters[int(l)] for l in label] for t in range(T)])
        assert torch.equal(codes[active], refc[active])                     # emitted symbols of active nodes
        err = (torch.stack(out['img']).cpu() - torch.stack(ref['img'])).abs().max().item()
        assert err < 1e-3
        # a frozen node's reconstruction stops changing exactly at its budget
        frozen = label >= 2
        rec = torch.stack(out['img']).cpu()
        assert torch.equal(rec[T // 2 - 1][frozen], rec[T - 1][frozen])
        assert not torch.equal(rec[T // 2 - 1][~frozen], rec[T - 1][~frozen])
        assert abs(out['loss'].item() - ref['loss'].item()) < 1e-5
    finally:
        config.PARAM.pop('node_iters', None)


def test_config4_heterogeneous_rates_training_gradients():
    import config
    M, T, B = 4, 4, 18
    m = build_model('dist_codec', 'CIFAR10', M, T, seed=6)
    iters = [4, 4, 2, 2]
    config.PARAM['node_iters'] = iters
    try:
        sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()}
        img, label, g = _batch(B, 3, M, 51)
        noise = torch.rand(T, B, 8, 4, 4, generator=g)
        ref = O.codec_forward(sd, img, label, T, M, training=True, noise=[n for n in noise], node_iters=iters)
        ref['loss'].backward()
        m.train(True)
        m.zero_grad()
        out = m({'img': img.cuda(), 'label': label.cuda()}, noise=noise.cuda())
        out['loss'].backward()
        assert abs(out['loss'].item() - ref['loss'].item()) < 1e-5
        worst = 0.0
        for k, p in m.named_parameters():
            ga, gb = p.grad.cpu(), sd[k].grad
            if gb is None:
                assert float(ga.abs().max()) == 0.0
                continue
            rel = float((ga - gb).norm() / (gb.norm() + 1e-20))
            worst = max(worst, rel)
        assert worst < 5e-3, worst
    finally:
        config.PARAM.pop('node_iters', None)


@pytest.mark.parametrize('name,M,B,T', [('dist_codec', 8, 37, 6), ('sep_codec', 3, 14, 4)])
def test_decode_from_bitstream_prefix(name, M, B, T):
    """N1: the 'bits' output is a real progressive bitstream; decoding any prefix with the decoder half
    alone reproduces the encoder-side reconstructions bit for bit, and matches the oracle's decoder."""
    m = build_model(name, 'CIFAR10', M, T, seed=8)
    img, label, _ = _batch(B, 3, M, 60)
    label[label == 1] = 0                                                    # an empty node
    m.train(False)
    with torch.no_grad():
        out = m({'img': img.cuda(), 'label': label.cuda()})
    enc = m.encode({'img': img.cuda(), 'label': label.cuda()})
    assert torch.equal(enc['bits'], out['bits']) and enc['bits'].dtype == torch.uint8
    assert tuple(enc['bits'].shape) == (T, B, 8 * 4 * 4 // 8)
    # unpack is the inverse of the device bit-pack
    assert torch.equal(m.unpack_bits(enc['bits'], (8, 4, 4)), out['code_pm1'])
    for t in (1, T // 2, T):
        rec = m.decode(enc['bits'][:t], label.cuda(), size=(32, 32))
        assert len(rec) == t
        for i in range(t):
            assert torch.equal(rec[i], out['img'][i]), (t, i)
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    with torch.no_grad():
        ref = O.codec_decode(sd, [c for c in out['code_pm1'].cpu()], label, M, separate=(name == 'sep_codec'))
    assert (torch.stack(rec).cpu() - torch.stack(ref)).abs().max().item() < 1e-3


def test_on_device_metrics_match_reference_definitions():
    """N3: PSNR per rate from the squared-error sums reduced inside the tail kernel; BPP from the code
    container -- same numbers as the reference's metrics.py:75-89 evaluated on the outputs."""
    import metrics
    M, T, B = 2, 5, 12
    m = build_model('dist_codec', 'CIFAR10', M, T, seed=9)
    img, label, _ = _batch(B, 3, M, 70)
    m.train(False)
    inp = {'img': img.cuda(), 'label': label.cuda()}
    with torch.no_grad():
        out = m(inp)
    ev = metrics.Metric().evaluate(['Loss', 'PSNR', 'BPP'], inp, out)
    want_psnr = [O.psnr(r.cpu(), img) for r in out['img']]
    np.testing.assert_allclose(ev['PSNR'], want_psnr, rtol=0, atol=1e-3)
    assert ev['BPP'] == [O.bpp(c, img) for c in out['code']]
    assert abs(ev['Loss'] - out['loss'].item()) == 0.0
    # plain definitions on tensors
    assert abs(metrics.PSNR(out['img'][-1], inp['img']) - want_psnr[-1]) < 1e-3
    with pytest.raises(ValueError):
        metrics.Metric().evaluate(['SSIM'], inp, out)


def test_untiled_full_resolution_inference_vs_oracle():
    """N4: the reference evaluates large images (Kodak) un-tiled and fully convolutionally (data.py:51-56).
    A 128x192 image (planes 64x96 / 32x48 / 16x24: pixel boxes 32x4, 16x8, 8x16; 384 code pixels = 12
    bottleneck tiles per image) against the oracle on the same tensor."""
    M, T, B = 2, 3, 3
    m = build_model('dist_codec', 'CIFAR10', M, T, seed=21)
    g = torch.Generator().manual_seed(80)
    img = torch.rand(B, 3, 128, 192, generator=g)
    label = torch.tensor([0, 1, 0])
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    with torch.no_grad():
        ref = O.codec_forward(sd, img, label, T, M)
    m.train(False)
    with torch.no_grad():
        out = m({'img': img.cuda(), 'label': label.cuda()})
    codes, refc, refz = out['code_pm1'].cpu(), torch.stack(ref['code_pm1']), torch.stack(ref['z'])
    assert tuple(codes.shape) == (T, B, 8, 16, 24)
    bad = codes != refc
    # a bit may differ only where the oracle's own margin is ~0 (none expected at this size)
    assert float(refz[bad].abs().max()) < 2e-6 if bad.any() else True
    if not bad.any():
        err = (torch.stack(out['img']).cpu() - torch.stack(ref['img'])).abs().max().item()
        assert err < 1e-3
        assert abs(out['loss'].item() - ref['loss'].item()) < 1e-5
    want = np.stack([O.pack_bits(c) for c in refc])
    if not bad.any():
        assert np.array_equal(out['bits'].cpu().numpy(), want)
    # decoder-only from the bitstream reproduces the reconstructions bit for bit
    rec = m.decode(out['bits'], label.cuda(), size=(128, 192))
    assert all(torch.equal(a, b) for a, b in zip(rec, out['img']))


def test_kodak_size_image_properties():
    """768x512 un-tiled (the size-independent properties: progressive decode == encode-side
    reconstructions, PSNR finite and non-decreasing over the first iterations on a smooth image)"""
    M, T = 1, 3
    m = build_model('dist_codec', 'CIFAR10', M, T, seed=22)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, 512), torch.linspace(0, 1, 768), indexing='ij')
    img = torch.stack([yy, xx, (yy + xx) / 2]).unsqueeze(0).contiguous()
    label = torch.zeros(1, dtype=torch.long)
    m.train(False)
    with torch.no_grad():
        out = m({'img': img.cuda(), 'label': label.cuda()})
    assert tuple(out['bits'].shape) == (T, 1, 8 * 64 * 96 // 8)
    rec = m.decode(out['bits'][:2], label.cuda(), size=(512, 768))
    assert torch.equal(rec[1], out['img'][1])
    mse = out['mse'].cpu().numpy()
    assert np.all(np.isfinite(mse)) and mse[0] > 0


def test_host_resident_labels_give_the_same_results():
    """input['label'] may stay a CPU tensor (the node ids are consumed on the host: batch plan, grid sizes): eval, a
    training step and a graph replay give exactly what CUDA labels give"""
    import config
    M, T, B = 3, 3, 20
    m = build_model('dist_codec', 'CIFAR10', M, T, seed=11)
    img, label, g = _batch(B, 3, M, 71)
    noise = torch.rand(T, B, 8, 4, 4, generator=g).cuda()
    img = img.cuda()
    m.train(False)
    outs = []
    for graph in (False, True):
        config.PARAM['cuda_graph'] = graph
        try:
            with torch.no_grad():
                outs.append([m({'img': img, 'label': lab}) for lab in (label.cuda(), label)])
        finally:
            config.PARAM['cuda_graph'] = False
    for on_dev, on_host in outs:
        assert torch.equal(on_dev['code_pm1'], on_host['code_pm1']) and on_dev['loss'].item() == on_host['loss'].item()
        assert all(np.array_equal(a, b) for a, b in zip(on_dev['code'], on_host['code']))
    assert torch.equal(outs[0][0]['code_pm1'], outs[1][1]['code_pm1'])
    m.train(True)
    res = []
    for lab in (label.cuda(), label):
        m.zero_grad(set_to_none=True)
        out = m({'img': img, 'label': lab}, noise=noise)
        out['loss'].backward()
        res.append((out['code_pm1'].clone(), out['loss'].item()))
        assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in m.parameters())
    assert torch.equal(res[0][0], res[1][0]) and res[0][1] == res[1][1]
