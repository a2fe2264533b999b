"""The library's multi-tensor Adam (engine.optim.FusedAdam -> drasic_adam_step) against torch.optim.Adam on the
CPU, the optimizer of the reference's training loop (train_model.py:155-161, :116)."""
import pytest
import torch

from conftest import build_model
from engine import binding as B
from engine.optim import FusedAdam
from oracle import drasic_oracle as O

pytestmark = pytest.mark.gpu

# fp32 arithmetic in a different association order than torch's CPU loop: a few ulp of the update per step
TOL = 2e-6


def _tensors(gen):
    """sizes around the vector width, the chunk size (16384) and the 64-tensors-per-launch batch"""
    sizes = [(1,), (3,), (7, 5), (16384,), (16385,), (4, 3, 3, 3), (100000,)] + [(17 + i,) for i in range(70)]
    return [torch.randn(*s, generator=gen) * (0.1 + 0.01 * i) for i, s in enumerate(sizes)]


@pytest.mark.parametrize('weight_decay', [0.0, 5e-4])
def test_fused_adam_matches_torch_adam(weight_decay):
    gen = torch.Generator().manual_seed(80)
    init = _tensors(gen)
    ref = [torch.nn.Parameter(t.clone()) for t in init]
    ours = [torch.nn.Parameter(t.clone().cuda()) for t in init]
    # an unaligned view as a parameter (scalar path of the kernel): 4-byte offset into a larger buffer
    buf = torch.zeros(1001, device='cuda')
    odd = buf[1:].detach().requires_grad_(True)
    odd_ref = torch.nn.Parameter(torch.zeros(1000))
    kw = dict(lr=3e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=weight_decay)
    opt_ref = torch.optim.Adam(ref + [odd_ref], **kw)
    opt = FusedAdam(ours + [odd], **kw)
    for step in range(4):
        for r, o in zip(ref + [odd_ref], ours + [odd]):
            g = torch.randn(r.shape, generator=gen) * 0.5
            r.grad = g.clone()
            o.grad = g.cuda()
        if step == 2:                                   # a parameter without a gradient is skipped, its count stays behind
            ref[5].grad = None
            ours[5].grad = None
        opt_ref.step()
        opt.step()
    assert opt.launches == 3                            # 77 tensors at step 4 -> two launches, the straggler one more
    for r, o in zip(ref + [odd_ref], ours + [odd]):
        assert (o.detach().cpu() - r.detach()).abs().max().item() <= TOL * max(1.0, r.detach().abs().max().item())
        sr, so = opt_ref.state[r], opt.state[o]
        assert float(sr['step']) == float(so['step'])
        assert torch.allclose(so['exp_avg'].cpu(), sr['exp_avg'], rtol=1e-5, atol=1e-8)
        assert torch.allclose(so['exp_avg_sq'].cpu(), sr['exp_avg_sq'], rtol=1e-5, atol=1e-10)
    assert float(opt.state[ours[5]]['step']) == 3.0 and float(opt.state[ours[0]]['step']) == 4.0
    assert float(buf[0]) == 0.0                         # nothing written outside the view


def test_fused_adam_state_dict_interchanges_with_torch_adam():
    """utils.py:173-180 resume: a checkpoint written by torch.optim.Adam continues under FusedAdam (and back)"""
    gen = torch.Generator().manual_seed(81)
    init = [torch.randn(33, 5, generator=gen), torch.randn(1000, generator=gen)]
    grads = [[torch.randn(t.shape, generator=gen) for t in init] for _ in range(4)]
    a = [torch.nn.Parameter(t.clone().cuda()) for t in init]
    b = [torch.nn.Parameter(t.clone().cuda()) for t in init]
    opt_a = torch.optim.Adam(a, lr=1e-2, weight_decay=1e-4)
    opt_b = torch.optim.Adam(b, lr=1e-2, weight_decay=1e-4, fused=True)     # keeps 'step' on the device
    for gs in grads[:2]:
        for opt, ps in ((opt_a, a), (opt_b, b)):
            for p, g in zip(ps, gs):
                p.grad = g.cuda()
            opt.step()
    ours = [FusedAdam(a, lr=1e-2, weight_decay=1e-4), FusedAdam(b, lr=1e-2, weight_decay=1e-4)]
    ours[0].load_state_dict(opt_a.state_dict())
    ours[1].load_state_dict(opt_b.state_dict())
    ref = [torch.nn.Parameter(t.clone()) for t in init]
    opt_ref = torch.optim.Adam(ref, lr=1e-2, weight_decay=1e-4)
    for gs in grads:
        for p, g in zip(ref, gs):
            p.grad = g.clone()
        opt_ref.step()
    for gs in grads[2:]:
        for opt, ps in zip(ours, (a, b)):
            for p, g in zip(ps, gs):
                p.grad = g.cuda()
            opt.step()
    for ps in (a, b):
        for p, r in zip(ps, ref):
            assert (p.detach().cpu() - r.detach()).abs().max().item() <= 5e-6
    back = torch.optim.Adam(a, lr=1e-2, weight_decay=1e-4)
    back.load_state_dict(ours[0].state_dict())          # and the other way round
    assert float(back.state[a[0]]['step']) == 4.0


def test_fused_adam_rejects_what_it_cannot_do():
    p = torch.nn.Parameter(torch.zeros(4))
    with pytest.raises(ValueError):
        FusedAdam([p], lr=-1.0)
    with pytest.raises(ValueError):
        FusedAdam([p], amsgrad=True)
    opt = FusedAdam([p])
    p.grad = torch.ones(4)
    with pytest.raises(B.DrasicError):                  # CPU parameter: no fallback
        opt.step()
    q = torch.nn.Parameter(torch.zeros(4, device='cuda', dtype=torch.float64))
    q.grad = torch.ones_like(q)
    with pytest.raises(ValueError):
        FusedAdam([q]).step()
    L = B.lib()
    assert L.drasic_adam_step(None, 1, 1e-3, 0.9, 0.999, 1e-8, 0.0, 0.1, 0.001, None) == -1
    arr = (B.AdamTensor * 1)()
    arr[0].numel = 4                                    # null pointers
    assert L.drasic_adam_step(arr, 1, 1e-3, 0.9, 0.999, 1e-8, 0.0, 0.1, 0.001, None) == -1
    assert L.drasic_adam_step(arr, 0, 1e-3, 0.9, 0.999, 1e-8, 0.0, 0.1, 0.001, None) == 0
    assert L.drasic_adam_step(arr, 1, 1e-3, 1.0, 0.999, 1e-8, 0.0, 0.1, 0.001, None) == -1


def test_training_with_fused_adam_tracks_the_oracle():
    """three optimizer steps of the codec (train_model.py:112-116) against the oracle + torch.optim.Adam on the CPU;
    the updates must reach the engine's packed weight planes (the kernel bumps the parameters' versions)"""
    m = build_model('dist_codec', 'CIFAR10', 2, 3, seed=82)
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    opt = FusedAdam(m.parameters(), lr=1e-3, weight_decay=5e-4)
    opt_ref = torch.optim.Adam(list(sd.values()), lr=1e-3, weight_decay=5e-4)
    g = torch.Generator().manual_seed(83)
    img = torch.rand(8, 3, 32, 32, generator=g)
    label = torch.randint(0, 2, (8,), generator=g)
    inp = {'img': img.cuda(), 'label': label.cuda()}
    v0 = [p._version for p in m.parameters()]
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
        assert abs(out['loss'].item() - ref['loss'].item()) < 5e-4
    assert all(a > b for a, b in zip([p._version for p in m.parameters()], v0))
    worst = max((p.detach().cpu() - sd[k].detach()).abs().max().item() for k, p in m.named_parameters())
    assert worst < 2e-3        # three Adam steps of lr 1e-3: the sign of a tiny fp16-backward gradient may differ
    m.train(False)
    with torch.no_grad():
        out = m(inp)
    m2 = build_model('dist_codec', 'CIFAR10', 2, 3, seed=99)
    m2.load_state_dict(m.state_dict())
    m2.train(False)
    with torch.no_grad():
        out2 = m2(inp)
    assert torch.equal(out['code_pm1'], out2['code_pm1']) and torch.equal(out['img'][-1], out2['img'][-1])
