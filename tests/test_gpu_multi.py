"""SURVEY 8(e) parity check on real GPUs: gradients and loss of a global batch sharded over two ranks
(NCCL all-reduce weighted by n_local / N_global) equal the single-GPU gradients of the same global batch.
Needs two visible GPUs (skipped on the single-GPU test tier; run with `gpurun --gpus 2`)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, shards, results):
    import torch.distributed as dist
    from conftest import build_model
    from engine.dist import allreduce_gradients, allreduce_scalar_mean
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        M, T, B = 4, 3, sum(shards)
        m = build_model('dist_codec', 'CIFAR10', M, T, seed=31, device=dev)
        g = torch.Generator().manual_seed(32)
        img = torch.rand(B, 3, 32, 32, generator=g)
        label = torch.randint(0, M, (B,), generator=g)
        noise = torch.rand(T, B, 8, 4, 4, generator=g)
        lo = sum(shards[:rank])
        hi = lo + shards[rank]
        m.train(True)
        m.zero_grad()
        out = m({'img': img[lo:hi].to(dev), 'label': label[lo:hi].to(dev)}, noise=noise[:, lo:hi].to(dev))
        out['loss'].backward()
        params = [p for p in m.parameters()]
        allreduce_gradients(params, shards[rank], B)
        loss = allreduce_scalar_mean(out['loss'], shards[rank], B)
        if rank == 0:
            # the same global batch on this one GPU
            m1 = build_model('dist_codec', 'CIFAR10', M, T, seed=31, device=dev)
            m1.train(True)
            m1.zero_grad()
            o1 = m1({'img': img.to(dev), 'label': label.to(dev)}, noise=noise.to(dev))
            o1['loss'].backward()
            worst = 0.0
            for (k, p), (_, q) in zip(m.named_parameters(), m1.named_parameters()):
                ga, gb = p.grad, q.grad
                den = float(gb.norm()) + 1e-20
                worst = max(worst, float((ga - gb).norm()) / den if float(gb.abs().max()) > 0 else float(ga.abs().max()))
            results['worst'] = worst
            results['dloss'] = abs(float(loss) - float(o1['loss']))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('shards', [(8, 8), (11, 5)])
def test_two_gpu_gradients_equal_single_gpu(shards):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    import torch.multiprocessing as mp
    port = _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_worker, args=(2, port, shards, results), nprocs=2, join=True)
        # fp16 backward GEMMs scale each rank's gradient planes by its own power of two: equal within fp16-grade noise
        assert results['dloss'] < 1e-6, dict(results)
        assert results['worst'] < 2e-3, dict(results)


def _source_worker(rank, world, port, results):
    import torch.distributed as dist
    from conftest import build_model
    from engine.dist import allreduce_scalar_mean, reduce_source_sharded, source_shard, sync_node_parameters
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        M, T, B = 4, 3, 21
        m = build_model('dist_codec', 'CIFAR10', M, T, seed=41, device=dev)
        g = torch.Generator().manual_seed(42)
        img = torch.rand(B, 3, 32, 32, generator=g)
        label = torch.randint(0, M, (B,), generator=g)
        noise = torch.rand(T, B, 8, 4, 4, generator=g)
        idx = source_shard(label, rank, world)                       # this rank: every sample of nodes j = rank mod 2
        opt = torch.optim.SGD(m.parameters(), lr=0.05)
        m.train(True)
        opt.zero_grad()
        out = m({'img': img[idx].to(dev), 'label': label[idx].to(dev)}, noise=noise[:, idx].to(dev))
        out['loss'].backward()
        reduce_source_sharded(m, idx.numel(), B)
        loss = allreduce_scalar_mean(out['loss'], idx.numel(), B)
        m1 = build_model('dist_codec', 'CIFAR10', M, T, seed=41, device=dev)
        opt1 = torch.optim.SGD(m1.parameters(), lr=0.05)
        m1.train(True)
        opt1.zero_grad()
        o1 = m1({'img': img.to(dev), 'label': label.to(dev)}, noise=noise.to(dev))
        o1['loss'].backward()
        worst, foreign = 0.0, 0
        for (k, p), (_, q) in zip(m.named_parameters(), m1.named_parameters()):
            if k.startswith('model.encoder.') and int(k.split('.')[2]) % world != rank:
                foreign += int(p.grad is not None)
                continue
            worst = max(worst, float((p.grad - q.grad).norm()) / (float(q.grad.norm()) + 1e-20))
        opt.step()
        opt1.step()
        sync_node_parameters(m)                                       # owners publish their encoders
        dw = max(float((p - q).abs().max()) for p, q in zip(m.parameters(), m1.parameters()))
        m.train(False)
        m1.train(False)
        with torch.no_grad():                                         # the refreshed weights reach the packed copies
            a = m({'img': img.to(dev), 'label': label.to(dev)})
            b = m1({'img': img.to(dev), 'label': label.to(dev)})
        results[rank] = dict(worst=worst, foreign=foreign, dloss=abs(float(loss) - float(o1['loss'])), dw=dw,
                             flips=int((a['code_pm1'] != b['code_pm1']).sum()),
                             drecon=float((a['img'][-1] - b['img'][-1]).abs().max()))
    finally:
        dist.destroy_process_group()


def test_two_gpu_source_sharded_step_equals_single_gpu():
    """SURVEY 8(e) partitioning for M >= G: rank r owns the encoders (and samples) of nodes j = r mod G; only the
    joint decoder's gradients are all-reduced"""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    import torch.multiprocessing as mp
    port = _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_source_worker, args=(2, port, results), nprocs=2, join=True)
        for r in (0, 1):
            res = results[r]
            assert res['foreign'] == 0 and res['dloss'] < 1e-6, res
            assert res['worst'] < 2e-3, res
            assert res['dw'] < 1e-5 and res['drecon'] < 1e-3, res
            assert res['flips'] <= 2, res      # weights equal to ~1e-7: at most an exactly-marginal sign may differ


def test_one_process_two_devices():
    """a model on cuda:1 in a process whose current device is cuda:0 (and then one on cuda:0): the engine makes the
    tensors' device current for its launches and configures its kernels per device"""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    from conftest import build_model
    g = torch.Generator().manual_seed(70)
    img = torch.rand(9, 3, 32, 32, generator=g)
    label = torch.randint(0, 2, (9,), generator=g)
    outs = []
    for dev in ('cuda:1', 'cuda:0'):
        m = build_model('dist_codec', 'CIFAR10', 2, 3, seed=71, device=dev)
        m.train(False)
        assert torch.cuda.current_device() == 0
        with torch.no_grad():
            out = m({'img': img.to(dev), 'label': label.to(dev)})
        outs.append((out['code_pm1'].cpu(), out['img'][-1].cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
