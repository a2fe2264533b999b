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
