"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: batch sharding and the weighted gradient
all-reduce that replaces nn.DataParallel's reduce (train_model.py:59-60)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from engine.dist import allreduce_gradients, allreduce_scalar_mean, shard_range


def test_shard_range_covers_batch():
    for n, w in [(100, 8), (7, 2), (8, 8), (5, 8)]:
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_local, results):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        n_global = sum(n_local)
        torch.manual_seed(0)
        # a model whose loss is a mean over the *global* batch, like dist_codec's l1_loss
        w = torch.nn.Parameter(torch.randn(5, 3))
        b = torch.nn.Parameter(torch.randn(5))
        unused = torch.nn.Parameter(torch.randn(2))          # e.g. the encoder of a node absent on this rank
        g = torch.Generator().manual_seed(1)
        x_all = torch.randn(n_global, 3, generator=g)
        lo = sum(n_local[:rank])
        x = x_all[lo:lo + n_local[rank]]
        loss = (x @ w.t() + b).abs().mean()
        loss.backward()
        if rank == 1:
            unused.grad = torch.ones(2)
        n_coll = allreduce_gradients([w, b, unused], n_local[rank], n_global, bucket_bytes=32)
        gl = allreduce_scalar_mean(loss, n_local[rank], n_global)
        # single-process reference on the whole batch
        w2, b2 = w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
        ref = (x_all @ w2.t() + b2).abs().mean()
        ref.backward()
        ok = torch.allclose(w.grad, w2.grad, rtol=1e-5, atol=1e-7) and torch.allclose(b.grad, b2.grad, rtol=1e-5, atol=1e-7)
        ok = ok and abs(gl.item() - ref.item()) < 1e-6 and n_coll >= 2
        ok = ok and torch.allclose(unused.grad, torch.full((2,), n_local[1] / n_global))
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_local', [(4, 4), (5, 3)])
def test_weighted_gradient_allreduce_matches_single_process(n_local):
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_worker, args=(world, port, n_local, results), nprocs=world, join=True)
        assert results[0] and results[1]
