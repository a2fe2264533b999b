"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: batch sharding and the weighted gradient
all-reduce that replaces nn.DataParallel's reduce (train_model.py:59-60)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from engine.dist import (allreduce_gradients, allreduce_scalar_mean, owned_nodes, reduce_source_sharded, shard_range,
                         source_shard, split_parameters, sync_node_parameters)


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


class _ToyCodec(torch.nn.Module):
    """the reference's parameter tree in miniature: per-node encoders, a joint (or per-node) decoder"""

    def __init__(self, num_node, node_decoders):
        super().__init__()
        enc = torch.nn.ModuleList([torch.nn.Sequential(torch.nn.Linear(3, 4)) for _ in range(num_node)])
        if node_decoders:
            dec = torch.nn.ModuleList([torch.nn.Sequential(torch.nn.Linear(4, 3)) for _ in range(num_node)])
        else:
            dec = torch.nn.Sequential(torch.nn.Linear(4, 3))
        self.model = torch.nn.ModuleDict({'encoder': enc, 'decoder': dec})
        self.node_decoders = node_decoders

    def forward(self, x, label):
        out = torch.zeros_like(x)
        for j in range(len(self.model['encoder'])):
            idx = torch.nonzero(label == j).reshape(-1)
            if idx.numel() == 0:
                continue
            dec = self.model['decoder'][j] if self.node_decoders else self.model['decoder']
            out = out.index_copy(0, idx, dec(torch.tanh(self.model['encoder'][j](x[idx]))))
        return (out - x).abs().mean()


def test_source_shard_helpers():
    assert owned_nodes(8, 1, 4) == [1, 5] and owned_nodes(3, 2, 8) == [2] and owned_nodes(3, 5, 8) == []
    label = torch.tensor([3, 0, 1, 2, 1, 0, 3])
    assert source_shard(label, 1, 2).tolist() == [0, 2, 4, 6] and source_shard(label, 0, 2).tolist() == [1, 3, 5]
    shared, per_node = split_parameters(_ToyCodec(3, False))
    assert len(shared) == 2 and sorted(per_node) == [0, 1, 2] and all(len(v) == 2 for v in per_node.values())
    shared, per_node = split_parameters(_ToyCodec(3, True))
    assert shared == [] and all(len(v) == 4 for v in per_node.values())


def _source_worker(rank, world, port, node_decoders, results):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        M = 5
        torch.manual_seed(3)
        m = _ToyCodec(M, node_decoders)
        ref = _ToyCodec(M, node_decoders)
        ref.load_state_dict(m.state_dict())
        g = torch.Generator().manual_seed(4)
        x_all = torch.randn(23, 3, generator=g)
        lab_all = torch.randint(0, M, (23,), generator=g)
        idx = source_shard(lab_all, rank, world)
        loss = m(x_all[idx], lab_all[idx])
        loss.backward()
        n_coll = reduce_source_sharded(m, idx.numel(), 23)
        ref(x_all, lab_all).backward()
        ok = (n_coll >= 1) != node_decoders                      # sep_codec shares nothing: no collective
        for (k, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
            node = int(k.split('.')[2]) if (k.startswith('model.encoder') or node_decoders) else None
            if node is None or node % world == rank:
                ok = ok and torch.allclose(p.grad, q.grad, rtol=1e-5, atol=1e-7)
            else:
                ok = ok and p.grad is None
        # one SGD step on the owned / shared parameters, then the owners publish their nodes
        with torch.no_grad():
            for p in m.parameters():
                if p.grad is not None:
                    p -= 0.1 * p.grad
            for q in ref.parameters():
                q -= 0.1 * q.grad
        sync_node_parameters(m)
        for p, q in zip(m.parameters(), ref.parameters()):
            ok = ok and torch.allclose(p, q, rtol=1e-5, atol=1e-7)
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('node_decoders', [False, True])
def test_source_sharded_step_matches_single_process(node_decoders):
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_source_worker, args=(world, port, node_decoders, results), nprocs=world, join=True)
        assert results[0] and results[1]


def _ckpt_worker(rank, world, port, tmpdir, results):
    from engine.dist import gather_sharded_state
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        M = 5
        torch.manual_seed(5)
        m, ref = _ToyCodec(M, False), _ToyCodec(M, False)
        ref.load_state_dict(m.state_dict())
        opt, opt_ref = torch.optim.Adam(m.parameters(), lr=0.05), torch.optim.Adam(ref.parameters(), lr=0.05)
        g = torch.Generator().manual_seed(6)

        def train(model, optimizer, sharded, steps):
            for _ in range(steps):
                x = torch.randn(23, 3, generator=g)
                lab = torch.randint(0, M, (23,), generator=g)
                optimizer.zero_grad(set_to_none=True)
                if sharded:
                    idx = source_shard(lab, rank, world)
                    model(x[idx], lab[idx]).backward()
                    reduce_source_sharded(model, idx.numel(), 23)
                else:
                    model(x, lab).backward()
                optimizer.step()
        state = g.get_state()
        train(m, opt, True, 3)
        g.set_state(state)
        train(ref, opt_ref, False, 3)
        # before the gather this rank's copy of the other rank's encoders is untrained and has no Adam state
        foreign = [p for k, p in m.named_parameters() if k.startswith('model.encoder.') and int(k.split('.')[2]) % world != rank]
        stale = all(len(opt.state[p]) == 0 for p in foreign)
        gather_sharded_state(m, opt)                             # MANDATORY before a source-sharded checkpoint
        ok = stale
        for p, q in zip(m.parameters(), ref.parameters()):
            ok = ok and torch.allclose(p, q, rtol=1e-5, atol=1e-7)
            for k in ('exp_avg', 'exp_avg_sq'):
                ok = ok and torch.allclose(opt.state[p][k], opt_ref.state[q][k], rtol=1e-5, atol=1e-9)
            ok = ok and float(opt.state[p]['step']) == float(opt_ref.state[q]['step'])
        # the reference's checkpoint, written by rank 0 only (train_model.py:87-98), resumes on ONE process to the
        # same trajectory as the never-sharded run
        path = os.path.join(tmpdir, 'ckpt.pt')
        if rank == 0:
            torch.save({'model_dict': m.state_dict(), 'optimizer_dict': opt.state_dict()}, path)
        dist.barrier()
        if rank == 0:
            ck = torch.load(path, weights_only=False)
            m2 = _ToyCodec(M, False)
            m2.load_state_dict(ck['model_dict'])
            opt2 = torch.optim.Adam(m2.parameters(), lr=0.05)
            opt2.load_state_dict(ck['optimizer_dict'])
            state = g.get_state()
            train(m2, opt2, False, 2)
            g.set_state(state)
            train(ref, opt_ref, False, 2)
            for p, q in zip(m2.parameters(), ref.parameters()):
                ok = ok and torch.allclose(p, q, rtol=1e-4, atol=1e-6)
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_source_sharded_checkpoint_round_trip(tmp_path):
    """a checkpoint of a source-sharded run must carry every node's trained encoder and Adam moments"""
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_ckpt_worker, args=(world, port, str(tmp_path), results), nprocs=world, join=True)
        assert results[0] and results[1]


def test_flat_gradient_view():
    """adjacent views of one buffer are recognised (one in-place collective); anything else is not"""
    from engine.dist import flat_gradient_view
    flat = torch.arange(20, dtype=torch.float32)
    ps = [torch.nn.Parameter(torch.zeros(2, 3)), torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(5, 2))]
    ps[0].grad, ps[1].grad, ps[2].grad = flat[0:6].view(2, 3), flat[6:10], flat[10:20].view(5, 2)
    v = flat_gradient_view(ps)
    assert v is not None and v.numel() == 20 and v.data_ptr() == flat.data_ptr()
    v.mul_(2.0)
    assert float(ps[2].grad[4, 1]) == 38.0
    assert flat_gradient_view(ps[1:]).numel() == 14                         # a contiguous sub-range (the joint decoder)
    assert flat_gradient_view([ps[0], ps[2]]) is None                       # a gap
    ps[1].grad = torch.zeros(4)
    assert flat_gradient_view(ps) is None                                   # another buffer
    ps[1].grad = None
    assert flat_gradient_view(ps) is None
