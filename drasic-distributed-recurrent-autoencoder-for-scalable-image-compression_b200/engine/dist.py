"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink / NVSwitch).

Samples are independent through the whole coding loop (no batch statistics; "joint decoding" means
shared decoder weights, models/dist_codec.py:55), so the global batch is sharded across ranks and the
only exchange step of a training iteration is one all-reduce (sum) of the parameter gradients -- the
replacement of nn.DataParallel's per-step broadcast / scatter / gather / reduce_add
(train_model.py:59-60,114).  The reference loss is a mean over the global batch (dist_codec.py:58), so a
rank holding n_local of N_global samples contributes its mean-loss gradient with weight n_local/N_global.
Evaluation needs no collective at all.

With M >= world sources the batch can instead be sharded *by source* (`source_shard`): rank r holds every
sample of the nodes j with j % world == r, so an encoder's gradient exists on exactly one rank and only the
shared decoder's gradients (99 MB of the 326 MB at M=8) are all-reduced (`reduce_source_sharded`); the
per-node parameters of the other ranks' nodes stay untouched locally and are refreshed from their owners
before an evaluation on the whole model (`sync_node_parameters`) and, together with the owners' optimizer state,
before a checkpoint (`gather_sharded_state`: the reference's checkpoint is single-process and complete,
train_model.py:87-98)."""
import re

import torch
import torch.distributed as dist


def shard_range(n_global, rank, world):
    """contiguous shard [lo, hi) of rank; sizes differ by at most one"""
    base, rem = divmod(n_global, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def flat_gradient_view(params):
    """The gradients of `params` as ONE 1-D tensor when they are adjacent views of one buffer (the engine's
    backward lays every gradient out in a single flat fp32 buffer in state_dict order, engine/codec.py), else None."""
    grads = [p.grad for p in params]
    if not grads or any(g is None or not g.is_contiguous() or g.dtype != grads[0].dtype for g in grads):
        return None
    store = grads[0].untyped_storage()
    if any(g.untyped_storage().data_ptr() != store.data_ptr() for g in grads):
        return None
    order = sorted(grads, key=lambda g: g.storage_offset())
    off = order[0].storage_offset()
    for g in order:
        if g.storage_offset() != off:
            return None
        off += g.numel()
    first = order[0].storage_offset()
    return torch.empty(0, dtype=grads[0].dtype, device=grads[0].device).set_(store, first, (off - first,))


def allreduce_gradients(params, n_local, n_global, group=None, bucket_bytes=1 << 28):
    """In place: p.grad <- sum_r (n_r / N) * p.grad_r for every parameter (missing grads count as zero).
    When the gradients are views of one flat buffer -- what the engine's backward produces -- this is ONE in-place
    collective on that buffer, no flatten / unflatten copies.  Otherwise they travel in a few large fp32 buckets:
    NVSwitch makes collective cost latency-, not link-bound, so buckets are sized for launch count."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    w = float(n_local) / float(n_global)
    params = [p for p in params if p.requires_grad]
    flat = flat_gradient_view(params)
    if flat is not None:
        flat.mul_(w)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        return 1
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    state = {'bucket': [], 'size': 0, 'n': 0}

    def flush():
        if not state['bucket']:
            return
        flat = torch.cat([p.grad.reshape(-1) for p in state['bucket']])
        flat.mul_(w)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        off = 0
        for p in state['bucket']:
            n = p.grad.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n
        state['n'] += 1
        state['bucket'], state['size'] = [], 0
    for p in params:
        state['bucket'].append(p)
        state['size'] += p.grad.numel() * 4
        if state['size'] >= bucket_bytes:
            flush()
    flush()
    return state['n']


def allreduce_scalar_mean(value, n_local, n_global, group=None):
    """global mean of per-rank means weighted by the rank's sample count"""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return value
    t = value.detach().clone() * (float(n_local) / float(n_global))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


# ---------------------------------------------------------------- sharding by source
def owned_nodes(num_node, rank, world):
    """nodes whose encoder (and, for sep_codec, decoder) lives on this rank"""
    return [j for j in range(num_node) if j % world == rank]


def source_shard(label, rank, world):
    """indices (ascending) of the samples whose node is owned by rank"""
    return torch.nonzero(label % world == rank).reshape(-1)


_NODE_KEY = re.compile(r'^(?:module\.)?model\.(encoder|decoder)\.(\d+)\.')


def split_parameters(model):
    """(shared, per_node): per_node[j] are the parameters only node j's samples touch -- its encoder
    (model.encoder.{j}, dist_codec.py:50) and, where decoders are per node too (sep_codec.py:51), its decoder;
    shared is everything else (dist_codec's joint decoder)."""
    core = model.module if hasattr(model, 'module') else model
    node_decoders = isinstance(core.model['decoder'], torch.nn.ModuleList)
    shared, per_node = [], {}
    for name, p in model.named_parameters():
        m = _NODE_KEY.match(name)
        if m and (m.group(1) == 'encoder' or node_decoders):
            per_node.setdefault(int(m.group(2)), []).append(p)
        else:
            shared.append(p)
    return shared, per_node


def reduce_source_sharded(model, n_local, n_global, rank=None, world=None, group=None):
    """Gradient exchange of a source-sharded step: all-reduce the shared parameters' gradients, weight the owned
    nodes' gradients by n_local / N_global (the loss is a mean over the global batch), and leave the other
    nodes' parameters without a gradient so the optimizer skips them.  Returns the number of collectives."""
    if not dist.is_available() or not dist.is_initialized():
        return 0
    rank = dist.get_rank(group) if rank is None else rank
    world = dist.get_world_size(group) if world is None else world
    shared, per_node = split_parameters(model)
    n = allreduce_gradients(shared, n_local, n_global, group=group)
    w = float(n_local) / float(n_global)
    mine = []
    for j, ps in per_node.items():
        for p in ps:
            if j % world != rank:
                p.grad = None
            elif p.grad is not None and world > 1:
                mine.append(p.grad)
    if mine:
        torch._foreach_mul_(mine, w)
    return n


def sync_node_parameters(model, group=None):
    """broadcast every node's parameters from the rank that owns (trains) them"""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    world = dist.get_world_size(group)
    _, per_node = split_parameters(model)
    for j in sorted(per_node):
        flat = torch.cat([p.data.reshape(-1) for p in per_node[j]])
        dist.broadcast(flat, src=dist.get_global_rank(group, j % world) if group is not None else j % world, group=group)
        off = 0
        with torch.no_grad():
            for p in per_node[j]:
                p.copy_(flat[off:off + p.numel()].view_as(p))      # in place on the parameter: bumps its _version
                off += p.numel()


def gather_sharded_state(model, optimizer=None, group=None):
    """Make every rank's model -- and optimizer -- complete before a checkpoint of a source-sharded run
    (`{'model_dict': model.state_dict(), 'optimizer_dict': optimizer.state_dict()}`, train_model.py:87-98 /
    utils.resume): each node's parameters and the Adam moments / step counts of those parameters are broadcast
    from the rank that trains the node.  Without it rank 0's checkpoint would hold untrained encoders and no
    optimizer state for the other ranks' nodes, and a resume would silently restart them.  MANDATORY before
    `state_dict()` whenever `reduce_source_sharded` is the gradient exchange.  Works for optimizers whose per-
    parameter state is {'step', 'exp_avg', 'exp_avg_sq'} (torch.optim.Adam, engine.optim.FusedAdam)."""
    sync_node_parameters(model, group)
    if optimizer is None or not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    _, per_node = split_parameters(model)
    for j in sorted(per_node):
        owner = j % world
        src = dist.get_global_rank(group, owner) if group is not None else owner
        ps = per_node[j]
        dev = ps[0].device
        # does the owner hold any state yet (it has none before its first step)
        have = torch.tensor([1.0 if (rank == owner and all(len(optimizer.state[p]) for p in ps)) else 0.0], device=dev)
        dist.broadcast(have, src=src, group=group)
        if float(have) == 0.0:
            continue
        if rank == owner:
            steps = [float(optimizer.state[p]['step']) for p in ps]
            flat = torch.cat([optimizer.state[p][k].reshape(-1).float() for p in ps for k in ('exp_avg', 'exp_avg_sq')] +
                             [torch.tensor(steps, dtype=torch.float32, device=dev)])
        else:
            flat = torch.empty(2 * sum(p.numel() for p in ps) + len(ps), dtype=torch.float32, device=dev)
        dist.broadcast(flat, src=src, group=group)
        if rank == owner:
            continue
        off = 0
        steps = flat[-len(ps):].tolist()
        for i, p in enumerate(ps):
            st = optimizer.state[p]
            for k in ('exp_avg', 'exp_avg_sq'):
                st[k] = flat[off:off + p.numel()].view_as(p).clone()
                off += p.numel()
            st['step'] = torch.tensor(steps[i], dtype=torch.float32)
