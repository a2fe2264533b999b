"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink / NVSwitch).

Samples are independent through the whole coding loop (no batch statistics; "joint decoding" means
shared decoder weights, models/dist_codec.py:55), so the global batch is sharded across ranks and the
only exchange step of a training iteration is one all-reduce (sum) of the parameter gradients -- the
replacement of nn.DataParallel's per-step broadcast / scatter / gather / reduce_add
(train_model.py:59-60,114).  The reference loss is a mean over the global batch (dist_codec.py:58), so a
rank holding n_local of N_global samples contributes its mean-loss gradient with weight n_local/N_global.
Evaluation needs no collective at all."""
import torch
import torch.distributed as dist


def shard_range(n_global, rank, world):
    """contiguous shard [lo, hi) of rank; sizes differ by at most one"""
    base, rem = divmod(n_global, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_gradients(params, n_local, n_global, group=None, bucket_bytes=1 << 28):
    """In place: p.grad <- sum_r (n_r / N) * p.grad_r for every parameter (missing grads count as zero).
    Gradients travel in a few large fp32 buckets: NVSwitch makes collective cost latency-, not
    link-bound, so buckets are sized for launch count."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    w = float(n_local) / float(n_global)
    params = [p for p in params if p.requires_grad]
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
