"""Adam on the device in a few launches of the library's own multi-tensor kernel (csrc/optim.cu through
drasic_adam_step): the `optimizer.step()` of the reference's training loop (train_model.py:116) for the
optimizer `make_optimizer` builds at train_model.py:155-161 (optim.Adam with lr and L2 weight_decay).

Drop-in for torch.optim.Adam on the codec's parameters: same constructor arguments, same update rule, and the
same per-parameter state ('step', 'exp_avg', 'exp_avg_sq'), so `optimizer.state_dict()` checkpoints written by
either (train_model.py:90, utils.py:173-180) load into the other."""

import torch

from . import binding as B


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if not 0.0 <= lr:
            raise ValueError('Invalid learning rate: {}'.format(lr))
        if not 0.0 <= eps:
            raise ValueError('Invalid epsilon value: {}'.format(eps))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError('Invalid beta parameter at index 0: {}'.format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError('Invalid beta parameter at index 1: {}'.format(betas[1]))
        if not 0.0 <= weight_decay:
            raise ValueError('Invalid weight_decay value: {}'.format(weight_decay))
        if amsgrad:
            raise ValueError('Not valid optimizer option: amsgrad (the reference never sets it)')
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False))
        self.launches = 0     # kernels launched by the last step()

    @staticmethod
    def _step_count(state):
        s = state['step']
        if torch.is_tensor(s):
            if s.is_cuda:      # a checkpoint of torch's fused / capturable Adam keeps the count on the device
                s = state['step'] = s.detach().to('cpu', torch.float32)
            return s
        state['step'] = torch.tensor(float(s), dtype=torch.float32)
        return state['step']

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = B.lib()
        self.launches = 0
        for group in self.param_groups:
            # options a loaded torch.optim.Adam state_dict may carry that this kernel does not implement
            for opt in ('amsgrad', 'maximize', 'decoupled_weight_decay', 'differentiable'):
                if group.get(opt):
                    raise ValueError('Not valid optimizer option for FusedAdam: {}'.format(opt))
            beta1, beta2 = group['betas']
            buckets = {}      # (device, step count) -> [(param, grad, exp_avg, exp_avg_sq)]
            for p in group['params']:
                g = p.grad
                if g is None:
                    continue
                if g.is_sparse:
                    raise RuntimeError('Adam does not support sparse gradients')
                if not p.is_cuda:
                    raise B.DrasicError('FusedAdam needs CUDA parameters; there is no CPU path')
                if p.dtype != torch.float32 or g.dtype != torch.float32 or not p.is_contiguous():
                    raise ValueError('Not valid parameter for FusedAdam: contiguous fp32 tensors only')
                if not g.is_contiguous():
                    g = g.contiguous()
                state = self.state[p]
                if len(state) == 0:
                    state['step'] = torch.tensor(0.0, dtype=torch.float32)
                    state['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                step = self._step_count(state)
                step += 1
                buckets.setdefault((p.device, float(step)), []).append((p, g, state['exp_avg'], state['exp_avg_sq']))
            for (dev, step), items in buckets.items():
                arr = (B.AdamTensor * len(items))()
                for a, (p, g, m, v) in zip(arr, items):
                    a.param, a.grad, a.exp_avg, a.exp_avg_sq = p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr()
                    a.numel = p.numel()
                with torch.cuda.device(dev):
                    B.check(L.drasic_adam_step(arr, len(items), group['lr'], beta1, beta2, group['eps'],
                                               group['weight_decay'], 1.0 - beta1 ** step, 1.0 - beta2 ** step,
                                               torch.cuda.current_stream(dev).cuda_stream), 'adam_step')
                self.launches += (len(items) + 63) // 64
                # the kernel wrote the parameters behind autograd's back: bump their version counters so that
                # anything keyed on (data_ptr, _version) -- the engine's packed fp16 weight planes -- sees the update
                torch.autograd.graph.increment_version([p for p, _, _, _ in items])
        return loss
