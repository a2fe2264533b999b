"""+-1 binarizer with straight-through gradient (reference functions/quantize.py:4-22).
Training: +1 with probability (x+1)/2; eval: sign with -0.0 -> +1, values outside [-1,1] and NaN
passed through.  Runs in the drasic_quantize_fwd CUDA kernel; the uniforms are drawn with torch's
generator (as the reference does) and handed to the kernel, or can be passed in for reproducibility."""
import torch
from torch.autograd.function import Function

from engine import binding as B

__all__ = ['Quantize']


class Quantize(Function):
    @staticmethod
    def forward(ctx, input, is_training, noise=None):
        if not input.is_cuda:
            raise B.DrasicError('DRASIC B200 Quantize needs a CUDA tensor; there is no CPU path')
        x = input.detach().contiguous().float()
        u = None
        if is_training:
            u = noise if noise is not None else x.new(x.size()).uniform_()
            u = u.detach().contiguous().float()
        y = torch.empty_like(x)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        B.check(B.lib().drasic_quantize_fwd(x.data_ptr(), B.ptr(u), y.data_ptr(), x.numel(), stream),
                'quantize_fwd')
        return y.to(input.dtype)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output, None, None
