"""Stand-alone ConvLSTM cell step on the gate kernel, for modules.ConvLSTMCell.forward when a cell is
called outside the fused loop (NCHW fp32 in / out like the reference, modules/cell.py:110-144).
Layout conversion between the caller's NCHW fp32 tensors and the kernel's NHWC fp16 planes is torch
plumbing; the convolution, gates and state update run in the CUDA kernel."""
import torch

from . import binding as B

ACT_SCALE = 256.0


def to_planes(x_nchw, split, pad_to=1, order=B.ORDER_IDENT):
    """NCHW fp32 -> NHWC fp16 planes {hi, lo} scaled by 2^8, batch padded to a multiple of pad_to."""
    x = x_nchw.detach().float().permute(0, 2, 3, 1)
    n, h, w, c = x.shape
    if order == B.ORDER_QUAD:
        x = x.reshape(n, h, w, c // 4, 4).permute(0, 1, 2, 4, 3).reshape(n, h, w, c)
    n_pad = (n + pad_to - 1) // pad_to * pad_to
    s = torch.zeros((n_pad, h, w, c), dtype=torch.float32, device=x.device)
    s[:n] = x * ACT_SCALE
    hi = s.half()
    lo = (s - hi.float()).half() if split else None
    return hi.contiguous(), (lo.contiguous() if lo is not None else None)


def from_nhwc(t, n, order=B.ORDER_IDENT):
    """NHWC fp32 (physical order) -> NCHW fp32 logical order, first n images."""
    t = t[:n]
    _, h, w, c = t.shape
    if order == B.ORDER_QUAD:
        t = t.reshape(n, h, w, 4, c // 4).permute(0, 1, 2, 4, 3).reshape(n, h, w, c)
    return t.permute(0, 3, 1, 2).contiguous()


class CellRunner(object):
    """Holds packed weights and recurrent state of one ConvLSTM cell."""

    def __init__(self, precision='parity'):
        self.split = 1 if precision == 'parity' else 0
        self.precise = self.split
        self._wkey = None
        self._w = None
        self.reset()

    def reset(self):
        self.h = None      # [ping, pong] of (hi, lo)
        self.c = None      # NHWC fp32
        self.h_f32 = None  # NHWC fp32 copy of the latest h
        self.t = 0
        self.n = 0

    def step(self, x, w_in, w_h):
        if not x.is_cuda:
            raise B.DrasicError('DRASIC B200 cells need CUDA tensors; there is no CPU path')
        L = B.lib()
        n, cin, H, W = x.shape
        co = w_h.shape[1]
        if tuple(w_in.shape) != (4 * co, cin, 3, 3) or tuple(w_h.shape) != (4 * co, co, 3, 3):
            raise ValueError('Not valid ConvLSTM weight shape')
        if cin % 64 or co % 64:
            raise ValueError('Not valid channel count for the B200 ConvLSTM kernel (multiples of 64)')
        bw = min(W, 128)
        if 128 % bw or W % bw:
            raise ValueError('Not valid width for the B200 ConvLSTM kernel')
        bh = min(H, 128 // bw)
        if (128 // bw) % bh or H % bh:
            raise ValueError('Not valid height for the B200 ConvLSTM kernel')
        pad_to = 128 // (bw * bh)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        dev = x.device
        n_pad = (n + pad_to - 1) // pad_to * pad_to
        n_mtiles = n_pad * H * W // 128
        sms = L.drasic_device_sms()
        block_n = 64
        for bn in (256, 128, 64):
            if (4 * co) % bn == 0 and n_mtiles * (4 * co // bn) >= sms:
                block_n = bn
                break
        key = (w_in.data_ptr(), w_in._version, w_h.data_ptr(), w_h._version, block_n)
        # re-packed at the start of every sequence as well: fused optimizers update parameters in place without
        # bumping Tensor._version, so the version key alone cannot be trusted
        if key != self._wkey or self.h is None:
            K = 9 * (cin + co)
            hi = torch.empty((4 * co, K), dtype=torch.float16, device=dev)
            lo = torch.empty((4 * co, K), dtype=torch.float16, device=dev)
            inv = torch.empty((1,), dtype=torch.float32, device=dev)
            scratch = torch.zeros((1,), dtype=torch.int32, device=dev)
            wi, wh = w_in.detach().float().contiguous(), w_h.detach().float().contiguous()
            B.check(L.drasic_pack_lstm_weights(wi.data_ptr(), wh.data_ptr(), cin, co, B.ORDER_IDENT, B.ORDER_IDENT,
                                               block_n, hi.data_ptr(), lo.data_ptr(), inv.data_ptr(),
                                               scratch.data_ptr(), stream), 'pack_lstm_weights')
            self._w, self._wkey = (hi, lo, inv, wi, wh, scratch), key
        hi, lo, inv = self._w[:3]
        if self.h is not None and (self.n != n or tuple(self.c.shape) != (n_pad, H, W, co)):
            raise RuntimeError('hidden state shape mismatch: call free_hidden() first')  # as the reference
        if self.h is None:
            def planes():
                return (torch.zeros((n_pad, H, W, co), dtype=torch.float16, device=dev),
                        torch.zeros((n_pad, H, W, co), dtype=torch.float16, device=dev) if self.split else None)
            self.h = [planes(), planes()]
            self.c = torch.zeros((n_pad, H, W, co), dtype=torch.float32, device=dev)
            self.h_f32 = torch.zeros((n_pad, H, W, co), dtype=torch.float32, device=dev)
            self.t, self.n = 0, n
        x_hi, x_lo = to_planes(x, self.split, pad_to)
        cur, prv = self.h[self.t & 1], self.h[(self.t & 1) ^ 1]
        a = B.ConvLSTMArgs()
        P = B.ptr
        a.x_hi, a.x_lo = P(x_hi), P(x_lo)
        a.h_hi, a.h_lo = P(prv[0]), P(prv[1])
        a.w_hi, a.w_lo, a.w_inv_scale = P(hi), P(lo), P(inv)
        a.c_state = P(self.c)
        a.h_out_hi, a.h_out_lo = P(cur[0]), P(cur[1])
        a.next_hi = P(self.h_f32)
        a.B_pad, a.H, a.W, a.Cin, a.Co = n_pad, H, W, cin, co
        a.n_groups = 1
        a.n_units = n_mtiles              # one work unit per M tile (single CTAs, one weight group)
        a.has_state = 1 if self.t > 0 else 0
        a.next_mode, a.next_f32 = B.NEXT_SAME, 1
        a.split, a.precise, a.block_n = self.split, self.precise, block_n
        B.check(L.drasic_convlstm_fwd(a, stream), 'convlstm_fwd')
        self.t += 1
        self._keep = (x_hi, x_lo)
        return from_nhwc(self.h_f32, n)

    def hidden_nchw(self):
        """[[h],[c]] like ConvLSTMCell.hidden (modules/cell.py:142)"""
        if self.c is None:
            return None
        return [[from_nhwc(self.h_f32, self.n)], [from_nhwc(self.c, self.n)]]
