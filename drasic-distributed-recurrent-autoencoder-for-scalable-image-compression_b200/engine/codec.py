"""Host side of the fused residual coding loop (models/dist_codec.py:36-65, sep_codec.py:35-66).

Per forward call: sort the batch by node once (labels are constant over the T iterations; the
reference re-gathers 2*M times per iteration with a host sync each, dist_codec.py:45-52), pack the
ConvLSTM weights if they changed, then launch per iteration
    E1 E2 E3 (gate kernels, grouped by node) -> bottleneck -> D1 D2 D3 (gate kernels) -> tail/head
on the caller's CUDA stream through the C ABI.  torch is used for device memory and streams only.
"""
import numpy as np
import torch

from . import binding as B

PAD = 8  # slots per node are padded to a multiple of 8 images (one 128-pixel M tile at 4x4)

# (name, Cin, Co, spatial divisor, in_order, out_order, next_mode, next_f32)
ENC_LAYERS = (('E1', 128, 128, 2, B.ORDER_QUAD, B.ORDER_IDENT, B.NEXT_DOWN, 0),
              ('E2', 512, 128, 4, B.ORDER_QUAD, B.ORDER_IDENT, B.NEXT_DOWN, 0),
              ('E3', 512, 128, 8, B.ORDER_QUAD, B.ORDER_IDENT, B.NEXT_SAME, 1))
DEC_LAYERS = (('D1', 128, 512, 8, B.ORDER_IDENT, B.ORDER_QUAD, B.NEXT_UP, 0),
              ('D2', 128, 512, 4, B.ORDER_IDENT, B.ORDER_QUAD, B.NEXT_UP, 0),
              ('D3', 128, 128, 2, B.ORDER_IDENT, B.ORDER_QUAD, B.NEXT_UP, 1))
ENC_IDX = {'E1': 2, 'E2': 4, 'E3': 6}   # nn.Sequential indices, dist_codec.py:70-86
DEC_IDX = {'D1': 1, 'D2': 3, 'D3': 5}   # dist_codec.py:88-104


class BatchPlan(object):
    """Node-sorted, padded slot layout of one batch."""

    def __init__(self, label_host, num_node, device, pad_total=None):
        label_host = np.asarray(label_host).astype(np.int64).reshape(-1)
        if label_host.size and (label_host.min() < 0 or label_host.max() >= num_node):
            # the reference mis-indexes and raises IndexError for label >= num_node (dist_codec.py:52)
            raise IndexError('label out of range for num_node={}'.format(num_node))
        self.B = int(label_host.size)
        order = np.argsort(label_host, kind='stable')          # keeps the reference's within-node order
        counts = np.bincount(label_host, minlength=num_node)
        padded = (counts + PAD - 1) // PAD * PAD
        ends = np.cumsum(padded)
        self.B_pad = int(ends[-1]) if pad_total is None else int(pad_total)
        if self.B_pad < int(ends[-1]) or self.B_pad % PAD:
            raise ValueError('Not valid pad_total')
        perm = np.full(self.B_pad, -1, dtype=np.int32)
        starts = ends - padded
        src = 0
        for j in range(num_node):
            n = int(counts[j])
            perm[starts[j]:starts[j] + n] = order[src:src + n]
            src += n
        self.counts = counts
        self.group_img_end_host = ends.astype(np.int32)
        self.perm_host = perm
        self.perm = torch.from_numpy(perm).to(device)
        self.group_img_end = torch.from_numpy(self.group_img_end_host).to(device)
        self._mtile_end = {}
        self.device = device

    def mtile_end(self, hw):
        """exclusive end M-tile (128 pixels) of each node at a layer with hw pixels per image"""
        if hw not in self._mtile_end:
            e = (self.group_img_end_host.astype(np.int64) * hw) // 128
            self._mtile_end[hw] = torch.from_numpy(e.astype(np.int32)).to(self.device)
        return self._mtile_end[hw]


class _Layer(object):
    pass


class CodecEngine(object):
    """Runs the T-iteration loop for one codec (dist: per-node encoders + shared decoder; sep: per-node
    encoders and decoders).  `precision`: 'parity' = fp16 hi+lo operands with IEEE transcendental
    functions (fp32-grade, bit-exact codes); 'fast' = single fp16 operands with tanh.approx."""

    def __init__(self, num_channel, code_size, num_node, separate=False, precision='parity'):
        if precision not in ('parity', 'fast'):
            raise ValueError('Not valid precision')
        self.C, self.code, self.M, self.separate = num_channel, code_size, num_node, separate
        self.split = 1 if precision == 'parity' else 0
        self.precise = self.split
        self.Md = num_node if separate else 1
        self._arena = None
        self._arena_key = None
        self._wcache = {}
        self._sms = None
        self.launches = 0   # kernels launched by the last forward() (bench.py reports it)
        self.profile = None  # set to [] to collect (label, flops, start_event, end_event) per launch

    # ------------------------------------------------------------------ helpers
    def sms(self):
        if self._sms is None:
            n = B.lib().drasic_device_sms()
            if n <= 0:
                raise B.DrasicError('no CUDA device: ' + B.lib().drasic_last_error().decode())
            self._sms = n
        return self._sms

    def _timed(self, label, flops, fn):
        if self.profile is None:
            return fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        self.profile.append((label, flops, e0, e1))
        return r

    def profile_summary(self):
        """{label: (launches, total ms, total algorithmic flops)} after a synchronize"""
        out = {}
        for label, flops, e0, e1 in self.profile or []:
            n, ms, fl = out.get(label, (0, 0.0, 0.0))
            out[label] = (n + 1, ms + e0.elapsed_time(e1), fl + flops)
        return out

    def _block_n(self, n_mtiles, co):
        for bn in (256, 128, 64):
            if n_mtiles * (4 * co // bn) >= self.sms():
                return bn
        return 64

    def _get_arena(self, B_pad, H, W, device):
        key = (B_pad, H, W, str(device))
        if self._arena_key == key:
            return self._arena
        A = {}
        f16 = dict(dtype=torch.float16, device=device)
        f32 = dict(dtype=torch.float32, device=device)

        def planes(*shape):
            return [torch.zeros(shape, **f16), torch.zeros(shape, **f16) if self.split else None]
        A['e1_in'] = planes(B_pad, H // 2, W // 2, 128)
        A['d1_in'] = planes(B_pad, H // 8, W // 8, 128)
        for (name, cin, co, div, _, _, nmode, nf32) in ENC_LAYERS + DEC_LAYERS:
            h, w = H // div, W // div
            L = {}
            L['h'] = [planes(B_pad, h, w, co), planes(B_pad, h, w, co)]   # ping-pong
            L['c'] = torch.zeros((B_pad, h, w, co), **f32)
            if nmode == B.NEXT_DOWN:
                L['next'] = planes(B_pad, h // 2, w // 2, 4 * co)
            elif nmode == B.NEXT_UP:
                L['next'] = [torch.zeros((B_pad, 2 * h, 2 * w, co // 4), **f32), None] if nf32 else \
                    planes(B_pad, 2 * h, 2 * w, co // 4)
            else:
                L['next'] = [torch.zeros((B_pad, h, w, co), **f32), None] if nf32 else planes(B_pad, h, w, co)
            A[name] = L
        self._arena, self._arena_key = A, key
        return A

    def _packed(self, name, weights, cin, co, in_order, out_order, block_n, stream):
        """weights: list over groups of (w_in, w_h) fp32 tensors. Cached on tensor identity+version."""
        key = (name, block_n) + tuple((w.data_ptr(), w._version) for pair in weights for w in pair)
        hit = self._wcache.get(name)
        if hit is not None and hit[0] == key:
            return hit[1]
        L = B.lib()
        G = len(weights)
        dev = weights[0][0].device
        K = 9 * (cin + co)
        hi = torch.empty((G * 4 * co, K), dtype=torch.float16, device=dev)
        lo = torch.empty((G * 4 * co, K), dtype=torch.float16, device=dev)
        inv = torch.empty((G,), dtype=torch.float32, device=dev)
        scratch = torch.zeros((G,), dtype=torch.int32, device=dev)
        plane = 4 * co * K * 2
        for g, (w_in, w_h) in enumerate(weights):
            if tuple(w_in.shape) != (4 * co, cin, 3, 3) or tuple(w_h.shape) != (4 * co, co, 3, 3):
                raise ValueError('Not valid ConvLSTM weight shape for {}'.format(name))
            w_in = w_in.detach().contiguous()
            w_h = w_h.detach().contiguous()
            B.check(L.drasic_pack_lstm_weights(w_in.data_ptr(), w_h.data_ptr(), cin, co, in_order, out_order,
                                               block_n, hi.data_ptr() + g * plane, lo.data_ptr() + g * plane,
                                               inv.data_ptr() + 4 * g, scratch.data_ptr() + 4 * g, stream),
                    'pack_lstm_weights')
            self.launches += 2
        val = (hi, lo, inv, (w_in, w_h))
        self._wcache[name] = (key, val)
        return val

    # ------------------------------------------------------------------ the loop
    def forward(self, sd, img, label, num_iter, training=False, noise=None, loss_kind='mae',
                label_host=None, want_z=False):
        """sd: mapping reference state_dict key -> fp32 CUDA tensor.  img [B,C,H,W] fp32 in [0,1],
        label [B] node ids.  noise: None (eval / deterministic sign) or [T,B,code,H/8,W/8] uniforms.
        Returns dict(recon [T,B,C,H,W], codes [T,B,code,H/8,W/8], bits [T,B,code*HW/64/8] uint8,
        loss_acc [T,2] fp64 {sum loss terms, sum squared error}, z (optional))."""
        if loss_kind not in B.LOSS_KINDS:
            raise ValueError('Not valid loss function')
        if not img.is_cuda:
            raise B.DrasicError('DRASIC B200 engine needs CUDA tensors; there is no CPU path')
        L = B.lib()
        dev = img.device
        img = img.detach().contiguous().float()
        Bn, Cc, H, W = img.shape
        if Cc != self.C:
            raise ValueError('Not valid input channels')
        if H % 8 or W % 8 or (H * W) % 256:
            raise ValueError('Not valid image size {}x{} for the fused path'.format(H, W))
        T = int(num_iter)
        if label_host is None:
            label_host = label.detach().cpu().numpy()          # the one host sync per forward
        plan = BatchPlan(label_host, self.M, dev)
        Bp = plan.B_pad
        stream = torch.cuda.current_stream(dev).cuda_stream
        self.launches = 0
        A = self._get_arena(Bp, H, W, dev)
        code, hc, wc = self.code, H // 8, W // 8
        if training and noise is None:
            noise = torch.rand((T, Bn, code, hc, wc), device=dev, dtype=torch.float32)
        if not training:
            noise = None
        if noise is not None:
            noise = noise.detach().contiguous().float()

        # ---- parameters
        def lstm_w(prefix, idx):
            base = '{}.{}.cell.cell.0.'.format(prefix, idx)
            return sd[base + 'in.cell.cell.main.weight'], sd[base + 'hidden.cell.cell.main.weight']
        enc_pref = ['model.encoder.{}'.format(j) for j in range(self.M)]
        dec_pref = ['model.decoder.{}'.format(j) for j in range(self.M)] if self.separate else ['model.decoder']
        layers = []
        for (name, cin, co, div, io, oo, nmode, nf32) in ENC_LAYERS + DEC_LAYERS:
            is_enc = name[0] == 'E'
            prefs = enc_pref if is_enc else dec_pref
            idx = ENC_IDX[name] if is_enc else DEC_IDX[name]
            h, w = H // div, W // div
            n_mtiles = Bp * h * w // 128
            ly = _Layer()
            ly.name, ly.cin, ly.co, ly.h, ly.w = name, cin, co, h, w
            ly.next_mode, ly.next_f32 = nmode, nf32
            ly.block_n = self._block_n(n_mtiles, co)
            ly.groups = len(prefs)
            ly.hi, ly.lo, ly.inv, _ = self._packed(name, [lstm_w(p, idx) for p in prefs], cin, co, io, oo,
                                                   ly.block_n, stream)
            ly.mtile_end = plan.mtile_end(h * w) if ly.groups > 1 else None
            layers.append(ly)

        def stack(keys):
            return torch.stack([sd[k].detach().float().reshape(sd[k].shape[0], -1) for k in keys]).contiguous()
        w_e0 = stack([p + '.0.cell.cell.main.weight' for p in enc_pref])             # [M,32,C]
        b_key = enc_pref[0] + '.0.cell.cell.main.bias'
        b_e0 = stack([p + '.0.cell.cell.main.bias' for p in enc_pref]) if b_key in sd else None
        w_e4 = stack([p + '.7.cell.cell.main.weight' for p in enc_pref])             # [M,code,128]
        w_d0 = stack([p + '.0.cell.cell.main.weight' for p in dec_pref])             # [Md,128,code]
        w_d4 = stack([p + '.7.cell.cell.main.weight' for p in dec_pref])             # [Md,C,32]

        # ---- outputs
        recon = torch.empty((T, Bn, Cc, H, W), dtype=torch.float32, device=dev)
        codes = torch.empty((T, Bn, code, hc, wc), dtype=torch.float32, device=dev)
        zbuf = torch.empty((T, Bn, code, hc, wc), dtype=torch.float32, device=dev) if want_z else None
        bits = torch.empty((T, Bn, code * hc * wc // 8), dtype=torch.uint8, device=dev)
        loss_acc = torch.zeros((T, 2), dtype=torch.float64, device=dev)

        P = B.ptr

        def tail(mode, t):
            a = B.TailHeadArgs()
            last = (mode != B.TAIL_INIT and t == T - 1)
            if mode != B.TAIL_INIT:
                a.d3_next = P(A['D3']['next'][0])
                a.w_d4 = P(w_d4)
                a.recon_out = recon[t].data_ptr()
                a.recon_prev = recon[t - 1].data_ptr() if mode == B.TAIL_STEP else None
                a.loss_acc = loss_acc[t].data_ptr()
            a.w_e0, a.b_e0, a.img, a.perm = P(w_e0), P(b_e0), P(img), P(plan.perm)
            if not last:
                a.e1_in_hi, a.e1_in_lo = P(A['e1_in'][0]), P(A['e1_in'][1])
            a.group_img_end = P(plan.group_img_end)
            a.B_pad, a.C, a.H, a.W, a.mode = Bp, Cc, H, W, mode
            a.loss_kind = B.LOSS_KINDS[loss_kind]
            a.n_enc_groups, a.n_dec_groups = self.M, self.Md
            a.precise, a.split = self.precise, self.split
            self._timed('tail_head', 0.0, lambda: B.check(L.drasic_tail_head_fwd(a, stream), 'tail_head_fwd'))
            self.launches += 1

        def gate(ly, x_planes, t):
            Lb = A[ly.name]
            cur, prv = Lb['h'][t & 1], Lb['h'][(t & 1) ^ 1]
            a = B.ConvLSTMArgs()
            a.x_hi, a.x_lo = P(x_planes[0]), P(x_planes[1])
            a.h_hi, a.h_lo = P(prv[0]), P(prv[1])
            a.w_hi, a.w_lo, a.w_inv_scale = P(ly.hi), P(ly.lo), P(ly.inv)
            a.c_state = P(Lb['c'])
            a.h_out_hi, a.h_out_lo = P(cur[0]), P(cur[1])
            a.next_hi, a.next_lo = P(Lb['next'][0]), P(Lb['next'][1])
            a.group_mtile_end = P(ly.mtile_end)
            a.B_pad, a.H, a.W, a.Cin, a.Co = Bp, ly.h, ly.w, ly.cin, ly.co
            a.n_groups = ly.groups
            a.has_state = 1 if t > 0 else 0
            a.next_mode, a.next_f32 = ly.next_mode, ly.next_f32
            a.split, a.precise, a.block_n = self.split, self.precise, ly.block_n
            # algorithmic FLOPs of this launch: real images only, hidden conv skipped on zero state
            k = 9 * (ly.cin + (ly.co if t > 0 else 0))
            flops = 2.0 * Bn * ly.h * ly.w * 4 * ly.co * k
            self._timed('gates_' + ly.name, flops, lambda: B.check(L.drasic_convlstm_fwd(a, stream), 'convlstm_fwd'))
            self.launches += 1
            return Lb['next']

        tail(B.TAIL_INIT, -1)
        for t in range(T):
            x = A['e1_in']
            for ly in layers[:3]:
                x = gate(ly, x, t)
            a = B.BottleneckArgs()
            a.h_e3, a.w_e4, a.w_d0 = P(x[0]), P(w_e4), P(w_d0)
            a.noise = noise[t].data_ptr() if noise is not None else None
            a.perm = P(plan.perm)
            a.code_out = codes[t].data_ptr()
            a.z_out = zbuf[t].data_ptr() if zbuf is not None else None
            a.bits_out = bits[t].data_ptr()
            a.d1_in_hi, a.d1_in_lo = P(A['d1_in'][0]), P(A['d1_in'][1])
            a.group_img_end = P(plan.group_img_end)
            a.B_pad, a.HW, a.Ch, a.code = Bp, hc * wc, 128, code
            a.n_enc_groups, a.n_dec_groups = self.M, self.Md
            a.precise, a.split = self.precise, self.split
            self._timed('bottleneck', 0.0, lambda a=a: B.check(L.drasic_bottleneck_fwd(a, stream), 'bottleneck_fwd'))
            self.launches += 1
            x = A['d1_in']
            for ly in layers[3:]:
                x = gate(ly, x, t)
            tail(B.TAIL_FIRST if t == 0 else B.TAIL_STEP, t)
        out = {'recon': recon, 'codes': codes, 'bits': bits, 'loss_acc': loss_acc, 'plan': plan}
        if zbuf is not None:
            out['z'] = zbuf
        # keep the small stacked weights alive until the stream has consumed them
        out['_keep'] = (w_e0, b_e0, w_e4, w_d0, w_d4, noise, img)
        return out

    @staticmethod
    def loss_from_acc(loss_acc, numel):
        """mean over T of the per-iteration mean loss (dist_codec.py:58,61)"""
        T = loss_acc.shape[0]
        return (loss_acc[:, 0].sum() / (float(numel) * T)).float()
