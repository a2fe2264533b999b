"""Shared implementation of the two codec Models (reference models/dist_codec.py:10-65 and
models/sep_codec.py:10-66): the architecture spec builders and the fused forward that hands the whole
T-iteration residual loop to engine.CodecEngine."""
from collections.abc import Sequence

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

import config
from engine.codec import CodecEngine, GraphedCodec
from modules.cell import ConvCell, ConvLSTMCell, PixelShuffleCell
from .utils import make_model


def _conv(cin, cout, bias):
    return {'cell': 'ConvCell', 'input_size': cin, 'output_size': cout, 'kernel_size': 1, 'stride': 1,
            'padding': 0, 'bias': bias, 'activation': config.PARAM['activation'],
            'normalization': config.PARAM['normalization']}


def _lstm(cin, cout):
    return {'cell': 'ConvLSTMCell', 'input_size': cin, 'output_size': cout, 'kernel_size': 3, 'stride': 1,
            'padding': 1, 'bias': False, 'num_layers': 1, 'activation': config.PARAM['activation']}


def _shuffle(mode):
    return {'cell': 'PixelShuffleCell', 'mode': mode, 'scale_factor': 2}


def encoder_spec(head_bias):
    """C -> 32 @HxW | down | LSTM 128->128 | down | LSTM 512->128 | down | LSTM 512->128 | 128 -> code
    (dist_codec.py:70-86; the head conv has a bias in dist_codec and none in sep_codec.py:73)"""
    return (_conv(config.PARAM['num_channel'], 32, head_bias), _shuffle('down'), _lstm(128, 128),
            _shuffle('down'), _lstm(512, 128), _shuffle('down'), _lstm(512, 128),
            _conv(128, config.PARAM['code_size'], False))


def decoder_spec():
    """code -> 128 | LSTM 128->512 | up | LSTM 128->512 | up | LSTM 128->128 | up | 32 -> C
    (dist_codec.py:88-104)"""
    return (_conv(config.PARAM['code_size'], 128, False), _lstm(128, 512), _shuffle('up'), _lstm(128, 512),
            _shuffle('up'), _lstm(128, 128), _shuffle('up'), _conv(32, config.PARAM['num_channel'], False))


_ENC_KINDS = (ConvCell, PixelShuffleCell, ConvLSTMCell, PixelShuffleCell, ConvLSTMCell, PixelShuffleCell,
              ConvLSTMCell, ConvCell)
_DEC_KINDS = (ConvCell, ConvLSTMCell, PixelShuffleCell, ConvLSTMCell, PixelShuffleCell, ConvLSTMCell,
              PixelShuffleCell, ConvCell)


class CodeList(Sequence):
    """output['code'] of the reference (dist_codec.py:42,61-63): a list of T cumulative packed-code arrays.  The
    reference fills it with one device->host copy per iteration; here the arrays are built from ONE copy at the
    first access (len() needs none), so a training step that never looks at them -- only metrics.BPP does,
    metrics.py:84-89 -- has no host synchronisation between its forward and its backward."""

    def __init__(self, n, make):
        self._n, self._make, self._items = n, make, None

    def _get(self):
        if self._items is None:
            self._items = self._make()
            self._make = None
        return self._items

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        return self._get()[i]

    def __iter__(self):
        return iter(self._get())

    def __repr__(self):
        return 'CodeList({})'.format('pending' if self._items is None else repr(self._items))


class _LoopFunction(torch.autograd.Function):
    """The whole T-iteration loop as one autograd node: forward launches the fused kernels, backward
    (backward-through-time) is provided by the engine when it has been built with training support."""

    @staticmethod
    def forward(ctx, model, img, label, noise, *params):
        ctx.model = model
        ctx.names = model._index()[0]
        if config.PARAM.get('cuda_graph', False) and config.PARAM.get('node_iters') is None:
            # one graph holds forward AND backward-through-time; backward() only hands the gradients out
            out, ctx.grads = model._run_graphed(img, label, noise, with_backward=True)
        else:
            out = model._run_engine(img, label, noise, save_for_backward=True)
            ctx.grads = None
        ctx.saved = out
        ctx.mark_non_differentiable(out['recon'], out['codes'], out['bits'])
        loss = CodecEngine.loss_from_acc(out['loss_acc'], img.numel())
        return loss, out['recon'], out['codes'], out['bits']

    @staticmethod
    def backward(ctx, g_loss, g_recon, g_codes, g_bits):
        if ctx.grads is not None:
            grads = {k: v * g_loss for k, v in ctx.grads.items()}
        else:
            grads = ctx.model._engine.backward(ctx.saved, g_loss)
        return (None, None, None, None) + tuple(grads[k] for k in ctx.names)


class CodecModel(nn.Module):
    separate = False

    def __init__(self):
        super(CodecModel, self).__init__()
        self.model = make_model(config.PARAM['model'])
        self._engine = None
        self._engine_key = None
        self._tree = None

    def _index(self):
        """(parameter names, parameters, modules with free_hidden) of the module tree, walked once: the tree has
        ~1800 modules and three walks per step cost more host time than issuing a small batch's kernels.  The
        Parameter objects survive .to() / load_state_dict(); both drop the index anyway."""
        if self._tree is None:
            named = list(self.named_parameters())
            cells = [m for m in self.modules() if m is not self and hasattr(m, 'free_hidden')]
            self._tree = ([k for k, _ in named], [p for _, p in named], cells)
        return self._tree

    def _apply(self, fn, *args, **kwargs):
        self._tree = None
        return super(CodecModel, self)._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._tree = None
        return super(CodecModel, self).load_state_dict(*args, **kwargs)

    def _free_hidden(self):
        """apply_fn(self, 'free_hidden') (dist_codec.py:64) over the indexed cells"""
        for m in self._index()[2]:
            m.free_hidden()

    # ---- reference API -------------------------------------------------------------------------
    def pack(self, code):
        """dist_codec.py:16-18.  (-1 and +1 are both non-zero, so every byte is 0xFF; only the size of
        the result is consumed downstream, metrics.py:84-89.)"""
        return np.packbits(code.detach().cpu().numpy().astype(np.int8)).reshape(-1, 1)

    def unpack(self, code):
        """dist_codec.py:20-22"""
        return torch.from_numpy(code.astype(np.float32)).to(config.PARAM['device'])

    def loss_fn(self, output, target):
        """dist_codec.py:24-34"""
        kind = config.PARAM['loss_fn']
        if kind == 'bce':
            return F.binary_cross_entropy(output, target)
        if kind == 'mse':
            return F.mse_loss(output, target)
        if kind == 'mae':
            return F.l1_loss(output, target)
        raise ValueError('Not valid loss function')

    # ---- fused path ----------------------------------------------------------------------------
    def _check_architecture(self):
        M = config.PARAM['num_node']
        enc = self.model['encoder']
        decs = list(self.model['decoder']) if self.separate else [self.model['decoder']]
        if len(enc) != M or (self.separate and len(decs) != M):
            raise ValueError('Not valid model: num_node does not match the module tree')
        for seq, kinds in [(e, _ENC_KINDS) for e in enc] + [(d, _DEC_KINDS) for d in decs]:
            if len(seq) != len(kinds) or any(not isinstance(c.cell, k) for c, k in zip(seq, kinds)):
                raise ValueError('Not valid model: the B200 path implements the dist_codec/sep_codec architecture')
            for c in seq:
                if isinstance(c.cell, ConvLSTMCell):
                    c.cell._check_supported()
                elif isinstance(c.cell, ConvCell) and c.cell.activation_mode != 'tanh':
                    raise ValueError('Not valid activation for the fused B200 path (tanh)')
        if config.PARAM.get('normalization', 'none') != 'none':
            raise ValueError('Not valid normalization for the B200 path (only none)')

    def _ensure_engine(self):
        key = (config.PARAM['num_channel'], config.PARAM['code_size'], config.PARAM['num_node'],
               config.PARAM.get('precision', 'parity'), config.PARAM.get('grad_precision', 'fp16'),
               config.PARAM.get('cta_pair', 'auto'), config.PARAM.get('stream_k', 'auto'),
               bool(config.PARAM.get('deterministic', False)))
        if self._engine_key != key:
            self._check_architecture()
            self._engine = CodecEngine(key[0], key[1], key[2], separate=self.separate, precision=key[3],
                                       grad_precision=key[4], cta_pair=key[5], stream_k=key[6], deterministic=key[7])
            self._engine_key = key
            self._graphs = {}

    def _run_graphed(self, img, label, noise, with_backward=False):
        self._ensure_engine()
        if not img.is_cuda:
            from engine.binding import DrasicError
            raise DrasicError('DRASIC B200 engine needs CUDA tensors; there is no CPU path')
        key = (tuple(img.shape), config.PARAM['num_iter'], self.training, with_backward, config.PARAM['loss_fn'])
        g = self._graphs.get(key)
        if g is None:
            sd = dict(zip(*self._index()[:2]))
            g = GraphedCodec(self._engine, sd, img.shape[0], tuple(img.shape[2:]), config.PARAM['num_iter'],
                             training=self.training, with_backward=with_backward, loss_kind=config.PARAM['loss_fn'])
            self._graphs[key] = g
        return g(img.detach(), label.detach().cpu().numpy(), noise)

    def _run_engine(self, img, label, noise, save_for_backward=False):
        self._ensure_engine()
        sd = dict(zip(*self._index()[:2]))
        return self._engine.forward(sd, img, label, config.PARAM['num_iter'], training=self.training,
                                    noise=noise, loss_kind=config.PARAM['loss_fn'],
                                    save_for_backward=save_for_backward, node_iters=config.PARAM.get('node_iters'))

    def release(self):
        """Drop the engine's device arenas (recurrent state, per-iteration saved tensors, backward work space,
        packed weight planes, captured graphs); they are rebuilt on the next call.  The training arena at
        B=1024, T=16 is ~44 GB (2.6 MB per image and iteration)."""
        self._graphs = {}
        if self._engine is not None:
            self._engine.release()

    # ---- progressive bitstream ---------------------------------------------------
    def encode(self, input):
        """Run the coding loop and return the real bitstream: {'bits': uint8 [T, B, code*H*W/512] (iteration-major,
        so bits[:t] is the rate-t prefix), 'size': (H, W), 'label': node ids}.  1 bit per symbol, numpy packbits
        order over the NCHW code tensor of each image."""
        was_training = self.training
        self.train(False)
        try:
            with torch.no_grad():
                res = self._run_engine(input['img'], input['label'], None)
        finally:
            self.train(was_training)
            self._free_hidden()
        return {'bits': res['bits'], 'size': tuple(input['img'].shape[2:]), 'label': input['label']}

    def decode(self, bits, label, size=(32, 32)):
        """Decoder-only reconstruction from a bitstream prefix bits[:t] (t <= T): list of t cumulative
        reconstructions, bit-identical to forward()['img'][:t] for the same weights."""
        self._ensure_engine()
        sd = dict(zip(*self._index()[:2]))
        with torch.no_grad():
            res = self._engine.decode(sd, bits, label, tuple(size))
        self._free_hidden()
        return [res['recon'][t] for t in range(res['recon'].shape[0])]

    @staticmethod
    def unpack_bits(bits, code_shape):
        """uint8 [..., n/8] -> float +-1 symbols [..., *code_shape] (inverse of the device bit-pack)"""
        shifts = torch.arange(7, -1, -1, device=bits.device, dtype=torch.uint8)
        b = (bits.unsqueeze(-1) >> shifts) & 1
        return (b.reshape(bits.shape[:-1] + tuple(code_shape)).float() * 2 - 1)

    def forward(self, input, noise=None):
        """input: {'img': [B,C,H,W] in [0,1], 'label': [B] node ids}; returns {'loss', 'img': list of T
        cumulative reconstructions, 'code': list of T cumulative packed codes} like the reference
        (dist_codec.py:36-65), plus 'code_pm1' ([T,B,code,H/8,W/8] raw +-1 symbols) and 'bits'
        ([T,B,code*H*W/512] uint8, the real 1-bit/symbol stream)."""
        if config.PARAM['loss_fn'] not in ('bce', 'mse', 'mae'):
            raise ValueError('Not valid loss function')
        img, label = input['img'], input['label']
        params = self._index()[1]
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        mse = None
        use_graph = config.PARAM.get('cuda_graph', False) and config.PARAM.get('node_iters') is None
        if need_grad:
            loss, recon, codes, bits = _LoopFunction.apply(self, img, label, noise, *params)
        else:
            if use_graph:
                res, _ = self._run_graphed(img, label, noise)
            else:
                res = self._run_engine(img, label, noise)
            loss = CodecEngine.loss_from_acc(res['loss_acc'], img.numel())
            recon, codes, bits = res['recon'], res['codes'], res['bits']
            mse = (res['loss_acc'][:, 1] / float(img.numel())).float()   # per-iteration MSE, stays on the device
        if config.PARAM['loss_fn'] == 'bce' and bool(torch.isnan(loss.detach())):
            # F.binary_cross_entropy on the un-clamped cumulative reconstruction (dist_codec.py:26,58) raises as soon
            # as one element leaves [0, 1]; the kernels poison the loss sum instead and the error surfaces here
            self._free_hidden()
            raise RuntimeError('all elements of input should be between 0 and 1')
        T = recon.shape[0]
        if use_graph:
            # a graph's output buffers are rewritten by its next replay: the lazily built code list below reads copies
            # (device to device on the launching stream -- no host synchronisation, unlike materialising it here)
            codes, bits = codes.clone(), bits.clone()
        output = {'loss': loss, 'img': [recon[t] for t in range(T)], 'code_pm1': codes, 'bits': bits}
        if mse is not None:
            output['mse'] = mse                                     # metrics.psnr_per_iteration(output)
        if config.PARAM.get('pack_mode', 'reference') == 'bits':
            def make():
                b = bits.cpu().numpy()                              # one D2H per forward
                return [np.ascontiguousarray(b[:t + 1].transpose(1, 2, 0).reshape(-1, t + 1)) for t in range(T)]
        else:
            def make():
                c = codes.detach().cpu().numpy().astype(np.int8)    # one D2H per forward (reference: one per iteration)
                per_iter = [np.packbits(c[t]).reshape(-1, 1) for t in range(T)]
                out = []
                for t in range(T):
                    out.append(per_iter[t] if t == 0 else np.concatenate((out[t - 1], per_iter[t]), axis=1))
                return out
        output['code'] = CodeList(T, make)                          # materialised at the first access
        self._free_hidden()                                         # dist_codec.py:64
        return output
