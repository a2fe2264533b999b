"""Cell zoo with the reference's class names, constructor dicts and state_dict layout
(reference src/modules/cell.py), computing on the B200 kernels.

Parameters live in nn.Conv2d containers under the reference's attribute names, so state_dict keys,
shapes, dtypes and the default initialisation order are identical and checkpoints move both ways;
the nn.Conv2d forward is never called.  When a cell is called on its own (outside the fused loop of
models.Model.forward) it takes and returns NCHW fp32 tensors like the reference and runs
    ConvLSTMCell     -> tcgen05 gate kernel   (engine.cell.CellRunner, drasic_convlstm_fwd)
    ConvCell (1x1)   -> drasic_conv1x1_fwd
    QuantizationCell -> drasic_quantize_fwd
    PixelShuffleCell -> an index permutation (torch views; no arithmetic)
Configurations the kernels do not cover raise ValueError -- there is no PyTorch/CPU fallback.
"""
import copy

import torch
import torch.nn as nn

import config
from engine import binding as B
from engine.cell import CellRunner
from modules.quantization import Quantization
from modules.shuffle import PixelShuffle, PixelUnShuffle

__all__ = ['Normalization', 'Activation', 'ConvCell', 'ConvLSTMCell', 'PixelShuffleCell', 'QuantizationCell',
           'Cell']

_ACTIVATIONS = ('none', 'tanh', 'hardtanh', 'relu', 'elu', 'selu', 'celu', 'sigmoid')


class _Tag(nn.Module):
    """Parameter-free marker standing where the reference puts an nn activation / normalization
    module; the math itself is fused into the kernels."""

    def __init__(self, mode):
        super(_Tag, self).__init__()
        self.mode = mode

    def forward(self, input):
        raise B.DrasicError('activation/normalization are fused into the B200 kernels and cannot be called alone')

    def extra_repr(self):
        return 'mode={}'.format(self.mode)


def Normalization(cell_info):
    """cell.py:9-20.  Only 'none' is usable even in the reference (its bn/in/ln pick 1-D norms that
    reject 4-D input); anything else is rejected here."""
    if cell_info['mode'] == 'none':
        return nn.Sequential()
    if cell_info['mode'] in ('bn', 'in', 'ln'):
        raise ValueError('Not valid normalization for the B200 path (only none)')
    raise ValueError('Not valid normalization')


def Activation(cell_info):
    """cell.py:23-46."""
    mode = cell_info['mode']
    if mode == 'none':
        return nn.Sequential()
    if mode in _ACTIVATIONS:
        return _Tag(mode)
    if mode in ('prelu', 'softmax'):
        raise ValueError('Not valid activation for the B200 path')
    raise ValueError('Not valid activation')


def _mode_of(m):
    return m.mode if isinstance(m, _Tag) else 'none'


class ConvCell(nn.Module):
    """conv -> activation -> normalization (cell.py:49-72; the reference stores the Activation under
    the key 'normalization' and vice versa, which is kept so module trees print alike)."""

    def __init__(self, cell_info):
        super(ConvCell, self).__init__()
        self.cell_info = cell_info
        self.cell = self.make_cell()

    def make_cell(self):
        info = copy.deepcopy(self.cell_info)
        cell = nn.ModuleDict({})
        cell['main'] = nn.Conv2d(info['input_size'], info['output_size'], kernel_size=info['kernel_size'],
                                 stride=info['stride'], padding=info['padding'], bias=info['bias'])
        cell['activation'] = Cell({'cell': 'Normalization', 'input_size': info['output_size'],
                                   'mode': info['normalization']})
        cell['normalization'] = Cell({'cell': 'Activation', 'mode': info['activation']})
        return cell

    @property
    def activation_mode(self):
        return _mode_of(self.cell['normalization'].cell)

    def forward(self, input):
        conv = self.cell['main']
        if conv.kernel_size != (1, 1) or conv.stride != (1, 1) or conv.padding != (0, 0):
            raise ValueError('Not valid ConvCell for the stand-alone B200 path (kernel 1, stride 1, padding 0)')
        if not input.is_cuda:
            raise B.DrasicError('DRASIC B200 cells need CUDA tensors; there is no CPU path')
        x = input.detach().contiguous().float()
        n, cin, h, w = x.shape
        cout = conv.out_channels
        y = torch.empty((n, cout, h, w), dtype=torch.float32, device=x.device)
        wt = conv.weight.detach().contiguous().float()
        bs = conv.bias.detach().contiguous().float() if conv.bias is not None else None
        stream = torch.cuda.current_stream(x.device).cuda_stream
        B.check(B.lib().drasic_conv1x1_fwd(x.data_ptr(), wt.data_ptr(), B.ptr(bs), y.data_ptr(), n, cin, cout,
                                           h * w, B.ACT_KINDS[self.activation_mode], stream), 'conv1x1_fwd')
        return y


class ConvLSTMCell(nn.Module):
    """Stateful convolutional LSTM (cell.py:75-144): gates = W_in*x + W_h*h (3x3, no bias), order
    i,f,g,o; c = f*c + i*g; h = o*tanh(c).  State persists across calls until free_hidden()."""

    def __init__(self, cell_info):
        super(ConvLSTMCell, self).__init__()
        self.cell_info = cell_info
        self.cell = self.make_cell()
        self._runner = None

    def make_cell(self):
        info = copy.deepcopy(self.cell_info)
        cell = nn.ModuleList([nn.ModuleDict({}) for _ in range(info['num_layers'])])
        for i in range(info['num_layers']):
            common = {'cell': 'ConvCell', 'output_size': 4 * info['output_size'], 'kernel_size': info['kernel_size'],
                      'stride': info['stride'], 'padding': info['padding'], 'bias': info['bias'],
                      'normalization': 'none', 'activation': 'none'}
            cell[i]['in'] = Cell(dict(common, input_size=info['input_size']))
            cell[i]['hidden'] = Cell(dict(common, input_size=info['output_size']))
            act = {'cell': 'Activation', 'mode': info['activation']}
            cell[i]['activation'] = nn.ModuleList([Cell(act), Cell(act)])
        return cell

    def _check_supported(self):
        info = self.cell_info
        if info['num_layers'] != 1 or info['kernel_size'] != 3 or info['stride'] != 1 or info['padding'] != 1 \
                or info['bias'] or info['activation'] != 'tanh':
            raise ValueError('Not valid ConvLSTMCell for the B200 path (1 layer, 3x3, stride 1, padding 1, '
                             'no bias, tanh)')

    def weights(self):
        return self.cell[0]['in'].cell.cell['main'].weight, self.cell[0]['hidden'].cell.cell['main'].weight

    @property
    def hidden(self):
        """[[h],[c]] NCHW fp32 like the reference attribute (cell.py:142), or None"""
        return None if self._runner is None else self._runner.hidden_nchw()

    def init_hidden(self, hidden_size):
        z = torch.zeros(hidden_size, device=config.PARAM['device'])
        return [[z], [z.clone()]]

    def free_hidden(self):
        if self._runner is not None:
            self._runner.reset()

    def forward(self, input, hidden=None):
        self._check_supported()
        if input.dim() != 4:
            raise ValueError('Not valid input for the B200 ConvLSTMCell (expects [N,C,H,W])')
        if self._runner is None:
            self._runner = CellRunner(config.PARAM.get('precision', 'parity'))
        w_in, w_h = self.weights()
        return self._runner.step(input, w_in, w_h)


class PixelShuffleCell(nn.Module):
    """cell.py:147-165."""

    def __init__(self, cell_info):
        super(PixelShuffleCell, self).__init__()
        self.cell_info = cell_info
        self.cell = self.make_cell()

    def make_cell(self):
        if self.cell_info['mode'] == 'down':
            return PixelUnShuffle(self.cell_info['scale_factor'])
        if self.cell_info['mode'] == 'up':
            return PixelShuffle(self.cell_info['scale_factor'])
        raise ValueError('Not valid model mode')

    def forward(self, input):
        return self.cell(input)


class QuantizationCell(nn.Module):
    """cell.py:168-182."""

    def __init__(self, cell_info):
        super(QuantizationCell, self).__init__()
        self.cell_info = cell_info
        self.cell = self.make_cell()

    def make_cell(self):
        return Quantization()

    def forward(self, input, noise=None):
        return self.cell(input, noise)


_FACTORY = {'Normalization': Normalization, 'Activation': Activation, 'ConvCell': ConvCell,
            'ConvLSTMCell': ConvLSTMCell, 'PixelShuffleCell': PixelShuffleCell,
            'QuantizationCell': QuantizationCell}


class Cell(nn.Module):
    """dict -> module dispatch (cell.py:185-212)."""

    def __init__(self, cell_info):
        super(Cell, self).__init__()
        self.cell_info = cell_info
        self.cell = self.make_cell()

    def make_cell(self):
        kind = self.cell_info['cell']
        if kind == 'none':
            return nn.Sequential()
        if kind not in _FACTORY:
            raise ValueError('Not valid {} model'.format(kind))
        return _FACTORY[kind](self.cell_info)

    def forward(self, *input):
        return self.cell(*input)
