"""GPU parity of the stand-alone cells (reference src/modules/cell.py) against the golden fixtures the
reference produced and against the CPU oracle.  These go through the C ABI (ctypes) via the modules."""
import os

import numpy as np
import pytest
import torch

from conftest import sd_digest
from oracle import drasic_oracle as O

pytestmark = pytest.mark.gpu

# stated tolerances (max-abs on h, c in [-1,1]) of the two precisions against the fp32 reference
TOL = {'parity': 2e-5, 'fast': 2e-3}


def make_cell(cin, co, precision, seed):
    import config
    import modules
    config.init()
    config.PARAM['precision'] = precision
    torch.manual_seed(seed)
    cell = modules.Cell({'cell': 'ConvLSTMCell', 'input_size': cin, 'output_size': co, 'kernel_size': 3,
                         'stride': 1, 'padding': 1, 'bias': False, 'num_layers': 1, 'activation': 'tanh'})
    return cell.cuda()


@pytest.mark.parametrize('precision', ['parity', 'fast'])
@pytest.mark.parametrize('tag,cin,co,hw', [('lstm_a', 128, 128, 4), ('lstm_b', 64, 64, 8)])
def test_convlstm_cell_golden(golden_dir, tag, cin, co, hw, precision):
    g = np.load(os.path.join(golden_dir, 'cells.npz'))
    cell = make_cell(cin, co, precision, 21)
    assert sd_digest(cell.state_dict()) == str(g[tag + '_digest'])
    gi = torch.Generator().manual_seed(22)
    for t in range(3):
        x = torch.rand(2, cin, hw, hw, generator=gi) * 2 - 1
        h = cell(x.cuda())
        assert h.shape == (2, co, hw, hw) and h.dtype == torch.float32
        np.testing.assert_allclose(h.cpu().numpy(), g[tag + '_h'][t], rtol=0, atol=TOL[precision])
        hid = cell.cell.hidden                      # [[h],[c]] like the reference attribute
        np.testing.assert_allclose(hid[1][0].cpu().numpy(), g[tag + '_c'][t], rtol=0, atol=2 * TOL[precision])
        np.testing.assert_allclose(hid[0][0].cpu().numpy(), h.cpu().numpy(), rtol=0, atol=0)
    cell.cell.free_hidden()
    assert cell.cell.hidden is None


@pytest.mark.parametrize('cin,co,hw,n', [(128, 128, 16, 3), (512, 128, 8, 5), (512, 128, 4, 9), (128, 512, 4, 16),
                                         (128, 512, 8, 2), (128, 128, 32, 1)])
def test_convlstm_cell_all_layer_shapes_vs_oracle(cin, co, hw, n):
    """the six ConvLSTM shapes of the codec plus a 32x32 map; ragged batch sizes"""
    cell = make_cell(cin, co, 'parity', 5)
    w_in, w_h = [w.detach().cpu() for w in cell.cell.weights()]
    gi = torch.Generator().manual_seed(6)
    state = None
    for t in range(3):
        x = torch.rand(n, cin, hw, hw, generator=gi) * 2 - 1
        h_ref, state = O.convlstm_cell(x, state, w_in, w_h)
        h = cell(x.cuda())
        assert (h.cpu() - h_ref).abs().max().item() < TOL['parity']
        assert (cell.cell.hidden[1][0].cpu() - state[1]).abs().max().item() < 2 * TOL['parity']


def test_convlstm_state_is_per_sample_and_resettable():
    cell = make_cell(64, 64, 'parity', 1)
    x = torch.rand(4, 64, 8, 8).cuda()
    a = cell(x).clone()
    cell.cell.free_hidden()
    b = cell(x[:2])                                  # batch composition must not matter
    assert torch.equal(a[:2], b)
    with pytest.raises(RuntimeError):                # reference: stale state of another shape crashes
        cell(x)
    cell.cell.free_hidden()
    assert torch.equal(cell(x), a)                   # deterministic


def test_convlstm_rejects_unsupported_configs():
    import modules
    bad = {'cell': 'ConvLSTMCell', 'input_size': 64, 'output_size': 64, 'kernel_size': 5, 'stride': 1,
           'padding': 2, 'bias': False, 'num_layers': 1, 'activation': 'tanh'}
    with pytest.raises(ValueError, match='Not valid'):
        modules.Cell(bad).cuda()(torch.rand(1, 64, 8, 8).cuda())
    cell = make_cell(96, 64, 'parity', 0)            # channels not a multiple of 64
    with pytest.raises(ValueError, match='Not valid'):
        cell(torch.rand(1, 96, 8, 8).cuda())


def test_conv_cell_golden(golden_dir):
    import config
    import modules
    config.init()
    g = np.load(os.path.join(golden_dir, 'cells.npz'))
    torch.manual_seed(31)
    cc = modules.Cell({'cell': 'ConvCell', 'input_size': 3, 'output_size': 32, 'kernel_size': 1, 'stride': 1,
                       'padding': 0, 'bias': True, 'activation': 'tanh', 'normalization': 'none'}).cuda()
    np.testing.assert_array_equal(cc.state_dict()['cell.cell.main.weight'].cpu().numpy(), g['convcell_w'])
    y = cc(torch.from_numpy(g['convcell_in']).cuda())
    np.testing.assert_allclose(y.cpu().numpy(), g['convcell_out'], rtol=0, atol=3e-7)


def test_quantization_cell_golden(golden_dir):
    import modules
    g = np.load(os.path.join(golden_dir, 'cells.npz'))
    q = modules.Cell({'cell': 'QuantizationCell'}).cuda()
    q.train(False)
    out = q(torch.from_numpy(g['quant_in']).cuda())
    assert np.array_equal(out.cpu().numpy(), g['quant_eval'], equal_nan=True)     # incl. -0.0 -> +1, passthrough
    q.train(True)
    out = q(torch.from_numpy(g['quant_train_in']).cuda(), torch.from_numpy(g['quant_train_u']).cuda())
    assert np.array_equal(out.cpu().numpy(), g['quant_train'])
    # internally drawn noise: +1 with probability (x+1)/2
    x = torch.full((200000,), 0.5).cuda()
    frac = (q(x) > 0).float().mean().item()
    assert abs(frac - 0.75) < 0.01
    # straight-through gradient
    x = torch.linspace(-0.9, 0.9, 7).cuda().requires_grad_(True)
    q.train(False)
    q(x).backward(torch.arange(7.0).cuda())
    assert torch.equal(x.grad.cpu(), torch.arange(7.0))


def test_pixel_shuffle_cells_golden(golden_dir):
    import modules
    g = np.load(os.path.join(golden_dir, 'cells.npz'))
    x = torch.from_numpy(g['shuffle_in']).cuda()
    down = modules.Cell({'cell': 'PixelShuffleCell', 'mode': 'down', 'scale_factor': 2})
    up = modules.Cell({'cell': 'PixelShuffleCell', 'mode': 'up', 'scale_factor': 2})
    assert np.array_equal(down(x).cpu().numpy(), g['unshuffle_out'])
    assert np.array_equal(up(x).cpu().numpy(), g['shuffle_out'])
