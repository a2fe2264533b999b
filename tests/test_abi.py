"""The C-ABI library loads without a GPU, exports every symbol include/drasic_b200.h declares, and
rejects bad arguments with the documented status before touching a device.  CPU only."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from engine import binding as B


def header_symbols():
    text = open(os.path.join(ROOT, 'include', 'drasic_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(drasic_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    syms = header_symbols()
    assert len(syms) >= 10
    lib = ctypes.CDLL(B.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(B.SYMBOLS) == syms          # the binding covers the whole header


def test_version_and_sizes():
    L = B.lib()
    assert L.drasic_version() == 1
    assert L.drasic_packed_weight_bytes(128, 512) == 4 * 512 * 9 * (128 + 512) * 2


def test_struct_layouts_match_header():
    # pointer and int counts of the argument structs in include/drasic_b200.h
    assert ctypes.sizeof(B.ConvLSTMArgs) == 15 * 8 + 14 * 4 + 3 * 8 + 4 + 4
    assert ctypes.sizeof(B.DgradArgs) == 9 * 8 + 9 * 4 + 4 + 3 * 8 + 4 + 4
    assert ctypes.sizeof(B.WgradArgs) == 9 * 8 + 10 * 4
    assert ctypes.sizeof(B.BottleneckBwdArgs) == 11 * 8 + 6 * 4
    assert ctypes.sizeof(B.TailBwdArgs) == 9 * 8 + 4 + 8 * 4 + 4 + 8 + 2 * 4
    assert ctypes.sizeof(B.HeadBwdArgs) == 10 * 8 + 7 * 4 + 4
    assert ctypes.sizeof(B.BottleneckArgs) == 12 * 8 + 8 * 4 + 8
    assert ctypes.sizeof(B.AdamTensor) == 4 * 8 + 8
    assert ctypes.sizeof(B.TailHeadArgs) == 13 * 8 + 10 * 4 + 8 + 4 + 4   # trailing pad to 8


def test_invalid_arguments_map_to_value_error():
    L = B.lib()
    a = B.ConvLSTMArgs()
    a.Cin, a.Co, a.H, a.W, a.B_pad, a.block_n, a.n_groups = 100, 128, 4, 4, 8, 256, 1   # Cin not multiple of 64
    rc = L.drasic_convlstm_fwd(a, None)
    assert rc == -1
    assert b'multiples of 64' in L.drasic_last_error()
    with pytest.raises(ValueError, match='Not valid'):
        B.check(rc, 'convlstm_fwd')
    a.Cin = 128
    a.B_pad = 7                                                                          # 4x4 needs 8 images per tile
    assert L.drasic_convlstm_fwd(a, None) == -1
    t = B.TailHeadArgs()
    assert L.drasic_tail_head_fwd(t, None) == -1
    b = B.BottleneckArgs()
    assert L.drasic_bottleneck_fwd(b, None) == -1
    assert L.drasic_pack_lstm_weights(None, None, 128, 128, 0, 0, 256, None, None, None, None, None) == -1
    assert L.drasic_quantize_fwd(None, None, None, 4, None) == -1
    assert L.drasic_conv1x1_fwd(None, None, None, None, 1, 1, 1, 1, 1, None) == -1
