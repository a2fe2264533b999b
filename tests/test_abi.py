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
    assert ctypes.sizeof(B.ConvLSTMArgs) == 15 * 8 + 14 * 4 + 3 * 8 + 4 + 4 + 8 + 4 + 4 + 2 * 8 + 8   # (pos_pairs, pad) (n_units, pad) (pos_order) (pos_imgs, stream_k) (sk_*)
    assert ctypes.sizeof(B.DgradArgs) == 9 * 8 + 9 * 4 + 4 + 3 * 8 + 4 + 4 + 8 + 4 + 4 + 2 * 8 + 8
    assert ctypes.sizeof(B.WgradArgs) == 9 * 8 + 10 * 4
    assert ctypes.sizeof(B.BottleneckBwdArgs) == 11 * 8 + 6 * 4 + 8                      # + fix_scale (double)
    assert ctypes.sizeof(B.TailBwdArgs) == 9 * 8 + 4 + 8 * 4 + 4 + 8 + 2 * 4 + 8
    assert ctypes.sizeof(B.HeadBwdArgs) == 10 * 8 + 7 * 4 + 4 + 8
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


def test_header_is_plain_c_and_links(tmp_path):
    """include/drasic_b200.h compiles as C (no C++ or torch types in the signatures) and a C program links the
    library through it; bad arguments come back as DRASIC_ERR_INVALID with a message, without touching a device"""
    import shutil
    import subprocess
    if shutil.which('gcc') is None:
        pytest.skip('no gcc')
    src = tmp_path / 'use_abi.c'
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "drasic_b200.h"
int main(void) {
  drasic_convlstm_args a;
  drasic_adam_tensor t;
  memset(&a, 0, sizeof a);
  memset(&t, 0, sizeof t);
  t.numel = 4;
  if (drasic_version() != 1) return 1;
  if (drasic_convlstm_fwd(&a, NULL) != -1) return 2;
  if (strlen(drasic_last_error()) == 0) return 3;
  if (drasic_adam_step(&t, 1, 1e-3, 0.9, 0.999, 1e-8, 0.0, 0.1, 0.001, NULL) != -1) return 4;
  if (drasic_adam_step(&t, 0, 1e-3, 0.9, 0.999, 1e-8, 0.0, 0.1, 0.001, NULL) != 0) return 5;
  printf("%d %d\n", (int)sizeof a, (int)sizeof t);
  return 0;
}
''')
    exe = tmp_path / 'use_abi'
    libdir = os.path.dirname(B.LIB_PATH)
    subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe),
                    '-L', libdir, '-ldrasic_b200', '-Wl,-rpath,' + libdir], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert int(out[0]) == ctypes.sizeof(B.ConvLSTMArgs) and int(out[1]) == ctypes.sizeof(B.AdamTensor)


def header_structs():
    """{typedef name: [(kind, field name)]} parsed from the header; kind is 'ptr', 'int', 'float', 'size_t' or 'll'"""
    text = open(os.path.join(ROOT, 'include', 'drasic_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    out = {}
    for body, name in re.findall(r'typedef\s+struct\s*\{(.*?)\}\s*(drasic_[a-z0-9_]+)\s*;', text, flags=re.S):
        fields = []
        for decl in body.split(';'):
            decl = ' '.join(decl.split())
            if not decl:
                continue
            m = re.match(r'^(const\s+)?(void|float|double|int|uint8_t|size_t|long long|unsigned int)\s*(.*)$', decl)
            assert m, decl
            base = m.group(2)
            for item in m.group(3).split(','):
                item = item.strip()
                ptr = item.startswith('*')
                fname = item.lstrip('* ').strip()
                kind = 'ptr' if ptr else {'int': 'int', 'float': 'float', 'double': 'double', 'size_t': 'size_t', 'long long': 'll'}[base]
                fields.append((kind, fname))
        out[name] = fields
    return out


def test_binding_structs_follow_the_header_field_for_field():
    structs = header_structs()
    pairs = {'drasic_convlstm_args': B.ConvLSTMArgs, 'drasic_bottleneck_args': B.BottleneckArgs,
             'drasic_tail_head_args': B.TailHeadArgs, 'drasic_lstm_bwd_fused_args': B.LstmBwdFusedArgs,
             'drasic_dgrad_args': B.DgradArgs, 'drasic_wgrad_args': B.WgradArgs,
             'drasic_bottleneck_bwd_args': B.BottleneckBwdArgs, 'drasic_tail_bwd_args': B.TailBwdArgs,
             'drasic_head_bwd_args': B.HeadBwdArgs, 'drasic_adam_tensor': B.AdamTensor}
    assert sorted(structs) == sorted(pairs)
    ctype_of = {'ptr': ctypes.c_void_p, 'int': ctypes.c_int, 'float': ctypes.c_float, 'double': ctypes.c_double,
                'size_t': ctypes.c_size_t, 'll': ctypes.c_longlong}
    for name, cls in pairs.items():
        want = [(f, ctype_of[k]) for k, f in structs[name]]
        got = [(f, t) for f, t in cls._fields_]
        assert got == want, name


def test_binding_prototypes_follow_the_header():
    """argument count and kind of every ctypes prototype against the header's declaration"""
    text = open(os.path.join(ROOT, 'include', 'drasic_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    protos = re.findall(r'\b(int|size_t|const char\*)\s+(drasic_[a-z0-9_]+)\s*\(([^)]*)\)\s*;', text)
    assert len(protos) == len(B.SYMBOLS)
    L = B.lib()
    scalar = {'int': ctypes.c_int, 'size_t': ctypes.c_size_t, 'float': ctypes.c_float, 'double': ctypes.c_double}
    restype = {'int': ctypes.c_int, 'size_t': ctypes.c_size_t, 'const char*': ctypes.c_char_p}
    for ret, name, args in protos:
        fn = getattr(L, name)
        assert fn.restype is restype[ret], name
        params = [] if args.strip() in ('', 'void') else [' '.join(a.split()) for a in args.split(',')]
        got = list(fn.argtypes or [])
        assert len(got) == len(params), name
        for p, t in zip(params, got):
            if '*' in p:
                m = re.search(r'(drasic_[a-z0-9_]+)\s*\*', p)
                if m:       # pointer to an argument struct
                    assert issubclass(t, ctypes._Pointer) and ctypes.sizeof(t._type_) > 0, (name, p)
                else:
                    assert t is ctypes.c_void_p, (name, p)
            else:
                base = re.match(r'^(?:const\s+)?(int|size_t|float|double)\b', p).group(1)
                assert t is scalar[base], (name, p)


def test_library_is_sm100a_tensor_core_and_tma_code():
    """the shipped library really is tcgen05 / TMA code for sm_100a: its SASS holds UTCHMMA (tcgen05.mma), LDTM
    (tcgen05.ld), UTMALDG (cp.async.bulk.tensor), UBLKCP (cp.async.bulk), SYNCS (mbarriers), the two-CTA variants and
    cluster barriers -- and no legacy HMMA (mma.sync) anywhere"""
    import shutil
    import subprocess
    tool = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(tool):
        pytest.skip('no cuobjdump')
    sass = subprocess.run([tool, '-sass', B.LIB_PATH], check=True, capture_output=True, text=True).stdout
    assert 'arch = sm_100a' in sass and 'arch = sm_90' not in sass
    for mnemonic in ('UTCHMMA', 'LDTM', 'UTMALDG', 'UBLKCP', 'SYNCS', 'UCGABAR_ARV', 'UTCBAR'):
        assert mnemonic in sass, mnemonic
    assert '.2CTA' in sass                                   # cta_group::2 MMAs / TMA loads of the CTA-pair kernels
    assert not re.search(r'\bHMMA\b', sass)
