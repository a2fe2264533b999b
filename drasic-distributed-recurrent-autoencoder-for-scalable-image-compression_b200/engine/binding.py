"""ctypes binding of the C ABI in include/drasic_b200.h (libdrasic_b200.so, built by
__graft_entry__.build()).  There is no fallback: a missing library or a failing call raises."""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG, 'libdrasic_b200.so')

ORDER_IDENT, ORDER_QUAD = 0, 1
NEXT_NONE, NEXT_SAME, NEXT_DOWN, NEXT_UP = 0, 1, 2, 3
LOSS_KINDS = {'mae': 0, 'mse': 1, 'bce': 2}
TAIL_INIT, TAIL_FIRST, TAIL_STEP = 0, 1, 2
ACT_KINDS = {'none': 0, 'tanh': 1, 'relu': 2, 'sigmoid': 3, 'hardtanh': 4, 'elu': 5, 'selu': 6, 'celu': 7}

vp, ip, fp = C.c_void_p, C.c_int, C.c_void_p


class ConvLSTMArgs(C.Structure):
    _fields_ = [(n, vp) for n in ('x_hi', 'x_lo', 'h_hi', 'h_lo', 'w_hi', 'w_lo', 'w_inv_scale', 'c_state',
                                  'c_prev', 'gates_out', 'h_out_hi', 'h_out_lo', 'next_hi', 'next_lo',
                                  'group_mtile_end')] + \
               [(n, ip) for n in ('B_pad', 'H', 'W', 'Cin', 'Co', 'n_groups', 'has_state', 'next_mode',
                                  'next_f32', 'split', 'precise', 'block_n', 'gates_f32', 'cta_pair')] + \
               [('group_mtile_start', vp), ('group_unit_end', vp), ('group_img_end', vp), ('n_units', ip),
                ('pos_order', vp), ('pos_imgs', ip), ('stream_k', ip), ('sk_scratch', vp), ('sk_flags', vp), ('pos_pairs', ip)]


class BottleneckArgs(C.Structure):
    _fields_ = [(n, vp) for n in ('h_e3', 'w_e4', 'w_d0', 'noise', 'perm', 'code_out', 'z_out', 'bits_out',
                                  'd1_in_hi', 'd1_in_lo', 'd0_out_f32', 'group_img_end')] + \
               [(n, ip) for n in ('B_pad', 'HW', 'Ch', 'code', 'n_enc_groups', 'n_dec_groups', 'precise', 'split')] + \
               [('bits_in', vp)]


class TailHeadArgs(C.Structure):
    _fields_ = [(n, vp) for n in ('d3_next', 'w_d4', 'w_e0', 'b_e0', 'img', 'recon_prev', 'recon_out',
                                  'dec_out', 'loss_acc', 'e1_in_hi', 'e1_in_lo', 'perm', 'group_img_end')] + \
               [(n, ip) for n in ('B_pad', 'C', 'H', 'W', 'mode', 'loss_kind', 'n_enc_groups', 'n_dec_groups',
                                  'precise', 'split')] + [('group_iters', vp), ('iter', ip)]


class DgradArgs(C.Structure):
    _fields_ = [(n, vp) for n in ('dg_hi', 'dg_lo', 'w_hi', 'w_lo', 'w_inv_scale', 'dg_inv_scale', 'dx_out',
                                  'dhrec_out', 'group_mtile_end')] + \
               [(n, ip) for n in ('B_pad', 'H', 'W', 'Cin', 'Co', 'n_groups', 'dx_mode', 'split', 'cta_pair')] + \
               [('group_mtile_start', vp), ('group_unit_end', vp), ('group_img_end', vp), ('n_units', ip),
                ('pos_order', vp), ('pos_imgs', ip), ('stream_k', ip), ('sk_scratch', vp), ('sk_flags', vp), ('pos_pairs', ip)]


class WgradArgs(C.Structure):
    _fields_ = [(n, vp) for n in ('dg_hi', 'dg_lo', 'x_hi', 'x_lo', 'h_hi', 'h_lo', 'gw', 'dg_inv_scale',
                                  'group_kb_end')] + \
               [(n, ip) for n in ('B_pad', 'H', 'W', 'Cin', 'Co', 'n_groups', 'has_h', 'split', 'k_splits', 'cta_pair')]


class LstmBwdFusedArgs(C.Structure):
    _fields_ = [(n, vp) for n in ('gates', 'c_t', 'c_prev', 'dh_above', 'dh_rec', 'dc_in', 'dc_out', 'dg_hi', 'dg_lo',
                                  'prev_amax_bits', 'amax_bits', 'inv_scale_out')] + [('n_pix', C.c_size_t)] + \
               [(n, ip) for n in ('Co', 'gates_f32', 'verify')]


class BottleneckBwdArgs(C.Structure):
    _fields_ = [(n, vp) for n in ('dd0', 'h_e3', 'z', 'codes', 'w_e4', 'w_d0', 'perm', 'group_img_end', 'dh_e3',
                                  'dw_e4', 'dw_d0')] + \
               [(n, ip) for n in ('B_pad', 'HW', 'Ch', 'code', 'n_enc_groups', 'n_dec_groups')] + [('fix_scale', C.c_double)]


class TailBwdArgs(C.Structure):
    _fields_ = [(n, vp) for n in ('d3_next', 'w_d4', 'img', 'recon_t', 'gacc', 'dh_d3', 'dw_d4', 'perm',
                                  'group_img_end')] + [('loss_scale', C.c_float)] + \
               [(n, ip) for n in ('B_pad', 'C', 'H', 'W', 'first', 'loss_kind', 'n_dec_groups', 'has_acc')] + \
               [('group_iters', vp), ('iter', ip), ('n_enc_groups', ip), ('fix_scale', C.c_double)]


class HeadBwdArgs(C.Structure):
    _fields_ = [(n, vp) for n in ('du1', 'w_e0', 'b_e0', 'img', 'recon_prev', 'gacc', 'dw_e0', 'db_e0', 'perm',
                                  'group_img_end')] + \
               [(n, ip) for n in ('B_pad', 'C', 'H', 'W', 'first', 'propagate', 'n_enc_groups')] + [('fix_scale', C.c_double)]


class AdamTensor(C.Structure):
    _fields_ = [('param', vp), ('grad', vp), ('exp_avg', vp), ('exp_avg_sq', vp), ('numel', C.c_longlong)]


DX_SAME, DX_INV_DOWN, DX_INV_UP = 0, 1, 2

# every symbol include/drasic_b200.h declares (tests check the library exports all of them)
SYMBOLS = ('drasic_version', 'drasic_last_error', 'drasic_device_sms', 'drasic_packed_weight_bytes', 'drasic_packed_dgrad_weight_bytes',
           'drasic_stream_k_scratch_bytes', 'drasic_stream_k_flag_bytes', 'drasic_set_pdl',
           'drasic_pack_lstm_weights', 'drasic_pack_lstm_weights_batch', 'drasic_pack_dgrad_weights_batch', 'drasic_convlstm_fwd', 'drasic_bottleneck_fwd', 'drasic_tail_head_fwd',
           'drasic_conv1x1_fwd', 'drasic_quantize_fwd', 'drasic_lstm_bwd_fused',
           'drasic_pack_dgrad_weights', 'drasic_convlstm_bwd_data', 'drasic_convlstm_bwd_weight',
           'drasic_unpack_wgrad', 'drasic_bottleneck_bwd', 'drasic_tail_bwd', 'drasic_head_bwd', 'drasic_adam_step')

_lib = None


class DrasicError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DrasicError('{} not found: build the CUDA extension first (python -c "import __graft_entry__ as g; '
                          'g.build()"); there is no CPU fallback'.format(LIB_PATH))
    L = C.CDLL(LIB_PATH)
    L.drasic_version.restype = C.c_int
    L.drasic_last_error.restype = C.c_char_p
    L.drasic_device_sms.restype = C.c_int
    L.drasic_packed_weight_bytes.restype = C.c_size_t
    L.drasic_packed_weight_bytes.argtypes = [C.c_int, C.c_int]
    L.drasic_packed_dgrad_weight_bytes.restype = C.c_size_t
    L.drasic_packed_dgrad_weight_bytes.argtypes = [C.c_int, C.c_int]
    L.drasic_set_pdl.restype = C.c_int
    L.drasic_set_pdl.argtypes = [C.c_int]
    L.drasic_stream_k_scratch_bytes.restype = C.c_size_t
    L.drasic_stream_k_scratch_bytes.argtypes = []
    L.drasic_stream_k_flag_bytes.restype = C.c_size_t
    L.drasic_stream_k_flag_bytes.argtypes = []
    L.drasic_pack_lstm_weights.restype = C.c_int
    L.drasic_pack_lstm_weights.argtypes = [vp, vp, ip, ip, ip, ip, ip, vp, vp, vp, vp, vp]
    L.drasic_pack_lstm_weights_batch.restype = C.c_int
    L.drasic_pack_lstm_weights_batch.argtypes = [vp, vp, vp, ip, ip, ip, ip, ip, ip, ip, vp, vp, vp, vp, vp]
    L.drasic_pack_dgrad_weights_batch.restype = C.c_int
    L.drasic_pack_dgrad_weights_batch.argtypes = [vp, vp, vp, ip, ip, ip, ip, ip, ip, vp, vp, vp, vp, vp]
    L.drasic_convlstm_fwd.restype = C.c_int
    L.drasic_convlstm_fwd.argtypes = [C.POINTER(ConvLSTMArgs), vp]
    L.drasic_bottleneck_fwd.restype = C.c_int
    L.drasic_bottleneck_fwd.argtypes = [C.POINTER(BottleneckArgs), vp]
    L.drasic_tail_head_fwd.restype = C.c_int
    L.drasic_tail_head_fwd.argtypes = [C.POINTER(TailHeadArgs), vp]
    L.drasic_conv1x1_fwd.restype = C.c_int
    L.drasic_conv1x1_fwd.argtypes = [vp, vp, vp, vp, ip, ip, ip, ip, ip, vp]
    L.drasic_quantize_fwd.restype = C.c_int
    L.drasic_quantize_fwd.argtypes = [vp, vp, vp, C.c_size_t, vp]
    L.drasic_lstm_bwd_fused.restype = C.c_int
    L.drasic_lstm_bwd_fused.argtypes = [C.POINTER(LstmBwdFusedArgs), vp]
    L.drasic_pack_dgrad_weights.restype = C.c_int
    L.drasic_pack_dgrad_weights.argtypes = [vp, vp, ip, ip, ip, ip, vp, vp, vp, vp, vp]
    L.drasic_convlstm_bwd_data.restype = C.c_int
    L.drasic_convlstm_bwd_data.argtypes = [C.POINTER(DgradArgs), vp]
    L.drasic_convlstm_bwd_weight.restype = C.c_int
    L.drasic_convlstm_bwd_weight.argtypes = [C.POINTER(WgradArgs), vp]
    L.drasic_unpack_wgrad.restype = C.c_int
    L.drasic_unpack_wgrad.argtypes = [vp, ip, ip, ip, ip, vp, vp, vp]
    L.drasic_bottleneck_bwd.restype = C.c_int
    L.drasic_bottleneck_bwd.argtypes = [C.POINTER(BottleneckBwdArgs), vp]
    L.drasic_tail_bwd.restype = C.c_int
    L.drasic_tail_bwd.argtypes = [C.POINTER(TailBwdArgs), vp]
    L.drasic_head_bwd.restype = C.c_int
    L.drasic_head_bwd.argtypes = [C.POINTER(HeadBwdArgs), vp]
    L.drasic_adam_step.restype = C.c_int
    L.drasic_adam_step.argtypes = [C.POINTER(AdamTensor), ip, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                   C.c_double, C.c_double, vp]
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        msg = lib().drasic_last_error().decode(errors='replace')
        if rc == -1:
            raise ValueError('Not valid {}: {}'.format(what, msg))   # mirrors the reference's ValueError
        raise DrasicError('{} failed (code {}): {}'.format(what, rc, msg))


def ptr(t):
    """device pointer of a torch tensor (or None)"""
    return None if t is None else t.data_ptr()


def pointer_arrays(groups):
    """host arrays for the *_batch packers: groups = [(slot, w_in tensor, w_h tensor), ...]"""
    n = len(groups)
    a_in = (C.c_void_p * n)(*[g[1].data_ptr() for g in groups])
    a_h = (C.c_void_p * n)(*[g[2].data_ptr() for g in groups])
    slots = (C.c_int * n)(*[int(g[0]) for g in groups])
    return a_in, a_h, slots, n
