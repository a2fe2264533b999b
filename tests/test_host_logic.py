"""Host-side logic of the drop-in packages: batch plan, module tree / state_dict contract, error
behaviour, layout helpers.  CPU only (no kernel launches)."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, build_model, sd_digest
from engine import binding as B
from engine.cell import from_nhwc, to_planes
from engine.codec import BatchPlan, PAD


def test_batch_plan_sorts_pads_and_keeps_order():
    label = np.array([2, 0, 2, 2, 0, 3, 2])
    p = BatchPlan(label, 4, 'cpu')
    assert PAD == 4 and p.B == 7 and p.B_pad == 16             # 4 + 0 + 4 + 4 slots, total rounded up to 8
    assert list(p.group_img_end_host) == [4, 4, 8, 16]          # node 1 is empty: no slots; the last node owns the alignment slots
    assert list(p.perm_host[:2]) == [1, 4] and list(p.perm_host[4:8]) == [0, 2, 3, 6] and p.perm_host[8] == 5
    assert (p.perm_host[[2, 3, 9, 11, 12, 15]] == -1).all()
    valid = p.perm_host[p.perm_host >= 0]
    assert sorted(valid.tolist()) == list(range(7))
    # 4x4 layer (8 images per 128-pixel tile): nodes 0 and 2 share tile 0, node 3 walks tile 1
    u, total = p.m_units(16, False)
    assert u.numpy().tolist() == [[0, 0, 0, 1], [1, 0, 1, 2], [1, 1, 2, 3]] and total == 3
    u, total = p.m_units(16, True)
    assert u.numpy()[2].tolist() == [1, 1, 2, 3] and total == 3   # every node: one half-filled pair
    # 16x16 layer: two tiles per image, ranges are disjoint
    u, total = p.m_units(256, False)
    assert u.numpy().tolist() == [[0, 8, 8, 16], [8, 8, 16, 32], [8, 8, 16, 32]] and total == 32
    assert p.m_units(256, True)[1] == 16
    assert list(p.kb_end(16).numpy()) == [1, 1, 2, 4]           # wgrad K blocks of 64 pixels = 4 images at 4x4


def test_inference_plan_packs_nodes_back_to_back():
    """pad=1 (no backward): no per-node padding; tiles shared by neighbouring nodes are walked by both"""
    label = np.array([2, 0, 2, 2, 0, 3, 2])
    p = BatchPlan(label, 4, 'cpu', pad=1)
    assert p.B_pad == 8 and list(p.group_img_end_host) == [2, 2, 6, 8]
    assert list(p.perm_host) == [1, 4, 0, 2, 3, 6, 5, -1]
    u, total = p.m_units(64, False)                              # 8x8 layer: 2 images per tile, all boundaries aligned
    assert u.numpy().tolist() == [[0, 1, 1, 3], [1, 1, 3, 4], [1, 1, 3, 4]] and total == 4
    u, total = p.m_units(16, False)                              # 4x4 layer: one tile holds all 8 slots, three nodes walk it
    assert u.numpy().tolist() == [[0, 0, 0, 0], [1, 0, 1, 1], [1, 1, 2, 3]] and total == 3
    q = BatchPlan(np.array([0, 1, 1]), 2, 'cpu', pad=1)          # odd boundary inside an 8x8 tile
    u, total = q.m_units(64, False)
    assert u.numpy().tolist() == [[0, 0], [1, 4], [1, 5]] and total == 5


def test_static_plan_update_rewrites_every_table_in_place():
    """what a captured graph reads: the tables of a static plan are views of ONE buffer; update() refreshes all of them
    (one copy) and they then equal the tables of a fresh plan of the same padded size"""
    from engine.codec import worst_case_pad
    rng = np.random.RandomState(0)
    for pad in (1, 4):
        Bn, M, hws = 48, 4, (16, 64, 256)
        total = worst_case_pad(Bn, M, pad, 8)
        p = BatchPlan(np.arange(Bn) % M, M, 'cpu', pad_total=total, static_hw=hws, pad=pad)
        views = [p.perm, p.group_img_end] + [p.kb_end(hw) for hw in hws] + [p.m_units(hw, pr)[0] for hw in hws for pr in (False, True)]
        ptrs = [v.data_ptr() for v in views]
        for _ in range(3):
            label = rng.randint(0, M, Bn)
            label[label == 1] = 0                                    # an empty node
            p.update(label)
            q = BatchPlan(label, M, 'cpu', pad_total=total, pad=pad)
            assert p.perm.numpy().tolist() == q.perm_host.tolist() == p.perm_host.tolist()
            assert p.group_img_end.numpy().tolist() == q.group_img_end_host.tolist()
            for hw in hws:
                assert p.kb_end(hw).numpy().tolist() == q.kb_end(hw).numpy().tolist()
                for pr in (False, True):
                    assert p.m_units(hw, pr)[0].numpy().tolist() == q.m_units(hw, pr)[0].numpy().tolist()
                    assert p.m_units(hw, pr)[1] >= q.m_units(hw, pr)[1]      # the static launch bound covers any labelling
            assert [v.data_ptr() for v in [p.perm, p.group_img_end] + [p.kb_end(hw) for hw in hws]
                    + [p.m_units(hw, pr)[0] for hw in hws for pr in (False, True)]] == ptrs
        with pytest.raises(ValueError):
            p.update(np.zeros(Bn + 1, dtype=np.int64))
        with pytest.raises(RuntimeError):
            p.kb_end(4)                                              # a frozen static plan has no buffer for a new size


def test_batch_plan_rejects_bad_labels_like_the_reference():
    with pytest.raises(IndexError):
        BatchPlan(np.array([0, 4]), 4, 'cpu')


def test_batch_plan_large_ragged():
    rng = np.random.RandomState(2)
    label = rng.randint(0, 8, size=1000)
    p = BatchPlan(label, 8, 'cpu')
    assert p.B_pad % 8 == 0 and p.B_pad >= 1000 and p.B_pad <= 1000 + 3 * 8 + 7
    assert (p.group_img_end_host % PAD == 0).all() and p.group_img_end_host[-1] == p.B_pad
    for j in range(8):
        lo = 0 if j == 0 else p.group_img_end_host[j - 1]
        idx = p.perm_host[lo:p.group_img_end_host[j]]
        idx = idx[idx >= 0]
        assert (label[idx] == j).all() and (np.diff(idx) > 0).all()


@pytest.mark.parametrize('name,data,M,seed,golden', [('dist_codec', 'MNIST', 1, 0, 'loop_mnist_m1.npz'),
                                                     ('dist_codec', 'CIFAR10', 2, 3, 'loop_cifar_m2.npz'),
                                                     ('sep_codec', 'CIFAR10', 2, 6, 'loop_sep_m2.npz')])
def test_state_dict_contract_matches_reference(golden_dir, name, data, M, seed, golden):
    g = np.load(os.path.join(golden_dir, golden))
    m = build_model(name, data, M, 4, seed=seed, device='cpu')
    sd = m.state_dict()
    if 'keys' in g:
        assert list(sd.keys()) == [str(k) for k in g['keys']]
    assert sd_digest(sd) == str(g['digest'])               # same keys, shapes, dtypes and seeded values
    assert all(v.dtype == torch.float32 for v in sd.values())
    # load_state_dict round trip (test_model.py:46)
    m2 = build_model(name, data, M, 4, seed=seed + 1, device='cpu')
    m2.load_state_dict(sd)
    assert sd_digest(m2.state_dict()) == str(g['digest'])


def test_module_tree_and_reference_api_surface():
    import config
    import modules
    m = build_model('dist_codec', 'CIFAR10', 3, 16, device='cpu')
    assert set(config.PARAM['model'].keys()) == {'encoder', 'quantizer', 'decoder'}   # factory writes PARAM['model']
    assert len(m.model['encoder']) == 3 and len(m.model['encoder'][0]) == 8 and len(m.model['decoder']) == 8
    lstm = m.model['encoder'][1][2].cell
    assert isinstance(lstm, modules.ConvLSTMCell) and lstm.hidden is None
    lstm.free_hidden()
    for attr in ('pack', 'unpack', 'loss_fn', 'forward'):
        assert hasattr(m, attr)
    code = torch.tensor([[1.0, -1.0] * 8])
    packed = m.pack(code)
    assert packed.dtype == np.uint8 and (packed == 0xFF).all()                        # degenerate, as the reference
    n_par = sum(p.numel() for p in m.parameters())
    assert n_par == 3 * 7079040 + 24773728


def test_error_behaviour_mirrors_reference():
    import config
    import modules
    config.init()
    with pytest.raises(ValueError, match='Not valid'):
        modules.Cell({'cell': 'NoSuchCell'})
    with pytest.raises(ValueError, match='Not valid activation'):
        modules.Cell({'cell': 'Activation', 'mode': 'bogus'})
    with pytest.raises(ValueError, match='Not valid normalization'):
        modules.Cell({'cell': 'Normalization', 'mode': 'bogus', 'input_size': 4})
    with pytest.raises(ValueError, match='Not valid model mode'):
        modules.Cell({'cell': 'PixelShuffleCell', 'mode': 'sideways', 'scale_factor': 2})
    from models.utils import make_model
    with pytest.raises(ValueError, match='Not valid model format'):
        make_model(3)
    m = build_model('dist_codec', 'MNIST', 1, 2, device='cpu', loss_fn='huber')
    with pytest.raises(ValueError, match='Not valid loss function'):
        m({'img': torch.rand(1, 1, 32, 32), 'label': torch.zeros(1, dtype=torch.long)})


def test_no_cpu_fallback():
    m = build_model('dist_codec', 'MNIST', 1, 2, device='cpu')
    with pytest.raises(B.DrasicError, match='no CPU path'):
        m({'img': torch.rand(2, 1, 32, 32), 'label': torch.zeros(2, dtype=torch.long)})
    import modules
    cell = modules.Cell({'cell': 'ConvLSTMCell', 'input_size': 64, 'output_size': 64, 'kernel_size': 3, 'stride': 1,
                         'padding': 1, 'bias': False, 'num_layers': 1, 'activation': 'tanh'})
    with pytest.raises(B.DrasicError, match='no CPU path'):
        cell(torch.rand(1, 64, 8, 8))
    q = modules.Cell({'cell': 'QuantizationCell'})
    with pytest.raises(B.DrasicError, match='no CPU path'):
        q(torch.rand(4))


def test_shuffle_functions_match_golden(golden_dir):
    import functions
    g = np.load(os.path.join(golden_dir, 'cells.npz'))
    x = torch.from_numpy(g['shuffle_in'])
    assert np.array_equal(functions.pixel_unshuffle(x, 2).numpy(), g['unshuffle_out'])
    assert np.array_equal(functions.pixel_shuffle(x, 2).numpy(), g['shuffle_out'])


@pytest.mark.parametrize('order', [B.ORDER_IDENT, B.ORDER_QUAD])
def test_plane_layout_round_trip(order):
    x = torch.rand(3, 16, 4, 4) * 2 - 1
    hi, lo = to_planes(x, 1, pad_to=8, order=order)
    assert hi.shape == (8, 4, 4, 16) and hi.dtype == torch.float16
    rec = (hi.float() + lo.float()) / 256.0
    back = from_nhwc(rec, 3, order=order)
    assert (back - x).abs().max() < 2e-7          # hi + lo carries ~22 significand bits
    if order == B.ORDER_QUAD:                     # physical p = q*(C/4)+c  <->  logical c*4+q
        assert torch.allclose(rec[0, 1, 2, 1 * 4 + 3], x[0, 3 * 4 + 1, 1, 2], atol=2e-7)


def test_oracle_decoder_half_reproduces_forward_reconstructions():
    """codec_decode (decoder half alone, fed the emitted symbols) == codec_forward's reconstructions; the
    real bitstream round-trips through pack_bits / unpack_bits"""
    import torch
    from oracle import drasic_oracle as O
    for separate, M in ((False, 3), (True, 2)):
        sd = O.init_state_dict(3, 8, M, separate=separate, seed=1)
        g = torch.Generator().manual_seed(2)
        img = torch.rand(5, 3, 32, 32, generator=g)
        label = torch.randint(0, M, (5,), generator=g)
        with torch.no_grad():
            ref = O.codec_forward(sd, img, label, 3, M, separate=separate)
            dec = O.codec_decode(sd, ref['code_pm1'], label, M, separate=separate)
        for a, b in zip(dec, ref['img']):
            assert torch.allclose(a, b, atol=1e-6)
        bits = O.pack_bits(ref['code_pm1'][0])
        assert bits.shape == (5, 16)
        assert torch.equal(O.unpack_bits(bits, (8, 4, 4)), ref['code_pm1'][0])


def test_metrics_module_definitions_on_cpu_tensors():
    import torch
    import metrics
    from oracle import drasic_oracle as O
    g = torch.Generator().manual_seed(3)
    a, b = torch.rand(4, 3, 32, 32, generator=g), torch.rand(4, 3, 32, 32, generator=g)
    assert abs(metrics.PSNR(a, b) - O.psnr(a, b)) < 1e-4
    code = np.zeros((4 * 16, 3), dtype=np.uint8)
    assert metrics.BPP(code, a) == O.bpp(code, a)
    out = {'mse': torch.tensor([0.01, 0.001])}
    assert np.allclose(metrics.psnr_per_iteration(out), [20.0, 30.0], atol=1e-4)


def test_caller_shims_torch2_scheduler_data_and_logger(tmp_path):
    """SURVEY 8f N2 host pieces: config.init() makes ReduceLROnPlateau(verbose=True) (train_model.py:179-182)
    constructible on torch 2.x; the synthetic data module and the logger keep the reference's interfaces"""
    import pickle
    import torch
    import config
    config.init()
    p = torch.nn.Parameter(torch.zeros(2))
    opt = torch.optim.Adam([p], lr=1.0)
    s = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode='min', factor=0.1, patience=0, verbose=True,
                                                   threshold=1e-3, threshold_mode='rel')
    s.step(1.0)
    s.step(2.0)
    assert abs(opt.param_groups[0]['lr'] - 0.1) < 1e-12
    import data
    import utils
    config.PARAM.update(data_size={'train': 6, 'test': 3}, batch_size={'train': 4, 'test': 3})
    config.PARAM['control']['num_node'] = '3'
    ds = data.fetch_dataset('MNIST', 'label')
    assert len(ds['train']) == 6 and len(ds['test']) == 3
    a, b = ds['train'][2], ds['train'][2]
    assert torch.equal(a['img'], b['img']) and tuple(a['img'].shape) == (1, 32, 32) and 0 <= int(a['label']) < 3
    batch = utils.collate(next(iter(data.make_data_loader(ds)['test'])))
    assert tuple(batch['img'].shape) == (3, 1, 32, 32) and batch['label'].dtype == torch.int64
    with pytest.raises(ValueError):
        data.fetch_dataset('NoSuchData', 'label')
    from logger import Logger
    lg = Logger(str(tmp_path / 'log'))
    lg.append({'info': ['m', 'e']}, 'test', mean=False)
    lg.append({'Loss': 1.0, 'PSNR': [10.0, 20.0]}, 'test', n=2)
    lg.append({'Loss': 3.0, 'PSNR': [20.0, 40.0]}, 'test', n=2)
    assert lg.mean['test/Loss'] == 2.0 and lg.mean['test/PSNR'] == [15.0, 30.0] and lg.tracker['test/Loss'] == 3.0
    lg.write('test', ['Loss', 'PSNR'])
    lg.reset()
    assert lg.history['test/Loss'] == [2.0] and lg.tracker['test/Loss'] == 0
    with pytest.raises(ValueError):
        lg.append({'x': 'str'}, 'test')
    pickle.loads(pickle.dumps(lg))


def test_code_list_is_materialised_once_and_only_on_access():
    """output['code'] (dist_codec.py:42,61-63) is built from one device->host copy at the first access"""
    from models.codec_base import CodeList
    import utils
    calls = []

    def make():
        calls.append(1)
        return [np.full((4, t + 1), 255, dtype=np.uint8) for t in range(3)]
    c = CodeList(3, make)
    assert len(c) == 3 and calls == [] and 'pending' in repr(c)
    assert c[-1].nbytes == 12 and calls == [1]
    assert [a.shape for a in c] == [(4, 1), (4, 2), (4, 3)] and len(list(reversed(c))) == 3 and calls == [1]
    assert utils.recur(lambda a: a.nbytes, c) == [4, 8, 12]              # metrics.BPP's walk (metrics.py:43)
    assert np.concatenate(c, axis=1).shape == (4, 6)                      # C-level consumers iterate it correctly
    with pytest.raises(ValueError):
        utils.recur(lambda a: a, 'abc')


def test_module_tree_index_follows_the_parameters():
    m = build_model('dist_codec', 'CIFAR10', 2, 2, seed=5, device='cpu')
    names, params, cells = m._index()
    assert names == [k for k, _ in m.named_parameters()] and all(a is b for a, b in zip(params, m.parameters()))
    assert len(cells) >= 9 and all(hasattr(c, 'free_hidden') for c in cells)
    assert m._index() is m._index()
    m.load_state_dict(m.state_dict())
    assert m._tree is None
    m._index()
    m.double()
    assert m._tree is None and all(a is b for a, b in zip(m._index()[1], m.parameters()))


def test_batch_plan_invariants_property():
    """any labelling, any padding unit: perm is a bijection onto the batch, every slot lies in exactly one node's
    range, node order is kept inside a node, every tile a node's pixels touch is walked by that node, and wgrad
    K blocks never straddle two nodes when the plan is a training plan (pad 4)"""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 10).flatmap(lambda m: st.tuples(st.just(m), st.lists(st.integers(0, m - 1), min_size=1, max_size=70))),
           st.sampled_from([1, 4]))
    def check(case, pad):
        M, labels = case
        label = np.array(labels)
        p = BatchPlan(label, M, 'cpu', pad=pad, hws=(256, 64, 16))
        assert p.B_pad % 8 == 0 and p.B_pad >= label.size
        valid = p.perm_host[p.perm_host >= 0]
        assert sorted(valid.tolist()) == list(range(label.size))
        ends = p.group_img_end_host
        assert (np.diff(np.concatenate(([0], ends))) >= 0).all() and ends[-1] == p.B_pad
        for j in range(M):
            lo = 0 if j == 0 else ends[j - 1]
            idx = p.perm_host[lo:ends[j]]
            idx = idx[idx >= 0]
            assert (label[idx] == j).all() and (np.diff(idx) > 0).all() and idx.size == (label == j).sum()
            if pad == 4:
                assert lo % 4 == 0 and ends[j] % 4 == 0
        for hw in (256, 64, 16):
            for pair in (False, True):
                u, total = p.m_units(hw, pair)
                u = u.numpy()
                assert total == u[2, -1] and (np.diff(np.concatenate(([0], u[2]))) >= 0).all()
                for j in range(M):
                    lo = 0 if j == 0 else ends[j - 1]
                    if ends[j] > lo:
                        assert u[0, j] * 128 <= lo * hw and u[1, j] * 128 >= ends[j] * hw    # tiles cover the node's pixels
                        n = u[1, j] - u[0, j]
                        got = u[2, j] - (u[2, j - 1] if j else 0)
                        assert got == ((n + 1) // 2 if pair else n)
            kb = p.kb_end(hw).numpy()
            if pad == 4:
                assert (kb * 64 == ends.astype(np.int64) * hw).all()                           # K blocks end exactly at node ends
    check()


def test_wgrad_k_split_picker():
    """_pick_k_splits: within the K blocks a node offers, aim for >= 3 waves of work on the persistent grid and
    prefer the split whose last wave is fullest"""
    from engine.codec import _pick_k_splits
    assert 1 <= _pick_k_splits(tiles=1000, units=74, k_max=64) <= 4       # plenty of tiles already: only wave balance
    for tiles, units, k_max in [(9, 74, 32), (18, 74, 16), (90, 74, 16), (180, 74, 65), (45, 148, 8), (1, 74, 1), (7, 74, 2)]:
        k = _pick_k_splits(tiles, units, k_max)
        assert 1 <= k <= k_max
        n = tiles * k
        waves = -(-n // units)
        # no admissible split with at least as many waves fills its last wave better
        k_lo = max(1, min(k_max, (3 * units + tiles - 1) // tiles))
        assert k >= k_lo or k == k_max
        eff = n / float(waves * units)
        for other in range(k_lo, max(k_lo, min(k_max, 4 * k_lo)) + 1):
            m = tiles * other
            assert m / float(-(-m // units) * units) <= eff + 1e-3 or other > k


def test_position_order_heaviest_first():
    """engine.codec.position_order: every position once, in-bounds tap counts non-increasing (9, 6, 4)"""
    from engine.codec import position_order
    for h, w in [(4, 4), (8, 8), (16, 16), (2, 2), (4, 6), (1, 1)]:
        order = position_order(h, w)
        assert sorted(order.tolist()) == list(range(h * w))
        y, x = np.divmod(order, w)
        taps = (3 - (y == 0) - (y == h - 1)) * (3 - (x == 0) - (x == w - 1))
        assert np.all(np.diff(taps) <= 0)
    taps44 = sorted(((3 - (p // 4 == 0) - (p // 4 == 3)) * (3 - (p % 4 == 0) - (p % 4 == 3))) for p in range(16))
    assert sum(taps44) == 100                                      # 100 of 144 (position, tap) pairs are in bounds at 4x4


def test_plan_units_position_major():
    """CodecEngine._plan_units: position-major tiles cover whole units of 128 (pairs: 256) slots, pixel-major the rest"""
    from engine.codec import CodecEngine, _Layer
    eng = CodecEngine(3, 8, 1)
    for Bp, h, w, pair, want_pos, want_units in [(1024, 4, 4, True, 1024, 16 * 4), (1040, 8, 8, True, 1024, 64 * 4 + 4),
                                                 (128, 4, 4, False, 128, 16), (200, 16, 16, False, 128, 256 + 144),
                                                 (64, 8, 8, False, 0, 32), (128, 4, 4, True, 128, 8), (200, 8, 8, True, 128, 32 + 18),
                                                 (64, 8, 8, True, 0, 16)]:
        ly = _Layer()
        ly.h, ly.w, ly.pair, ly.groups = h, w, pair, 1
        eng._plan_units(ly, Bp, None, 'cpu')
        assert (ly.pos_imgs, ly.n_units) == (want_pos, want_units), (Bp, h, w, pair, ly.pos_imgs, ly.n_units)
        assert (ly.pos_order is not None) == (want_pos > 0)
        assert ly.pos_pairs == int(pair and 128 <= Bp < 256)       # a pair on fewer than 256 slots walks two positions
    eng.pos_major = False
    ly = _Layer()
    ly.h, ly.w, ly.pair, ly.groups = 4, 4, True, 1
    eng._plan_units(ly, 1024, None, 'cpu')
    assert (ly.pos_imgs, ly.n_units) == (0, 64)


def test_position_pairs_share_their_tap_sets():
    """engine.codec.position_pairs: every position once; 9-tap pairs first, then 6-tap unions (edges, corner pairs)"""
    from engine.codec import position_pairs

    def taps(pos, h, w):
        y, x = divmod(int(pos), w)
        return sum(1 << t for t in range(9) if 0 <= y + t // 3 - 1 < h and 0 <= x + t % 3 - 1 < w)
    for h, w in [(4, 4), (8, 8), (16, 16), (4, 8)]:
        pairs = position_pairs(h, w).reshape(-1, 2)
        assert sorted(pairs.reshape(-1).tolist()) == list(range(h * w))
        union = [bin(taps(a, h, w) | taps(b, h, w)).count('1') for a, b in pairs]
        n9 = (h - 2) * (w - 2) // 2
        assert union[:n9] == [9] * n9 and union[n9:] == [6] * ((h - 2) + (w - 2) + 2)


def test_pads_for_follow_the_smallest_layer():
    """engine.codec.pads_for: node ranges are whole wgrad K blocks (64 pixels) of the smallest layer when a backward will
    run, the batch whole M tiles (128 pixels); inference packs the nodes back to back"""
    from engine.codec import pads_for, plan_arrays
    assert pads_for(32, 32, True) == (4, 8) and pads_for(32, 32, False) == (1, 8)
    assert pads_for(16, 16, True) == (16, 32) and pads_for(16, 16, False) == (1, 32)
    assert pads_for(512, 768, True) == (4, 8)
    # the 28 / 36 split of 16x16 inputs that used to put two nodes into one K block
    pad, total = pads_for(16, 16, True)
    perm, ends, counts = plan_arrays(np.array([0] * 28 + [1] * 36), 2, pad=pad, total_pad=total)
    assert list(ends) == [32, 96] and perm.size == 96 and np.all((ends * 4) % 64 == 0)


def test_active_groups_and_single_node_collapse():
    """CodecEngine._active_groups: nodes without samples are not packed; a per-node layer with ONE active node runs as
    one group on that node's planes"""
    from engine.codec import BatchPlan, CodecEngine, _Layer
    ly = _Layer()
    ly.groups = 4
    CodecEngine._active_groups(ly, BatchPlan(np.array([2, 2, 0, 2]), 4, 'cpu', pad=1))
    assert (ly.active, ly.g0, ly.gemm_groups) == ([0, 2], 0, 4)
    CodecEngine._active_groups(ly, BatchPlan(np.array([2, 2, 2]), 4, 'cpu', pad=1))
    assert (ly.active, ly.g0, ly.gemm_groups) == ([2], 2, 1)
    ly.groups = 1
    CodecEngine._active_groups(ly, BatchPlan(np.array([0, 0]), 1, 'cpu', pad=1))
    assert (ly.active, ly.g0, ly.gemm_groups) == (None, 0, 1)


def test_bench_traffic_comes_from_the_committed_capture_of_the_same_configuration():
    import importlib.util
    spec = importlib.util.spec_from_file_location('bench_mod', os.path.join(ROOT, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    t, note = bench.captured_traffic('eval', 1024, 8, 'parity')
    assert t is not None and 2e8 < t < 5e9 and 'r2_ncu_full_gates_eval_B1024_summary.csv' in note
    tt, note = bench.captured_traffic('train', 1024, 8, 'parity')          # the training forward keeps every gate: more writes
    assert tt is not None and t < tt < 5e9 and 'r2_ncu_full_gates_train_fwd_B1024_summary.csv' in note
    for other in [('train', 128, 8, 'parity'), ('eval', 128, 8, 'parity'), ('eval', 1024, 1, 'parity'), ('eval', 1024, 8, 'fast')]:
        assert bench.captured_traffic(*other)[0] is None
    a, b = bench.workload_config('train', 1024, 1), bench.workload_config('train', 1024, 1)
    assert a == b and set(a) >= {'workload', 'batch_per_gpu', 'global_batch', 'num_node', 'num_iter'}
