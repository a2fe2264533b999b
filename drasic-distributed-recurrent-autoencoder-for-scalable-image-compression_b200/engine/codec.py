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

PAD = 4        # slots per node are padded to a multiple of 4 images (one 64-pixel wgrad K block at the 4x4 layers)
TOTAL_PAD = 8  # ... and the whole batch to 8 (one 128-pixel M tile at the 4x4 layers)


def pads_for(H, W, training):
    """(slots-per-node multiple, whole-batch multiple) for H x W images: the smallest layer (E3 / D1, H/8 x W/8) has
    hw pixels per image, a wgrad K block is 64 pixels and an M tile 128.  32x32 images give (4, 8); 16x16 images
    (4 pixels per image there) need node ranges of 16 images when a backward will run."""
    hw = max(1, (H // 8) * (W // 8))
    return (max(PAD, -(-64 // hw)) if training else 1), max(TOTAL_PAD, -(-128 // hw))

# (name, Cin, Co, spatial divisor, in_order, out_order, next_mode, next_f32)
ENC_LAYERS = (('E1', 128, 128, 2, B.ORDER_QUAD, B.ORDER_IDENT, B.NEXT_DOWN, 0),
              ('E2', 512, 128, 4, B.ORDER_QUAD, B.ORDER_IDENT, B.NEXT_DOWN, 0),
              ('E3', 512, 128, 8, B.ORDER_QUAD, B.ORDER_IDENT, B.NEXT_SAME, 1))
DEC_LAYERS = (('D1', 128, 512, 8, B.ORDER_IDENT, B.ORDER_QUAD, B.NEXT_UP, 0),
              ('D2', 128, 512, 4, B.ORDER_IDENT, B.ORDER_QUAD, B.NEXT_UP, 0),
              ('D3', 128, 128, 2, B.ORDER_IDENT, B.ORDER_QUAD, B.NEXT_SAME, 1))  # fp32 copy in D3's own quad-major order for the tail
ENC_IDX = {'E1': 2, 'E2': 4, 'E3': 6}   # nn.Sequential indices, dist_codec.py:70-86
DEC_IDX = {'D1': 1, 'D2': 3, 'D3': 5}   # dist_codec.py:88-104


def plan_arrays(label_host, num_node, pad_total=None, pad=None, total_pad=TOTAL_PAD):
    """node-sort one batch: (perm [B_pad] slot -> original index or -1, exclusive end slot per node)"""
    label_host = np.asarray(label_host).astype(np.int64).reshape(-1)
    if label_host.size and (label_host.min() < 0 or label_host.max() >= num_node):
        # the reference mis-indexes and raises IndexError for label >= num_node (dist_codec.py:52)
        raise IndexError('label out of range for num_node={}'.format(num_node))
    order = np.argsort(label_host, kind='stable')              # keeps the reference's within-node order
    counts = np.bincount(label_host, minlength=num_node)
    pad = PAD if pad is None else int(pad)
    padded = (counts + pad - 1) // pad * pad
    ends = np.cumsum(padded)
    need = (int(ends[-1]) + total_pad - 1) // total_pad * total_pad
    B_pad = need if pad_total is None else int(pad_total)
    if B_pad < need or B_pad % total_pad:
        raise ValueError('Not valid pad_total')
    perm = np.full(B_pad, -1, dtype=np.int32)
    starts = ends - padded
    src = 0
    for j in range(num_node):
        n = int(counts[j])
        perm[starts[j]:starts[j] + n] = order[src:src + n]
        src += n
    # the alignment slots at the end belong to the last node: every slot is inside exactly one node's range, so
    # every row of every tensor is (re)written each pass -- padding rows carry zeros through backward
    ends = ends.copy()
    ends[-1] = B_pad
    return perm, ends.astype(np.int32), counts


def worst_case_pad(B, num_node, pad=PAD, total_pad=TOTAL_PAD):
    """slot count that fits any assignment of B samples to num_node nodes whose ranges are multiples of `pad`"""
    return (B + (pad - 1) * min(num_node, B) + total_pad - 1) // total_pad * total_pad


class BatchPlan(object):
    """Node-sorted, padded slot layout of one batch.  With `static_hw` the plan owns fixed device
    buffers (worst-case slot count) that update() refreshes in place -- what a captured CUDA graph reads."""

    def __init__(self, label_host, num_node, device, pad_total=None, static_hw=None, pad=None, hws=None,
                 total_pad=TOTAL_PAD):
        """pad: slots per node are a multiple of this; PAD (4) when backward will run (wgrad K blocks of 4 images
        at the 4x4 layers must not mix two nodes), 1 for inference (shared M tiles are handled by row masks).
        hws: pixels per image of the layers that will read this plan; their unit / K-block tables travel to the
        device together with perm and the node ends in ONE asynchronous copy from pinned memory (a table asked for
        later costs a synchronising copy of its own)."""
        self.num_node, self.device = num_node, device
        self.pad_total = pad_total
        self.pad = PAD if pad is None else int(pad)
        self.total_pad = int(total_pad)
        perm, ends, counts = plan_arrays(label_host, num_node, pad_total, self.pad, self.total_pad)
        self.B, self.B_pad = int(np.asarray(label_host).size), int(perm.size)
        self.counts, self.group_img_end_host, self.perm_host = counts, ends, perm
        self._kb_end, self._units = {}, {}
        self.static = static_hw is not None
        pre = tuple(dict.fromkeys(static_hw if static_hw is not None else (hws or ())))
        self._pre = pre
        flat = self._flat_host()
        on_gpu = torch.device(device).type == 'cuda'
        host = torch.empty(flat.size, dtype=torch.int32, pin_memory=on_gpu)
        host.numpy()[:] = flat
        buf = host.to(device, non_blocking=True) if on_gpu else host
        self._buf, self._stage, self._n_update = buf, [], 0
        M, off = num_node, 0

        def take(n):
            nonlocal off
            off += n
            return buf[off - n:off]
        self.perm = take(perm.size)
        self.group_img_end = take(M)
        for hw in pre:
            self._kb_end[hw] = take(M)
            for pair in (False, True):
                self._units[(hw, pair)] = take(3 * M).view(3, M)

    def _flat_host(self):
        """perm, node ends and the K-block / unit tables of the pre-registered layer sizes as one int32 array (the layout
        of the device buffer the tables are views of)"""
        parts = [self.perm_host, self.group_img_end_host]
        for hw in self._pre:
            parts.append(self._kb_host(hw))
            parts.extend(self._units_host(hw, pair).reshape(-1) for pair in (False, True))
        return np.concatenate(parts).astype(np.int32, copy=False)

    def kb_end(self, hw):
        """exclusive end K block (64 pixels) of each node, for the wgrad GEMM (node ranges are multiples of 4
        images, so K blocks never mix two nodes)"""
        if hw not in self._kb_end:
            if self.static and getattr(self, '_frozen', False):
                raise RuntimeError('static batch plan has no buffer for hw={}'.format(hw))
            self._kb_end[hw] = torch.from_numpy(self._kb_host(hw)).to(self.device)
        return self._kb_end[hw]

    def _kb_host(self, hw):
        px = self.group_img_end_host.astype(np.int64) * hw
        if self.pad > 1 and np.any(px % 64):
            # a 64-pixel K block of the wgrad GEMM would mix the samples of two nodes (their weights differ)
            raise ValueError('Not valid batch plan: node ranges of {} pixels per image are not K-block aligned'.format(hw))
        return (px // 64).astype(np.int32)

    def _units_host(self, hw, pair):
        """per node: first / end M tile (128 pixels) overlapping its slot range -- neighbours may share a tile
        where an image is smaller than 32 pixels -- and the running count of work units (tiles, or pairs of tiles)"""
        ends = self.group_img_end_host.astype(np.int64)
        starts = np.concatenate(([0], ends[:-1]))
        t0 = (starts * hw) // 128
        t1 = -((-ends * hw) // 128)
        t1 = np.where(ends > starts, t1, t0)                    # empty nodes walk nothing
        n = t1 - t0
        units = (n + 1) // 2 if pair else n
        return np.stack([t0, t1, np.cumsum(units)]).astype(np.int32)

    def m_units(self, hw, pair):
        """(device [3, num_node]: tile start, tile end, running unit end; total units) for the conv GEMMs"""
        host = self._units_host(hw, pair)
        key = (hw, bool(pair))
        if key not in self._units:
            if self.static and getattr(self, '_frozen', False):
                raise RuntimeError('static batch plan has no buffer for hw={}'.format(hw))
            self._units[key] = torch.from_numpy(host).to(self.device)
        total = int(host[2, -1])
        if self.static:      # a captured graph launches a fixed unit count: the bound for any labelling (extra units are no-ops)
            tiles = self.B_pad * hw // 128 + self.num_node
            total = (tiles + self.num_node + 1) // 2 if pair else tiles
        return self._units[key], total

    def update(self, label_host):
        """refresh the device buffers in place for a new batch of the same size"""
        perm, ends, counts = plan_arrays(label_host, self.num_node, self.B_pad, self.pad, self.total_pad)
        if int(np.asarray(label_host).size) != self.B:
            raise ValueError('Not valid batch size for this static plan')
        self._frozen = True
        self.counts, self.group_img_end_host, self.perm_host = counts, ends, perm
        flat = self._flat_host()
        for hw, t in self._kb_end.items():                      # (tables asked for outside the pre-registered sizes)
            if hw not in self._pre:
                t.copy_(torch.from_numpy(self._kb_host(hw)))
        for (hw, pair), t in self._units.items():
            if hw not in self._pre:
                t.copy_(torch.from_numpy(self._units_host(hw, pair)))
        if not self._buf.is_cuda:
            self._buf.numpy()[:] = flat
            return
        # ONE asynchronous copy from pinned memory (the tables are views of one device buffer).  Two staging buffers
        # take turns; one is rewritten only after the copy that last read it has completed.
        if len(self._stage) < 2:
            self._stage.append([torch.empty(flat.size, dtype=torch.int32, pin_memory=True), None])
        stage = self._stage[self._n_update % len(self._stage)]
        self._n_update += 1
        if stage[1] is not None:
            stage[1].synchronize()
        stage[0].numpy()[:] = flat
        self._buf.copy_(stage[0], non_blocking=True)
        stage[1] = torch.cuda.Event()
        stage[1].record(torch.cuda.current_stream(self._buf.device))


class _Layer(object):
    pass


def position_pairs(h, w):
    """the h*w/2 pairs of pixel positions a CTA pair walks on ONE block of 128 images (even planes of at least 4x4):
    partners share their set of in-bounds 3x3 taps wherever they can -- interior positions (9 taps) pair with their
    right neighbour, edge positions (6 taps) with the next one along the same edge -- and the four corners pair
    left-right (4 taps each, 6 in the union).  Order: the 9-tap pairs first (the long K loops), then the 6-tap ones."""
    pairs = []
    for y in range(1, h - 1):
        pairs += [(y * w + x, y * w + x + 1) for x in range(1, w - 1, 2)]
    for y in (0, h - 1):
        pairs += [(y * w + x, y * w + x + 1) for x in range(1, w - 1, 2)]
    for x in (0, w - 1):
        pairs += [(y * w + x, (y + 1) * w + x) for y in range(1, h - 1, 2)]
    pairs += [(0, w - 1), ((h - 1) * w, (h - 1) * w + w - 1)]
    return np.asarray(pairs, dtype=np.int32).reshape(-1)


def position_order(h, w):
    """pixel positions y*w + x of an h x w plane, most in-bounds 3x3 taps first (interior 9, edges 6, corners 4):
    the order in which the conv GEMMs walk their position-major tiles, so that the persistent CTAs -- dealt the
    units round robin -- get the same mix of long and short K loops and finish together"""
    y, x = np.divmod(np.arange(h * w), w)
    taps = (3 - (y == 0) - (y == h - 1)) * (3 - (x == 0) - (x == w - 1))
    return np.argsort(-taps, kind='stable').astype(np.int32)


def _pick_k_splits(tiles, units, k_max):
    """K splits of the wgrad GEMM: the persistent grid walks tiles*k work items in waves of `units`;
    pick the split count (at least ~3 waves of work when K allows) whose last wave is fullest."""
    best_k, best_eff = 1, 0.0
    k_lo = max(1, min(k_max, (3 * units + tiles - 1) // tiles))
    for k in range(k_lo, max(k_lo, min(k_max, 4 * k_lo)) + 1):
        n = tiles * k
        eff = n / float(-(-n // units) * units)
        if eff > best_eff + 1e-3:
            best_k, best_eff = k, eff
    return int(best_k)


class CodecEngine(object):
    """Runs the T-iteration loop for one codec (dist: per-node encoders + shared decoder; sep: per-node
    encoders and decoders).  `precision`: 'parity' = fp16 hi+lo operands with IEEE transcendental
    functions (fp32-grade, bit-exact codes); 'fast' = single fp16 operands with tanh.approx."""

    def __init__(self, num_channel, code_size, num_node, separate=False, precision='parity', grad_precision='fp16',
                 cta_pair='auto', stream_k='auto', deterministic=False):
        """grad_precision: 'fp16' = backward GEMMs on single fp16 operands, gates saved in fp16 (gradients
        within ~1e-3 relative of the fp32 reference); 'parity' = split-fp16 backward GEMMs and fp32 saved
        gates (fp32-grade gradients; needs precision='parity')."""
        if precision not in ('parity', 'fast'):
            raise ValueError('Not valid precision')
        if grad_precision not in ('fp16', 'parity') or (grad_precision == 'parity' and precision != 'parity'):
            raise ValueError('Not valid grad_precision')
        if cta_pair not in ('auto', 'on', 'off'):
            raise ValueError('Not valid cta_pair')
        if stream_k not in ('auto', 'on', 'off'):
            raise ValueError('Not valid stream_k')
        self._ctor = dict(num_channel=num_channel, code_size=code_size, num_node=num_node, separate=separate,
                          precision=precision, grad_precision=grad_precision, cta_pair=cta_pair, stream_k=stream_k,
                          deterministic=deterministic)
        # deterministic=True: gradients are bit-identical from run to run, like the reference's CPU path -- every
        # weight-gradient element of the wgrad GEMM has a single writer per launch (no K split), and the pointwise
        # backward kernels accumulate their weight gradients in int64 fixed point instead of with float atomics.  (The
        # loss value itself is a sum of fp64 atomics: order dependent in its last bits, identical after rounding to fp32.)
        self.deterministic = bool(deterministic)
        # counts the forwards that saved state for a backward, over this engine and the engines cloned from it (one
        # per captured graph): whoever holds packed weight planes re-packs when it has moved (see _stale below)
        self.train_count = [0]
        self.cta_pair = cta_pair  # CTA pairs (cta_group::2) for the gate / dgrad / wgrad GEMMs
        # wgrad pair tiles: 1 = 256 columns (default), 2 = 512 columns where a node has >= 256 K blocks.  The wide tiles
        # move fewer bytes per MMA but leave 9..180 work units for 74 CTA pairs; measured per step at B=1024 they are
        # slower on every layer (D1 7.8 vs 4.9 ms, E1 10.0 vs 8.5, D2 20.1 vs 17.9, D3 8.7 vs 8.5)
        self.wgrad_pair_mode = 1      # (an attribute, not an environment knob: tests set it to 2)
        self.wgrad_min_kb = 256       # K blocks (of 64 pixels) every K split of a wgrad tile keeps at least
        self.gsplit = 1 if grad_precision == 'parity' else 0
        self.C, self.code, self.M, self.separate = num_channel, code_size, num_node, separate
        self.split = 1 if precision == 'parity' else 0
        self.precise = self.split
        self.Md = num_node if separate else 1
        self._arena = None
        self._arena_key = None
        self._wcache = {}
        self._sms = None
        self.launches = 0   # kernels launched by the last forward() (bench.py reports it)
        self.force_pack = False  # True while capturing a CUDA graph: always re-pack into the same buffers
        # Fused optimizers (torch.optim.Adam(fused=True), the multi-tensor kernels) update parameters in place
        # WITHOUT bumping Tensor._version, so (data_ptr, _version) alone cannot validate the packed-weight
        # cache: a forward that saves state for backward always re-packs (the weights are about to change
        # anyway) and leaves the cache marked stale for whatever forward comes next.
        self._stale = False
        self._seen_train = 0
        self._repack = False
        self.overlap_wgrad = True   # wgrad GEMMs on a side stream
        self.profile = None  # set to [] to collect (label, flops, start_event, end_event) per launch
        # position-major M tiles for the layers whose weights are shared by the whole batch (the joint decoder; the
        # encoder too with a single source): padding taps are never issued (15 % of the MACs of an iteration)
        self.pos_major = True
        self._pos_tables = {}
        # stream-K schedule of the gate / dgrad GEMMs: 'auto' = launches with fewer than SK_WAVES whole work items
        # per CTA (pair), where dealing whole items leaves most SMs idle (the 4x4 and 8x8 layers of a small batch);
        # measured, it does not pay for launches of several waves: those are bound by L2 -> SMEM fill, a shared
        # resource, so a last wave on half of the SMs already runs faster.  'on' / 'off' force it (tests)
        self.stream_k = stream_k
        self._sk = {}

    def _pos_order(self, h, w, dev, pairs=False):
        key = (h, w, str(dev), pairs)
        if key not in self._pos_tables:
            self._pos_tables[key] = torch.from_numpy(position_pairs(h, w) if pairs else position_order(h, w)).to(dev)
        return self._pos_tables[key]

    SK_WAVES = 2

    def _sk_buffers(self, dev):
        key = str(dev)
        if key not in self._sk:
            L = B.lib()
            self._sk[key] = (torch.empty(L.drasic_stream_k_scratch_bytes() // 4, dtype=torch.float32, device=dev),
                             torch.zeros(L.drasic_stream_k_flag_bytes() // 4, dtype=torch.int32, device=dev))
        return self._sk[key]

    def _use_stream_k(self, ly, n_ntiles):
        if self.stream_k != 'auto':
            return self.stream_k == 'on'
        workers = self.sms() // 2 if ly.pair else self.sms()
        return ly.n_units * n_ntiles < self.SK_WAVES * workers and (ly.n_units * n_ntiles) % workers != 0

    @staticmethod
    def _active_groups(ly, plan):
        """which weight groups (nodes) of a layer this batch touches.  ly.active: None (all: a single group, or a
        static plan whose node counts change per replay) or the list of nodes that have samples.  When exactly one
        node of a per-node layer has samples -- a rank of a source-sharded run that owns one node, single-source
        data through a multi-node model -- the layer runs as ONE group on that node's weight planes (ly.g0), which
        opens the position-major tiles to it."""
        ly.active, ly.g0, ly.gemm_groups = None, 0, ly.groups
        if ly.groups > 1 and plan is not None and not plan.static:
            ly.active = [int(j) for j in np.nonzero(plan.counts)[0]]
            if len(ly.active) == 1:
                ly.g0, ly.gemm_groups = ly.active[0], 1

    def _plan_units(self, ly, Bp, plan, dev):
        """work units of the M dimension of one layer: per-node tile ranges (device tables of the batch plan) for
        grouped layers; for a layer shared by the whole batch, position-major tiles over as many slots as fill whole
        units (128 images, 256 with CTA pairs) followed by pixel-major tiles over the rest"""
        hw = ly.h * ly.w
        n_mtiles = Bp * hw // 128
        ly.units, ly.pos_imgs, ly.pos_order, ly.pos_pairs = None, 0, None, 0
        if getattr(ly, 'gemm_groups', ly.groups) > 1:
            ly.units, ly.n_units = plan.m_units(hw, ly.pair)
            return
        unit_imgs = 256 if ly.pair else 128
        pos_units = 0
        if self.pos_major:
            ly.pos_imgs = Bp // unit_imgs * unit_imgs
            pos_units = hw * (ly.pos_imgs // unit_imgs)
            if ly.pair and not ly.pos_imgs and Bp >= 128 and min(ly.h, ly.w) >= 4 and not ((ly.h | ly.w) & 1):
                # a CTA pair on fewer than 256 slots: two positions of the first 128 slots per unit
                ly.pos_imgs, ly.pos_pairs, pos_units = 128, 1, hw // 2
        rest = n_mtiles - ly.pos_imgs * hw // 128
        ly.n_units = pos_units + ((rest + 1) // 2 if ly.pair else rest)
        if ly.pos_imgs:
            ly.pos_order = self._pos_order(ly.h, ly.w, dev, bool(ly.pos_pairs))

    def clone(self):
        """an engine with the same configuration and its own arenas / packed weight planes (a captured CUDA graph
        bakes device addresses in: it must own every buffer it replays on)"""
        e = CodecEngine(**self._ctor)
        e.train_count = self.train_count
        return e

    def release(self):
        """free the device arenas and caches of this engine"""
        self._arena = self._arena_key = None
        self._bwd_buf = self._bwd_key = None
        self._wcache = {}
        self._sk = {}
        self._gen = getattr(self, '_gen', 0) + 1      # a pending backward() can no longer run
        self._stale = False

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
        """N tile of a single-CTA launch.  With the stream-K schedule available the tiles stay 256 columns wide
        (least L2 -> SMEM traffic per MMA) and a launch of few tiles is spread over the SMs along K instead;
        without it, narrower tiles until there is one per SM."""
        if self.stream_k != 'off':
            return 256
        for bn in (256, 128, 64):
            if n_mtiles * (4 * co // bn) >= self.sms():
                return bn
        return 64

    def _use_pair(self, n_mtiles, hw, co, groups=1, B_pad=0):
        """CTA pairs (256-column N tiles): pair units are formed inside each node (BatchPlan.m_units), a node
        with an odd number of M tiles leaves half of its last unit idle"""
        if self.cta_pair == 'off':
            return False
        if self.cta_pair == 'on':
            return True
        if self.stream_k != 'off':
            # pairs everywhere: the 256-column pair tile moves the fewest L2 -> SMEM bytes per MMA, launches of few
            # items are spread along K instead, and below 256 slots the pair walks two positions of one image block
            return True
        return (n_mtiles + 1) // 2 * (4 * co // 256) >= self.sms() // 2

    def _get_arena(self, B_pad, H, W, device, T=0):
        """T == 0: inference arena (h ping-pongs, c in place, one buffer per edge).  T > 0: training
        arena that keeps every iteration's h, c, gates and layer inputs for backward-through-time.
        Buffers are zero-initialised once and reused; padding slots therefore always hold finite data."""
        key = (B_pad, H, W, str(device), T)
        if self._arena_key == key:
            return self._arena
        self._arena = None          # release the previous arena before allocating the next
        A = {'T': T}
        n_t = max(T, 1)
        f16 = dict(dtype=torch.float16, device=device)
        f32 = dict(dtype=torch.float32, device=device)

        def planes(n, *shape):
            return [[torch.zeros(shape, **f16), torch.zeros(shape, **f16) if self.split else None] for _ in range(n)]
        A['e1_in'] = planes(n_t, B_pad, H // 2, W // 2, 128)
        A['d1_in'] = planes(n_t, B_pad, H // 8, W // 8, 128)
        for (name, cin, co, div, _, _, nmode, nf32) in ENC_LAYERS + DEC_LAYERS:
            h, w = H // div, W // div
            L = {}
            L['h'] = planes(T if T else 2, B_pad, h, w, co)
            L['c'] = [torch.zeros((B_pad, h, w, co), **f32) for _ in range(n_t)]
            L['gates'] = [torch.zeros((B_pad, h, w, 4, co), **(f32 if self.gsplit else f16)) for _ in range(T)]
            if nmode == B.NEXT_DOWN:
                nshape = (B_pad, h // 2, w // 2, 4 * co)
            elif nmode == B.NEXT_UP:
                nshape = (B_pad, 2 * h, 2 * w, co // 4)
            else:
                nshape = (B_pad, h, w, co)
            L['next'] = [[torch.zeros(nshape, **f32), None] for _ in range(n_t)] if nf32 else planes(n_t, *nshape)
            A[name] = L
        self._arena, self._arena_key = A, key
        return A

    def _packed(self, name, weights, cin, co, in_order, out_order, block_n, stream, needed=None):
        """weights: list over groups of (w_in, w_h) fp32 tensors.  Cached on tensor identity + version; within one
        cache generation only the groups a call needs (`needed`: nodes that have samples in this batch; None = all)
        are packed -- a rank of a source-sharded run never packs the encoders of the other ranks' nodes."""
        key = (name, block_n) + tuple((w.data_ptr(), w._version) for pair in weights for w in pair)
        hit = self._wcache.get(name)
        G = len(weights)
        need = set(range(G)) if needed is None else set(int(g) for g in needed)
        if hit is not None and hit[0] == key and not (self.force_pack or self._repack) and need <= hit[2]:
            return hit[1]
        L = B.lib()
        dev = weights[0][0].device
        K = 9 * (cin + co)
        fresh = set()
        if hit is not None and hit[0][:2] == key[:2] and hit[1][0].shape == (G * 4 * co, K):
            hi, lo, inv, _, scratch = hit[1]                   # repack into the same (graph-static) buffers
            if hit[0] == key and not (self.force_pack or self._repack):
                fresh = hit[2]                                 # same generation: only the missing groups
        else:
            hi = torch.empty((G * 4 * co, K), dtype=torch.float16, device=dev)
            lo = torch.empty((G * 4 * co, K), dtype=torch.float16, device=dev)
            inv = torch.empty((G,), dtype=torch.float32, device=dev)
            scratch = torch.zeros((G,), dtype=torch.int32, device=dev)
        todo, keep_alive = [], []
        for g, (w_in, w_h) in enumerate(weights):
            if tuple(w_in.shape) != (4 * co, cin, 3, 3) or tuple(w_h.shape) != (4 * co, co, 3, 3):
                raise ValueError('Not valid ConvLSTM weight shape for {}'.format(name))
            if g in fresh or g not in need:
                continue
            w_in, w_h = w_in.detach().contiguous(), w_h.detach().contiguous()
            keep_alive.append((w_in, w_h))
            todo.append((g, w_in, w_h))
        if todo:                                               # all groups of the layer in two launches
            a_in, a_h, slots, n = B.pointer_arrays(todo)
            self._timed('pack_weights', 0.0, lambda: B.check(
                L.drasic_pack_lstm_weights_batch(a_in, a_h, slots, n, G, cin, co, in_order, out_order, block_n,
                                                 hi.data_ptr(), lo.data_ptr(), inv.data_ptr(), scratch.data_ptr(), stream),
                'pack_lstm_weights'))
            self.launches += 2
        w_in, w_h = keep_alive[-1] if keep_alive else (None, None)
        val = (hi, lo, inv, (w_in, w_h), scratch)
        self._wcache[name] = (key, val, fresh | need)
        return val

    # ------------------------------------------------------------------ the loop
    def forward(self, sd, img, *args, **kwargs):
        """see _forward; runs with the tensors' device current (the C ABI launches on the current device)"""
        if not img.is_cuda:
            raise B.DrasicError('DRASIC B200 engine needs CUDA tensors; there is no CPU path')
        with torch.cuda.device(img.device):
            return self._forward(sd, img, *args, **kwargs)

    def backward(self, out, g_loss=None):
        """see _backward"""
        with torch.cuda.device(out['_bwd']['img'].device):
            return self._backward(out, g_loss)

    def decode(self, sd, bits, *args, **kwargs):
        """see _decode"""
        if not bits.is_cuda:
            raise B.DrasicError('DRASIC B200 engine needs CUDA tensors; there is no CPU path')
        with torch.cuda.device(bits.device):
            return self._decode(sd, bits, *args, **kwargs)

    def _forward(self, sd, img, label, num_iter, training=False, noise=None, loss_kind='mae',
                 label_host=None, want_z=False, save_for_backward=False, plan=None, pack_only=False,
                 node_iters=None):
        """sd: mapping reference state_dict key -> fp32 CUDA tensor.  img [B,C,H,W] fp32 in [0,1],
        label [B] node ids.  noise: None (eval / deterministic sign) or [T,B,code,H/8,W/8] uniforms.
        node_iters: None, or one iteration budget per node (heterogeneous rates, BASELINE config 4, inferred
        semantics restated in oracle.codec_forward): at iteration i >= node_iters[j] the samples of node j
        receive no further refinement (their reconstruction stays, the loss keeps counting it).
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
        if Bn == 0:
            # the reference fails in torch.cat over an empty list of per-node outputs (dist_codec.py:49)
            raise RuntimeError('torch.cat(): expected a non-empty list of Tensors (empty batch)')
        if Cc != self.C:
            raise ValueError('Not valid input channels')
        if H % 8 or W % 8 or (H * W) % 256:
            raise ValueError('Not valid image size {}x{} for the fused path'.format(H, W))
        T = int(num_iter)
        if plan is None:
            if label_host is None:
                label_host = label.detach().cpu().numpy()      # the one host sync per forward
            pad, total_pad = pads_for(H, W, save_for_backward)
            plan = BatchPlan(label_host, self.M, dev, pad=pad, total_pad=total_pad,
                             hws=((H // 2) * (W // 2), (H // 4) * (W // 4), (H // 8) * (W // 8)))
        Bp = plan.B_pad
        stream = torch.cuda.current_stream(dev).cuda_stream
        self.launches = 0
        group_iters = None
        if node_iters is not None:
            ni = np.asarray(node_iters, dtype=np.int32).reshape(-1)
            if ni.size != self.M or ni.min() < 0:
                raise ValueError('Not valid node_iters')
            group_iters = torch.from_numpy(ni).to(dev)
        keep = save_for_backward
        want_z = want_z or keep
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
        self._repack = bool(save_for_backward or self._stale or self._seen_train != self.train_count[0])
        for (name, cin, co, div, io, oo, nmode, nf32) in ENC_LAYERS + DEC_LAYERS:
            is_enc = name[0] == 'E'
            prefs = enc_pref if is_enc else dec_pref
            idx = ENC_IDX[name] if is_enc else DEC_IDX[name]
            h, w = H // div, W // div
            n_mtiles = Bp * h * w // 128
            ly = _Layer()
            ly.name, ly.cin, ly.co, ly.h, ly.w = name, cin, co, h, w
            ly.next_mode, ly.next_f32 = nmode, nf32
            ly.in_order, ly.out_order, ly.is_enc, ly.idx, ly.prefs = io, oo, is_enc, idx, prefs
            ly.groups = len(prefs)
            self._active_groups(ly, plan)
            ly.pair = self._use_pair(n_mtiles, h * w, co, ly.gemm_groups, Bp)
            ly.block_n = 256 if ly.pair else self._block_n(n_mtiles, co)
            ly.hi, ly.lo, ly.inv = self._packed(name, [lstm_w(p, idx) for p in prefs], cin, co, io, oo,
                                                   ly.block_n, stream, needed=ly.active)[:3]
            self._plan_units(ly, Bp, plan, dev)
            layers.append(ly)

        self._repack = False
        self._stale = bool(save_for_backward)
        self._last_layers = layers      # (tests / tools: which layers ran as CTA pairs)
        if save_for_backward:
            self.train_count[0] += 1
        self._seen_train = self.train_count[0]
        if pack_only:
            return None
        A = self._get_arena(Bp, H, W, dev, T if save_for_backward else 0)

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
                a.d3_next = P(A['D3']['next'][t if keep else 0][0])
                a.w_d4 = P(w_d4)
                a.recon_out = recon[t].data_ptr()
                a.recon_prev = recon[t - 1].data_ptr() if mode == B.TAIL_STEP else None
                a.loss_acc = loss_acc[t].data_ptr()
            a.w_e0, a.b_e0, a.img, a.perm = P(w_e0), P(b_e0), P(img), P(plan.perm)
            if not last:
                e1 = A['e1_in'][(t + 1) if keep else 0]
                a.e1_in_hi, a.e1_in_lo = P(e1[0]), P(e1[1])
            a.group_img_end = P(plan.group_img_end)
            a.group_iters, a.iter = P(group_iters), max(t, 0)
            a.B_pad, a.C, a.H, a.W, a.mode = Bp, Cc, H, W, mode
            a.loss_kind = B.LOSS_KINDS[loss_kind]
            a.n_enc_groups, a.n_dec_groups = self.M, self.Md
            a.precise, a.split = self.precise, self.split
            self._timed('tail_head', 0.0, lambda: B.check(L.drasic_tail_head_fwd(a, stream), 'tail_head_fwd'))
            self.launches += 1

        def gate(ly, x_planes, t):
            Lb = A[ly.name]
            a = B.ConvLSTMArgs()
            if keep:
                cur, prv = Lb['h'][t], Lb['h'][t - 1] if t > 0 else Lb['h'][t]
                a.c_state = P(Lb['c'][t])
                a.c_prev = P(Lb['c'][t - 1]) if t > 0 else None
                a.gates_out = P(Lb['gates'][t])
                a.gates_f32 = self.gsplit
                nxt = Lb['next'][t]
            else:
                cur, prv = Lb['h'][t & 1], Lb['h'][(t & 1) ^ 1]
                a.c_state = P(Lb['c'][0])
                nxt = Lb['next'][0]
            a.x_hi, a.x_lo = P(x_planes[0]), P(x_planes[1])
            if t > 0:
                a.h_hi, a.h_lo = P(prv[0]), P(prv[1])
            wplane = ly.g0 * 4 * ly.co * 9 * (ly.cin + ly.co) * 2        # (a collapsed layer reads its one node's planes)
            a.w_hi, a.w_lo, a.w_inv_scale = P(ly.hi) + wplane, (P(ly.lo) + wplane) if ly.lo is not None else None, P(ly.inv) + 4 * ly.g0
            a.h_out_hi, a.h_out_lo = P(cur[0]), P(cur[1])
            a.next_hi, a.next_lo = P(nxt[0]), P(nxt[1])
            a.group_mtile_start = ly.units[0].data_ptr() if ly.units is not None else None
            a.group_mtile_end = ly.units[1].data_ptr() if ly.units is not None else None
            a.group_unit_end = ly.units[2].data_ptr() if ly.units is not None else None
            a.group_img_end, a.n_units = P(plan.group_img_end), ly.n_units
            a.pos_order, a.pos_imgs, a.pos_pairs = P(ly.pos_order), ly.pos_imgs, ly.pos_pairs
            if self._use_stream_k(ly, 4 * ly.co // ly.block_n):
                a.stream_k, (sk_s, sk_f) = 1, self._sk_buffers(dev)
                a.sk_scratch, a.sk_flags = P(sk_s), P(sk_f)
            a.B_pad, a.H, a.W, a.Cin, a.Co = Bp, ly.h, ly.w, ly.cin, ly.co
            a.n_groups = ly.gemm_groups
            a.has_state = 1 if t > 0 else 0
            a.next_mode, a.next_f32 = ly.next_mode, ly.next_f32
            a.split, a.precise, a.block_n = self.split, self.precise, ly.block_n
            a.cta_pair = int(ly.pair)
            # algorithmic FLOPs of this launch: real images only, hidden conv skipped on zero state
            k = 9 * (ly.cin + (ly.co if t > 0 else 0))
            flops = 2.0 * Bn * ly.h * ly.w * 4 * ly.co * k
            self._timed('gates_' + ly.name, flops, lambda: B.check(L.drasic_convlstm_fwd(a, stream), 'convlstm_fwd'))
            self.launches += 1
            return nxt

        tail(B.TAIL_INIT, -1)
        for t in range(T):
            x = A['e1_in'][t if keep else 0]
            for ly in layers[:3]:
                x = gate(ly, x, t)
            a = B.BottleneckArgs()
            a.h_e3, a.w_e4, a.w_d0 = P(x[0]), P(w_e4), P(w_d0)
            a.noise = noise[t].data_ptr() if noise is not None else None
            a.perm = P(plan.perm)
            a.code_out = codes[t].data_ptr()
            a.z_out = zbuf[t].data_ptr() if zbuf is not None else None
            a.bits_out = bits[t].data_ptr()
            d1 = A['d1_in'][t if keep else 0]
            a.d1_in_hi, a.d1_in_lo = P(d1[0]), P(d1[1])
            a.group_img_end = P(plan.group_img_end)
            a.B_pad, a.HW, a.Ch, a.code = Bp, hc * wc, 128, code
            a.n_enc_groups, a.n_dec_groups = self.M, self.Md
            a.precise, a.split = self.precise, self.split
            self._timed('bottleneck', 0.0, lambda a=a: B.check(L.drasic_bottleneck_fwd(a, stream), 'bottleneck_fwd'))
            self.launches += 1
            x = d1
            for ly in layers[3:]:
                x = gate(ly, x, t)
            tail(B.TAIL_FIRST if t == 0 else B.TAIL_STEP, t)
        out = {'recon': recon, 'codes': codes, 'bits': bits, 'loss_acc': loss_acc, 'plan': plan}
        if zbuf is not None:
            out['z'] = zbuf
        # keep the small stacked weights alive until the stream has consumed them
        out['_keep'] = (w_e0, b_e0, w_e4, w_d0, w_d4, noise, img)
        if keep:
            self._gen = getattr(self, '_gen', 0) + 1
            out['_bwd'] = dict(gen=self._gen, A=A, layers=layers, sd=sd, img=img, T=T, loss_kind=loss_kind, w_e0=w_e0, b_e0=b_e0,
                               w_e4=w_e4, w_d0=w_d0, w_d4=w_d4, enc_pref=enc_pref, dec_pref=dec_pref,
                               shape=(Bn, Cc, H, W), group_iters=group_iters)
        return out

    # ------------------------------------------------------------------ decode-only
    def _decode(self, sd, bits, label, hw, label_host=None):
        """Reconstruct from the packed bitstream alone: bits [T', B, code*H*W/512] uint8 (the 'bits' output of
        forward(), or any prefix of its iterations -- the codec is progressive), label [B] node ids (which
        decoder a sample belongs to in the sep codec; unused by the shared decoder but kept for symmetry).
        Runs only the decoder half of the loop (dist_codec.py:55-57): D0 -> D1 D2 D3 -> D4, recon += decoded.
        Returns dict(recon [T',B,C,H,W], codes [T',B,code,H/8,W/8])"""
        if not bits.is_cuda:
            raise B.DrasicError('DRASIC B200 engine needs CUDA tensors; there is no CPU path')
        L = B.lib()
        dev = bits.device
        H, W = hw
        if H % 8 or W % 8 or (H * W) % 256:
            raise ValueError('Not valid image size {}x{} for the fused path'.format(H, W))
        code, hc, wc = self.code, H // 8, W // 8
        bits = bits.detach().contiguous()
        if bits.dtype != torch.uint8 or bits.dim() != 3 or bits.shape[2] != code * hc * wc // 8:
            raise ValueError('Not valid bitstream shape')
        T, Bn = int(bits.shape[0]), int(bits.shape[1])
        if label_host is None:
            label_host = label.detach().cpu().numpy()
        plan = BatchPlan(label_host, self.M, dev, pad=1, total_pad=pads_for(H, W, False)[1],
                         hws=(hc * wc * 16, hc * wc * 4, hc * wc))
        Bp = plan.B_pad
        stream = torch.cuda.current_stream(dev).cuda_stream
        self.launches = 0
        A = self._get_arena(Bp, H, W, dev, 0)
        dec_pref = ['model.decoder.{}'.format(j) for j in range(self.M)] if self.separate else ['model.decoder']

        def lstm_w(prefix, idx):
            base = '{}.{}.cell.cell.0.'.format(prefix, idx)
            return sd[base + 'in.cell.cell.main.weight'], sd[base + 'hidden.cell.cell.main.weight']

        def stack(keys):
            return torch.stack([sd[k].detach().float().reshape(sd[k].shape[0], -1) for k in keys]).contiguous()
        layers = []
        self._repack = bool(self._stale or self._seen_train != self.train_count[0])
        for (name, cin, co, div, io, oo, nmode, nf32) in DEC_LAYERS:
            h, w = H // div, W // div
            n_mtiles = Bp * h * w // 128
            ly = _Layer()
            ly.name, ly.cin, ly.co, ly.h, ly.w, ly.next_mode, ly.next_f32 = name, cin, co, h, w, nmode, nf32
            ly.groups = len(dec_pref)
            self._active_groups(ly, plan)
            ly.pair = self._use_pair(n_mtiles, h * w, co, ly.gemm_groups, Bp)
            ly.block_n = 256 if ly.pair else self._block_n(n_mtiles, co)
            ly.hi, ly.lo, ly.inv = self._packed(name, [lstm_w(p, DEC_IDX[name]) for p in dec_pref], cin, co, io, oo,
                                                   ly.block_n, stream, needed=ly.active)[:3]
            self._plan_units(ly, Bp, plan, dev)
            layers.append(ly)
        self._repack = False      # (the encoder planes stay stale: _stale is cleared only by a full forward)
        w_d0 = stack([p + '.0.cell.cell.main.weight' for p in dec_pref])
        w_d4 = stack([p + '.7.cell.cell.main.weight' for p in dec_pref])
        recon = torch.empty((T, Bn, self.C, H, W), dtype=torch.float32, device=dev)
        codes = torch.empty((T, Bn, code, hc, wc), dtype=torch.float32, device=dev)
        P = B.ptr
        for t in range(T):
            a = B.BottleneckArgs()
            a.bits_in, a.w_d0, a.perm = bits[t].data_ptr(), P(w_d0), P(plan.perm)
            a.code_out = codes[t].data_ptr()
            d1 = A['d1_in'][0]
            a.d1_in_hi, a.d1_in_lo = P(d1[0]), P(d1[1])
            a.group_img_end = P(plan.group_img_end)
            a.B_pad, a.HW, a.Ch, a.code = Bp, hc * wc, 128, code
            a.n_enc_groups, a.n_dec_groups = self.M, self.Md
            a.precise, a.split = self.precise, self.split
            B.check(L.drasic_bottleneck_fwd(a, stream), 'bottleneck_fwd (decode)')
            x = d1
            for ly in layers:
                Lb = A[ly.name]
                g = B.ConvLSTMArgs()
                cur, prv = Lb['h'][t & 1], Lb['h'][(t & 1) ^ 1]
                g.c_state = P(Lb['c'][0])
                nxt = Lb['next'][0]
                g.x_hi, g.x_lo = P(x[0]), P(x[1])
                if t > 0:
                    g.h_hi, g.h_lo = P(prv[0]), P(prv[1])
                wplane = ly.g0 * 4 * ly.co * 9 * (ly.cin + ly.co) * 2
                g.w_hi, g.w_lo, g.w_inv_scale = P(ly.hi) + wplane, P(ly.lo) + wplane, P(ly.inv) + 4 * ly.g0
                g.h_out_hi, g.h_out_lo = P(cur[0]), P(cur[1])
                g.next_hi, g.next_lo = P(nxt[0]), P(nxt[1])
                g.group_mtile_start = ly.units[0].data_ptr() if ly.units is not None else None
                g.group_mtile_end = ly.units[1].data_ptr() if ly.units is not None else None
                g.group_unit_end = ly.units[2].data_ptr() if ly.units is not None else None
                g.group_img_end, g.n_units = P(plan.group_img_end), ly.n_units
                g.pos_order, g.pos_imgs, g.pos_pairs = P(ly.pos_order), ly.pos_imgs, ly.pos_pairs
                if self._use_stream_k(ly, 4 * ly.co // ly.block_n):
                    g.stream_k, (sk_s, sk_f) = 1, self._sk_buffers(dev)
                    g.sk_scratch, g.sk_flags = P(sk_s), P(sk_f)
                g.B_pad, g.H, g.W, g.Cin, g.Co = Bp, ly.h, ly.w, ly.cin, ly.co
                g.n_groups, g.has_state = ly.gemm_groups, 1 if t > 0 else 0
                g.next_mode, g.next_f32 = ly.next_mode, ly.next_f32
                g.split, g.precise, g.block_n, g.cta_pair = self.split, self.precise, ly.block_n, int(ly.pair)
                B.check(L.drasic_convlstm_fwd(g, stream), 'convlstm_fwd (decode)')
                x = nxt
            th = B.TailHeadArgs()
            th.d3_next, th.w_d4 = P(A['D3']['next'][0][0]), P(w_d4)
            th.recon_out = recon[t].data_ptr()
            th.recon_prev = recon[t - 1].data_ptr() if t > 0 else None
            th.perm, th.group_img_end = P(plan.perm), P(plan.group_img_end)
            th.B_pad, th.C, th.H, th.W = Bp, self.C, H, W
            th.mode = B.TAIL_FIRST if t == 0 else B.TAIL_STEP
            th.n_enc_groups, th.n_dec_groups = self.M, self.Md
            th.precise, th.split = self.precise, self.split
            B.check(L.drasic_tail_head_fwd(th, stream), 'tail_head_fwd (decode)')
            self.launches += 5
        return {'recon': recon, 'codes': codes, 'plan': plan, '_keep': (w_d0, w_d4, bits)}

    # ------------------------------------------------------------------ backward through time
    def _packed_dgrad(self, ly, sd, stream):
        def lstm_w(prefix, idx):
            base = '{}.{}.cell.cell.0.'.format(prefix, idx)
            return sd[base + 'in.cell.cell.main.weight'], sd[base + 'hidden.cell.cell.main.weight']
        weights = [lstm_w(p, ly.idx) for p in ly.prefs]
        key = tuple((w.data_ptr(), w._version) for pair in weights for w in pair)
        hit = self._wcache.get('dgrad_' + ly.name)
        need = set(range(len(weights))) if ly.active is None else set(ly.active)
        if hit is not None and hit[0] == key and not (self.force_pack or self._repack) and need <= hit[2]:
            return hit[1][:3]
        L = B.lib()
        dev = weights[0][0].device
        G, K = len(weights), 36 * ly.co
        n_pad = (ly.cin + ly.co + 255) // 256 * 256            # a group's plane holds whole 256-row N tiles
        if hit is not None:
            hi, lo, inv, scratch = hit[1]
        else:
            hi = torch.zeros((G * n_pad, K), dtype=torch.float16, device=dev)
            lo = torch.zeros((G * n_pad, K), dtype=torch.float16, device=dev)
            inv = torch.empty((G,), dtype=torch.float32, device=dev)
            scratch = torch.zeros((G,), dtype=torch.int32, device=dev)
        todo = [(g, w_in.detach().contiguous(), w_h.detach().contiguous()) for g, (w_in, w_h) in enumerate(weights) if g in need]
        if todo:
            a_in, a_h, slots, n = B.pointer_arrays(todo)
            self._timed('pack_dgrad_weights', 0.0, lambda: B.check(
                L.drasic_pack_dgrad_weights_batch(a_in, a_h, slots, n, G, ly.cin, ly.co, ly.in_order, ly.out_order,
                                                  hi.data_ptr(), lo.data_ptr(), inv.data_ptr(), scratch.data_ptr(), stream),
                'pack_dgrad_weights'))
            self.launches += 2
        self._wcache['dgrad_' + ly.name] = (key, (hi, lo, inv, scratch), need)
        return hi, lo, inv

    def _bwd_buffers(self, layers, Bp, H, W, dev):
        key = (Bp, H, W, str(dev))
        if getattr(self, '_bwd_key', None) == key:
            return self._bwd_buf
        self._bwd_buf = None
        f32 = dict(dtype=torch.float32, device=dev)
        f16 = dict(dtype=torch.float16, device=dev)
        W_ = {}
        for ly in layers:
            n = (Bp, ly.h, ly.w, ly.co)
            W_[ly.name] = dict(dh_above=torch.zeros(n, **f32), dh_rec=torch.zeros(n, **f32),
                               dc=[torch.zeros(n, **f32), torch.zeros(n, **f32)],     # ping-pong: in (t+1) / out (t)
                               # gate-gradient planes and their scale ping-pong by iteration parity: wgrad of
                               # iteration t runs on a side stream while the main stream already produces t-1
                               dg_hi=[torch.zeros((Bp, ly.h, ly.w, 4 * ly.co), **f16) for _ in range(2)],
                               dg_lo=[torch.zeros((Bp, ly.h, ly.w, 4 * ly.co), **f16) if self.gsplit else None for _ in range(2)],
                               last_amax=torch.zeros((1,), dtype=torch.int32, device=dev),   # carried across steps
                               inv=[torch.ones((1,), **f32) for _ in range(2)])
        W_['du1'] = torch.zeros((Bp, H // 2, W // 2, 128), **f32)
        W_['dd0'] = torch.zeros((Bp, H // 8, W // 8, 128), **f32)
        self._bwd_buf, self._bwd_key = W_, key
        return W_

    def _backward(self, out, g_loss=None):
        """Backward-through-time for the forward that produced `out` (save_for_backward=True).
        Returns {state_dict key: fp32 gradient} for every parameter of the codec."""
        ctx = out['_bwd']
        if ctx['gen'] != getattr(self, '_gen', 0):
            raise RuntimeError('backward() must follow the forward() that saved its state: the per-iteration '
                               'arena of this engine has been overwritten by a later forward')
        plan = out['plan']
        A, layers, sd, img, T = ctx['A'], ctx['layers'], ctx['sd'], ctx['img'], ctx['T']
        Bn, Cc, H, W = ctx['shape']
        L = B.lib()
        dev = img.device
        stream = torch.cuda.current_stream(dev).cuda_stream
        Bp = plan.B_pad
        P = B.ptr
        code, hc, wc = self.code, H // 8, W // 8
        WB = self._bwd_buffers(layers, Bp, H, W, dev)
        f32 = dict(dtype=torch.float32, device=dev)
        gacc = torch.empty((Bn, Cc, H, W), **f32)
        acc_dtype = torch.int64 if self.deterministic else torch.float32     # (deterministic: fixed point, see below)
        dw_e0 = torch.zeros_like(ctx['w_e0'], dtype=acc_dtype)
        db_e0 = torch.zeros_like(ctx['b_e0'], dtype=acc_dtype) if ctx['b_e0'] is not None else None
        dw_e4, dw_d0, dw_d4 = (torch.zeros_like(ctx[k], dtype=acc_dtype) for k in ('w_e4', 'w_d0', 'w_d4'))
        gw = {ly.name: torch.zeros((ly.groups, 4 * ly.co, 9 * (ly.cin + ly.co)), **f32) for ly in layers}
        self._repack = True      # once per step; parameter versions cannot be trusted (fused optimizers)
        dpk = {ly.name: self._packed_dgrad(ly, sd, stream) for ly in layers}
        self._repack = False
        by = {ly.name: ly for ly in layers}
        sms = self.sms()
        # true amax (float bits) of every (layer, iteration) gate-gradient tensor of this backward pass
        amax_tab = {ly.name: torch.zeros((T,), dtype=torch.int32, device=dev) for ly in layers}
        no_amax = torch.zeros((1,), dtype=torch.int32, device=dev)
        # The weight-gradient GEMMs are off the critical path of backward-through-time (nothing but the final
        # unpack reads their result): they go to a side stream, so the HBM-bound cell / tail / head backward
        # kernels of the main stream run next to them instead of between the GEMMs.
        main = torch.cuda.current_stream(dev)
        overlap = self.overlap_wgrad and self.profile is None
        side = None
        if overlap:
            if getattr(self, '_side', None) is None or self._side.device != dev:
                self._side = torch.cuda.Stream(device=dev)
            side = self._side
            side.wait_stream(main)          # gw and the buffers above are initialised on the main stream
        wgrad_done = {}                     # (layer, t) -> event recorded on the side stream after wgrad(layer, t)

        def group_kb_end(ly):
            return plan.kb_end(ly.h * ly.w)

        # where each layer's input lives / where its input gradient goes
        x_of = {'E1': lambda t: A['e1_in'][t], 'E2': lambda t: A['E1']['next'][t], 'E3': lambda t: A['E2']['next'][t],
                'D1': lambda t: A['d1_in'][t], 'D2': lambda t: A['D1']['next'][t], 'D3': lambda t: A['D2']['next'][t]}
        dx_of = {'E1': (WB['du1'], B.DX_SAME), 'E2': (WB['E1']['dh_above'], B.DX_INV_DOWN),
                 'E3': (WB['E2']['dh_above'], B.DX_INV_DOWN), 'D1': (WB['dd0'], B.DX_SAME),
                 'D2': (WB['D1']['dh_above'], B.DX_INV_UP), 'D3': (WB['D2']['dh_above'], B.DX_INV_UP)}

        def lstm_layer_bwd(ly, t):
            Lb, wb = A[ly.name], WB[ly.name]
            n_pix = Bp * ly.h * ly.w
            # fused cell backward -> scaled fp16 planes; scale predicted from the previously processed iteration
            am = amax_tab[ly.name]
            # (deterministic mode: no memory of the previous step -- the first iteration processed starts from "no
            # prediction" and takes the exact scale through the verify pass, so the bits depend on this step's data only)
            prev = am[t + 1:t + 2] if t < T - 1 else (no_amax if self.deterministic else wb['last_amax'])
            f = B.LstmBwdFusedArgs()
            f.gates, f.gates_f32 = P(Lb['gates'][t]), self.gsplit
            f.c_t, f.c_prev = P(Lb['c'][t]), P(Lb['c'][t - 1]) if t > 0 else None
            f.dh_above, f.dh_rec = P(wb['dh_above']), P(wb['dh_rec']) if t < T - 1 else None
            f.dc_in = P(wb['dc'][(t + 1) & 1]) if t < T - 1 else None
            f.dc_out = P(wb['dc'][t & 1])
            pp = t & 1
            dg_hi, dg_lo, inv_dg = wb['dg_hi'][pp], wb['dg_lo'][pp], wb['inv'][pp]
            ev = wgrad_done.pop((ly.name, t + 2), None)
            if ev is not None:
                main.wait_event(ev)          # the planes of this parity were last read by wgrad(ly, t+2)
            f.dg_hi, f.dg_lo = P(dg_hi), P(dg_lo)
            f.prev_amax_bits, f.amax_bits, f.inv_scale_out = prev.data_ptr(), am[t:t + 1].data_ptr(), P(inv_dg)
            f.n_pix, f.Co = n_pix, ly.co
            for verify in (0, 1):
                f.verify = verify
                self._timed('lstm_bwd' if not verify else 'lstm_bwd_verify', 0.0,
                            lambda: B.check(L.drasic_lstm_bwd_fused(f, stream), 'lstm_bwd_fused'))
            planes_ready = None
            if overlap:
                planes_ready = torch.cuda.Event()
                planes_ready.record(main)
            hi, lo, inv = dpk[ly.name]
            a = B.DgradArgs()
            dplane = ly.g0 * ((ly.cin + ly.co + 255) // 256 * 256) * 36 * ly.co * 2
            a.dg_hi, a.dg_lo, a.w_hi, a.w_lo = P(dg_hi), P(dg_lo), P(hi) + dplane, P(lo) + dplane
            a.w_inv_scale, a.dg_inv_scale = P(inv) + 4 * ly.g0, P(inv_dg)
            a.dx_out = P(dx_of[ly.name][0])
            a.dhrec_out = P(wb['dh_rec']) if t > 0 else None
            a.group_mtile_start = ly.units[0].data_ptr() if ly.units is not None else None
            a.group_mtile_end = ly.units[1].data_ptr() if ly.units is not None else None
            a.group_unit_end = ly.units[2].data_ptr() if ly.units is not None else None
            a.group_img_end, a.n_units = P(plan.group_img_end), ly.n_units
            a.pos_order, a.pos_imgs, a.pos_pairs = P(ly.pos_order), ly.pos_imgs, ly.pos_pairs
            # (not next to the weight-gradient GEMMs of the side stream: stream-K workers wait for one another, and
            # would do so holding SMs the other stream's kernels are queued for)
            if not overlap and self._use_stream_k(ly, (ly.cin + (ly.co if t > 0 else 0) + 255) // 256):
                a.stream_k, (sk_s, sk_f) = 1, self._sk_buffers(dev)
                a.sk_scratch, a.sk_flags = P(sk_s), P(sk_f)
            a.B_pad, a.H, a.W, a.Cin, a.Co = Bp, ly.h, ly.w, ly.cin, ly.co
            a.n_groups, a.dx_mode, a.split = ly.gemm_groups, dx_of[ly.name][1], self.gsplit
            a.cta_pair = int(ly.pair)
            fl = 2.0 * Bn * ly.h * ly.w * 36 * ly.co * (ly.cin + (ly.co if t > 0 else 0))
            self._timed('dgrad_' + ly.name, fl, lambda: B.check(L.drasic_convlstm_bwd_data(a, stream), 'convlstm_bwd_data'))
            w = B.WgradArgs()
            xp = x_of[ly.name](t)
            w.dg_hi, w.dg_lo, w.x_hi, w.x_lo = P(dg_hi), P(dg_lo), P(xp[0]), P(xp[1])
            if t > 0:
                w.h_hi, w.h_lo = P(Lb['h'][t - 1][0]), P(Lb['h'][t - 1][1])
            # (a per-node layer whose samples all sit on one node accumulates straight into that node's gradient plane)
            w.gw, w.dg_inv_scale = P(gw[ly.name]) + ly.g0 * 4 * ly.co * 9 * (ly.cin + ly.co) * 4, P(inv_dg)
            w.group_kb_end = P(group_kb_end(ly)) if ly.gemm_groups > 1 else None
            w.B_pad, w.H, w.W, w.Cin, w.Co = Bp, ly.h, ly.w, ly.cin, ly.co
            w.n_groups, w.has_h, w.split = ly.gemm_groups, 1 if t > 0 else 0, self.gsplit
            # (groups that hold samples: the tiles of a node without samples find an empty K range and return at once)
            n_act = 1 if ly.gemm_groups == 1 else (len(ly.active) if ly.active else ly.groups)
            kb_per_group = max(1, Bp * ly.h * ly.w // 64 // n_act)
            # 4Co/128 M tiles: always even.  512-column pair tiles (one TMEM-wide accumulator, no epilogue overlap)
            # only where a tile's K loop is long enough to amortise the exposed epilogue
            w.cta_pair = 0 if self.cta_pair == 'off' else (self.wgrad_pair_mode if kb_per_group >= 256 else 1)
            upt = 8 if w.cta_pair == 2 else 4
            tiles = n_act * (4 * ly.co // 128) * ((9 * ly.cin // 64 + upt - 1) // upt + ((9 * ly.co // 64 + upt - 1) // upt if t > 0 else 0))
            units = sms
            if w.cta_pair:
                tiles, units = tiles // 2, sms // 2
            # a K split costs its own accumulate pass over the tile (a 128 KB red.add per CTA): every split keeps at
            # least wgrad_min_kb K blocks (256 x 64 pixels), which at small batch means no split at all -- the wgrad GEMMs
            # run next to the main stream's kernels, where their memory traffic counts for more than their own latency
            # (measured, training img/s at B=128 / B=1024 for 16, 64, 256 K blocks: 2 951 / 4 438, 2 997 / 4 435,
            # 3 110 / 4 424; the stand-alone wgrad rate at B=1024 drops from 1 245 to 1 169 TFLOP/s)
            w.k_splits = 1 if self.deterministic else _pick_k_splits(tiles, units, max(1, kb_per_group // self.wgrad_min_kb))
            if overlap:
                side.wait_event(planes_ready)
                B.check(L.drasic_convlstm_bwd_weight(w, side.cuda_stream), 'convlstm_bwd_weight')
                done = torch.cuda.Event()
                done.record(side)
                wgrad_done[(ly.name, t)] = done
            else:
                self._timed('wgrad_' + ly.name, fl, lambda: B.check(L.drasic_convlstm_bwd_weight(w, stream), 'convlstm_bwd_weight'))
            self.launches += 4

        loss_scale = 1.0 / (T * float(Bn * Cc * H * W))
        # fixed-point unit of the deterministic accumulators: 2^-30 of the loss scale (block partial sums are O(1e-3..1e3)
        # loss scales: ~1e-9 relative resolution, 2^62 / 2^30 = 4e9 loss scales of head room)
        fix_scale = float(2 ** 30) / loss_scale if self.deterministic else 0.0
        for t in range(T - 1, -1, -1):
            a = B.TailBwdArgs()
            a.d3_next, a.w_d4, a.img = P(A['D3']['next'][t][0]), P(ctx['w_d4']), P(img)
            a.recon_t, a.gacc, a.dh_d3, a.dw_d4 = out['recon'][t].data_ptr(), P(gacc), P(WB['D3']['dh_above']), P(dw_d4)
            a.perm, a.group_img_end = P(plan.perm), P(plan.group_img_end)
            a.loss_scale = loss_scale
            a.B_pad, a.C, a.H, a.W = Bp, Cc, H, W
            a.first, a.loss_kind, a.n_dec_groups, a.has_acc = int(t == 0), B.LOSS_KINDS[ctx['loss_kind']], self.Md, int(t < T - 1)
            a.group_iters, a.iter, a.n_enc_groups = P(ctx['group_iters']), t, self.M
            a.fix_scale = fix_scale
            self._timed('tail_bwd', 0.0, lambda: B.check(L.drasic_tail_bwd(a, stream), 'tail_bwd'))
            for name in ('D3', 'D2', 'D1'):
                lstm_layer_bwd(by[name], t)
            b = B.BottleneckBwdArgs()
            b.dd0, b.h_e3 = P(WB['dd0']), P(A['E3']['next'][t][0])
            b.z, b.codes = out['z'][t].data_ptr(), out['codes'][t].data_ptr()
            b.w_e4, b.w_d0, b.perm, b.group_img_end = P(ctx['w_e4']), P(ctx['w_d0']), P(plan.perm), P(plan.group_img_end)
            b.dh_e3, b.dw_e4, b.dw_d0 = P(WB['E3']['dh_above']), P(dw_e4), P(dw_d0)
            b.B_pad, b.HW, b.Ch, b.code, b.n_enc_groups, b.n_dec_groups = Bp, hc * wc, 128, code, self.M, self.Md
            b.fix_scale = fix_scale
            self._timed('bottleneck_bwd', 0.0, lambda: B.check(L.drasic_bottleneck_bwd(b, stream), 'bottleneck_bwd'))
            for name in ('E3', 'E2', 'E1'):
                lstm_layer_bwd(by[name], t)
            h = B.HeadBwdArgs()
            h.du1, h.w_e0, h.b_e0, h.img = P(WB['du1']), P(ctx['w_e0']), P(ctx['b_e0']), P(img)
            h.recon_prev = out['recon'][t - 1].data_ptr() if t > 0 else None
            h.gacc, h.dw_e0, h.db_e0 = P(gacc), P(dw_e0), P(db_e0)
            h.perm, h.group_img_end = P(plan.perm), P(plan.group_img_end)
            h.B_pad, h.C, h.H, h.W = Bp, Cc, H, W
            h.first, h.propagate, h.n_enc_groups = int(t == 0), int((not self.separate) and t > 0), self.M
            h.fix_scale = fix_scale
            self._timed('head_bwd', 0.0, lambda: B.check(L.drasic_head_bwd(h, stream), 'head_bwd'))
            self.launches += 3

        if overlap:
            main.wait_stream(side)          # join: the unpack below reads the accumulated weight gradients
        if self.deterministic:              # fixed point -> fp32
            dw_e0, dw_e4, dw_d0, dw_d4 = ((v.double() / fix_scale).float() for v in (dw_e0, dw_e4, dw_d0, dw_d4))
            db_e0 = (db_e0.double() / fix_scale).float() if db_e0 is not None else None
        for ly in layers:
            WB[ly.name]['last_amax'].copy_(amax_tab[ly.name][T - 1:T])
        # Every gradient is a view of ONE flat fp32 buffer laid out in the state_dict's own order (model.encoder.* then
        # model.decoder.*): the all-reduce of a multi-GPU step is then a single in-place collective on the buffer --
        # or on the contiguous slice that holds the joint decoder when the batch is sharded by source (engine/dist.py).
        # Gradients of nodes without samples are zero, like autograd's for an unused encoder would be None -> zero.
        offs, total = {}, 0
        for k, v in sd.items():
            offs[k] = total
            total += v.numel()
        flat = torch.zeros((total,), **f32)

        def view(k):
            v = sd[k]
            return flat[offs[k]:offs[k] + v.numel()].view(v.shape)
        grads = {k: view(k) for k in sd}
        for ly in layers:
            for g, pref in enumerate(ly.prefs):
                if ly.active is not None and g not in ly.active:
                    continue                                          # no samples: the zeros of `flat` stand
                base = '{}.{}.cell.cell.0.'.format(pref, ly.idx)
                dwi, dwh = grads[base + 'in.cell.cell.main.weight'], grads[base + 'hidden.cell.cell.main.weight']
                self._timed('unpack_wgrad', 0.0, lambda: B.check(
                    L.drasic_unpack_wgrad(gw[ly.name][g].data_ptr(), ly.cin, ly.co, ly.in_order, ly.out_order,
                                          dwi.data_ptr(), dwh.data_ptr(), stream), 'unpack_wgrad'))
                self.launches += 1
        dst, src = [], []
        for j, pref in enumerate(ctx['enc_pref']):
            dst.append(grads[pref + '.0.cell.cell.main.weight']); src.append(dw_e0[j].reshape(dst[-1].shape))
            if db_e0 is not None:
                dst.append(grads[pref + '.0.cell.cell.main.bias']); src.append(db_e0[j].reshape(dst[-1].shape))
            dst.append(grads[pref + '.7.cell.cell.main.weight']); src.append(dw_e4[j].reshape(dst[-1].shape))
        for j, pref in enumerate(ctx['dec_pref']):
            dst.append(grads[pref + '.0.cell.cell.main.weight']); src.append(dw_d0[j].reshape(dst[-1].shape))
            dst.append(grads[pref + '.7.cell.cell.main.weight']); src.append(dw_d4[j].reshape(dst[-1].shape))
        torch._foreach_copy_(dst, src)
        if g_loss is not None:
            flat.mul_(g_loss)                                      # chain rule for d(anything)/d(loss), in place
        self.grad_flat, self.grad_offsets = flat, offs             # (engine/dist.py reduces the buffer in place)
        return grads

    @staticmethod
    def loss_from_acc(loss_acc, numel):
        """mean over T of the per-iteration mean loss (dist_codec.py:58,61)"""
        T = loss_acc.shape[0]
        return (loss_acc[:, 0].sum() / (float(numel) * T)).float()


class GraphedCodec(object):
    """The whole T-iteration loop -- and, for training, backward-through-time -- captured once as a CUDA
    graph and replayed per batch.  Inputs live in static buffers (image batch, node-sort metadata of
    worst-case size); weights are re-packed from the fp32 master parameters inside the graph, so
    optimizer steps between replays are honoured; recurrent state never leaves the device.
    Replay cost on the host: two small H2D copies and one cudaGraphLaunch."""

    def __init__(self, engine, sd, batch, hw, num_iter, training=False, with_backward=False, loss_kind='mae',
                 device=None):
        # The graph bakes in the addresses of the arena, the backward work buffers and the packed weight planes.  The
        # engine keeps one slot of each and replaces it when a call with another batch size / mode comes along, so the
        # graph replays on an engine of its own that nothing else ever calls.
        self.parent = engine
        self.eng, self.sd, self.T = engine.clone(), sd, int(num_iter)
        engine = self.eng
        self.training, self.with_backward, self.loss_kind = training, with_backward, loss_kind
        H, W = hw
        dev = device or next(iter(sd.values())).device
        self.B = int(batch)
        self.img = torch.zeros((self.B, engine.C, H, W), dtype=torch.float32, device=dev)
        self.noise = torch.rand((self.T, self.B, engine.code, H // 8, W // 8), dtype=torch.float32, device=dev) \
            if training else None
        hws = sorted({(H // d) * (W // d) for d in (2, 4, 8)})
        label0 = np.arange(self.B) % engine.M
        # inference plans pack the nodes back to back: no worst-case slack at all
        pad, total_pad = pads_for(H, W, with_backward)
        self.plan = BatchPlan(label0, engine.M, dev, pad_total=worst_case_pad(self.B, engine.M, pad, total_pad),
                              static_hw=hws, pad=pad, total_pad=total_pad)
        self.graph = None
        self.out = self.grads = None
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                          # warm-up: allocates arenas, sets attributes
            for _ in range(2):
                self._run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        # training re-packs the weights inside the graph (they change every step); inference keeps the
        # packed planes outside and refreshes them eagerly only when a parameter version changes
        engine.force_pack = bool(with_backward)
        try:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._run()
        finally:
            engine.force_pack = False
        self.launches = engine.launches

    def _versions(self):
        return tuple(v._version for v in self.sd.values())

    def _refresh_weights(self):
        # parameter versions, plus the family's count of training forwards: fused optimizers update parameters in
        # place without bumping Tensor._version, so any training step since the last pack makes the planes suspect
        if self._versions() == self._seen and self._seen_train == self.eng.train_count[0] and not self.eng._stale:
            return
        # a parameter changed since capture: re-pack eagerly into the buffers the graph reads
        self.eng.force_pack = True
        try:
            with torch.no_grad():
                self.eng.forward(self.sd, self.img, None, 0, training=False, plan=self.plan, pack_only=True)
        finally:
            self.eng.force_pack = False
        self._seen, self._seen_train = self._versions(), self.eng.train_count[0]

    def _run(self):
        self._seen, self._seen_train = self._versions(), self.eng.train_count[0]
        with torch.no_grad():
            self.out = self.eng.forward(self.sd, self.img, None, self.T, training=self.training, noise=self.noise,
                                        loss_kind=self.loss_kind, save_for_backward=self.with_backward,
                                        plan=self.plan)
            self.grads = self.eng.backward(self.out) if self.with_backward else None

    def __call__(self, img, label_host, noise=None):
        """img: [B,C,H,W] tensor (host or device); label_host: host array of node ids; noise: optional
        uniforms for the stochastic binarizer (drawn with torch's generator otherwise)"""
        self.img.copy_(img, non_blocking=True)
        if self.noise is not None:
            if noise is not None:
                self.noise.copy_(noise, non_blocking=True)
            else:
                self.noise.uniform_()
        self.plan.update(label_host)
        if not self.with_backward:
            self._refresh_weights()
        self.graph.replay()
        if self.with_backward:
            self.eng.train_count[0] += 1        # (the capture-time forward counted itself once; replays count here)
        self.eng.launches = self.parent.launches = self.launches
        return self.out, self.grads
