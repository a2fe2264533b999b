"""Benchmark of DRASIC's recurrent residual coding loop on B200 (see BASELINE.json / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--mode all|train|eval]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path (T=16 iterations of encode -> binarize -> decode over the
residual) over one batch of synthetic CIFAR-shaped images from 8 sources.  Rank 0 prints ONE JSON line.

The line's own `value` is the training step (forward + backward-through-time + gradient all-reduce + Adam) of
BASELINE configs[1] at --batch images per GPU.  With the default `--mode all` the same line also carries, each
with its own value / e2e / roofline:
    eval          encode+decode of the same workload (the other half of BASELINE.json's metric)
    b128          training step and encode+decode at 128 images per GPU (SURVEY 8d config 2's batch)
    config1       BASELINE configs[0]: MNIST-shaped single source, eval, batch 64
    config5       BASELINE configs[4]: 768x512 images tiled into 32x32 patches, eval, 4096 patches per batch and GPU
    grad_parity   the training step with fp32-grade (split-fp16) backward GEMMs
    parity        codes / reconstructions / PSNR of this build against the CPU oracle on 256 images (initial weights;
                  parity_after_training: the same on the weights the run's training steps leave behind)
    multi_gpu_parity   (N > 1) all-reduced gradients of a sharded batch against the same batch on one GPU
`--impl reference` times the CPU oracle (the reference's algorithm on torch CPU ops, all host threads)
on a bounded sample of the same workload: the SAME sample `cpu_baseline` of our arm is timed on.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'drasic-distributed-recurrent-autoencoder-for-scalable-image-compression_b200')
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

T_ITER = 16
M_NODES = 8
CHANNELS = 3
FLOP_PER_IMG_ITER = {3: 3567714304.0, 1: 3567452160.0}     # dense forward count per image and iteration
METRIC = 'images/s train & encode+decode at T=16 (CIFAR-10 32x32), 1/2/4/8 B200; PSNR parity'
# one CPU sample per mode, used by BOTH `cpu_baseline` (our arm) and `--impl reference`
CPU_SAMPLE = {'train': 64, 'eval': 256}
PARITY_SAMPLE = 256
PDL_ON = False
DETERMINISTIC = False     # --deterministic: bit-reproducible gradients (single-writer wgrad, fixed-point pointwise accumulation)
STREAM_K = 'auto'          # --stream-k: schedule of the gate / dgrad GEMMs (auto: stream-K for launches of few work items per SM)


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('bf16_tflops_sustained', 1401.4), d.get('bf16_tflops', 1679.4), d.get('hbm_gbs', 6544.3), 'measured'
    return 1400.0, 1590.0, 6650.0, 'fallback'


class ClockSampler(object):
    """samples nvidia-smi clocks / throttle reasons during the timed region"""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            f = [x.strip() for x in r.split(',')]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        sm.sort()
        # median over the upper half of the samples = clocks under load (the first samples are idle)
        load = sm[len(sm) // 2:] if sm else []
        return {'sm_mhz': load[len(load) // 2] if load else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


def captured_traffic(mode, B, nodes, precision):
    """DRAM bytes per launch of the dominant kernel (forward gate GEMM) from the committed `ncu --set full` capture of
    the SAME mode and configuration, or None.  profiles/r2_ncu_full_gates_eval_B1024_summary.csv holds the six gate
    launches of one eval iteration at B=1024, M=8, parity precision (tools/ncu_round2.sh gates);
    profiles/r2_ncu_full_gates_train_fwd_B1024_summary.csv the twelve of a training forward's first two iterations
    (tools/ncu_round2.sh train_fwd), of which the last six (t = 1: hidden convolution included, as in 15 of the 16
    iterations) are averaged."""
    if not (B == 1024 and nodes == 8 and precision == 'parity'):
        return None, 'no ncu --set full capture of this mode / configuration under profiles/'
    name = 'r2_ncu_full_gates_{}_B1024_summary.csv'.format('eval' if mode == 'eval' else 'train_fwd')
    path = os.path.join(ROOT, 'profiles', name)
    try:
        import csv
        rows = list(csv.reader(open(path)))
        unit = {'Tbyte': 1e12, 'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}
        cols = {}
        for i, h in enumerate(rows[0]):
            for key in ('dram_rd', 'dram_wr'):
                if h.startswith(key):
                    cols[key] = (i, unit[h[h.index('[') + 1:h.index(']')]])
        per = [sum(float(r[cols[k][0]]) * cols[k][1] for k in ('dram_rd', 'dram_wr')) for r in rows[1:] if 'conv_gemm' in r[0]]
        per = per[-6:]
        return sum(per) / len(per), ('dram__bytes_read.sum + dram__bytes_write.sum, mean over the {} forward gate launches of one '
                                     'iteration, profiles/{}').format(len(per), name)
    except Exception as e:          # the capture is evidence, not a dependency
        return None, 'capture unreadable: {}'.format(e)


def workload_name(mode, channels=CHANNELS, nodes=M_NODES):
    data = 'cifar10_32x32' if channels == 3 else 'mnist_32x32'
    src = 'm{}_random_subsets'.format(nodes) if nodes > 1 else 'single_source'
    return '{}_{}_T16_{}'.format(data, src, 'train_step_fwd_bptt_adam' if mode == 'train' else 'eval_encode_decode')


def workload_config(mode, batch, world, channels=CHANNELS, nodes=M_NODES):
    """what is computed -- shared verbatim by our arm and the reference arm"""
    return {'workload': workload_name(mode, channels, nodes), 'batch_per_gpu': batch, 'global_batch': world * batch,
            'num_node': nodes, 'num_iter': T_ITER, 'code_size': 8, 'image': [channels, 32, 32],
            'l2': 'per-step working set (activations+state, ~1 MB/image) exceeds the 126 MB L2'}


def synth_batch(batch, seed, channels=CHANNELS, nodes=M_NODES):
    import torch
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(batch, channels, 32, 32, generator=g)
    label = torch.randint(0, nodes, (batch,), generator=g)        # "split by random subsets"
    return img, label


# ------------------------------------------------------------------------------------------ CPU (oracle) timing
def cpu_time(mode, steps, warmup, sd=None, with_adam=True):
    """the reference's algorithm (oracle port: torch CPU fp32 ops in the reference's order) on all host cores,
    on CPU_SAMPLE[mode] images per step.  Returns (images/s, seconds per step, cores, description)."""
    import torch
    from oracle import drasic_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    train = mode == 'train'
    nb = CPU_SAMPLE[mode]
    if sd is None:
        sd = O.init_state_dict(CHANNELS, 8, M_NODES, seed=0)
    sd = {k: v.detach().cpu().clone().requires_grad_(train) for k, v in sd.items()}
    opt = torch.optim.Adam(list(sd.values()), lr=1e-3) if train and with_adam else None   # train_model.py:161

    def step(img, label, n_iter=T_ITER):
        if train:                                                    # forward + full BPTT, like train_model.py:113-116
            for v in sd.values():
                v.grad = None
            O.codec_forward(sd, img, label, n_iter, M_NODES, training=True)['loss'].backward()
            if opt is not None:
                opt.step()
        else:
            with torch.no_grad():
                O.codec_forward(sd, img, label, n_iter, M_NODES)
    img, label = synth_batch(8, 1)
    step(img, label, 2)                                              # warm the thread pool / oneDNN primitives
    img, label = synth_batch(nb, 7)
    for _ in range(warmup):
        step(img, label)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(img, label)
    dt = (time.perf_counter() - t0) / steps
    what = 'oracle/drasic_oracle.py codec_forward{}, B={} M=8 T=16 per step, {} step(s) of {:.1f} s'.format(
        ' + backward (stochastic binarizer, full BPTT)' + (' + Adam step' if opt is not None else '') if train else ' (eval)',
        nb, steps, dt)
    return nb / dt, dt, cores, what


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    main_mode = 'train' if args.mode in ('all', 'train') else 'eval'
    v, dt, cores, what = cpu_time(main_mode, args.steps, args.warmup)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'images/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(main_mode, args.batch, args.gpus),
        'cpu_baseline': {'value': v, 'unit': 'images/s', 'cores': cores, 'kind': 'port', 'sample': what},
        'e2e': {'value': v, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    if args.mode == 'all':
        ev, edt, _, ewhat = cpu_time('eval', min(args.steps, 2), 1)
        line['eval'] = {'value': ev, 'unit': 'images/s', 'ms_per_step': edt * 1e3,
                        'config': workload_config('eval', args.batch, args.gpus),
                        'cpu_baseline': {'value': ev, 'unit': 'images/s', 'cores': cores, 'kind': 'port', 'sample': ewhat}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ our arm
class Arm(object):
    """one model + optimizer on this rank's GPU, measured through the public API"""

    def __init__(self, dev, rank, world, channels=CHANNELS, nodes=M_NODES, precision='parity', grad_precision='fp16',
                 shard='batch', workload=None):
        import torch
        import config
        import utils
        self.torch, self.config = torch, config
        self.dev, self.rank, self.world = dev, rank, world
        self.channels, self.nodes, self.precision, self.grad_precision = channels, nodes, precision, grad_precision
        self.workload = workload
        self.by_source = shard == 'source' and world > 1 and nodes >= world
        self._apply_config()
        import models
        torch.manual_seed(0)
        self.model = models.dist_codec().to(dev)
        self.params = [p for p in self.model.parameters()]
        self.opt = None
        self.opt_prof = None      # [] while the per-kernel pass runs: (start, end) events around optimizer.step()

    def _apply_config(self):
        c = self.config
        c.PARAM.update(device='cuda', data_name='CIFAR10' if self.channels == 3 else 'MNIST', precision=self.precision,
                       pack_mode='reference', grad_precision=self.grad_precision, loss_fn='mae', stream_k=STREAM_K,
                       deterministic=DETERMINISTIC)
        c.PARAM['control'] = {'normalization': 'none', 'activation': 'tanh', 'num_iter': str(T_ITER),
                              'code_size': '8', 'num_node': str(self.nodes)}
        import utils
        utils.process_control_name()

    def batch(self, B, seed=1000):
        img_h, label_h = synth_batch(B, seed + self.rank, self.channels, self.nodes)
        if self.by_source:
            # SURVEY 8(e) partitioning for M >= G: this rank holds the samples (and trains the encoders) of its own nodes
            from engine.dist import owned_nodes
            mine = owned_nodes(self.nodes, self.rank, self.world)
            label_h = self.torch.tensor(mine)[label_h % len(mine)]
        return img_h, label_h

    def barrier(self):
        if self.world > 1:
            self.torch.distributed.barrier()
        self.torch.cuda.synchronize()

    def step(self, inp, train, B):
        """one pass of the hot path through the public API"""
        from engine.dist import allreduce_gradients, reduce_source_sharded
        torch = self.torch
        if train:
            if self.opt is None:
                from engine.optim import FusedAdam
                self.opt = FusedAdam(self.params, lr=1e-3)                # train_model.py:161 (optim.Adam)
            self.opt.zero_grad(set_to_none=True)
            out = self.model(inp)                                         # train_model.py:113
            out['loss'].backward()                                        # :115  (backward-through-time kernels)
            if self.by_source:
                reduce_source_sharded(self.model, B, self.world * B, self.rank, self.world)   # joint decoder only
            else:
                allreduce_gradients(self.params, B, self.world * B)       # replaces DataParallel's reduce (:60)
            if self.opt_prof is not None:                                 # per-kernel pass: the optimizer step on its own
                ev = (self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True))
                ev[0].record()
            self.opt.step()                                               # :116
            if self.opt_prof is not None:
                ev[1].record()
                self.opt_prof.append(ev)
        else:
            with torch.no_grad():
                out = self.model(inp)
        return out

    def measure(self, mode, B, steps, warmup, graph=False, fast=False, clocks=False):
        """device-timed steps on resident inputs, a per-kernel pass, and the end-to-end loop on host buffers"""
        torch, config = self.torch, self.config
        self._apply_config()
        train = mode == 'train'
        if self.rank == 0:
            print('[bench] {} B={} M={} grad={} graph={}'.format(mode, B, self.nodes, self.grad_precision, graph), file=sys.stderr, flush=True)
        self.model.train(train)
        img_h, label_h = self.batch(B)
        img, label = img_h.to(self.dev), label_h.to(self.dev)
        config.PARAM['cuda_graph'] = graph
        resident = {'img': img, 'label': label}
        for _ in range(max(warmup, 3)):
            self.step(resident, train, B)
        self.barrier()
        sampler = ClockSampler(self.dev.index) if (clocks and self.rank == 0) else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            self.step(resident, train, B)
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1) / steps
        eng = self.model._engine
        launches = eng.launches + (self.opt.launches if train else 0)
        res = {'clocks': sampler.stop() if sampler else None}

        # ---- per-kernel timing on the launching stream (separate eager pass, CUDA events around every launch)
        config.PARAM['cuda_graph'] = False
        eng.profile = []
        self.opt_prof = []
        for _ in range(2):
            self.step(resident, train, B)
        torch.cuda.synchronize()
        prof = eng.profile_summary()
        opt_ms = sum(a.elapsed_time(b) for a, b in self.opt_prof) / 2 if self.opt_prof else None
        eng.profile = None
        self.opt_prof = None
        config.PARAM['cuda_graph'] = graph

        # ---- end to end through the public API with host buffers (H2D of inputs, D2H of loss + codes)
        pin_img, pin_label = img_h.pin_memory(), label_h.pin_memory()

        def e2e_step():
            out = self.step({'img': pin_img.to(self.dev, non_blocking=True),
                             'label': pin_label.to(self.dev, non_blocking=True)}, train, B)
            return out['loss'].item(), out['code'][-1].nbytes
        for _ in range(2):
            e2e_step()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            e2e_step()
        self.barrier()
        e2e_ms = (time.perf_counter() - t0) / steps * 1e3

        # ---- the same step at 'fast' precision (single fp16 operands + tanh.approx, stated tolerance), for reference
        fast_ms = float('nan')
        if fast and self.precision == 'parity':
            self.precision = 'fast'
            self._apply_config()
            for _ in range(3):
                self.step(resident, train, B)
            self.barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(steps):
                self.step(resident, train, B)
            f1.record()
            self.barrier()
            fast_ms = f0.elapsed_time(f1) / steps
            self.precision = 'parity'
            self._apply_config()
        config.PARAM['cuda_graph'] = False

        t = torch.tensor([ms, e2e_ms, fast_ms], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms, e2e_ms, fast_ms = t.tolist()
        res.update(mode=mode, B=B, steps=steps, warmup=max(warmup, 3), ms=ms, e2e_ms=e2e_ms, fast_ms=fast_ms, prof=prof,
                   launches=launches, graph=graph, opt_ms=opt_ms)
        return res

    def record(self, r):
        """bench record (value / e2e / roofline ...) of one measurement"""
        sust, burst, hbm, how = measured_peaks()
        train, B, world, prof = r['mode'] == 'train', r['B'], self.world, r['prof']
        gate = [(k, v) for k, v in prof.items() if k.split('_')[0] in ('gates', 'dgrad', 'wgrad')]
        g_n = sum(v[0] for _, v in gate)
        g_ms = sum(v[1] for _, v in gate)
        g_fl = sum(v[2] for _, v in gate)
        all_ms = sum(v[1] for v in prof.values())
        achieved = g_fl / (g_ms * 1e-3) / 1e12 if g_ms else 0.0
        fwd_mma = 3 if self.precision == 'parity' else 1
        bwd_mma = 3 if self.grad_precision == 'parity' else 1
        issued = sum(v[2] * (fwd_mma if k.startswith('gates') else bwd_mma) for k, v in gate)
        flop_per_img = FLOP_PER_IMG_ITER[self.channels] * T_ITER * (3 if train else 1)

        def kind_tflops(kk):
            ms = sum(v[1] for k, v in gate if k.startswith(kk))
            return sum(v[2] for k, v in gate if k.startswith(kk)) / max(1e-9, ms) / 1e9
        fwd = kind_tflops('gates')
        traffic, traffic_note = captured_traffic(r['mode'], B, self.nodes, self.precision)
        rec = {
            'value': world * B / r['ms'] * 1e3, 'unit': 'images/s', 'ms_per_step': r['ms'], 'steps': r['steps'],
            'warmup': r['warmup'],
            'dtype': ('f16x2-split forward (fp32-grade, fp32 accumulate)' if self.precision == 'parity'
                      else 'f16 forward (fp32 accumulate)')
                     + ((', f16 backward GEMMs with exact per-tensor power-of-two scaling, fp32 gradients (tolerance 1e-2 relative '
                         'per tensor, tests/test_gpu_train.py)' if self.grad_precision == 'fp16' else
                         ', f16x2-split backward GEMMs, fp32 saved gates, fp32 gradients (tolerance 5e-4 relative per tensor, '
                         'tests/test_gpu_train.py)') if train else ''),
            'config': dict(workload_config(r['mode'], B, world, self.channels, self.nodes),
                           **({'workload': self.workload} if self.workload else {})),
            'impl_config': {'precision': self.precision, 'grad_precision': self.grad_precision if train else None,
                            'cuda_graph': r['graph'], 'stream_k': STREAM_K, 'pdl': PDL_ON, 'deterministic': DETERMINISTIC,
                            'parallelism': (('dp{}: sharded by source (rank r owns nodes r mod N), one NCCL all-reduce of the joint '
                                             "decoder's gradients per step" if self.by_source else
                                             'dp{}: batch sharded, one NCCL all-reduce of gradients per step').format(world) if train
                                            else 'batch-sharded replicas x{} (no collective)'.format(world))},
            'e2e': {'value': world * B / r['e2e_ms'] * 1e3, 'unit': 'images/s',
                    'h2d_bytes_per_step': B * self.channels * 32 * 32 * 4 + B * 8,
                    'd2h_bytes_per_step': T_ITER * B * 128 * 4 + 4,
                    'api': 'models.dist_codec() Model.forward (+ loss.backward(), gradient all-reduce, engine.optim.FusedAdam step when '
                           'training) on pinned host img/label; loss.item() and the emitted codes read back'},
            'gpu_launches': r['launches'] * r['steps'],
            'roofline': {'bound': 'tensor',
                         'kernel': ('conv_gemm_kernel<gates> + conv_gemm_kernel<dgrad> + wgrad_kernel (6 layers x 16 iterations)'
                                    if train else 'conv_gemm_kernel<gates> (6 layers x 16 iterations)'),
                         'achieved': achieved, 'peak': sust, 'unit': 'TFLOP/s', 'frac': achieved / sust,
                         'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained ({}); burst {}'.format(how, burst),
                         'traffic': traffic, 'traffic_note': traffic_note,
                         'launches': g_n, 'avg_launch_ms': g_ms / max(g_n, 1), 'share_of_step': g_ms / max(all_ms, 1e-9),
                         'dominant_kernel': {'name': 'conv_gemm_kernel<gates> (forward gate GEMMs)', 'achieved': fwd,
                                             'frac': fwd / sust, 'tensor_mma_per_algorithmic_mac': fwd_mma,
                                             'issued_mma_frac': fwd * fwd_mma / sust},
                         'per_kind_tflops': {kk: kind_tflops(kk) for kk in ('gates', 'dgrad', 'wgrad')
                                             if any(k.startswith(kk) for k, _ in gate)},
                         'issued_mma_tflops': issued / (g_ms * 1e-3) / 1e12 if g_ms else 0.0},
            'kernels_ms_per_step': {k: v[1] / 2 for k, v in sorted(prof.items())},
            'algorithmic_tflops': world * B / r['ms'] * 1e3 * flop_per_img / 1e12,
        }
        rec['roofline']['issued_mma_frac'] = rec['roofline']['issued_mma_tflops'] / sust
        if r.get('opt_ms') is not None:
            # SURVEY 8(d) reports the optimizer separately: the timed step INCLUDES it (value, e2e); this is its share
            rec['optimizer_ms_per_step'] = r['opt_ms']
            rec['optimizer'] = 'engine.optim.FusedAdam (multi-tensor Adam over the fp32 masters), inside the timed step'
        if r['fast_ms'] == r['fast_ms']:
            rec['fast_precision'] = {'value': world * B / r['fast_ms'] * 1e3, 'unit': 'images/s', 'ms_per_step': r['fast_ms'],
                                     'dtype': 'f16 forward (fp32 accumulate, tanh.approx)' + (', f16 backward GEMMs' if train else ''),
                                     'tolerance': 'code mismatch rate < 2e-3 in the first iterations, PSNR within 0.05 dB per rate '
                                                  '(tests/test_gpu_loop.py::test_fast_precision_stated_tolerance)'}
        if r.get('clocks'):
            rec['clocks'] = r['clocks']
        return rec

    def release(self):
        self.model.release()
        self.opt = None
        self.torch.cuda.empty_cache()


def parity_check(arm, n_img):
    """eval-mode parity of this build against the CPU oracle on n_img images of the bench workload (M=8, T=16)"""
    import torch
    from oracle import drasic_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    arm._apply_config()
    model = arm.model
    sd = {k: v.detach().cpu() for k, v in model.named_parameters()}
    pi, pl = synth_batch(n_img, 11, arm.channels, arm.nodes)
    with torch.no_grad():
        ref = O.codec_forward(sd, pi, pl, T_ITER, arm.nodes)
        was = model.training
        model.train(False)
        out = model({'img': pi.to(arm.dev), 'label': pl.to(arm.dev)})
        model.train(was)
    refc, refz = torch.stack(ref['code_pm1']), torch.stack(ref['z'])
    codes = out['code_pm1'].cpu()
    diff = codes != refc                                          # [T,B,code,h,w]
    # a flip cascades through the later iterations of its image: count the FIRST differing symbol(s) per image
    flips, margins = 0, []
    clean = torch.ones(n_img, dtype=torch.bool)
    for b in range(n_img):
        for t in range(T_ITER):
            bad = diff[t, b]
            if bad.any():
                flips += int(bad.sum())
                margins.append(float(refz[t, b][bad].abs().max()))
                clean[b] = False
                break
    rec = torch.stack(out['img']).cpu()
    refr = torch.stack(ref['img'])
    rec_err = float((rec[:, clean] - refr[:, clean]).abs().max()) if clean.any() else float('nan')
    dps = max(abs(O.psnr(a, pi) - O.psnr(b, pi)) for a, b in zip(rec, refr))
    return {'sample': 'eval, parity precision, B={} M={} T=16 vs oracle/drasic_oracle.py'.format(n_img, arm.nodes),
            'code_symbols': int(refc.numel()), 'code_mismatches': int(diff.sum()),
            'code_flips': flips, 'images_with_a_flip': int((~clean).sum()),
            'flip_margin_max_abs_z': max(margins) if margins else None,
            'min_margin': float(refz.abs().min()),
            'note': 'code_flips = symbols differing at the first differing iteration of an image (later iterations of that image '
                    'cascade and are counted only in code_mismatches); flip_margin = |z_oracle| of those symbols',
            'recon_max_abs_err': rec_err, 'recon_err_over': 'images without a flip', 'psnr_max_abs_diff_db': dps,
            'loss_abs_diff': abs(float(out['loss']) - float(ref['loss']))}


def multi_gpu_parity(arm):
    """all-reduced gradients and loss of a global batch sharded over the ranks against the same
    batch computed on one GPU (every rank computes both; rank 0 reports).  Batch sharding always; sharding by source
    (rank r holds the samples of the nodes j = r mod N, only the joint decoder is all-reduced) when the run uses it."""
    import torch
    from engine.dist import (allreduce_gradients, allreduce_scalar_mean, reduce_source_sharded, shard_range, source_shard,
                             sync_node_parameters)
    config = arm.config
    arm._apply_config()
    world, rank, dev = arm.world, arm.rank, arm.dev
    model = arm.model
    # after source-sharded training steps every rank holds stale copies of the other ranks' encoders: the owners
    # publish theirs first (what a checkpoint or an evaluation of the whole model needs too)
    sync_node_parameters(model)
    Bg, T = 16 * world, 4
    g = torch.Generator().manual_seed(77)
    img = torch.rand(Bg, arm.channels, 32, 32, generator=g)
    label = torch.randint(0, arm.nodes, (Bg,), generator=g)
    noise = torch.rand(T, Bg, 8, 4, 4, generator=g)
    model.train(True)
    config.PARAM['num_iter'] = T
    named = list(model.named_parameters())

    def run(idx):
        model.zero_grad(set_to_none=True)
        out = model({'img': img[idx].to(dev), 'label': label[idx].to(dev)}, noise=noise[:, idx].to(dev))
        out['loss'].backward()
        return out['loss'].detach()
    res = {}
    try:
        lo, hi = shard_range(Bg, rank, world)
        loss = run(torch.arange(lo, hi))
        allreduce_gradients(arm.params, hi - lo, Bg)
        loss_b = float(allreduce_scalar_mean(loss, hi - lo, Bg))
        batch_sharded = [p.grad.clone() for _, p in named]
        source_sharded = None
        if arm.by_source:
            idx = source_shard(label, rank, world)
            loss = run(idx)
            reduce_source_sharded(model, idx.numel(), Bg, rank, world)
            loss_s = float(allreduce_scalar_mean(loss, idx.numel(), Bg))
            source_sharded = [None if p.grad is None else p.grad.clone() for _, p in named]
        full = float(run(torch.arange(Bg)))
        ref = [p.grad for _, p in named]

        def worst(got):
            w = 0.0
            for a, b in zip(got, ref):
                if a is None:                                    # another rank's encoder: not this rank's to update
                    continue
                den = float(b.norm())
                w = max(w, float((a - b).norm()) / den if den > 0 else float(a.abs().max()))
            return w
        res['batch_sharded'] = {'worst_rel': worst(batch_sharded), 'dloss': abs(loss_b - full)}
        if source_sharded is not None:
            res['source_sharded'] = {'worst_rel': worst(source_sharded), 'dloss': abs(loss_s - full),
                                     'tensors_compared': sum(a is not None for a in source_sharded)}
        model.zero_grad(set_to_none=True)
    finally:
        config.PARAM['num_iter'] = T_ITER
    res['worst_rel'] = max(v['worst_rel'] for v in res.values())
    res['dloss'] = max(v['dloss'] for k, v in res.items() if isinstance(v, dict))
    res['sample'] = ('training mode (given binarizer noise), global batch {} over {} ranks, M={} T={}: NCCL all-reduced '
                     'gradients (weights n_local/N_global) vs the whole batch on one GPU'.format(Bg, world, arm.nodes, T))
    res['tolerance'] = ('fp16 backward GEMMs scale each rank\'s gradient planes by its own power of two: 2e-3 relative '
                        'per tensor (tests/test_gpu_multi.py)')
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    import config
    config.init()
    if args.pdl:
        from engine import binding
        binding.lib().drasic_set_pdl(1)
    arm = Arm(dev, rank, world, precision=args.precision, grad_precision=args.grad_precision, shard=args.shard)
    B, K, W = args.batch, args.steps, args.warmup
    main_mode = 'train' if args.mode in ('all', 'train') else 'eval'
    everything = args.mode == 'all'

    parity_init = None
    if world == 1 and not args.no_cpu:
        # on the seeded initial weights (deterministic); once more below on the weights the training steps leave behind
        parity_init = parity_check(arm, PARITY_SAMPLE)
        parity_init['weights'] = 'initial (torch.manual_seed(0))'
        arm.release()
    main = arm.measure(main_mode, B, K, W, graph=args.graph, fast=not args.no_fast, clocks=True)
    line = {'metric': METRIC, 'n_gpus': world, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'data': 'synthetic'}
    line.update(arm.record(main))
    extras = {}
    k2 = max(3, min(K, 10))
    if everything:
        arm.release()
        extras['eval'] = arm.record(arm.measure('eval', B, K, W, fast=not args.no_fast))
        b128 = {}
        for mode in ('train', 'eval'):
            arm.release()
            b128[mode] = arm.record(arm.measure(mode, 128, k2, W, graph=(mode == 'eval')))
        extras['b128'] = b128
        arm.release()
        if args.precision == 'parity':
            gp = Arm(dev, rank, world, precision='parity', grad_precision='parity', shard=args.shard)
            extras['grad_parity'] = gp.record(gp.measure('train', B, max(2, min(K, 5)), W))
            gp.release()
            del gp
    if world > 1 and (everything or args.multi_parity):
        arm.release()
        extras['multi_gpu_parity'] = multi_gpu_parity(arm)
    if world > 1 and everything and B % world == 0:
        # strong scaling: the N=1 global batch spread over the ranks (the line's own value is weak scaling)
        arm.release()
        st = arm.record(arm.measure('train', B // world, k2, W))
        extras['strong_scaling'] = {'global_batch': B, 'batch_per_gpu': B // world, 'value': st['value'], 'unit': 'images/s',
                                    'ms_per_step': st['ms_per_step'], 'e2e': st['e2e'], 'roofline': st['roofline'],
                                    'kernels_ms_per_step': st['kernels_ms_per_step'],
                                    'note': 'same global batch as the N=1 line; efficiency = value / (N x the N=1 value)'}
    if world == 1 and not args.no_cpu:
        arm.release()
        extras['parity'] = parity_init
        if main_mode == 'train':
            extras['parity_after_training'] = parity_check(arm, PARITY_SAMPLE)
            extras['parity_after_training']['weights'] = 'after the training steps of this run'
    if everything:
        arm.release()
        c1 = Arm(dev, rank, world, channels=1, nodes=1, precision=args.precision)
        extras['config1'] = c1.record(c1.measure('eval', 64, k2, W, graph=True))
        if world == 1 and not args.no_cpu:
            extras['config1']['parity'] = parity_check(c1, 64)
        c1.release()
        del c1
        # BASELINE configs[4]: 768x512 images tiled into 32x32 patches, 4096 patches per batch (10.67 images), one source
        c5 = Arm(dev, rank, world, channels=3, nodes=1, precision=args.precision,
                 workload='kodak_768x512_tiled_into_32x32_patches_single_source_T16_eval_encode_decode')
        extras['config5'] = c5.record(c5.measure('eval', 4096, max(2, min(K, 5)), W))
        extras['config5']['kodak_images_per_s'] = extras['config5']['value'] / 384.0
        c5.release()
        del c5
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line.update(extras)
    # ---- CPU baseline (oracle port) on the same bounded sample `--impl reference` times, rank 0 at N=1 only
    if world == 1 and not args.no_cpu:
        sd = {k: v for k, v in arm.model.named_parameters()}
        v, dt, cores, what = cpu_time(main_mode, 1, 0, sd=sd)
        line['cpu_baseline'] = {'value': v, 'unit': 'images/s', 'cores': cores, 'kind': 'port', 'sample': what}
        if everything:
            v, dt, cores, what = cpu_time('eval', 1, 0, sd=sd)
            line['eval']['cpu_baseline'] = {'value': v, 'unit': 'images/s', 'cores': cores, 'kind': 'port', 'sample': what}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=1024, help='images per GPU per step')
    ap.add_argument('--precision', default='parity', choices=['parity', 'fast'])
    ap.add_argument('--mode', default='all', choices=['all', 'train', 'eval'],
                    help='all: the training step (BASELINE configs[1]) as the line\'s value plus eval / b128 / config1 / '
                         'grad_parity / parity records; train: forward + backward-through-time + all-reduce + Adam step only; '
                         'eval: encode+decode only')
    ap.add_argument('--grad-precision', default='fp16', choices=['fp16', 'parity'])
    ap.add_argument('--graph', action='store_true', help='replay one captured CUDA graph per step instead of launching eagerly')
    ap.add_argument('--shard', default='auto', choices=['auto', 'batch', 'source'],
                    help="N>1 training: 'batch' = every rank sees all nodes, all gradients all-reduced; 'source' = rank r "
                         "owns nodes j = r mod N and their samples, only the joint decoder's gradients are all-reduced; "
                         "'auto' = source when there are at least as many nodes as ranks")
    ap.add_argument('--multi-parity', action='store_true', help='N>1: run the in-line gradient parity check in train / eval mode too')
    ap.add_argument('--stream-k', default='auto', choices=['auto', 'on', 'off'])
    ap.add_argument('--pdl', action='store_true', help='experimental: programmatic dependent launches (hangs intermittently in training)')
    ap.add_argument('--deterministic', action='store_true', help='bit-reproducible training step (config.PARAM["deterministic"])')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-fast', action='store_true', help='skip the secondary fast-precision measurement')
    args = ap.parse_args()
    if args.shard == 'auto':
        # with at least as many sources as ranks every rank owns whole nodes (their encoders and samples)
        # and only the joint decoder's gradients are exchanged
        world = int(os.environ.get('WORLD_SIZE', '1'))
        args.shard = 'source' if (world > 1 and M_NODES >= world) else 'batch'
    global STREAM_K, PDL_ON, DETERMINISTIC
    STREAM_K, PDL_ON, DETERMINISTIC = args.stream_k, args.pdl, args.deterministic
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
