"""Benchmark of DRASIC's recurrent residual coding loop on B200 (see BASELINE.json / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--mode eval|train]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path (T=16 iterations of encode -> binarize -> decode over the
residual) over one batch of synthetic CIFAR-shaped images from 8 sources.  Rank 0 prints ONE JSON line.
`--impl reference` times the CPU oracle (the reference's algorithm on torch CPU ops, all host threads)
on a bounded sample of the same workload.
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
FLOP_PER_IMG_FWD = 57.083e9        # SURVEY 8d / BASELINE.md 3: dense forward count at T=16, C=3
METRIC = 'images/s train & encode+decode at T=16 (CIFAR-10 32x32), 1/2/4/8 B200; PSNR parity'


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


def workload_name(mode):
    return ('cifar10_32x32_m8_random_subsets_T16_train_step_fwd_bptt_adam' if mode == 'train'
            else 'cifar10_32x32_m8_random_subsets_T16_eval_encode_decode')


def synth_batch(batch, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(batch, CHANNELS, 32, 32, generator=g)
    label = torch.randint(0, M_NODES, (batch,), generator=g)        # "split by random subsets"
    return img, label


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """the reference's algorithm (oracle port: torch CPU fp32 ops in the reference's order) on host cores"""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    from oracle import drasic_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    train = args.mode == 'train'
    sd = O.init_state_dict(CHANNELS, 8, M_NODES, seed=0)
    opt = None
    if train:
        sd = {k: v.requires_grad_(True) for k, v in sd.items()}
        opt = torch.optim.Adam(list(sd.values()), lr=1e-3)           # train_model.py:161

    def step(img, label, n_iter=T_ITER):
        if train:                                                    # forward + full BPTT, like train_model.py:113-115
            for v in sd.values():
                v.grad = None
            out = O.codec_forward(sd, img, label, n_iter, M_NODES, training=True)
            out['loss'].backward()
            opt.step()                                               # train_model.py:116
        else:
            with torch.no_grad():
                O.codec_forward(sd, img, label, n_iter, M_NODES)
    # size the per-step sample so the whole run stays within a few minutes
    img, label = synth_batch(8, 1)
    step(img, label, 2)                                              # warm the thread pool / oneDNN primitives
    t0 = time.perf_counter()
    step(img, label)
    t8 = time.perf_counter() - t0
    budget = 150.0 / max(1, args.steps + args.warmup)
    sample = int(max(8, min(64, 8 * budget / max(t8, 1e-3))) // 8 * 8)
    img, label = synth_batch(sample, 2)
    for _ in range(args.warmup):
        step(img, label)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(img, label)
    dt = (time.perf_counter() - t0) / args.steps
    v = sample / dt
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'images/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args.mode), 'sample_batch': sample,
                   'num_node': M_NODES, 'num_iter': T_ITER, 'code_size': 8},
        'cpu_baseline': {'value': v, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                         'sample': 'oracle/drasic_oracle.py codec_forward{}, B={} M=8 T=16 per step'.format(
                             ' + backward (stochastic binarizer, full BPTT) + Adam step' if train else ' (eval)', sample)},
        'e2e': {'value': v, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ our arm
def prof_gate_tflops(gate, fwd_mma_per_mac):
    """tensor-core work actually issued: forward gate GEMMs count fwd_mma_per_mac MMAs per algorithmic MAC"""
    fl = sum(v[2] * (fwd_mma_per_mac if k.startswith('gates') else 1) for k, v in gate)
    ms = sum(v[1] for _, v in gate)
    return fl / (ms * 1e-3) / 1e12 if ms else 0.0


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
    from oracle import drasic_oracle as O          # used only by the cpu_baseline leg at the end
    import config
    config.init()
    config.PARAM.update(device='cuda', data_name='CIFAR10', precision=args.precision, pack_mode='reference',
                        grad_precision=args.grad_precision)
    config.PARAM['control'] = {'normalization': 'none', 'activation': 'tanh', 'num_iter': str(T_ITER),
                               'code_size': '8', 'num_node': str(M_NODES)}
    import utils
    utils.process_control_name()
    import models
    torch.manual_seed(0)
    model = models.dist_codec().to(dev)
    train = args.mode == 'train'
    model.train(train)
    B = args.batch
    # each rank codes its own shard of the global batch (samples are independent: no data-path collective)
    img_h, label_h = synth_batch(B, 1000 + rank)
    from engine.dist import allreduce_gradients, owned_nodes, reduce_source_sharded
    by_source = args.shard == 'source' and world > 1
    if by_source:
        # SURVEY 8(e) partitioning for M >= G: this rank holds the samples (and trains the encoders) of its own nodes
        mine = owned_nodes(M_NODES, rank, world)
        if not mine:
            raise SystemExit('--shard source needs at least one node per rank ({} nodes, {} ranks)'.format(M_NODES, world))
        label_h = torch.tensor(mine)[label_h % len(mine)]
    img, label = img_h.to(dev), label_h.to(dev)
    params = [p for p in model.parameters()]
    from engine.optim import FusedAdam
    opt = FusedAdam(params, lr=1e-3) if train else None               # train_model.py:161 (optim.Adam)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(inp):
        """one pass of the hot path through the public API"""
        if train:
            opt.zero_grad(set_to_none=True)
            out = model(inp)                                          # train_model.py:113
            out['loss'].backward()                                    # :115  (backward-through-time kernels)
            if by_source:
                reduce_source_sharded(model, B, world * B, rank, world)   # joint decoder only
            else:
                allreduce_gradients(params, B, world * B)             # replaces DataParallel's reduce (:60)
            opt.step()                                                # :116
        else:
            with torch.no_grad():
                out = model(inp)
        return out

    config.PARAM['cuda_graph'] = args.graph
    resident = {'img': img, 'label': label}
    for _ in range(max(args.warmup, 3)):
        step(resident)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(resident)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    eng = model._engine
    launches = eng.launches + (opt.launches if train else 0)
    clocks = sampler.stop() if sampler else None

    # ---- per-kernel timing on the launching stream (separate eager pass, CUDA events around every launch)
    config.PARAM['cuda_graph'] = False
    eng.profile = []
    for _ in range(2):
        step(resident)
    torch.cuda.synchronize()
    prof = eng.profile_summary()
    eng.profile = None
    config.PARAM['cuda_graph'] = args.graph

    # ---- end to end through the public API with host buffers (H2D of inputs, D2H of loss + codes)
    pin_img, pin_label = img_h.pin_memory(), label_h.pin_memory()

    def e2e_step():
        out = step({'img': pin_img.to(dev, non_blocking=True), 'label': pin_label.to(dev, non_blocking=True)})
        return out['loss'].item(), out['code'][-1].nbytes
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps

    # ---- the same step at 'fast' precision (single fp16 operands + tanh.approx, stated tolerance), for reference
    fast_ms = float('nan')
    if args.precision == 'parity' and not args.no_fast:
        config.PARAM['precision'] = 'fast'
        for _ in range(3):
            step(resident)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            step(resident)
        f1.record()
        barrier()
        fast_ms = f0.elapsed_time(f1) / args.steps
        config.PARAM['precision'] = args.precision

    t_ms = torch.tensor([ms, e2e_s * 1e3, fast_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms, e2e_ms, fast_ms = t_ms.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    sust, burst, hbm, how = measured_peaks()
    gate = [(k, v) for k, v in prof.items() if k.split('_')[0] in ('gates', 'dgrad', 'wgrad')]
    g_n = sum(v[0] for _, v in gate)
    g_ms = sum(v[1] for _, v in gate)
    g_fl = sum(v[2] for _, v in gate)
    all_ms = sum(v[1] for v in prof.values())
    achieved = g_fl / (g_ms * 1e-3) / 1e12
    mma_per_mac = 3 if args.precision == 'parity' else 1
    flop_per_img = FLOP_PER_IMG_FWD * (3 if train else 1)
    # DRAM bytes per gate-kernel launch (dram__bytes_read.sum + dram__bytes_write.sum) from the ncu --set full capture
    # committed as profiles/r1_ncu_full_gates_parity_final_summary.csv (eval forward, 1024 slots): one iteration t >= 1 moves
    # E1 795 + E2 400 + E3 155 + D1 315 + D2 1192 + D3 763 MB = 3620 MB over 6 launches
    traffic, traffic_note = None, 'no ncu capture for this configuration'
    if args.precision == 'parity' and B == 1024:
        traffic = 3620e6 / 6
        traffic_note = ('forward gate kernels only, average per launch, from profiles/r1_ncu_full_gates_parity_final_summary.csv; '
                        'compulsory bytes of the same six launches: 3.1 GB (x, h hi+lo in; c in/out; h, consumer copy out; weights) -> 1.17x')
    sd = {k: v for k, v in model.named_parameters()}
    line = {
        'metric': METRIC, 'value': world * B / ms * 1e3, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None,
        'dtype': ('f16x2-split forward (fp32-grade, fp32 accumulate)' if args.precision == 'parity' else 'f16 forward (fp32 accumulate)')
                 + (', f16 backward GEMMs with exact per-tensor power-of-two scaling, fp32 gradients' if train else ''),
        'data': 'synthetic',
        'config': {'workload': workload_name(args.mode), 'batch_per_gpu': B,
                   'global_batch': world * B, 'num_node': M_NODES, 'num_iter': T_ITER, 'code_size': 8,
                   'precision': args.precision, 'cuda_graph': args.graph,
                   'cta_pair': 'cta_group::2 pairs on every gate / dgrad / wgrad GEMM (pair units formed inside each node)',
                   'parallelism': (('dp{}: sharded by source (rank r owns nodes r mod N), one NCCL all-reduce of the joint '
                                    "decoder's gradients per step" if by_source else
                                    'dp{}: batch sharded, one NCCL all-reduce of gradients per step').format(world) if train
                                   else 'batch-sharded replicas x{} (no collective)'.format(world)),
                   'l2': 'per-step working set (activations+state, ~1 MB/image) exceeds the 126 MB L2'},
        'e2e': {'value': world * B / e2e_ms * 1e3, 'unit': 'images/s',
                'h2d_bytes_per_step': B * CHANNELS * 32 * 32 * 4 + B * 8,
                'd2h_bytes_per_step': T_ITER * B * 128 * 4 + 4 + B * 8,
                'api': 'models.dist_codec() Model.forward (+ loss.backward(), engine.optim.FusedAdam step when training) on pinned host '
                       'img/label; loss.item() and the emitted codes read back'},
        'gpu_launches': launches * args.steps,
        'clocks': clocks,
        'roofline': {'bound': 'tensor', 'kernel': 'conv_gemm_kernel<gates> + conv_gemm_kernel<dgrad> + wgrad_kernel (6 layers x 16 iterations)'
                     if train else 'conv_gemm_kernel<gates> (6 layers x 16 iterations)',
                     'achieved': achieved, 'peak': sust, 'unit': 'TFLOP/s', 'frac': achieved / sust,
                     'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained ({}); burst {}'.format(how, burst),
                     'traffic': traffic, 'traffic_note': traffic_note, 'launches': g_n, 'avg_launch_ms': g_ms / max(g_n, 1),
                     'share_of_step': g_ms / all_ms, 'tensor_mma_per_algorithmic_mac': mma_per_mac,
                     'per_kind_tflops': {kk: sum(v[2] for k, v in gate if k.startswith(kk)) / max(1e-9, sum(v[1] for k, v in gate if k.startswith(kk))) / 1e9
                                         for kk in ('gates', 'dgrad', 'wgrad') if any(k.startswith(kk) for k, _ in gate)},
                     'forward_mma_per_algorithmic_mac': mma_per_mac},
        'kernels_ms_per_step': {k: v[1] / 2 for k, v in sorted(prof.items())},
        'algorithmic_tflops': world * B / ms * 1e3 * flop_per_img / 1e12,
    }
    line['roofline']['issued_mma_tflops'] = (prof_gate_tflops(gate, mma_per_mac))
    line['roofline']['issued_mma_frac'] = line['roofline']['issued_mma_tflops'] / sust
    if fast_ms == fast_ms:
        line['fast_precision'] = {'value': world * B / fast_ms * 1e3, 'unit': 'images/s', 'ms_per_step': fast_ms,
                                  'dtype': 'f16 forward (fp32 accumulate, tanh.approx)' + (', f16 backward GEMMs' if train else ''),
                                  'tolerance': 'code mismatch rate < 2e-3 in the first iterations, PSNR within 0.05 dB per rate '
                                               '(tests/test_gpu_loop.py::test_fast_precision_stated_tolerance)'}
    # ---- CPU baseline (oracle port) on a bounded sample, rank 0 at N=1 only
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        nb = 96 if train else 512                                   # ~10-15 s of CPU work on 16 cores
        sdc = {k: v.detach().cpu().clone().requires_grad_(train) for k, v in sd.items()}
        ci, cl = synth_batch(nb, 7)
        with torch.no_grad():
            O.codec_forward(sdc, ci[:8], cl[:8], 2, M_NODES)
        t0 = time.perf_counter()
        if train:
            O.codec_forward(sdc, ci, cl, T_ITER, M_NODES, training=True)['loss'].backward()
        else:
            with torch.no_grad():
                O.codec_forward(sdc, ci, cl, T_ITER, M_NODES)
        dt = time.perf_counter() - t0
        line['cpu_baseline'] = {'value': nb / dt, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                                'sample': 'oracle codec_forward{} B={} M=8 T=16, one pass ({:.1f} s)'.format(
                                    ' + backward (BPTT)' if train else ' eval', nb, dt)}
    # ---- parity of this build against the CPU oracle on a small sample of the same workload (eval mode, T=16)
    if world == 1 and not args.no_cpu:
        pi, pl = synth_batch(16, 11)
        sdc = {k: v.detach().cpu() for k, v in sd.items()}
        with torch.no_grad():
            ref = O.codec_forward(sdc, pi, pl, T_ITER, M_NODES)
            was = model.training
            model.train(False)
            prec = config.PARAM['precision']
            config.PARAM['precision'] = 'parity'
            out = model({'img': pi.to(dev), 'label': pl.to(dev)})
            config.PARAM['precision'] = prec
            model.train(was)
        refc = torch.stack(ref['code_pm1'])
        mism = int((out['code_pm1'].cpu() != refc).sum())
        rec_err = float((torch.stack(out['img']).cpu() - torch.stack(ref['img'])).abs().max())
        dps = max(abs(O.psnr(a.cpu(), pi) - O.psnr(b, pi)) for a, b in zip(out['img'], ref['img']))
        line['parity'] = {'sample': 'eval, parity precision, B=16 M=8 T=16 vs oracle/drasic_oracle.py', 'code_mismatches': mism,
                          'code_symbols': int(refc.numel()), 'recon_max_abs_err': rec_err, 'psnr_max_abs_diff_db': dps}
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
    ap.add_argument('--mode', default='train', choices=['train', 'eval'],
                    help='train: forward + backward-through-time + all-reduce + Adam step (BASELINE configs[1]); '
                         'eval: encode+decode only')
    ap.add_argument('--grad-precision', default='fp16', choices=['fp16', 'parity'])
    ap.add_argument('--graph', action='store_true',
                    help='replay one captured CUDA graph per step instead of launching eagerly (measured slower: the '
                         'loop is GPU-bound at >100 us per launch and a static graph needs worst-case node padding)')
    ap.add_argument('--shard', default='batch', choices=['batch', 'source'],
                    help="N>1 training: 'batch' = every rank sees all nodes, all gradients all-reduced; 'source' = rank r "
                         "owns nodes j = r mod N and their samples, only the joint decoder's gradients are all-reduced")
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-fast', action='store_true', help='skip the secondary fast-precision measurement')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
