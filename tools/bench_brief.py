"""print the numbers of a bench.py JSON line (and its nested records) in a few lines"""
import json
import sys


def brief(r, name):
    rf = r['roofline']
    print('{:<12} {:8.0f} img/s  e2e {:8.0f}  {:7.2f} ms  frac {:.3f}  gates {:.3f}  {}  graph={} fast={}'.format(
        name, r['value'], r['e2e']['value'], r['ms_per_step'], rf['frac'], rf['dominant_kernel']['frac'],
        {k: round(v) for k, v in rf['per_kind_tflops'].items()}, r['impl_config']['cuda_graph'],
        round(r.get('fast_precision', {}).get('value', 0))))


for path in sys.argv[1:]:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    print(path, d['config']['workload'], 'n_gpus', d['n_gpus'])
    brief(d, 'main')
    for k in ('eval', 'grad_parity', 'config1', 'config5'):
        if k in d:
            brief(d[k], k)
    for k, v in d.get('b128', {}).items():
        brief(v, 'b128 ' + k)
    print('   kernels', {k: round(v, 2) for k, v in d['kernels_ms_per_step'].items()})
    for k in ('parity', 'parity_after_training', 'multi_gpu_parity', 'cpu_baseline', 'clocks'):
        if k in d:
            print('  ', k, {a: b for a, b in d[k].items() if a not in ('note', 'sample', 'tolerance')})
