"""aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list: per (kernel, grid) launches, mean and total us"""
import csv
import re
import sys
from collections import OrderedDict

SEQ = '--seq' in sys.argv
for path in [a for a in sys.argv[1:] if a != '--seq']:
    rows = []
    with open(path, newline='') as f:
        lines = [ln for ln in f if not ln.startswith('==')]
    rd = csv.DictReader(lines)
    agg = OrderedDict()
    for r in rd:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'\(.*', '', r['Kernel Name'])
        name = re.sub(r'^.*::', '', name)
        key = (name, r.get('Grid Size', ''))
        v = float(r['Metric Value'].replace(',', ''))
        unit = r.get('Metric Unit', 'ns')
        us = v / 1e3 if unit in ('ns', 'nsecond') else (v if unit in ('us', 'usecond') else v * 1e3)
        if SEQ:
            rows.append('{} {} {:.1f}'.format(name[:40], r.get('Grid Size', ''), us))
        n, tot = agg.get(key, (0, 0.0))
        agg[key] = (n + 1, tot + us)
    print(path)
    if SEQ:
        print('  ' + '\n  '.join(rows))
        continue
    total = sum(t for _, t in agg.values())
    for (name, grid), (n, tot) in agg.items():
        print('  {:<60} grid {:<14} n {:4d}  mean {:8.1f} us  total {:9.1f} us  {:5.1f} %'.format(name[:60], grid, n, tot / n, tot, 100 * tot / total))
    print('  total {:.1f} us'.format(total))

# sequence view: python tools/ncu_launches.py --seq file.csv  (one line per launch, in launch order)
