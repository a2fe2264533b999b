"""one line per kernel launch of an .ncu-rep (ncu -i ... --page raw --csv): the metrics DESIGN.md / profiles quote"""
import csv
import io
import subprocess
import sys

WANT = [('gpu__time_duration.sum', 'us'), ('gpc__cycles_elapsed.avg.per_second', 'GHz'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor_active_%'),
        ('sm__inst_executed_pipe_tensor.sum', 'tensor_inst'), ('sm__inst_executed_pipe_uniform.sum', 'uniform_inst'),
        ('dram__bytes_read.sum', 'dram_rd'), ('dram__bytes_write.sum', 'dram_wr'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_%'),
        ('dram__bytes.sum.per_second', 'dram_rate'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2_%'), ('lts__t_sector_hit_rate.pct', 'l2_hit_%'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_%'), ('launch__registers_per_thread', 'regs'),
        ('launch__grid_size', 'grid'), ('smsp__inst_executed.sum', 'inst')]
for path in sys.argv[1:]:
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = csv.writer(sys.stdout)
    out.writerow(['kernel'] + ['{} [{}]'.format(n, units[idx[m]]) if m in idx else n for m, n in WANT])
    for r in rows[2:]:
        name = r[idx['Kernel Name']]
        out.writerow([name[:70]] + [r[idx[m]] if m in idx else '' for m, _ in WANT])
