#!/usr/bin/env python
"""Condense ncu output into the text summaries committed under profiles/.

  python tools/ncu_summary.py rep  gpurun_out/prof.ncu-rep  > profiles/<name>.txt
  python tools/ncu_summary.py list gpurun_out/launches.csv  > profiles/<name>.txt
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

WANT = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
    'launch__shared_mem_per_block_dynamic', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
]


def rep(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print('# ncu --set full --clock-control none, source: %s' % path)
    for r in rows[2:]:
        print('kernel: %s' % r[hdr.index('Kernel Name')])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print('  %-86s %18s %s' % (w, r[i], units[i]))
        dr, t = float(r[hdr.index('dram__bytes_read.sum')]), float(r[hdr.index('gpu__time_duration.sum')])
        ub, ut = units[hdr.index('dram__bytes_read.sum')], units[hdr.index('gpu__time_duration.sum')]
        scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}[ub] / {'ms': 1e-3, 'us': 1e-6, 's': 1.0, 'ns': 1e-9}[ut]
        print('  %-86s %18.1f GB/s' % ('=> dram read rate', dr / t * scale / 1e9))
        print()


def lst(path):
    agg = OrderedDict()
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rd = csv.DictReader(io.StringIO(''.join(lines)))
    total = 0.0
    for r in rd:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        u = r['Metric Unit']
        v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(u, 1e-6)
        name = r['Kernel Name'].split('(')[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
    print('# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES), source: %s' % path)
    print('%-72s %8s %12s %8s' % ('kernel', 'launches', 'total ms', 'share'))
    for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-72s %8d %12.3f %7.1f%%' % (name[:72], n, ms, 100 * ms / total))


if __name__ == '__main__':
    {'rep': rep, 'list': lst}[sys.argv[1]](sys.argv[2])
