#!/usr/bin/env python
"""profiles/traffic.json + profiles/r2_ncu_<workload>.txt from the ncu --set full captures made by tools/ncu_workloads.sh.

For every workload the captured launches are one steady-state sub-chunk of the tile path (37,888 rows x 256 queries in flight):
K2 (dense_tile_ts*) and / or K1t (lex_tile_kernel).  Recorded per workload: DRAM bytes (read + write) per launch of each kernel
and per sub-chunk, the issue-slot / shared-memory-wavefront / tensor-pipe / DRAM utilisation ncu reports, and what a launch covers,
so that bench.py can turn it into a physical DRAM rate for the timed run."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)
from tools.ncu_summary import WANT  # noqa: E402

UNIT = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}
TIME = {'ms': 1e-3, 'us': 1e-6, 's': 1.0, 'ns': 1e-9, 'usecond': 1e-6, 'msecond': 1e-3, 'nsecond': 1e-9, 'second': 1.0}
EXTRA = ['l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
         'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__thread_inst_executed_per_inst_executed.pct',
         'smsp__thread_inst_executed_pred_on_per_inst_executed.ratio', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg',
         'smsp__inst_executed.sum', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']


def load(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    traffic = {'_comment': 'per workload: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch of the scan kernels of one '
                           'steady-state sub-chunk of the tile path at full size (rows_per_launch rows x queries_in_flight queries; last chunk, tight threshold), from ncu --set full '
                           '(tools/ncu_workloads.sh, summaries in profiles/r2_ncu_<workload>.txt); utilisation figures are those of the dominant kernel'}
    for w in ['delade_cls', 'delade_cls_ref', 'bm25', 'bm25_ref', 'dense', 'delade_cls_zipf']:
        rep = os.path.join(ROOT, 'gpurun_out', 'r2_ncu_%s.ncu-rep' % w)
        if not os.path.exists(rep):
            continue
        hdr, units, rows = load(rep)
        col = {h: i for i, h in enumerate(hdr)}

        def val(r, name, scale=None):
            if name not in col or r[col[name]] in ('', 'n/a'):
                return None
            v = float(r[col[name]].replace(',', ''))
            u = units[col[name]]
            if scale == 'bytes':
                return v * UNIT.get(u, 1.0)
            if scale == 'time':
                return v * TIME.get(u, 1.0)
            return v
        kernels = {}
        lines = ['# ncu --set full --clock-control none --import-source on, steady-state launches of the last chunk of `python bench.py --workload %s --queries 256` '
                 '(8,841,823 rows; 37,888 rows x 256 queries per tile-path launch, the 7.78 M-row last chunk for the dense-only index); source: %s' % (w, os.path.relpath(rep, ROOT))]
        for r in rows:
            name = r[col['Kernel Name']].split('(')[0].replace('void ', '').replace('dhr::', '')
            kind = 'K1t' if 'lex_tile' in name else 'K2'
            lines.append('kernel: %s' % r[col['Kernel Name']])
            for m in WANT + EXTRA:
                if m in col:
                    lines.append('  %-86s %18s %s' % (m, r[col[m]], units[col[m]]))
            d = (val(r, 'dram__bytes_read.sum', 'bytes') or 0) + (val(r, 'dram__bytes_write.sum', 'bytes') or 0)
            t = val(r, 'gpu__time_duration.sum', 'time')
            lines.append('  %-86s %18.1f GB/s' % ('=> dram read+write rate', d / t / 1e9))
            lines.append('')
            k = kernels.setdefault(kind, {'kernel': name, 'launches': 0, 'dram_bytes': 0.0, 'time_us': 0.0, 'issue_active_pct': [], 'tensor_pipe_pct': [],
                                          'dram_pct': [], 'lsu_wavefronts_shared': 0.0, 'pred_on_threads_per_inst': []})
            k['launches'] += 1
            k['dram_bytes'] += d
            k['time_us'] += t * 1e6
            for key, m in (('issue_active_pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
                           ('tensor_pipe_pct', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
                           ('dram_pct', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
                           ('pred_on_threads_per_inst', 'smsp__thread_inst_executed_pred_on_per_inst_executed.ratio')):
                x = val(r, m)
                if x is not None:
                    k[key].append(x)
            k['lsu_wavefronts_shared'] += val(r, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum') or 0.0
        out = {'rows_per_launch': 37888 if w != 'dense' else 8841823 - 1062656, 'queries_in_flight': 256, 'source': 'profiles/r2_ncu_%s.txt' % w, 'kernels': {}}
        total = 0.0
        for kind, k in kernels.items():
            n = k['launches']
            per = {'kernel': k['kernel'], 'dram_bytes_per_launch': k['dram_bytes'] / n, 'time_us_per_launch': k['time_us'] / n}
            for key in ('issue_active_pct', 'tensor_pipe_pct', 'dram_pct', 'pred_on_threads_per_inst'):
                if k[key]:
                    per[key] = sum(k[key]) / len(k[key])
            if k['lsu_wavefronts_shared']:
                # one shared-memory wavefront per cycle and SM is the pipe's peak (148 SMs, 1.965 GHz)
                per['lsu_shared_wavefront_pct'] = 100.0 * (k['lsu_wavefronts_shared'] / n) / (148 * 1.965e3 * per['time_us_per_launch'])
            out['kernels'][kind] = per
            total += per['dram_bytes_per_launch']
        dom = 'K1t' if 'K1t' in out['kernels'] else 'K2'
        out['dram_bytes_per_launch'] = total
        out['launch_kind'] = 'sum over one launch of each scan kernel of a sub-chunk (%s)' % ' + '.join(sorted(out['kernels']))
        out['dominant'] = dom
        for key in ('issue_active_pct', 'tensor_pipe_pct', 'dram_pct', 'lsu_shared_wavefront_pct', 'pred_on_threads_per_inst'):
            if key in out['kernels'][dom]:
                out[key] = out['kernels'][dom][key]
        traffic[w] = out
        with open(os.path.join(ROOT, 'profiles', 'r2_ncu_%s.txt' % w), 'w') as f:
            f.write('\n'.join(lines) + '\n')
        print(w, json.dumps(out)[:300])
    with open(os.path.join(ROOT, 'profiles', 'traffic.json'), 'w') as f:
        json.dump(traffic, f, indent=1)


if __name__ == '__main__':
    main()
