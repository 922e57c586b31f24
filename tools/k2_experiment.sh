#!/bin/bash
# K2 bottleneck experiment: kernel durations under ncu's launch list with the DHR_K2_DBG bits (not a bench).
mkdir -p gpurun_out
for d in 8 9 12 13; do
  DHR_K2_DBG=$d timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dense_tile -s 20 -c 40 --csv --log-file gpurun_out/k2dbg_$d.csv \
    python bench.py --rows 2000000 --queries 512 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/k2dbg_$d.log 2>&1
  python - <<PY
import csv
v=[float(r['Metric Value'].replace(',','')) for r in csv.DictReader(l for l in open('gpurun_out/k2dbg_$d.csv') if not l.startswith('==')) if r.get('Metric Name')=='gpu__time_duration.sum']
print('dbg=$d', 'n=%d'%len(v), 'median us = %.1f' % (sorted(v)[len(v)//2]/1000 if v else -1))
PY
done
