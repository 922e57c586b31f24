#!/bin/bash
# every single-GPU bench line committed under profiles/ (round 2): default workload with the CPU arm, the other workloads, the reference arm
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_default_n1.json 2> gpurun_out/r2_bench_default_n1.err; echo "default rc=$?"
for w in delade_cls_ref bm25 bm25_ref dense delade_cls_zipf; do
  timeout 600 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n1_$w.json 2> gpurun_out/r2_bench_n1_$w.err; echo "$w rc=$?"
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; echo "reference rc=$?"
timeout 900 python bench.py --impl reference --cpu-full --steps 1 --warmup 0 > gpurun_out/r2_bench_reference_full.json 2> gpurun_out/r2_bench_reference_full.err; echo "reference full rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_bench_*n1*.json')) + ['gpurun_out/r2_bench_reference_arm.json', 'gpurun_out/r2_bench_reference_full.json']:
    try:
        d = json.load(open(f))
        print(f.split('/')[-1], round(d['value'], 3), d.get('verified', {}).get('ok'), d.get('roofline', {}).get('frac'))
    except Exception as e:
        print(f, 'ERR', e)
PY
