mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_default_n1.json 2> gpurun_out/r2_bench_default_n1.err; echo "default rc=$?"
for w in delade_cls_ref bm25 bm25_ref dense delade_cls_zipf; do
  timeout 600 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n1_$w.json 2> gpurun_out/r2_bench_n1_$w.err; echo "$w rc=$?"
done
timeout 300 python bench.py --unmasked --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n1_unmasked_ip.json 2>gpurun_out/r2_bench_n1_unmasked_ip.err; echo "unmasked rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_bench_*n1*.json')):
    try:
        d = json.load(open(f))
        print(f.split('/')[-1], round(d['value'], 3), round(d['e2e']['value'],1), d.get('verified', {}).get('ok'), d.get('roofline', {}).get('bound'), d.get('roofline', {}).get('frac'), d['clocks'].get('reasons'))
    except Exception as e:
        print(f, 'ERR', e)
PY
